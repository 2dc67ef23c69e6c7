// Band indices and small order statistics shared by the preprocessing kernels.
#pragma once
#include <cuda_runtime.h>
// ---- band indices, exactly the reference's float32 operation order ---------------------
__device__ __forceinline__ float clip01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

// All arithmetic uses the explicit round-to-nearest intrinsics so nvcc cannot contract
// mul+add into FMA: NumPy evaluates each float32 operation separately.
__device__ __forceinline__ float idx_evi(float b2, float b3, float b4, float b8) {
  // src/preprocessing/indices.py:15-27 : 2.5 * ((NIR-RED) / (NIR + 6*RED - 7.5*BLUE + 1))
  float BLUE = clip01(b2), RED = clip01(b4), NIR = clip01(b8);
  (void)b3;
  float den = __fadd_rn(__fsub_rn(__fadd_rn(NIR, __fmul_rn(6.f, RED)), __fmul_rn(7.5f, BLUE)), 1.f);
  float e = __fmul_rn(2.5f, __fdiv_rn(__fsub_rn(NIR, RED), den));
  return fminf(fmaxf(e, -1.5f), 1.5f);
}
__device__ __forceinline__ float idx_bi(float b2, float b4, float b8, float b11) {
  // src/preprocessing/indices.py:47-54
  float B11 = clip01(b11), B4 = clip01(b4), B8 = clip01(b8), B2 = clip01(b2);
  float p = __fadd_rn(B11, B4), q = __fadd_rn(B8, B2);
  float v = __fdiv_rn(__fsub_rn(p, q), __fadd_rn(__fadd_rn(p, q), 1e-5f));
  return fminf(fmaxf(v, -1.f), 1.f);
}
__device__ __forceinline__ float idx_msavi2(float b4, float b8) {
  // src/preprocessing/indices.py:30-44
  float RED = clip01(b4), NIR = clip01(b8);
  float t = __fadd_rn(__fmul_rn(2.f, NIR), 1.f);
  float s = __fsub_rn(__fmul_rn(t, t), __fmul_rn(8.f, __fsub_rn(NIR, RED)));
  if (s < 0.f) s = 0.f;
  float m = __fdiv_rn(__fsub_rn(t, __fsqrt_rn(s)), 2.f);
  return fminf(fmaxf(m, -1.f), 1.f);
}
__device__ __forceinline__ float idx_grndvi(float b3, float b4, float b8) {
  // src/preprocessing/indices.py:4-12
  float nir = clip01(b8), green = clip01(b3), red = clip01(b4);
  float gr = __fadd_rn(green, red);
  float den = __fadd_rn(__fadd_rn(nir, gr), 1e-5f);
  return __fdiv_rn(__fsub_rn(nir, gr), den);
}

__device__ __forceinline__ float med3(float a, float b, float c) {
  return fmaxf(fminf(a, b), fminf(fmaxf(a, b), c));
}

template <int NMAX>
__device__ __forceinline__ float median_n(float* v, int n) {
  // insertion sort in registers/local; np.median: even n -> mean of the two middle values
  for (int i = 1; i < n; ++i) {
    float x = v[i]; int j = i - 1;
    while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; }
    v[j + 1] = x;
  }
  if (n & 1) return v[n >> 1];
  return __fmul_rn(__fadd_rn(v[(n >> 1) - 1], v[n >> 1]), 0.5f);
}

__device__ __forceinline__ float median12(const float* in) {
  float v[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) v[i] = in[i];
  // fully unrolled odd-even transposition sort (12 passes), branch-free
#pragma unroll
  for (int pass = 0; pass < 12; ++pass) {
#pragma unroll
    for (int i = (pass & 1); i + 1 < 12; i += 2) {
      float a = v[i], b = v[i + 1];
      v[i] = fminf(a, b); v[i + 1] = fmaxf(a, b);
    }
  }
  return __fmul_rn(__fadd_rn(v[5], v[6]), 0.5f);
}


// Median of 12 by a pruned Batcher network: 34 compare-exchanges of which only the halves that reach ranks 5 and 6
// survive dead-code elimination (60 FMNMX instead of the 132 of the transposition sort above).  Verified against
// sorted() on all 4096 zero-one inputs (0-1 principle; tools/exp/median12_network.py).  Same value as median12().
__device__ __forceinline__ float median12_net(const float* in) {
  float v[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) v[i] = in[i];
#define STC_CE(a, b) { const float lo_ = fminf(v[a], v[b]), hi_ = fmaxf(v[a], v[b]); v[a] = lo_; v[b] = hi_; }
  STC_CE(0, 1) STC_CE(2, 3) STC_CE(0, 2) STC_CE(1, 3) STC_CE(1, 2) STC_CE(4, 5) STC_CE(6, 7) STC_CE(4, 6) STC_CE(5, 7)
  STC_CE(5, 6) STC_CE(0, 4) STC_CE(2, 6) STC_CE(2, 4) STC_CE(1, 5) STC_CE(3, 7) STC_CE(3, 5) STC_CE(1, 2) STC_CE(3, 4)
  STC_CE(5, 6) STC_CE(8, 9) STC_CE(10, 11) STC_CE(8, 10) STC_CE(9, 11) STC_CE(9, 10) STC_CE(0, 8) STC_CE(4, 8)
  STC_CE(2, 10) STC_CE(6, 10) STC_CE(6, 8) STC_CE(1, 9) STC_CE(5, 9) STC_CE(3, 11) STC_CE(3, 5) STC_CE(5, 6)
#undef STC_CE
  return __fmul_rn(__fadd_rn(v[5], v[6]), 0.5f);
}

// Index variants for the fused front end only: approximate division (<= 2 ulp).  Their results are clipped, normalised
// and rounded to fp16 before the model sees them, which hides the difference; the standalone kernels that return
// float32 indices (make_indices) keep the exact forms above.
__device__ __forceinline__ float idx_evi_fast(float b2, float b4, float b8) {
  float BLUE = clip01(b2), RED = clip01(b4), NIR = clip01(b8);
  float den = __fadd_rn(__fsub_rn(__fadd_rn(NIR, __fmul_rn(6.f, RED)), __fmul_rn(7.5f, BLUE)), 1.f);
  float e = __fmul_rn(2.5f, __fdividef(__fsub_rn(NIR, RED), den));
  return fminf(fmaxf(e, -1.5f), 1.5f);
}
__device__ __forceinline__ float idx_bi_fast(float b2, float b4, float b8, float b11) {
  float B11 = clip01(b11), B4 = clip01(b4), B8 = clip01(b8), B2 = clip01(b2);
  float p = __fadd_rn(B11, B4), q = __fadd_rn(B8, B2);
  float v = __fdividef(__fsub_rn(p, q), __fadd_rn(__fadd_rn(p, q), 1e-5f));
  return fminf(fmaxf(v, -1.f), 1.f);
}
__device__ __forceinline__ float idx_msavi2_fast(float b4, float b8) {
  float RED = clip01(b4), NIR = clip01(b8);
  float t = __fadd_rn(__fmul_rn(2.f, NIR), 1.f);
  float s = __fsub_rn(__fmul_rn(t, t), __fmul_rn(8.f, __fsub_rn(NIR, RED)));
  if (s < 0.f) s = 0.f;
  float m = __fmul_rn(__fsub_rn(t, __fsqrt_rn(s)), 0.5f);
  return fminf(fmaxf(m, -1.f), 1.f);
}
__device__ __forceinline__ float idx_grndvi_fast(float b3, float b4, float b8) {
  float nir = clip01(b8), green = clip01(b3), red = clip01(b4);
  float gr = __fadd_rn(green, red);
  return __fdividef(__fsub_rn(nir, gr), __fadd_rn(__fadd_rn(nir, gr), 1e-5f));
}
