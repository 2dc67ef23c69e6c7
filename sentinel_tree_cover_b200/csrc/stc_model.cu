// ConvGRU + U-Net forward of predict_graph-<H>.pb on the chunk-major fp16 layout.
// Reference call: src/download_and_predict_job.py:353-357 (sess.run(predict_logits, ...));
// architecture: pb:down_16/bidirectional_rnn/*, pb:conv_median ... pb:conv2d/Sigmoid
// (SURVEY.md section 8a rows M1-M5).  Convolutions run in stc_conv.cu; this file holds
// the HBM-bound elementwise stages (input packing, GroupNorm apply + gating, sSE,
// pooling / upsampling / concat placement, head) and the launch sequence.
#include "stc_common.cuh"
#include "stc_indices.cuh"
#include <cstring>
#include <cmath>

static constexpr float GN_EPS = 1e-5f;

// --------------------------------------------------------------------------------------
// device helpers
// --------------------------------------------------------------------------------------
__device__ __forceinline__ float sigm(float v) { return 1.0f / (1.0f + expf(-v)); }
// MUFU-based variants for the HBM-bound GRU gating kernels (32-64 transcendentals per pixel made them
// issue-bound): __expf / __fdividef are accurate to ~2 ulp on the ranges GroupNorm produces, far inside the
// fp16 activation rounding that follows.
__device__ __forceinline__ float sigm_fast(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
__device__ __forceinline__ float tanh_fast(float v) {
  const float e = __expf(-2.0f * fabsf(v));                  // in (0, 1]: no overflow
  return copysignf(__fdividef(1.0f - e, 1.0f + e), v);
}

__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 r;
  __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
  r.x = *reinterpret_cast<uint32_t*>(&h0); r.y = *reinterpret_cast<uint32_t*>(&h1);
  r.z = *reinterpret_cast<uint32_t*>(&h2); r.w = *reinterpret_cast<uint32_t*>(&h3);
  return r;
}
__device__ __forceinline__ void unpack8(uint4 u, float* v) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) { float2 f = __half22float2(h[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
}

// Build per-channel GroupNorm affine (a, b): z = x*a + b, from (sum, sumsq) in double.
// pb:*_norm/{moments,add,Sqrt,truediv,mul,add_1}: biased variance, eps inside the sqrt.
__device__ __forceinline__ void gn_affine(const double* st /*[G][2]*/, int g, float count, float gamma, float beta,
                                          float& a, float& b) {
  double mean = stat_get(st + 2 * g) / (double)count;
  double var = stat_get(st + 2 * g + 1) / (double)count - mean * mean;
  if (var < 0.0) var = 0.0;
  float rstd = (float)(1.0 / sqrt(var + (double)GN_EPS));
  a = rstd * gamma;
  b = beta - (float)mean * a;
}

// --------------------------------------------------------------------------------------
// input packing: x f32 [B,T1,H,W,17] -> X16 frames (32 channels = 4 chunks), optional
// normalize_subtile (src/download_and_predict_job.py:316-325).  Sequence frames get a
// reflect border (pb:.../gates/MirrorPad), the median frame a zero border (SAME conv).
// --------------------------------------------------------------------------------------
struct PrepParams {
  const float* x; int B, T1, H, W;
  uint4* dst; int64_t plane, frame_stride; int Hp, Wp;
  int normalize; float lo[17], hi[17], mid[17], half[17];
};

__global__ void __launch_bounds__(256) prep_input_kernel(PrepParams p) {
  const int t = blockIdx.y;
  const int64_t P = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t Ptot = (int64_t)p.B * p.Hp * p.Wp;
  if (P >= Ptot) return;
  int hw = p.Hp * p.Wp;
  int b = (int)(P / hw); int rem = (int)(P - (int64_t)b * hw);
  int yp = rem / p.Wp, xp = rem - yp * p.Wp;
  int ys = yp - 1, xs = xp - 1;
  bool border = (ys < 0 || ys >= p.H || xs < 0 || xs >= p.W);
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0.f;
  const bool seq = (t < p.T1 - 1);
  if (!border || seq) {
    if (ys < 0) ys = 1; if (ys >= p.H) ys = p.H - 2;
    if (xs < 0) xs = 1; if (xs >= p.W) xs = p.W - 2;
    const float* s = p.x + ((((int64_t)b * p.T1 + t) * p.H + ys) * p.W + xs) * 17;
#pragma unroll
    for (int c = 0; c < 17; ++c) {
      float x = s[c];
      if (p.normalize) {
        x = fminf(fmaxf(x, p.lo[c]), p.hi[c]);
        x = __fdiv_rn(__fsub_rn(x, p.mid[c]), p.half[c]);
      }
      v[c] = x;
    }
  }
  uint4* d = p.dst + (int64_t)t * p.frame_stride;
#pragma unroll
  for (int c = 0; c < 4; ++c) d[(int64_t)c * p.plane + P] = pack8(v + 8 * c);
}

// --------------------------------------------------------------------------------------
// fused tile front end: monthly [B,12,H,W,13] -> quarterly/annual medians + indices
// (assemble, stc_preproc.cu) -> normalize_subtile -> chunk-major fp16 frames with borders.
// Skips the f32 [B,5,H,W,17] intermediate of the separate assemble + prep kernels.
// --------------------------------------------------------------------------------------
__device__ __forceinline__ void write_border(uint4* base, int b, int yp, int xp, int Hp, int Wp,
                                             const uint4& v, bool reflect) {
  base[((int64_t)b * Hp + yp) * Wp + xp] = v;
  if (!reflect) return;
  int ys[3] = {yp, (yp == 2) ? 0 : -1, (yp == Hp - 3) ? Hp - 1 : -1};
  int xs[3] = {xp, (xp == 2) ? 0 : -1, (xp == Wp - 3) ? Wp - 1 : -1};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (ys[i] < 0) continue;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (xs[j] < 0 || (i == 0 && j == 0)) continue;
      base[((int64_t)b * Hp + ys[i]) * Wp + xs[j]] = v;
    }
  }
}

// grid (ceil(HW/128), B): a block owns 128 consecutive pixels of one sample; a warp's loads of one
// month cover one contiguous 32 x 13-element run (L1 absorbs the 13 channel passes).
// T = float, or uint16_t for the reference's integer storage convention (predict_subtile :345-347:
// integer input is divided by 65535; float32 division of the exactly representable operands gives
// the same float32 as NumPy's float64 division followed by astype(float32)).
__device__ __forceinline__ float load_monthly(float v) { return v; }
__device__ __forceinline__ float load_monthly(uint16_t v) { return __fdiv_rn((float)v, 65535.f); }

template <typename T>
__global__ void __launch_bounds__(128) assemble_prep_kernel(const T* __restrict__ in, PrepParams p) {
  // (a cp.async shared-memory staged variant was measured 11% slower -- 80 KB/block leaves 8 warps
  //  per SM and serialises load and sort phases; the direct form keeps 16 warps interleaving)
  const int HW = p.H * p.W;
  const int b = blockIdx.y;
  const int r = blockIdx.x * 128 + threadIdx.x;
  if (r >= HW) return;
  int y = r / p.W, x = r - y * p.W;
  const T* src = in + ((int64_t)b * 12 * HW + r) * 13;
  const int64_t fs_in = (int64_t)HW * 13;
  float bands[5][12];
  float fr[5][8];
  auto norm = [&](float v, int c) { v = fminf(fmaxf(v, p.lo[c]), p.hi[c]); return __fdiv_rn(__fsub_rn(v, p.mid[c]), p.half[c]); };
  auto flush = [&](int chunk) {
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      uint4 u = pack8(fr[f]);
      uint4* d = p.dst + (int64_t)f * p.frame_stride + (int64_t)chunk * p.plane;
      write_border(d, b, y + 1, x + 1, p.Hp, p.Wp, u, f < 4);
    }
  };
#pragma unroll
  for (int c = 0; c < 13; ++c) {
    float v[12];
#pragma unroll
    for (int t = 0; t < 12; ++t) v[t] = load_monthly(src[t * fs_in + c]);
    if (c < 4) {
#pragma unroll
      for (int t = 0; t < 12; ++t) bands[c][t] = v[t];
    } else if (c == 8) {
#pragma unroll
      for (int t = 0; t < 12; ++t) bands[4][t] = v[t];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) fr[q][c & 7] = norm(med3(v[3 * q], v[3 * q + 1], v[3 * q + 2]), c);
    fr[4][c & 7] = norm(median12(v), c);
    if ((c & 7) == 7) flush(c >> 3);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = 13 + k;
    float v[12];
#pragma unroll
    for (int t = 0; t < 12; ++t) {
      float b2 = bands[0][t], b3 = bands[1][t], b4 = bands[2][t], b8 = bands[3][t], b11 = bands[4][t];
      v[t] = (k == 0) ? idx_evi(b2, b3, b4, b8) : (k == 1) ? idx_bi(b2, b4, b8, b11)
           : (k == 2) ? idx_msavi2(b4, b8) : idx_grndvi(b3, b4, b8);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) fr[q][c & 7] = norm(med3(v[3 * q], v[3 * q + 1], v[3 * q + 2]), c);
    fr[4][c & 7] = norm(median12(v), c);
    if (c == 15) {
      flush(1);
#pragma unroll
      for (int f = 0; f < 5; ++f)
#pragma unroll
        for (int i = 0; i < 8; ++i) fr[f][i] = 0.f;
    }
  }
  flush(2);
#pragma unroll
  for (int f = 0; f < 5; ++f)
#pragma unroll
    for (int i = 0; i < 8; ++i) fr[f][i] = 0.f;
  flush(3);
}

// --------------------------------------------------------------------------------------
// Staged front end (default).  Same arithmetic as assemble_prep_kernel, different data movement: the direct kernel
// reads its 12 x 13 values per pixel with 4-byte loads 52 bytes apart and relies on L1 to absorb the 13 channel passes
// (16 warps x 20 KB in flight thrash it; measured 1.8 TB/s, 11 % of the step).  Here every WARP owns 32 consecutive
// pixels: one elected lane fetches the twelve contiguous 32 x 13-element runs with cp.async.bulk into the warp's own
// shared-memory tile (completion on the warp's mbarrier), then lane l reads pixel l's values at a 13-word stride
// (odd: bank-conflict free).  Persistent grid, 8 warps per CTA, warps drift apart so that the bulk loads of some
// overlap the sorting networks of the others.  median12_net (60 FMNMX) replaces the transposition sort (132), and
// the divisions that only feed fp16-rounded outputs use the approximate forms.
// --------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fe_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t fe_elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred px;\nelect.sync _|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ float fe_to_float(float v) { return v; }
__device__ __forceinline__ float fe_to_float(uint16_t v) { return __fdiv_rn((float)v, 65535.f); }

template <typename T>
__global__ void __launch_bounds__(256, 1) assemble_prep_staged_kernel(const T* __restrict__ in, PrepParams p, int groups_per_sample, int total_groups) {
  constexpr int GE = 32 * 13;                         // elements of one month's run for 32 pixels
  constexpr int TILE_BYTES = 12 * GE * (int)sizeof(T);
  extern __shared__ __align__(128) uint8_t fe_smem[];
  const int w = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  T* tile = reinterpret_cast<T*>(fe_smem + w * TILE_BYTES);
  const uint32_t bar = fe_smem_u32(fe_smem + 8 * TILE_BYTES) + 8u * w;
  const uint32_t tile_s = fe_smem_u32(tile);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint32_t leader = fe_elect_one();
  const int HW = p.H * p.W;
  auto norm = [&](float v, int c) { v = fminf(fmaxf(v, p.lo[c]), p.hi[c]); return __fmul_rn(__fsub_rn(v, p.mid[c]), p.half[c]); };   // half = 1/(range/2) here
  uint32_t phase = 0;
  for (int g = blockIdx.x * 8 + w; g < total_groups; g += gridDim.x * 8) {
    const int b = g / groups_per_sample;
    const int r0 = (g - b * groups_per_sample) * 32;
    const int npx = (HW - r0 < 32) ? HW - r0 : 32;
    const uint32_t run_bytes = (uint32_t)(npx * 13 * (int)sizeof(T));
    const T* src = in + ((int64_t)b * 12 * HW + r0) * 13;
    if (leader) {
      // the tile was last read through the generic proxy (previous iteration): order those reads before the async writes
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(12u * run_bytes) : "memory");
    }
#pragma unroll
    for (int t = 0; t < 12; ++t) {
      if (leader)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(tile_s + (uint32_t)(t * GE * (int)sizeof(T))), "l"(src + (int64_t)t * HW * 13), "r"(run_bytes), "r"(bar) : "memory");
    }
    {
      uint32_t ok = 0;
      long long t0 = clock64();
      while (!ok) {
        asm volatile("{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\nselp.u32 %0, 1, 0, q;\n}"
                     : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
        if (!ok && clock64() - t0 > 4000000000LL) __trap();          // never hang the box
      }
      phase ^= 1;
    }
    if (lane < npx) {
      const int r = r0 + lane;
      const int y = r / p.W, x = r - y * p.W;
      const T* mine = tile + lane * 13;
      float bands[5][12];
      float fr[5][8];
      auto flush = [&](int chunk) {
#pragma unroll
        for (int f = 0; f < 5; ++f) {
          uint4 u = pack8(fr[f]);
          uint4* d = p.dst + (int64_t)f * p.frame_stride + (int64_t)chunk * p.plane;
          write_border(d, b, y + 1, x + 1, p.Hp, p.Wp, u, f < 4);
        }
      };
#pragma unroll
      for (int c = 0; c < 13; ++c) {
        float v[12];
#pragma unroll
        for (int t = 0; t < 12; ++t) v[t] = fe_to_float(mine[t * GE + c]);
        if (c < 4) {
#pragma unroll
          for (int t = 0; t < 12; ++t) bands[c][t] = v[t];
        } else if (c == 8) {
#pragma unroll
          for (int t = 0; t < 12; ++t) bands[4][t] = v[t];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) fr[q][c & 7] = norm(med3(v[3 * q], v[3 * q + 1], v[3 * q + 2]), c);
        fr[4][c & 7] = norm(median12_net(v), c);
        if ((c & 7) == 7) flush(c >> 3);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = 13 + k;
        float v[12];
#pragma unroll
        for (int t = 0; t < 12; ++t) {
          float b2 = bands[0][t], b3 = bands[1][t], b4 = bands[2][t], b8 = bands[3][t], b11 = bands[4][t];
          v[t] = (k == 0) ? idx_evi_fast(b2, b4, b8) : (k == 1) ? idx_bi_fast(b2, b4, b8, b11)
               : (k == 2) ? idx_msavi2_fast(b4, b8) : idx_grndvi_fast(b3, b4, b8);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) fr[q][c & 7] = norm(med3(v[3 * q], v[3 * q + 1], v[3 * q + 2]), c);
        fr[4][c & 7] = norm(median12_net(v), c);
        if (c == 15) {
          flush(1);
#pragma unroll
          for (int f = 0; f < 5; ++f)
#pragma unroll
            for (int i = 0; i < 8; ++i) fr[f][i] = 0.f;
        }
      }
      flush(2);
#pragma unroll
      for (int f = 0; f < 5; ++f)
#pragma unroll
        for (int i = 0; i < 8; ++i) fr[f][i] = 0.f;
      flush(3);
    }
    __syncwarp();      // every lane is done with the tile before the next bulk copies overwrite it
  }
}

// --------------------------------------------------------------------------------------
// ConvGRU gating stages (pb:.../while/<d>/conv_gru_cell/*)
// --------------------------------------------------------------------------------------
struct GruParams {
  const uint4* rawG[2]; int64_t rawG_plane;    // gates conv output, 64 ch fp16 (8 planes of 8 ch)
  const uint4* rawY[2]; int64_t rawY_plane;    // candidate conv output, 32 ch fp16 (4 planes)
  float4* Hf[2]; int64_t Hf_plane;             // fp32 state, 32 ch (8 planes)
  uint4* Hh[2]; uint4* RH[2]; int64_t act_plane; // fp16 copies with reflect border
  uint4* cc[2]; int64_t cc_plane;              // final-step destination (CCin chunks), or null
  const double* stG[2]; const double* stY[2];  // [B][16][2], [B][8][2]
  const float* gam_r[2]; const float* bet_r[2]; const float* gam_u[2]; const float* bet_u[2];
  const float* gam_y[2]; const float* bet_y[2];
  int B, H, W, Hp, Wp;
  float count;                                 // H*W*4 elements per group
  int h_zero;                                  // first step: h == 0, do not read the state buffers
};

__device__ __forceinline__ void write_reflect(uint4* plane_base, int64_t plane, int chunks, int b, int yp, int xp,
                                              int Hp, int Wp, const uint4* vals) {
  int ys[3] = {yp, (yp == 2) ? 0 : -1, (yp == Hp - 3) ? Hp - 1 : -1};
  int xs[3] = {xp, (xp == 2) ? 0 : -1, (xp == Wp - 3) ? Wp - 1 : -1};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (ys[i] < 0) continue;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (xs[j] < 0) continue;
      int64_t P = ((int64_t)b * Hp + ys[i]) * Wp + xs[j];
      for (int c = 0; c < chunks; ++c) plane_base[(int64_t)c * plane + P] = vals[c];
    }
  }
}

// r = sigmoid(GN(g_r)); RH = r * h  (fp16, reflect border).
// h is read from the fp32 state.  Two cheaper variants were measured and rejected on accuracy, which is dominated by
// the fp16 rounding of the conv operands (tools/exp/precision_study.py): the state kept ONLY in fp16 (27.9 -> 27.0 ms
// per 256-tile step, reference-golden error 6.1e-4 -> 7.0e-4 of the 1e-3 budget) and h read here from the fp16 copy
// (r*h then rounds twice: tests/test_process_subtiles.py tips over its 1e-3 + rounding-step bound).
__global__ void __launch_bounds__(256) gru_apply1_kernel(GruParams p) {
  const int d = blockIdx.z, b = blockIdx.y;
  __shared__ float sa[32], sb[32];
  if (threadIdx.x < 32) {
    int c = threadIdx.x;
    gn_affine(p.stG[d] + (int64_t)b * 32, c >> 2, p.count, p.gam_r[d][c], p.bet_r[d][c], sa[c], sb[c]);
  }
  __syncthreads();
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.H * p.W) return;
  int y = idx / p.W, x = idx - y * p.W;
  int yp = y + 1, xp = x + 1;
  int64_t P = ((int64_t)b * p.Hp + yp) * p.Wp + xp;
  uint4 graw[4]; float4 hraw[8];                 // all twelve loads in flight before the first use
#pragma unroll
  for (int c = 0; c < 4; ++c) graw[c] = p.rawG[d][(int64_t)c * p.rawG_plane + P];
#pragma unroll
  for (int c = 0; c < 8; ++c) hraw[c] = p.Hf[d][(int64_t)c * p.Hf_plane + P];
  uint4 out[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float v[8], gr[8];
    unpack8(graw[c], gr);
    const float hs[8] = {hraw[2 * c].x, hraw[2 * c].y, hraw[2 * c].z, hraw[2 * c].w,
                         hraw[2 * c + 1].x, hraw[2 * c + 1].y, hraw[2 * c + 1].z, hraw[2 * c + 1].w};
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = sigm_fast(gr[k] * sa[8 * c + k] + sb[8 * c + k]) * hs[k];
    out[c] = pack8(v);
  }
  write_reflect(p.RH[d], p.act_plane, 4, b, yp, xp, p.Hp, p.Wp, out);
}

// u = sigmoid(GN(g_u)); h~ = u*h + (1-u)*tanh(GN(y)); h = 0.75 h + 0.25 h~ (zoneout, inference)
__global__ void __launch_bounds__(256, 3) gru_apply2_kernel(GruParams p) {
  const int d = blockIdx.z, b = blockIdx.y;
  __shared__ float ua[32], ub[32], ya[32], yb[32];
  if (threadIdx.x < 32) {
    int c = threadIdx.x;
    gn_affine(p.stG[d] + (int64_t)b * 32, 8 + (c >> 2), p.count, p.gam_u[d][c], p.bet_u[d][c], ua[c], ub[c]);
    gn_affine(p.stY[d] + (int64_t)b * 16, c >> 2, p.count, p.gam_y[d][c], p.bet_y[d][c], ya[c], yb[c]);
  }
  __syncthreads();
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.H * p.W) return;
  int y = idx / p.W, x = idx - y * p.W;
  int yp = y + 1, xp = x + 1;
  int64_t P = ((int64_t)b * p.Hp + yp) * p.Wp + xp;
  // all sixteen loads of the pixel in flight before the first use (the stores below would otherwise fence them)
  uint4 graw[4], yraw[4]; float4 hraw[8];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    graw[c] = p.rawG[d][(int64_t)(4 + c) * p.rawG_plane + P];
    yraw[c] = p.rawY[d][(int64_t)c * p.rawY_plane + P];
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) hraw[c] = p.h_zero ? make_float4(0.f, 0.f, 0.f, 0.f) : p.Hf[d][(int64_t)c * p.Hf_plane + P];
  uint4 out[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float v[8], gu[8], yv[8];
    unpack8(graw[c], gu); unpack8(yraw[c], yv);
    const float hv[8] = {hraw[2 * c].x, hraw[2 * c].y, hraw[2 * c].z, hraw[2 * c].w,
                         hraw[2 * c + 1].x, hraw[2 * c + 1].y, hraw[2 * c + 1].z, hraw[2 * c + 1].w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float u = sigm_fast(gu[k] * ua[8 * c + k] + ub[8 * c + k]);
      const float cand = tanh_fast(yv[k] * ya[8 * c + k] + yb[8 * c + k]);
      const float ht = u * hv[k] + (1.f - u) * cand;
      v[k] = 0.75f * hv[k] + 0.25f * ht;
    }
    p.Hf[d][(int64_t)(2 * c) * p.Hf_plane + P] = make_float4(v[0], v[1], v[2], v[3]);
    p.Hf[d][(int64_t)(2 * c + 1) * p.Hf_plane + P] = make_float4(v[4], v[5], v[6], v[7]);
    out[c] = pack8(v);
  }
  write_reflect(p.Hh[d], p.act_plane, 4, b, yp, xp, p.Hp, p.Wp, out);
  if (p.cc[d]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) p.cc[d][(int64_t)c * p.cc_plane + P] = out[c];
  }
}

// --------------------------------------------------------------------------------------
// conv block tail: GroupNorm(8) -> sSE -> {identity/crop | maxpool2 | nearest x2} placement
// into the consumer's fp16 buffer, or the 1x1 head + sigmoid (pb:<blk>_norm, csse_<blk>_*,
// max_pooling2d*, up_sampling2d*, cropping2d*, conv2d/Sigmoid).
// --------------------------------------------------------------------------------------
struct ApplyParams {
  const float4* raw; int64_t raw_plane; int C;
  int sHp, sWp, so;            // source padded geometry; so = padded offset of valid output (1 SAME, 2 VALID)
  const double* stats; float count;
  const float* gamma; const float* beta; const float* sse_w; const float* sse_b;
  uint4* dst; int64_t dst_plane; int dHp, dWp;
  int Hd, Wd; int mode; int off;  // mode 0 identity(+crop off) 1 maxpool2 2 upsample2 3 head
  const float* head_w; const float* head_b; float* head_out;
  float* feat_out;             // mode 3, optional: the block's sSE output (pb:csse_out_mul/mul) [B,Hd,Wd,C] float32
  // mode 1, optional second consumer of the same block output: the centre crop [off2, off2 + Hd2) x [off2, off2 + Wd2) of the un-pooled tensor
  // (skip connection into a decoder concat buffer), written by the thread that pools the pixel -- the block's raw output
  // is then read once instead of by two launches
  uint4* dst2; int64_t dst2_plane; int d2Hp, d2Wp, off2, Hd2, Wd2;
};

// Compiled per (C, MODE) so that the channel loops unroll and all plane loads of a pixel are in flight at once
// (a runtime-C loop keeps one 16-byte load per thread in flight: measured ~48 % of HBM peak).
// The sSE logit is folded to dot(v, a*w) + (sum b*w + bias); for C == 64 the raw values stay in registers
// between the logit pass and the output pass.
__device__ __forceinline__ float fast_sigm(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }

template <int C, int MODE>
__global__ void __launch_bounds__(256, 2) block_apply_t(ApplyParams p) {
  __shared__ float4 s_a[C / 4], s_b[C / 4], s_w[C / 4];     // GN scale, GN shift, logit (or head) weights a*w
  __shared__ float s_red[8];
  __shared__ float s_bias;
  const int b = blockIdx.y;
  constexpr int gs = C / 8;
  float part = 0.f;
  for (int c = threadIdx.x; c < C; c += 256) {
    float a, bb;
    gn_affine(p.stats + (int64_t)b * 16, c / gs, p.count, p.gamma[c], p.beta[c], a, bb);
    const float w = p.sse_w[c];
    reinterpret_cast<float*>(s_a)[c] = a; reinterpret_cast<float*>(s_b)[c] = bb; reinterpret_cast<float*>(s_w)[c] = a * w;
    part += bb * w;
  }
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) { float s = 0.f; for (int w = 0; w < 8; ++w) s += s_red[w]; s_bias = s + p.sse_b[0]; }
  __syncthreads();
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= p.Hd * p.Wd) return;
  const int yd = idx / p.Wd, xd = idx - yd * p.Wd;
  constexpr int NSRC = (MODE == 1) ? 4 : 1;
  constexpr bool KEEP = (C == 64) && (MODE != 1);
  int64_t SP[NSRC]; float sv[NSRC];
  float4 keep[KEEP ? C / 4 : 1];
#pragma unroll
  for (int k = 0; k < NSRC; ++k) {
    int ys, xs;
    if (MODE == 1) { ys = 2 * yd + (k >> 1); xs = 2 * xd + (k & 1); }
    else if (MODE == 2) { ys = yd >> 1; xs = xd >> 1; }
    else { ys = yd + p.off; xs = xd + p.off; }
    SP[k] = ((int64_t)b * p.sHp + ys + p.so) * p.sWp + xs + p.so;
    float dot = s_bias;
#pragma unroll
    for (int c0 = 0; c0 < C / 4; c0 += 16) {
      float4 v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = p.raw[(int64_t)(c0 + j) * p.raw_plane + SP[k]];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 w = s_w[c0 + j];
        dot += v[j].x * w.x + v[j].y * w.y + v[j].z * w.z + v[j].w * w.w;
        if (KEEP) keep[c0 + j] = v[j];
      }
    }
    sv[k] = fast_sigm(dot);
  }
  if (MODE == 3) {      // 1x1 head: sigmoid(sum_c z_c * sv * hw_c + hb)
    float acc = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < C / 4; ++c4) {
      const float4 v = KEEP ? keep[c4] : p.raw[(int64_t)c4 * p.raw_plane + SP[0]];
      const float4 a = s_a[c4], bb = s_b[c4];
      const float4 hw = *reinterpret_cast<const float4*>(p.head_w + 4 * c4);
      acc += (v.x * a.x + bb.x) * hw.x + (v.y * a.y + bb.y) * hw.y + (v.z * a.z + bb.z) * hw.z + (v.w * a.w + bb.w) * hw.w;
    }
    p.head_out[((int64_t)b * p.Hd + yd) * p.Wd + xd] = fast_sigm(acc * sv[0] + p.head_b[0]);
    if (p.feat_out) {     // --gen_feats late features (src/download_and_predict_job.py:1430, pb:csse_out_mul/mul)
      float4* fo = reinterpret_cast<float4*>(p.feat_out + (((int64_t)b * p.Hd + yd) * p.Wd + xd) * C);
#pragma unroll
      for (int c4 = 0; c4 < C / 4; ++c4) {
        const float4 v = KEEP ? keep[c4] : p.raw[(int64_t)c4 * p.raw_plane + SP[0]];
        const float4 a = s_a[c4], bb = s_b[c4];
        fo[c4] = make_float4((v.x * a.x + bb.x) * sv[0], (v.y * a.y + bb.y) * sv[0], (v.z * a.z + bb.z) * sv[0], (v.w * a.w + bb.w) * sv[0]);
      }
    }
    return;
  }
  const int64_t DP = ((int64_t)b * p.dHp + yd + 1) * p.dWp + xd + 1;
#pragma unroll
  for (int c8 = 0; c8 < C / 8; ++c8) {
    float o[8];
    const float4 a0 = s_a[2 * c8], a1 = s_a[2 * c8 + 1], b0 = s_b[2 * c8], b1 = s_b[2 * c8 + 1];
    const float sa8[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, sb8[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int k = 0; k < NSRC; ++k) {
      const float4 v0 = KEEP ? keep[2 * c8] : p.raw[(int64_t)(2 * c8) * p.raw_plane + SP[k]];
      const float4 v1 = KEEP ? keep[2 * c8 + 1] : p.raw[(int64_t)(2 * c8 + 1) * p.raw_plane + SP[k]];
      const float t[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      float zk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        zk[i] = (t[i] * sa8[i] + sb8[i]) * sv[k];
        o[i] = (k == 0) ? zk[i] : fmaxf(o[i], zk[i]);
      }
      if (MODE == 1 && p.dst2) {
        const int y2 = 2 * yd + (k >> 1) - p.off2, x2 = 2 * xd + (k & 1) - p.off2;
        if (y2 >= 0 && y2 < p.Hd2 && x2 >= 0 && x2 < p.Wd2)
          p.dst2[(int64_t)c8 * p.dst2_plane + ((int64_t)b * p.d2Hp + y2 + 1) * p.d2Wp + x2 + 1] = pack8(zk);
      }
    }
    p.dst[(int64_t)c8 * p.dst_plane + DP] = pack8(o);
  }
}

template <int C>
static void launch_apply_c(const ApplyParams& ap, dim3 grid, cudaStream_t s) {
  switch (ap.mode) {
    case 0: block_apply_t<C, 0><<<grid, 256, 0, s>>>(ap); break;
    case 1: block_apply_t<C, 1><<<grid, 256, 0, s>>>(ap); break;
    case 2: block_apply_t<C, 2><<<grid, 256, 0, s>>>(ap); break;
    default: block_apply_t<C, 3><<<grid, 256, 0, s>>>(ap); break;
  }
}

// --gen_feats early features (src/download_and_predict_job.py:1431, pb:gru_drop/drop_block2d/cond/Merge = the
// bidirectional ConvGRU output, DropBlock is the identity at inference): channels 0..63 of the concat buffer,
// centre-cropped by `crop` like predict_subtile does (:360-362), float32 [B,Hd,Wd,64].
__global__ void feat_early_kernel(const uint4* src, int64_t plane, int B, int Hp, int Wp, int crop, int Hd, int Wd, float* out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t n = (int64_t)B * Hd * Wd * 8;
  if (idx >= n) return;
  int c = (int)(idx & 7); int64_t r = idx >> 3;
  int x = (int)(r % Wd); r /= Wd; int y = (int)(r % Hd); int b = (int)(r / Hd);
  int64_t P = ((int64_t)b * Hp + y + crop + 1) * Wp + x + crop + 1;
  float v[8];
  unpack8(src[(int64_t)c * plane + P], v);
  float4* o = reinterpret_cast<float4*>(out + (((int64_t)b * Hd + y) * Wd + x) * 64 + c * 8);
  o[0] = make_float4(v[0], v[1], v[2], v[3]); o[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// decode an fp16 activation buffer interior to fp32 NHWC (debug / tests)
__global__ void act_decode_kernel(const uint4* src, int64_t plane, int chunks, int B, int Hp, int Wp, float* out) {
  int H = Hp - 2, W = Wp - 2;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t n = (int64_t)B * H * W * chunks;
  if (idx >= n) return;
  int c = (int)(idx % chunks); int64_t r = idx / chunks;
  int x = (int)(r % W); r /= W; int y = (int)(r % H); int b = (int)(r / H);
  int64_t P = ((int64_t)b * Hp + y + 1) * Wp + x + 1;
  float v[8];
  unpack8(src[(int64_t)c * plane + P], v);
  float* o = out + (((int64_t)b * H + y) * W + x) * (chunks * 8) + c * 8;
  for (int i = 0; i < 8; ++i) o[i] = v[i];
}

// --------------------------------------------------------------------------------------
// host side: packed weights + scratch plan
// --------------------------------------------------------------------------------------
static const char* BLK[8] = {"conv_median", "conv_concat", "conv1", "conv2", "up2", "up2_out", "up3", "out"};
static const int BLK_CIN[8] = {17, 128, 64, 128, 256, 256, 128, 128};
static const int BLK_COUT[8] = {64, 64, 128, 256, 128, 128, 64, 64};

struct DevAct {
  Act a; size_t bytes = 0;
};

struct ModelState {
  // packed weights
  uint4* w_gates[2] = {nullptr, nullptr}; uint4* w_cand[2] = {nullptr, nullptr};
  uint4* w_blk[8] = {nullptr};
  float* fparams = nullptr;   // all small f32 vectors, contiguous
  std::map<std::string, const float*> fp;
  // scratch plan
  int Bc = 0, H = 0, W = 0, T1 = 0;
  void* arena = nullptr; size_t arena_bytes = 0;
  Act X16; int64_t x_frame_stride = 0;
  Act Hh[2], RH[2], CCin, P1, CAT1, P2, U2in, U3in, CAT2;
  Raw Hf[2], rawG[2], rawY[2], rawB;
  double* stats = nullptr; size_t stats_bytes = 0;
  double* stG = nullptr; double* stY = nullptr; double* stB = nullptr;
  int lastB = 0;
  bool weights_ready = false;
};

// pack HWIO f32 conv weights into [Ksteps][9][2][Npad][8] fp16 according to chanmap
static std::vector<__half> pack_conv(const float* w, int Cin, int Cout, const std::vector<int>& chanmap, int Npad) {
  int Ksteps = (int)chanmap.size() / 16;
  std::vector<__half> out((size_t)Ksteps * 9 * 2 * Npad * 8);
  for (int ks = 0; ks < Ksteps; ++ks)
    for (int tap = 0; tap < 9; ++tap)
      for (int ch = 0; ch < 2; ++ch)
        for (int n = 0; n < Npad; ++n)
          for (int k = 0; k < 8; ++k) {
            int cin = chanmap[ks * 16 + ch * 8 + k];
            float v = (cin >= 0 && n < Cout) ? w[((size_t)tap * Cin + cin) * Cout + n] : 0.f;
            out[((((size_t)ks * 9 + tap) * 2 + ch) * Npad + n) * 8 + k] = __float2half_rn(v);
          }
  return out;
}

static int upload_packed(stc_ctx* ctx, const std::vector<__half>& h, uint4** dptr) {
  if (*dptr) cudaFree(*dptr);
  STC_CUDA(cudaMalloc((void**)dptr, h.size() * sizeof(__half)));
  STC_CUDA(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(__half), cudaMemcpyHostToDevice));
  return STC_OK;
}

int pack_and_upload_conv(stc_ctx* ctx, const float* w, int Cin, int Cout, const std::vector<int>& chanmap, int Npad, uint4** dptr) {
  return upload_packed(ctx, pack_conv(w, Cin, Cout, chanmap, Npad), dptr);
}

static const std::vector<float>* getw(stc_ctx* ctx, const std::string& name, size_t n) {
  auto it = ctx->host_w.find(name);
  if (it == ctx->host_w.end() || it->second.size() != n) return nullptr;
  return &it->second;
}

static int finalize_slot(stc_ctx* ctx, void** slot);

int model_num_slots() {
  static const int n = [] {
    const char* e = getenv("STC_SLOTS");
    int v = e ? atoi(e) : 4;
    return v < 1 ? 1 : (v > stc_ctx::MAX_SLOTS ? stc_ctx::MAX_SLOTS : v);
  }();
  return n;
}

cudaStream_t model_slot_stream(stc_ctx* ctx, int slot) {
  if (slot == 0) return ctx->stream;
  if (!ctx->slot_stream[slot]) cudaStreamCreateWithFlags(&ctx->slot_stream[slot], cudaStreamNonBlocking);
  return ctx->slot_stream[slot];
}

int model_finalize_weights(stc_ctx* ctx) {
  for (int i = 0; i < model_num_slots(); ++i) {
    int rc = finalize_slot(ctx, &ctx->slots[i]);
    if (rc) return rc;
  }
  return STC_OK;
}

static int finalize_slot(stc_ctx* ctx, void** slot) {
  ModelState* m = (ModelState*)*slot;
  if (!m) { m = new ModelState(); *slot = m; }
  // ---- conv kernels ----
  std::vector<int> gru_map(64, -1);     // A channels: [x 0..16 | pad | h 0..31] -> cin [0..16 | - | 17..48]
  for (int c = 0; c < 17; ++c) gru_map[c] = c;
  for (int c = 0; c < 32; ++c) gru_map[32 + c] = 17 + c;
  const char* dn[2] = {"fw", "bw"};
  for (int d = 0; d < 2; ++d) {
    auto* g = getw(ctx, std::string("gru.") + dn[d] + ".gates_w", 9 * 49 * 64);
    auto* c = getw(ctx, std::string("gru.") + dn[d] + ".cand_w", 9 * 49 * 32);
    if (!g || !c) STC_FAIL(STC_ERR_STATE, "missing GRU conv weights");
    int rc = upload_packed(ctx, pack_conv(g->data(), 49, 64, gru_map, 64), &m->w_gates[d]); if (rc) return rc;
    rc = upload_packed(ctx, pack_conv(c->data(), 49, 32, gru_map, 32), &m->w_cand[d]); if (rc) return rc;
  }
  for (int i = 0; i < 8; ++i) {
    int cin = BLK_CIN[i], cout = BLK_COUT[i];
    auto* w = getw(ctx, std::string(BLK[i]) + ".w", (size_t)9 * cin * cout);
    if (!w) STC_FAIL(STC_ERR_STATE, std::string("missing weights for block ") + BLK[i]);
    int cpad = (cin + 15) / 16 * 16;
    if (i == 0) cpad = 32;
    std::vector<int> map(cpad, -1);
    for (int c = 0; c < cin; ++c) map[c] = c;
    int rc = upload_packed(ctx, pack_conv(w->data(), cin, cout, map, cout), &m->w_blk[i]); if (rc) return rc;
  }
  // ---- small f32 vectors ----
  std::vector<std::pair<std::string, size_t>> names;
  for (int d = 0; d < 2; ++d) {
    std::string p = std::string("gru.") + dn[d] + ".";
    for (const char* s : {"r_gamma", "r_beta", "u_gamma", "u_beta", "y_gamma", "y_beta", "cand_sse_w"}) names.push_back({p + s, 32});
  }
  for (int i = 0; i < 8; ++i) {
    std::string p = std::string(BLK[i]) + ".";
    names.push_back({p + "gamma", (size_t)BLK_COUT[i]}); names.push_back({p + "beta", (size_t)BLK_COUT[i]});
    names.push_back({p + "sse_w", (size_t)BLK_COUT[i]}); names.push_back({p + "sse_b", 1});
  }
  names.push_back({"head.w", 64}); names.push_back({"head.b", 1});
  std::vector<float> flat; std::map<std::string, size_t> offs;
  for (auto& nm : names) {
    auto* v = getw(ctx, nm.first, nm.second);
    if (!v) STC_FAIL(STC_ERR_STATE, "missing or mis-sized weight " + nm.first);
    offs[nm.first] = flat.size();
    flat.insert(flat.end(), v->begin(), v->end());
    while (flat.size() % 4) flat.push_back(0.f);
  }
  if (m->fparams) cudaFree(m->fparams);
  STC_CUDA(cudaMalloc((void**)&m->fparams, flat.size() * sizeof(float)));
  STC_CUDA(cudaMemcpy(m->fparams, flat.data(), flat.size() * sizeof(float), cudaMemcpyHostToDevice));
  m->fp.clear();
  for (auto& o : offs) m->fp[o.first] = m->fparams + o.second;
  m->weights_ready = true;
  return STC_OK;
}

void model_destroy(stc_ctx* ctx) {
  for (int i = 0; i < stc_ctx::MAX_SLOTS; ++i) {
    ModelState* m = (ModelState*)ctx->slots[i];
    if (m) {
      for (int d = 0; d < 2; ++d) { cudaFree(m->w_gates[d]); cudaFree(m->w_cand[d]); }
      for (int k = 0; k < 8; ++k) cudaFree(m->w_blk[k]);
      cudaFree(m->fparams); cudaFree(m->arena); cudaFree(m->stats);
      delete m; ctx->slots[i] = nullptr;
    }
    if (ctx->slot_stream[i]) { cudaStreamDestroy(ctx->slot_stream[i]); ctx->slot_stream[i] = nullptr; }
    if (ctx->ev_join[i]) { cudaEventDestroy(ctx->ev_join[i]); ctx->ev_join[i] = nullptr; }
  }
  if (ctx->ev_fork) { cudaEventDestroy(ctx->ev_fork); ctx->ev_fork = nullptr; }
}

// One resolution level of the network: H x W images (the released graphs are square; the border re-segmentation pass of
// src/resegment_tiles_wide.py:478 predicts 220 x 684 seams), padded by one pixel on every side.
struct Geo { int H, W, Hp, Wp; int64_t P; };
static Geo geo(int Bc, int H, int W) { Geo g; g.H = H; g.W = W; g.Hp = H + 2; g.Wp = W + 2; g.P = (int64_t)Bc * g.Hp * g.Wp; return g; }

static size_t act_units(int chunks, const Geo& g, int& guard) {
  guard = ((g.Wp + 2 + 544 + 7) / 8) * 8;   // >= Wp + 1 + (NT*128 + 8) staged rows past the last tile
  return (size_t)chunks * (size_t)(g.P + 2 * guard);
}

static int ensure_plan(stc_ctx* ctx, ModelState* m, int Bc, int H, int W, int T1) {
  if (m->arena && m->Bc == Bc && m->H == H && m->W == W && m->T1 == T1) return STC_OK;
  if (m->arena) { cudaFree(m->arena); m->arena = nullptr; }
  if (m->stats) { cudaFree(m->stats); m->stats = nullptr; }
  const int p1 = H / 2, c1 = p1 - 2, p2 = c1 / 2, c2 = p2 - 2, u2 = 2 * c2, u3 = 2 * u2;
  const int q1 = W / 2, d1 = q1 - 2, q2 = d1 / 2, d2 = q2 - 2, v2 = 2 * d2, v3 = 2 * v2;       // the same chain along x
  Geo g0 = geo(Bc, H, W), g1 = geo(Bc, p1, q1), g2 = geo(Bc, p2, q2), gu2 = geo(Bc, u2, v2), gu3 = geo(Bc, u3, v3);
  struct Item { Act* a; int chunks; Geo g; };
  std::vector<Item> acts = {
      {&m->X16, 4 * T1, g0}, {&m->Hh[0], 4, g0}, {&m->Hh[1], 4, g0}, {&m->RH[0], 4, g0}, {&m->RH[1], 4, g0},
      {&m->CCin, 16, g0}, {&m->P1, 8, g1}, {&m->CAT1, 32, gu2}, {&m->P2, 16, g2}, {&m->U2in, 32, gu2},
      {&m->U3in, 16, gu3}, {&m->CAT2, 16, gu3}};
  size_t total = 0;
  std::vector<size_t> offs;
  for (auto& it : acts) { int guard; size_t u = act_units(it.chunks, it.g, guard); offs.push_back(total); total += u * 16; }
  // raw buffers (float4 planes): plane length = P0 rounded up to 512
  int64_t rp = (g0.P + 511) / 512 * 512;
  struct RItem { Raw* r; int planes; };
  std::vector<RItem> raws = {{&m->Hf[0], 8}, {&m->Hf[1], 8}, {&m->rawG[0], 8}, {&m->rawG[1], 8},
                             {&m->rawY[0], 4}, {&m->rawY[1], 4}, {&m->rawB, 16}};   // rawG/rawY hold fp16 (uint4 = 8 ch)
  std::vector<size_t> roffs;
  for (auto& it : raws) { roffs.push_back(total); total += (size_t)it.planes * rp * 16; }
  STC_CUDA(cudaMalloc(&m->arena, total));
  STC_CUDA(cudaMemsetAsync(m->arena, 0, total, ctx->stream));
  m->arena_bytes = total;
  for (size_t i = 0; i < acts.size(); ++i) {
    Act& a = *acts[i].a; int guard; act_units(acts[i].chunks, acts[i].g, guard);
    a.base = (uint4*)((char*)m->arena + offs[i]); a.guard = guard; a.plane = acts[i].g.P + 2 * guard;
    a.chunks = acts[i].chunks; a.B = Bc; a.Hp = acts[i].g.Hp; a.Wp = acts[i].g.Wp;
  }
  m->x_frame_stride = 4 * m->X16.plane;
  for (size_t i = 0; i < raws.size(); ++i) {
    Raw& r = *raws[i].r; r.base = (float4*)((char*)m->arena + roffs[i]); r.plane = rp; r.N = raws[i].planes * 4;
  }
  // stats: gates [T][2][Bc][16][2], cand [T][2][Bc][8][2], blocks [8][Bc][8][2]
  int T = T1 - 1;
  size_t nG = (size_t)T * 2 * Bc * 32, nY = (size_t)T * 2 * Bc * 16, nB = (size_t)8 * Bc * 16;
  m->stats_bytes = (nG + nY + nB) * sizeof(double);
  STC_CUDA(cudaMalloc((void**)&m->stats, m->stats_bytes));
  m->stG = m->stats; m->stY = m->stats + nG; m->stB = m->stats + nG + nY;
  m->Bc = Bc; m->H = H; m->W = W; m->T1 = T1;
  return STC_OK;
}

static void set_valid(ConvParams& cp, const Act& a, bool same) {
  cp.B = a.B; cp.Hp = a.Hp; cp.Wp = a.Wp; cp.Ptot = a.Ptot();
  int lo = same ? 1 : 2;
  cp.vy0 = lo; cp.vy1 = a.Hp - lo; cp.vx0 = lo; cp.vx1 = a.Wp - lo;
}

static int run_block_conv(stc_ctx* ctx, ModelState* m, int blk, const Act& in, int in_chunks, bool same, int B) {
  ConvParams cp; memset(&cp, 0, sizeof(cp));
  cp.a0[0] = in.at(0); cp.a0_plane = in.plane; cp.k0steps = in_chunks / 2; cp.k1steps = 0;
  cp.a1[0] = nullptr; cp.a1_plane = 0;
  cp.w[0] = m->w_blk[blk];
  cp.out[0] = m->rawB.base; cp.out_plane = (int64_t)((in.Ptot() + 511) / 512 * 512);
  cp.stats[0] = m->stB + (size_t)blk * m->Bc * 16;
  cp.N = BLK_COUT[blk]; cp.G = 8;
  set_valid(cp, in, same);
  cp.B = B; cp.Ptot = (int64_t)B * in.Hp * in.Wp;
  cp.mode = same ? MODE_PSCALE_SWISH : MODE_SWISH;
  return launch_conv(ctx, cp, 1);
}

static int run_apply(stc_ctx* ctx, ModelState* m, int blk, const Act& src_geo, bool same, int mode, int off,
                     Act* dst, int dst_chunk_off, int Hd, int Wd, int B, float* head_out,
                     Act* dst2 = nullptr, int dst2_chunk_off = 0, int off2 = 0, int Hd2 = 0, int Wd2 = 0) {
  ApplyParams ap; memset(&ap, 0, sizeof(ap));
  ap.raw = m->rawB.base; ap.raw_plane = (int64_t)((src_geo.Ptot() + 511) / 512 * 512); ap.C = BLK_COUT[blk];
  ap.sHp = src_geo.Hp; ap.sWp = src_geo.Wp; ap.so = same ? 1 : 2;
  ap.stats = m->stB + (size_t)blk * m->Bc * 16;
  int Hv = src_geo.Hp - 2 * ap.so, Wv = src_geo.Wp - 2 * ap.so;
  ap.count = (float)Hv * (float)Wv * (float)(ap.C / 8);
  std::string n = BLK[blk];
  ap.gamma = m->fp[n + ".gamma"]; ap.beta = m->fp[n + ".beta"]; ap.sse_w = m->fp[n + ".sse_w"]; ap.sse_b = m->fp[n + ".sse_b"];
  if (dst) { ap.dst = dst->at(dst_chunk_off); ap.dst_plane = dst->plane; ap.dHp = dst->Hp; ap.dWp = dst->Wp; }
  if (dst2 && mode == 1) { ap.dst2 = dst2->at(dst2_chunk_off); ap.dst2_plane = dst2->plane; ap.d2Hp = dst2->Hp; ap.d2Wp = dst2->Wp; ap.off2 = off2; ap.Hd2 = Hd2; ap.Wd2 = Wd2; }
  ap.Hd = Hd; ap.Wd = Wd; ap.mode = mode; ap.off = off;
  ap.head_w = m->fp["head.w"]; ap.head_b = m->fp["head.b"]; ap.head_out = head_out;
  ap.feat_out = (mode == 3) ? ctx->feat_late_chunk : nullptr;
  dim3 grid(cdiv((int64_t)Hd * Wd, 256), B);
  trace_begin(ctx, "block_apply");
  if (ap.C == 64) launch_apply_c<64>(ap, grid, ctx->stream);
  else if (ap.C == 128) launch_apply_c<128>(ap, grid, ctx->stream);
  else if (ap.C == 256) launch_apply_c<256>(ap, grid, ctx->stream);
  else STC_FAIL(STC_ERR_ARG, "block tail: unsupported channel count");
  trace_end(ctx);
  STC_CUDA(cudaGetLastError());
  ctx->launches++;
  return STC_OK;
}

static int forward_chunk(stc_ctx* ctx, ModelState* m, const float* x_dev, const float* monthly_dev, int B, int T, int H, int W, int length,
                         int normalize, const double* mn, const double* mx, float* out_dev, cudaEvent_t input_consumed) {
  const int T1 = T + 1;
  const int p1 = H / 2, c1 = p1 - 2, p2 = c1 / 2, c2 = p2 - 2, u2 = 2 * c2, u3 = 2 * u2;
  const int q1 = W / 2, d1 = q1 - 2, q2 = d1 / 2, d2 = q2 - 2, v2 = 2 * d2, v3 = 2 * v2;
  (void)c2; (void)d2;
  // ---- reset state ----
  // (the GRU state needs no clearing: step 0 runs with h == 0 folded in -- no h K-steps, no r*h stage)
  STC_CUDA(cudaMemsetAsync(m->stats, 0, m->stats_bytes, ctx->stream));
  // ---- pack input ----
  {
    PrepParams pp; memset(&pp, 0, sizeof(pp));
    pp.x = x_dev; pp.B = B; pp.T1 = T1; pp.H = H; pp.W = W;
    pp.dst = m->X16.at(0); pp.plane = m->X16.plane; pp.frame_stride = m->x_frame_stride;
    pp.Hp = H + 2; pp.Wp = W + 2; pp.normalize = normalize;
    for (int c = 0; c < 17; ++c) {
      // normalize_subtile: python-float (double) mins/maxs, float32 array arithmetic
      double lo = normalize ? mn[c] : 0.0, hi = normalize ? mx[c] : 1.0;
      pp.lo[c] = (float)lo; pp.hi[c] = (float)hi; pp.mid[c] = (float)((hi + lo) / 2); pp.half[c] = (float)((hi - lo) / 2);
    }
    trace_begin(ctx, "front");
    static const bool fe_direct = getenv("STC_FRONT_DIRECT") != nullptr;       // A/B switch: the direct-load kernel
    if (monthly_dev && !fe_direct) {
      // staged front end: half[] carries the reciprocal (the product is rounded to fp16 right after)
      for (int c = 0; c < 17; ++c) pp.half[c] = 1.0f / pp.half[c];
      const int gps = cdiv((int64_t)H * W, 32), total = gps * B;
      const int grid = total < 8 * ctx->num_sms ? cdiv(total, 8) : ctx->num_sms;
      if (ctx->monthly_u16) {
        const int smem = 8 * 12 * 32 * 13 * 2 + 64;
        static bool cfg = false;
        if (!cfg) { STC_CUDA(cudaFuncSetAttribute(assemble_prep_staged_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); cfg = true; }
        assemble_prep_staged_kernel<uint16_t><<<grid, 256, smem, ctx->stream>>>(reinterpret_cast<const uint16_t*>(monthly_dev), pp, gps, total);
      } else {
        const int smem = 8 * 12 * 32 * 13 * 4 + 64;
        static bool cfg = false;
        if (!cfg) { STC_CUDA(cudaFuncSetAttribute(assemble_prep_staged_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); cfg = true; }
        assemble_prep_staged_kernel<float><<<grid, 256, smem, ctx->stream>>>(monthly_dev, pp, gps, total);
      }
    } else if (monthly_dev) {
      dim3 agrid(cdiv((int64_t)H * W, 128), B);
      if (ctx->monthly_u16)
        assemble_prep_kernel<uint16_t><<<agrid, 128, 0, ctx->stream>>>(reinterpret_cast<const uint16_t*>(monthly_dev), pp);
      else
        assemble_prep_kernel<float><<<agrid, 128, 0, ctx->stream>>>(monthly_dev, pp);
    } else {
      dim3 grid(cdiv((int64_t)B * pp.Hp * pp.Wp, 256), T1);
      prep_input_kernel<<<grid, 256, 0, ctx->stream>>>(pp);
    }
    STC_CUDA(cudaGetLastError()); ctx->launches++;
    trace_end(ctx);
    if (input_consumed) STC_CUDA(cudaEventRecord(input_consumed, ctx->stream));   // the caller may refill its input buffer
  }
  // ---- bidirectional ConvGRU ----
  const int steps = length < T ? length : T;
  const char* dn[2] = {"fw", "bw"};
  for (int t = 0; t < steps; ++t) {
    int frame[2] = {t, length - 1 - t};
    ConvParams cp; memset(&cp, 0, sizeof(cp));
    for (int d = 0; d < 2; ++d) {
      cp.a0[d] = m->X16.at(0) + (int64_t)frame[d] * m->x_frame_stride;
      cp.a1[d] = m->Hh[d].at(0);
      cp.w[d] = m->w_gates[d];
      cp.out[d] = m->rawG[d].base;
      cp.stats[d] = m->stG + ((size_t)t * 2 + d) * m->Bc * 32;
    }
    cp.a0_plane = m->X16.plane; cp.k0steps = 2; cp.a1_plane = m->Hh[0].plane; cp.k1steps = (t == 0) ? 0 : 2;
    cp.out_plane = m->rawG[0].plane; cp.N = 64; cp.G = 16; cp.mode = MODE_PLAIN;
    cp.out_fp16 = 1;   // GRU pre-norm tensors in fp16: +2e-5 on the probability map (CPU study)
    set_valid(cp, m->Hh[0], true); cp.B = B; cp.Ptot = (int64_t)B * cp.Hp * cp.Wp;
    int rc = launch_conv(ctx, cp, 2); if (rc) return rc;

    GruParams gp; memset(&gp, 0, sizeof(gp));
    for (int d = 0; d < 2; ++d) {
      std::string pre = std::string("gru.") + dn[d] + ".";
      gp.rawG[d] = (const uint4*)m->rawG[d].base; gp.rawY[d] = (const uint4*)m->rawY[d].base; gp.Hf[d] = m->Hf[d].base;
      gp.Hh[d] = m->Hh[d].at(0); gp.RH[d] = m->RH[d].at(0);
      gp.cc[d] = (t == steps - 1) ? m->CCin.at(4 * d) : nullptr;
      gp.stG[d] = m->stG + ((size_t)t * 2 + d) * m->Bc * 32;
      gp.stY[d] = m->stY + ((size_t)t * 2 + d) * m->Bc * 16;
      gp.gam_r[d] = m->fp[pre + "r_gamma"]; gp.bet_r[d] = m->fp[pre + "r_beta"];
      gp.gam_u[d] = m->fp[pre + "u_gamma"]; gp.bet_u[d] = m->fp[pre + "u_beta"];
      gp.gam_y[d] = m->fp[pre + "y_gamma"]; gp.bet_y[d] = m->fp[pre + "y_beta"];
    }
    gp.rawG_plane = m->rawG[0].plane; gp.rawY_plane = m->rawY[0].plane; gp.Hf_plane = m->Hf[0].plane;
    gp.act_plane = m->Hh[0].plane; gp.cc_plane = m->CCin.plane;
    gp.B = B; gp.H = H; gp.W = W; gp.Hp = H + 2; gp.Wp = W + 2; gp.count = (float)H * (float)W * 4.f;
    gp.h_zero = (t == 0);
    dim3 ggrid(cdiv((int64_t)H * W, 256), B, 2);
    if (t > 0) {
      trace_begin(ctx, "apply1");
      gru_apply1_kernel<<<ggrid, 256, 0, ctx->stream>>>(gp);
      trace_end(ctx);
      STC_CUDA(cudaGetLastError()); ctx->launches++;
    }

    for (int d = 0; d < 2; ++d) {
      cp.a1[d] = m->RH[d].at(0);
      cp.w[d] = m->w_cand[d];
      cp.out[d] = m->rawY[d].base;
      cp.stats[d] = m->stY + ((size_t)t * 2 + d) * m->Bc * 16;
      cp.sse_w[d] = m->fp[std::string("gru.") + dn[d] + ".cand_sse_w"];
    }
    cp.out_plane = m->rawY[0].plane; cp.N = 32; cp.G = 8; cp.mode = MODE_CAND;
    rc = launch_conv(ctx, cp, 2); if (rc) return rc;

    trace_begin(ctx, "apply2");
    gru_apply2_kernel<<<ggrid, 256, 0, ctx->stream>>>(gp);
    trace_end(ctx);
    STC_CUDA(cudaGetLastError()); ctx->launches++;
  }
  // ---- optional feature taps (--gen_feats) ----
  const int Ho_f = u3 - 2, Wo_f = v3 - 2;
  if (ctx->feat_early_chunk) {
    int64_t work = (int64_t)B * Ho_f * Wo_f * 8;
    feat_early_kernel<<<cdiv(work, 256), 256, 0, ctx->stream>>>(m->CCin.at(0), m->CCin.plane, B, m->CCin.Hp, m->CCin.Wp, (H - Ho_f) / 2, Ho_f, Wo_f, ctx->feat_early_chunk);
    STC_CUDA(cudaGetLastError()); ctx->launches++;
  }
  // ---- U-Net ----
  int rc;
  Act xmed = m->X16; xmed.base = m->X16.base + (int64_t)T * m->x_frame_stride; xmed.chunks = 4;
  rc = run_block_conv(ctx, m, 0, xmed, 4, true, B); if (rc) return rc;                       // conv_median
  rc = run_apply(ctx, m, 0, m->CCin, true, 0, 0, &m->CCin, 8, H, W, B, nullptr); if (rc) return rc;
  rc = run_block_conv(ctx, m, 1, m->CCin, 16, true, B); if (rc) return rc;                   // conv_concat
  // maxpool -> P1 and the centre crop (skip connection) -> CAT2 from one read of the block's raw output
  rc = run_apply(ctx, m, 1, m->CCin, true, 1, 0, &m->P1, 0, p1, q1, B, nullptr, &m->CAT2, 8, 6, u3, v3); if (rc) return rc;
  rc = run_block_conv(ctx, m, 2, m->P1, 8, false, B); if (rc) return rc;                     // conv1 (VALID)
  rc = run_apply(ctx, m, 2, m->P1, false, 1, 0, &m->P2, 0, p2, q2, B, nullptr, &m->CAT1, 16, 2, u2, v2); if (rc) return rc;
  rc = run_block_conv(ctx, m, 3, m->P2, 16, false, B); if (rc) return rc;                    // conv2 (VALID)
  rc = run_apply(ctx, m, 3, m->P2, false, 2, 0, &m->U2in, 0, u2, v2, B, nullptr); if (rc) return rc;
  rc = run_block_conv(ctx, m, 4, m->U2in, 32, true, B); if (rc) return rc;                   // up2
  rc = run_apply(ctx, m, 4, m->U2in, true, 0, 0, &m->CAT1, 0, u2, v2, B, nullptr); if (rc) return rc;
  rc = run_block_conv(ctx, m, 5, m->CAT1, 32, true, B); if (rc) return rc;                   // up2_out
  rc = run_apply(ctx, m, 5, m->CAT1, true, 2, 0, &m->U3in, 0, u3, v3, B, nullptr); if (rc) return rc;
  rc = run_block_conv(ctx, m, 6, m->U3in, 16, true, B); if (rc) return rc;                   // up3
  rc = run_apply(ctx, m, 6, m->U3in, true, 0, 0, &m->CAT2, 0, u3, v3, B, nullptr); if (rc) return rc;
  rc = run_block_conv(ctx, m, 7, m->CAT2, 16, false, B); if (rc) return rc;                  // out (VALID)
  rc = run_apply(ctx, m, 7, m->CAT2, false, 3, 0, nullptr, 0, u3 - 2, v3 - 2, B, out_dev); if (rc) return rc;
  return STC_OK;
}

// Chunked forward over a device-resident batch.  Consecutive chunks rotate over model_num_slots() scratch slots /
// streams so the HBM-bound elementwise stages of some chunks run in the shadow of the tensor-bound convolutions of
// others (STC_SINGLE_STREAM=1 disables this; STC_SLOTS sets the number of slots).
static int run_chunks(stc_ctx* ctx, const float* x_dev, const float* monthly_dev, int B, int T, int H, int W, int length,
                      int normalize, const double* mn, const double* mx, float* out_dev) {
  if (B <= 0) return STC_OK;
  const char* env = getenv("STC_CHUNK");
  int chunk = env ? atoi(env) : 32;
  if (chunk < 1) chunk = 1;
  const int Bc = B < chunk ? B : chunk;
  const int nchunks = (B + Bc - 1) / Bc;
  int ns = getenv("STC_SINGLE_STREAM") ? 1 : model_num_slots();
  if (ns > nchunks) ns = nchunks;
  for (int i = 0; i < ns; ++i) {
    ModelState* m = (ModelState*)ctx->slots[i];
    if (!m || !m->weights_ready) STC_FAIL(STC_ERR_STATE, "predict: weights not finalized");
  }
  cudaStream_t main_stream = ctx->stream;
  if (ns > 1) {
    if (!ctx->ev_fork) STC_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    STC_CUDA(cudaEventRecord(ctx->ev_fork, main_stream));
    for (int i = 1; i < ns; ++i) {
      if (!ctx->ev_join[i]) STC_CUDA(cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming));
      STC_CUDA(cudaStreamWaitEvent(model_slot_stream(ctx, i), ctx->ev_fork, 0));
    }
  }
  const int Ho = H - 14, Wo = W - 14;
  const size_t per_in = monthly_dev ? (size_t)12 * H * W * 13 : (size_t)(T + 1) * H * W * 17;
  int rc = STC_OK, k = 0;
  for (int b0 = 0; b0 < B && !rc; b0 += Bc, ++k) {
    const int nb = (B - b0) < Bc ? (B - b0) : Bc;
    const int slot = k % ns;
    ModelState* m = (ModelState*)ctx->slots[slot];
    ctx->stream = slot ? model_slot_stream(ctx, slot) : main_stream;
    ctx->cur_slot = slot;
    rc = ensure_plan(ctx, m, Bc, H, W, T + 1);
    const float* mchunk = nullptr;
    if (monthly_dev) mchunk = reinterpret_cast<const float*>(reinterpret_cast<const char*>(monthly_dev) + b0 * per_in * (ctx->monthly_u16 ? 2 : 4));
    ctx->feat_early_chunk = ctx->feat_early_dev ? ctx->feat_early_dev + (size_t)b0 * Ho * Wo * 64 : nullptr;
    ctx->feat_late_chunk = ctx->feat_late_dev ? ctx->feat_late_dev + (size_t)b0 * Ho * Wo * 64 : nullptr;
    if (!rc) rc = forward_chunk(ctx, m, x_dev ? x_dev + b0 * per_in : nullptr, mchunk,
                                nb, T, H, W, length, normalize, mn, mx, out_dev + (size_t)b0 * Ho * Wo, nullptr);
    m->lastB = nb; ctx->last_slot = slot;
  }
  ctx->stream = main_stream;
  ctx->cur_slot = 0;
  ctx->feat_early_chunk = ctx->feat_late_chunk = nullptr;
  if (rc) return rc;
  for (int i = 1; i < ns; ++i) {
    STC_CUDA(cudaEventRecord(ctx->ev_join[i], ctx->slot_stream[i]));
    STC_CUDA(cudaStreamWaitEvent(main_stream, ctx->ev_join[i], 0));
  }
  return STC_OK;
}

int model_predict_dev(stc_ctx* ctx, const float* x_dev, int B, int T, int H, int W, int length,
                      int normalize, const double* min17, const double* max17, float* out_dev) {
  // The network is fully convolutional: any H x W whose sides survive the two pool / crop levels (multiples of 4, >= 28).
  // Square for the released graphs; the border re-segmentation pass predicts 220 x 684 seams (resegment_tiles_wide.py:478).
  if (H % 4 != 0 || H < 28 || W % 4 != 0 || W < 28) STC_FAIL(STC_ERR_ARG, "predict: H and W must be multiples of 4 and >= 28");
  if (T < 1 || T > 12 || length < 1) STC_FAIL(STC_ERR_ARG, "predict: bad T/length");
  if (length > T) STC_FAIL(STC_ERR_ARG, "predict: length exceeds the number of sequence frames (the backward direction would read past them)");
  if (normalize && (!min17 || !max17)) STC_FAIL(STC_ERR_ARG, "predict: normalize needs min/max");
  return run_chunks(ctx, x_dev, nullptr, B, T, H, W, length, normalize, min17, max17, out_dev);
}

int64_t model_debug_read(stc_ctx* ctx, const char* name, float* out_host) {
  ModelState* m = (ModelState*)ctx->slots[ctx->last_slot];
  for (int i = 1; i < stc_ctx::MAX_SLOTS; ++i) if (ctx->slot_stream[i]) cudaStreamSynchronize(ctx->slot_stream[i]);
  if (!m || !m->arena) { ctx->err = "debug_read: no forward pass yet"; return STC_ERR_STATE; }
  std::map<std::string, Act*> tab = {{"ccin", &m->CCin}, {"cat2", &m->CAT2}, {"p1", &m->P1}, {"cat1", &m->CAT1},
                                     {"p2", &m->P2}, {"u2in", &m->U2in}, {"u3in", &m->U3in}, {"hh_fw", &m->Hh[0]},
                                     {"hh_bw", &m->Hh[1]}, {"x16", &m->X16}};
  auto it = tab.find(name);
  if (it == tab.end()) { ctx->err = "debug_read: unknown buffer"; return STC_ERR_ARG; }
  Act& a = *it->second;
  int B = m->lastB, H = a.Hp - 2, W = a.Wp - 2;
  int64_t n = (int64_t)B * H * W * a.chunks * 8;
  if (!out_host) return n;
  float* d = nullptr;
  if (cudaMalloc((void**)&d, n * sizeof(float)) != cudaSuccess) { ctx->err = "debug_read: cudaMalloc"; return STC_ERR_NOMEM; }
  int64_t work = (int64_t)B * H * W * a.chunks;
  act_decode_kernel<<<cdiv(work, 256), 256, 0, ctx->stream>>>(a.at(0), a.plane, a.chunks, B, a.Hp, a.Wp, d);
  cudaError_t e = cudaMemcpyAsync(out_host, d, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) { ctx->err = std::string("debug_read: ") + cudaGetErrorString(e); return STC_ERR_CUDA; }
  return n;
}

// Fused tile path: monthly [B,12,H,W,13] (device) -> probabilities [B,H-14,W-14] (device).
int model_predict_patches_dev(stc_ctx* ctx, const float* monthly_dev, int B, int H, int W,
                              const double* min17, const double* max17, float* out_dev) {
  if (H != W || H % 4 != 0 || H < 28) STC_FAIL(STC_ERR_ARG, "predict_patches: H must equal W, be a multiple of 4 and >= 28");
  if (!min17 || !max17) STC_FAIL(STC_ERR_ARG, "predict_patches: min/max required");
  return run_chunks(ctx, nullptr, monthly_dev, B, 4, H, W, 4, 1, min17, max17, out_dev);
}

// One chunk on scratch slot `slot`, enqueued on that slot's stream.  Used by the host-buffer tile path, which
// interleaves its own H2D copies with the slots; `input_consumed` (optional) is recorded once the front end has read
// monthly_dev.
int model_forward_slot(stc_ctx* ctx, int slot, const float* monthly_dev, int nb, int Bc, int H,
                       const double* min17, const double* max17, float* out_dev, cudaEvent_t input_consumed) {
  if (slot < 0 || slot >= model_num_slots()) STC_FAIL(STC_ERR_ARG, "predict_patches: bad slot");
  ModelState* m = (ModelState*)ctx->slots[slot];
  if (!m || !m->weights_ready) STC_FAIL(STC_ERR_STATE, "predict_patches: weights not finalized");
  if (H % 4 != 0 || H < 28 || nb < 1 || nb > Bc) STC_FAIL(STC_ERR_ARG, "predict_patches: bad shape");
  cudaStream_t saved = ctx->stream;
  ctx->stream = model_slot_stream(ctx, slot);
  ctx->cur_slot = slot;
  int rc = ensure_plan(ctx, m, Bc, H, H, 5);
  if (!rc) rc = forward_chunk(ctx, m, nullptr, monthly_dev, nb, 4, H, H, 4, 1, min17, max17, out_dev, input_consumed);
  ctx->stream = saved;
  ctx->cur_slot = 0;
  if (rc) return rc;
  m->lastB = nb; ctx->last_slot = slot;
  return STC_OK;
}
