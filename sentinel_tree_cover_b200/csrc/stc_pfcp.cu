// Potential false-positive cloud pixels over built-up land: detect_pfcp (src/preprocessing/cloud_removal.py:1109-1212,
// the Fmask-4.0 parallax test) with the urban raster handed over as arrays (stc_set_ancillary_masks_host).
//   pfps  = median_t(NDBI > 0 and NDBI > NDVI) * (median_t NDWI < 0); urban core -> 1; farther than the dilated raster -> 0;
//           dem / 90 > 0.10 -> 0                                                                     (:1124-1141)
//   cdi_t = (V(B7/B8A) - V(B8/B8A)) / (V(B7/B8A) + V(B8/B8A)) at 20 m, V = 7x7 variance (E[x^2] - E[x]^2), B8 first
//           blurred with a Gaussian (sigma 0.5, truncate 3: 5 taps) at 10 m; cdis_t = (cdi >= -0.4) up-sampled, * (NDVI < 0.4) (:1143-1203)
//   both dilated 6 times with the 3x3 element; fcps = pfps * cdis                                    (:1205-1212)
// The float arithmetic follows what NumPy / SciPy execute in the reference environment, operation by operation:
//   * scipy.ndimage.gaussian_filter on float32: per axis (0 then 1) a float64 correlate1d that adds the symmetric pairs first,
//     w2*x0 + (x-2 + x+2)*w0 + (x-1 + x+1)*w1, rounded to float32 after each axis; 'reflect' = edge-repeating mirror;
//   * np.mean(x.reshape(h, 2, w, 2), axis=(1, 3)) on float32 = ((a + b) + (c + d)) / 4;
//   * scipy.signal.convolve2d(x, ones(7,7)/49, 'same', 'symm') in float64: kernel rows bottom-up, columns right-to-left,
//     per row ((p0 + p1) + p2) + p3 added to the running sum, then p4, p5, p6 one by one (the compiled inner loop is
//     4-wide; established against SciPy 1.18.1 bit for bit, tools/make_golden_cloud_anc.py / tests/test_cloud_masks.py).
// Odd tile sides: the reference first grows the band to even size with an order-0 resize and shrinks the up-sampled
// flags back the same way; the index maps of those two resizes come from the host (nn_index).
#include "stc_common.cuh"
#include <vector>
#include <cmath>

namespace {

__device__ __forceinline__ int symm(int i, int n) {          // edge-repeating mirror, |overshoot| < n
  if (i < 0) i = -1 - i;
  if (i >= n) i = 2 * n - 1 - i;
  return i;
}

// ---- pfps before the dilation: one thread per 10 m pixel ----
__global__ void __launch_bounds__(128) k_pfps_base(const float* __restrict__ img, const float* __restrict__ dem,
                                                   const unsigned char* __restrict__ core, const unsigned char* __restrict__ near_, int T,
                                                   int HW, unsigned char* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  float w[32];
  int cnt = 0; bool nan_w = false;
  for (int t = 0; t < T; ++t) {
    const float* x = img + ((int64_t)t * HW + p) * 10;
    const float ndvi = __fdiv_rn(__fsub_rn(x[3], x[2]), __fadd_rn(x[3], x[2]));
    const float ndbi = __fdiv_rn(__fsub_rn(x[8], x[3]), __fadd_rn(x[8], x[3]));
    const float ndwi = __fdiv_rn(__fsub_rn(x[1], x[3]), __fadd_rn(x[1], x[3]));
    cnt += (ndbi > 0.f) && (ndbi > ndvi);
    nan_w = nan_w || isnan(ndwi);
    w[t] = ndwi;
  }
  bool v = cnt >= (T + 1) / 2;                               // np.median of the T booleans is > 0
  if (v) {                                                   // * (np.median(ndwi) < 0): NaN propagates -> False
    if (nan_w) v = false;
    else {
      for (int i = 1; i < T; ++i) { float xx = w[i]; int j = i - 1; while (j >= 0 && w[j] > xx) { w[j + 1] = w[j]; --j; } w[j + 1] = xx; }
      const float med = (T & 1) ? w[T >> 1] : __fmul_rn(__fadd_rn(w[(T >> 1) - 1], w[T >> 1]), 0.5f);
      v = med < 0.f;
    }
  }
  if (core[p] == 1) v = true;
  if (near_[p] == 0) v = false;
  if (__fdiv_rn(dem[p], 90.f) > 0.10f) v = false;
  out[p] = v;
}

// scipy.ndimage.gaussian_filter(sigma=0.5, truncate=3) weights: exp(-2 k^2) / sum, float64 (k = -2..2)
__device__ __constant__ double GW[3] = {0x1.14aebe6a24088p-12, 0x1.b405b9842b206p-4, 0x1.92b965ef5aaeep-1};   // |k| = 2, 1, 0

// axis-0 pass of the blur on band B8 grown to (H2, W2): tmp[t][y][x] float32
__global__ void __launch_bounds__(256) k_gauss_rows(const float* __restrict__ img, int T, int H, int W, int H2, int W2,
                                                    const int* __restrict__ ru, const int* __restrict__ cu, float* __restrict__ tmp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)T * H2 * W2) return;
  const int x = (int)(i % W2); const int64_t r = i / W2; const int y = (int)(r % H2); const int t = (int)(r / H2);
  const float* b = img + (int64_t)t * H * W * 10 + 3;
  const int sx = cu[x];
  auto at = [&](int yy) { return (double)b[((int64_t)ru[symm(yy, H2)] * W + sx) * 10]; };
  double acc = __dmul_rn(at(y), GW[2]);
  acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(at(y - 2), at(y + 2)), GW[0]));
  acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(at(y - 1), at(y + 1)), GW[1]));
  tmp[i] = (float)acc;
}

// axis-1 pass + 2x2 mean pools of B8 (blurred), B8A, B7 -> the two ratio images at 20 m (float32)
__global__ void __launch_bounds__(256) k_pool_ratios(const float* __restrict__ img, const float* __restrict__ tmp, int T, int H, int W, int H2,
                                                     int W2, const int* __restrict__ ru, const int* __restrict__ cu,
                                                     float* __restrict__ r8a, float* __restrict__ r87) {
  const int h2 = H2 / 2, w2 = W2 / 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)T * h2 * w2) return;
  const int x = (int)(i % w2); const int64_t r = i / w2; const int y = (int)(r % h2); const int t = (int)(r / h2);
  const float* g = tmp + (int64_t)t * H2 * W2;
  const float* b = img + (int64_t)t * H * W * 10;
  auto blur = [&](int yy, int xx) {
    const float* row = g + (int64_t)yy * W2;
    double acc = __dmul_rn((double)row[xx], GW[2]);
    acc = __dadd_rn(acc, __dmul_rn(__dadd_rn((double)row[symm(xx - 2, W2)], (double)row[symm(xx + 2, W2)]), GW[0]));
    acc = __dadd_rn(acc, __dmul_rn(__dadd_rn((double)row[symm(xx - 1, W2)], (double)row[symm(xx + 1, W2)]), GW[1]));
    return (float)acc;
  };
  auto raw = [&](int yy, int xx, int band) { return b[((int64_t)ru[yy] * W + cu[xx]) * 10 + band]; };
  auto pool = [&](float a, float bb, float c, float d) { return __fdiv_rn(__fadd_rn(__fadd_rn(a, bb), __fadd_rn(c, d)), 4.f); };
  const int y0 = 2 * y, x0 = 2 * x;
  const float b8 = pool(blur(y0, x0), blur(y0, x0 + 1), blur(y0 + 1, x0), blur(y0 + 1, x0 + 1));
  const float b8a = pool(raw(y0, x0, 7), raw(y0, x0 + 1, 7), raw(y0 + 1, x0, 7), raw(y0 + 1, x0 + 1, 7));
  const float b7 = pool(raw(y0, x0, 6), raw(y0, x0 + 1, 6), raw(y0 + 1, x0, 6), raw(y0 + 1, x0 + 1, 6));
  r8a[i] = __fdiv_rn(b8, b8a);
  r87[i] = __fdiv_rn(b7, b8a);
}

// 7x7 box mean of f(x) in SciPy's accumulation order (see the header); sq: f = x*x in float32 first
__device__ __forceinline__ double box49(const float* __restrict__ a, int y, int x, int h, int w, bool sq) {
  const double c = 1.0 / 49.0;
  double d = 0.0;
  for (int j = 0; j < 7; ++j) {
    const float* row = a + (int64_t)symm(y + 3 - j, h) * w;
    double p[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      float v = row[symm(x + 3 - k, w)];
      if (sq) v = __fmul_rn(v, v);
      p[k] = __dmul_rn((double)v, c);
    }
    d = __dadd_rn(d, __dadd_rn(__dadd_rn(__dadd_rn(p[0], p[1]), p[2]), p[3]));
    d = __dadd_rn(d, p[4]); d = __dadd_rn(d, p[5]); d = __dadd_rn(d, p[6]);
  }
  return d;
}

__global__ void __launch_bounds__(128) k_cdi_flag(const float* __restrict__ r8a, const float* __restrict__ r87, int T, int h2, int w2,
                                                  unsigned char* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)T * h2 * w2) return;
  const int x = (int)(i % w2); const int64_t r = i / w2; const int y = (int)(r % h2); const int t = (int)(r / h2);
  const float* a = r8a + (int64_t)t * h2 * w2; const float* b = r87 + (int64_t)t * h2 * w2;
  const double ma = box49(a, y, x, h2, w2, false), mb = box49(b, y, x, h2, w2, false);
  const double va = __dsub_rn(box49(a, y, x, h2, w2, true), __dmul_rn(ma, ma));
  const double vb = __dsub_rn(box49(b, y, x, h2, w2, true), __dmul_rn(mb, mb));
  const double cdi = __ddiv_rn(__dsub_rn(vb, va), __dadd_rn(vb, va));
  flag[i] = cdi >= -0.4;
}

// cdis_t = up-sampled flag * (NDVI_t < 0.4)
__global__ void __launch_bounds__(256) k_cdis(const float* __restrict__ img, const unsigned char* __restrict__ flag, int T, int H, int W,
                                              int h2, int w2, const int* __restrict__ rd, const int* __restrict__ cd,
                                              unsigned char* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)T * H * W) return;
  const int x = (int)(i % W); const int64_t r = i / W; const int y = (int)(r % H); const int t = (int)(r / H);
  const float* q = img + i * 10;
  const float ndvi = __fdiv_rn(__fsub_rn(q[3], q[2]), __fadd_rn(q[3], q[2]));
  out[i] = flag[((int64_t)t * h2 + rd[y] / 2) * w2 + cd[x] / 2] && (ndvi < 0.4f);
}

__global__ void __launch_bounds__(256) k_and_frame(const unsigned char* __restrict__ cd, const unsigned char* __restrict__ pf, int HW, int64_t N,
                                                   unsigned char* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] = cd[i] && pf[i % HW];
}

// source index of an order-0 resize from n_in to n_out samples (scipy.ndimage.zoom(order=0, grid_mode=True,
// mode='nearest'), the stand-in for skimage.transform.resize(x, shape, 0)): nearest sample to (i + 0.5) * n_in / n_out - 0.5
std::vector<int> nn_index(int n_out, int n_in) {
  std::vector<int> m(n_out);
  const double zoom = (double)n_in / (double)n_out;
  for (int i = 0; i < n_out; ++i) {
    double c = ((double)i + 0.5) * zoom - 0.5;
    int k = (int)std::floor(c + 0.5);
    m[i] = k < 0 ? 0 : (k >= n_in ? n_in - 1 : k);
  }
  return m;
}

}  // namespace

// fcps_out [T][HW] uint8 = pfps * cdis (both after the 6-fold 3x3 dilation), pfps_out [HW] uint8 (the same for every date)
int pfcp_detect_dev(stc_ctx* ctx, const float* img, const float* dem, const unsigned char* urban_core, const unsigned char* urban_near, int T,
                    int H, int W, unsigned char* fcps_out, unsigned char* pfps_out) {
  const int HW = H * W, H2 = H + (H & 1), W2 = W + (W & 1), h2 = H2 / 2, w2 = W2 / 2;
  const int64_t N = (int64_t)T * HW;
  std::vector<int> maps;
  for (const std::vector<int>& m : {nn_index(H2, H), nn_index(W2, W), nn_index(H, H2), nn_index(W, W2)}) maps.insert(maps.end(), m.begin(), m.end());
  PoolBuf d_maps, d_tmp, d_r8a, d_r87, d_flag, d_pf0, d_cd;
  STC_CUDA(d_maps.alloc(maps.size() * 4)); STC_CUDA(d_tmp.alloc((size_t)T * H2 * W2 * 4));
  STC_CUDA(d_r8a.alloc((size_t)T * h2 * w2 * 4)); STC_CUDA(d_r87.alloc((size_t)T * h2 * w2 * 4)); STC_CUDA(d_flag.alloc((size_t)T * h2 * w2));
  STC_CUDA(d_pf0.alloc((size_t)HW)); STC_CUDA(d_cd.alloc((size_t)N));
  const void* hm = ctx_stage(ctx, maps.data(), maps.size() * 4);
  if (!hm) STC_FAIL(STC_ERR_NOMEM, "pfcp: pinned staging");
  STC_CUDA(cudaMemcpyAsync(d_maps.p, hm, maps.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  const int* ru = d_maps.as<int>(); const int* cu = ru + H2; const int* rd = cu + W2; const int* cd = rd + H;
  { TraceScope ts_(ctx, "k_pfps_base"); k_pfps_base<<<cdiv(HW, 128), 128, 0, ctx->stream>>>(img, dem, urban_core, urban_near, T, HW, d_pf0.as<unsigned char>()); }
  { TraceScope ts_(ctx, "k_gauss_rows"); k_gauss_rows<<<cdiv((int64_t)T * H2 * W2, 256), 256, 0, ctx->stream>>>(img, T, H, W, H2, W2, ru, cu, d_tmp.as<float>()); }
  { TraceScope ts_(ctx, "k_pool_ratios"); k_pool_ratios<<<cdiv((int64_t)T * h2 * w2, 256), 256, 0, ctx->stream>>>(img, d_tmp.as<float>(), T, H, W, H2, W2, ru, cu,
                                                                                               d_r8a.as<float>(), d_r87.as<float>()); }
  { TraceScope ts_(ctx, "k_cdi_flag"); k_cdi_flag<<<cdiv((int64_t)T * h2 * w2, 128), 128, 0, ctx->stream>>>(d_r8a.as<float>(), d_r87.as<float>(), T, h2, w2, d_flag.as<unsigned char>()); }
  { TraceScope ts_(ctx, "k_cdis"); k_cdis<<<cdiv(N, 256), 256, 0, ctx->stream>>>(img, d_flag.as<unsigned char>(), T, H, W, h2, w2, rd, cd, fcps_out); }
  ctx->launches += 5;
  int rc;
  if ((rc = morph_dilate_dev(ctx, fcps_out, d_cd.as<unsigned char>(), T, H, W, 6, 2, 0, 0, 0))) return rc;
  if ((rc = morph_dilate_dev(ctx, d_pf0.as<unsigned char>(), pfps_out, 1, H, W, 6, 2, 0, 0, 0))) return rc;
  { TraceScope ts_(ctx, "k_and_frame"); k_and_frame<<<cdiv(N, 256), 256, 0, ctx->stream>>>(d_cd.as<unsigned char>(), pfps_out, HW, N, fcps_out); }
  ctx->launches++;
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}
