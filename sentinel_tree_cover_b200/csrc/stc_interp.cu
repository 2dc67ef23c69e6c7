// Missing-pixel bookkeeping and temporal median fill of the smoothing front end:
// id_missing_px (src/preprocessing/interpolation.py:5-23) and the fill loops of
// deal_w_missing_px (src/download_and_predict_job.py:1039-1047).  Byte/compare work on
// [n,H,W,C] float32 cubes: HBM-bound, one pass per call.
#include "stc_common.cuh"

#define FILL_MAX_DATES 96

namespace {

// per date: number of pixels whose first-10-band count of (==0) + (>=1) exceeds 1, and number of NaN values
__global__ void __launch_bounds__(256) k_missing_counts(const float* __restrict__ arr, int HW, int C, int* __restrict__ bad_px,
                                                        int* __restrict__ nan_vals) {
  const int t = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  int bad = 0, nans = 0;
  if (p < HW) {
    const float* q = arr + ((int64_t)t * HW + p) * C;
    int cnt = 0;
    const int nb = C < 10 ? C : 10;
    for (int c = 0; c < C; ++c) {
      float v = q[c];
      if (c < nb) cnt += (v == 0.0f) + (v >= 1.0f);
      nans += isnan(v);
    }
    bad = cnt > 1;
  }
  unsigned bal = __ballot_sync(0xffffffffu, bad);
  for (int o = 16; o > 0; o >>= 1) nans += __shfl_xor_sync(0xffffffffu, nans, o);
  if ((threadIdx.x & 31) == 0) {
    if (bal) atomicAdd(bad_px + t, __popc(bal));
    if (nans) atomicAdd(nan_vals + t, nans);
  }
}

__device__ __forceinline__ float median_of(const float* v, int n, float* tmp) {   // np.median: NaN propagates
  for (int i = 0; i < n; ++i) { float x = v[i]; if (isnan(x)) return x; tmp[i] = x; }
  for (int i = 1; i < n; ++i) { float x = tmp[i]; int j = i - 1; while (j >= 0 && tmp[j] > x) { tmp[j + 1] = tmp[j]; --j; } tmp[j + 1] = x; }
  return (n & 1) ? tmp[n >> 1] : __fdiv_rn(__fadd_rn(tmp[(n >> 1) - 1], tmp[n >> 1]), 2.f);
}

// One thread per (pixel, channel) column.  The reference recomputes np.median(arr, axis=0) inside the
// date loop, so date i sees dates < i already filled: the loop below is sequential in i on purpose.
// The all-zero pass runs before the all-one pass (two separate loops in the reference).
__global__ void __launch_bounds__(128) k_median_fill(float* __restrict__ arr, int n, int64_t cols) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float v[FILL_MAX_DATES], tmp[FILL_MAX_DATES];
  bool any0 = false, any1 = false;
  for (int i = 0; i < n; ++i) { v[i] = arr[(int64_t)i * cols + c]; any0 |= (v[i] == 0.0f); any1 |= (v[i] == 1.0f); }
  if (!any0 && !any1) return;
  for (int pass = 0; pass < 2; ++pass) {
    const float sentinel = pass ? 1.0f : 0.0f;
    for (int i = 0; i < n; ++i)
      if (v[i] == sentinel) v[i] = median_of(v, n, tmp);
  }
  for (int i = 0; i < n; ++i) arr[(int64_t)i * cols + c] = v[i];
}

}  // namespace

// device-level entry points shared with stc_tilefuse.cu
int interp_missing_counts_dev(stc_ctx* ctx, const float* arr_dev, int n, int HW, int C, int* bad_px_dev, int* nan_vals_dev) {
  { TraceScope ts_(ctx, "k_missing_counts"); k_missing_counts<<<dim3(cdiv((int64_t)HW, 256), n), 256, 0, ctx->stream>>>(arr_dev, HW, C, bad_px_dev, nan_vals_dev); }
  ctx->launches++;
  return STC_OK;
}
int interp_median_fill_dev(stc_ctx* ctx, float* arr_dev, int n, int64_t cols) {
  if (n > FILL_MAX_DATES) STC_FAIL(STC_ERR_ARG, "median_fill: more than 96 dates");
  { TraceScope ts_(ctx, "k_median_fill"); k_median_fill<<<cdiv(cols, 128), 128, 0, ctx->stream>>>(arr_dev, n, cols); }
  ctx->launches++;
  return STC_OK;
}

extern "C" int stc_missing_px_host(stc_ctx* ctx, const float* arr_host, int n, int H, int W, int C, int32_t* bad_px_host,
                                   int32_t* nan_vals_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!arr_host || !bad_px_host || !nan_vals_host || n < 1 || H < 1 || W < 1 || C < 1) STC_FAIL(STC_ERR_ARG, "missing_px: bad argument");
  const int64_t bytes = (int64_t)n * H * W * C * 4;
  float* d = nullptr; int* cnt = nullptr;
  STC_CUDA(stc_dmalloc(&d, bytes)); STC_CUDA(stc_dmalloc(&cnt, 2 * n * 4));
  cudaMemcpyAsync(d, arr_host, bytes, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemsetAsync(cnt, 0, 2 * n * 4, ctx->stream);
  { TraceScope ts_(ctx, "k_missing_counts"); k_missing_counts<<<dim3(cdiv((int64_t)H * W, 256), n), 256, 0, ctx->stream>>>(d, H * W, C, cnt, cnt + n); }
  ctx->launches++;
  cudaMemcpyAsync(bad_px_host, cnt, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
  cudaMemcpyAsync(nan_vals_host, cnt + n, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  stc_dfree(d); stc_dfree(cnt);
  STC_CUDA(e);
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

extern "C" int stc_median_fill_host(stc_ctx* ctx, float* arr_host, int n, int H, int W, int C, int32_t* nan_vals_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!arr_host || !nan_vals_host || n < 1 || n > FILL_MAX_DATES || H < 1 || W < 1 || C < 1)
    STC_FAIL(STC_ERR_ARG, "median_fill: bad argument (1 <= n <= 96)");
  const int64_t cols = (int64_t)H * W * C, bytes = cols * n * 4;
  float* d = nullptr; int* cnt = nullptr;
  STC_CUDA(stc_dmalloc(&d, bytes)); STC_CUDA(stc_dmalloc(&cnt, 2 * n * 4));
  cudaMemcpyAsync(d, arr_host, bytes, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemsetAsync(cnt, 0, 2 * n * 4, ctx->stream);
  { TraceScope ts_(ctx, "k_median_fill"); k_median_fill<<<cdiv(cols, 128), 128, 0, ctx->stream>>>(d, n, cols); }
  { TraceScope ts_(ctx, "k_missing_counts"); k_missing_counts<<<dim3(cdiv((int64_t)H * W, 256), n), 256, 0, ctx->stream>>>(d, H * W, C, cnt, cnt + n); }
  ctx->launches += 2;
  cudaMemcpyAsync(arr_host, d, bytes, cudaMemcpyDeviceToHost, ctx->stream);
  cudaMemcpyAsync(nan_vals_host, cnt + n, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  stc_dfree(d); stc_dfree(cnt);
  STC_CUDA(e);
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

// ---------------------------------------------------------------------------------------------
// 20 m -> 10 m bilinear upsampling, process_tile (src/download_and_predict_job.py:743-782).
// skimage.transform.resize(img, shape, 1) for float input when upsampling is
// scipy.ndimage.zoom(order=1, mode='mirror', grid_mode=True): per axis cc = (k+0.5)*(in/out)-0.5,
// mirrored about 0, taps floor(cc) and floor(cc)+1 (mirrored about in-1), weights w0 = 1-frac,
// w1 = 1-w0, value = sum over the 2x2 taps of ((a*wy)*wx) accumulated row-major in float64, cast
// to float32.  The two 40 m bands are 2x2 mean-pooled first ((a+b)+(c+d))/4 in float32, with the
// odd-shape cases of :760-782 (first row / column copied with repeat(2)).
// ---------------------------------------------------------------------------------------------
namespace {

struct AxisTap { int i0, i1; double w0, w1; };
__device__ __forceinline__ AxisTap axis_tap(int k, int n_in, int n_out) {
  double zoom = (double)n_in / (double)n_out;
  double cc = __dsub_rn(__dmul_rn((double)k + 0.5, zoom), 0.5);
  if (n_in <= 1) cc = 0.0;            // map_coordinate(mirror) with a single sample
  else if (cc < 0.0) cc = -cc;
  double fl = floor(cc);
  AxisTap t; t.i0 = (int)fl; t.i1 = t.i0 + 1;
  if (n_in <= 1) { t.i0 = t.i1 = 0; }
  else {
    if (t.i1 >= n_in) t.i1 = 2 * n_in - 2 - t.i1;
    if (t.i0 >= n_in) t.i0 = 2 * n_in - 2 - t.i0;
  }
  double y = __dsub_rn(cc, fl);
  t.w0 = __dsub_rn(1.0, y); t.w1 = __dsub_rn(1.0, t.w0);
  return t;
}

__global__ void __launch_bounds__(256) k_build_sentinel2(const float* __restrict__ s10, const float* __restrict__ s20, int n, int h, int w,
                                                         float* __restrict__ out) {
  const int H = 2 * h, W = 2 * w;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * H * W) return;
  const int x = (int)(idx % W); const int y = (int)((idx / W) % H); const int t = (int)(idx / ((int64_t)W * H));
  float* o = out + idx * 10;
  const float* a = s10 + idx * 4;
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = a[3];
  const float* src = s20 + (int64_t)t * h * w * 6;
  {
    AxisTap ty = axis_tap(y, h, H), tx = axis_tap(x, w, W);
    for (int b = 0; b < 4; ++b) {
      double acc = 0.0;
      acc = __dadd_rn(acc, __dmul_rn(__dmul_rn((double)src[((int64_t)ty.i0 * w + tx.i0) * 6 + b], ty.w0), tx.w0));
      acc = __dadd_rn(acc, __dmul_rn(__dmul_rn((double)src[((int64_t)ty.i0 * w + tx.i1) * 6 + b], ty.w0), tx.w1));
      acc = __dadd_rn(acc, __dmul_rn(__dmul_rn((double)src[((int64_t)ty.i1 * w + tx.i0) * 6 + b], ty.w1), tx.w0));
      acc = __dadd_rn(acc, __dmul_rn(__dmul_rn((double)src[((int64_t)ty.i1 * w + tx.i1) * 6 + b], ty.w1), tx.w1));
      o[4 + b] = (float)acc;
    }
  }
  const int oy = h & 1, ox = w & 1, ph = (h - oy) / 2, pw = (w - ox) / 2;
  for (int b = 4; b < 6; ++b) {
    float v;
    if (ox && x == 0) v = src[((int64_t)(y >> 1) * w + 0) * 6 + b];              // sentinel2[:, 0] = mid[:, 0].repeat(2) (written last)
    else if (oy && y == 0) v = src[((int64_t)0 * w + (x >> 1)) * 6 + b];         // sentinel2[0, :] = mid[0, :].repeat(2)
    else {
      AxisTap ty = axis_tap(y - oy, ph, H - oy), tx = axis_tap(x - ox, pw, W - ox);
      auto pooled = [&](int i, int j) {
        const float* m = src + ((int64_t)(2 * i + oy) * w + (2 * j + ox)) * 6 + b;
        // np.mean(axis=(1,3)) of the (ph,2,pw,2) view: row pairs first; with a single pooled column NumPy
        // coalesces the 2x2 block into one contiguous run of four and adds it left to right
        const float a = m[0], b2 = m[6], c2 = m[(int64_t)w * 6], d = m[(int64_t)w * 6 + 6];
        const float s4 = (pw == 1) ? __fadd_rn(__fadd_rn(__fadd_rn(a, b2), c2), d) : __fadd_rn(__fadd_rn(a, b2), __fadd_rn(c2, d));
        return __fdiv_rn(s4, 4.f);
      };
      double acc = 0.0;
      acc = __dadd_rn(acc, __dmul_rn(__dmul_rn((double)pooled(ty.i0, tx.i0), ty.w0), tx.w0));
      acc = __dadd_rn(acc, __dmul_rn(__dmul_rn((double)pooled(ty.i0, tx.i1), ty.w0), tx.w1));
      acc = __dadd_rn(acc, __dmul_rn(__dmul_rn((double)pooled(ty.i1, tx.i0), ty.w1), tx.w0));
      acc = __dadd_rn(acc, __dmul_rn(__dmul_rn((double)pooled(ty.i1, tx.i1), ty.w1), tx.w1));
      v = (float)acc;
    }
    o[4 + b] = v;
  }
}

}  // namespace

int interp_build_sentinel2_dev(stc_ctx* ctx, const float* s2_10_dev, const float* s2_20_dev, int n, int h, int w, float* out_dev) {
  const int64_t px = (int64_t)n * 4 * h * w;
  { TraceScope ts_(ctx, "k_build_sentinel2"); k_build_sentinel2<<<cdiv(px, 256), 256, 0, ctx->stream>>>(s2_10_dev, s2_20_dev, n, h, w, out_dev); }
  STC_CUDA(cudaGetLastError()); ctx->launches++;
  return STC_OK;
}

extern "C" int stc_build_sentinel2_host(stc_ctx* ctx, const float* s2_10_host, const float* s2_20_host, int n, int h, int w,
                                        float* out_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!s2_10_host || !s2_20_host || !out_host || n < 1 || h < 2 || w < 2) STC_FAIL(STC_ERR_ARG, "build_sentinel2: bad argument");
  const int64_t px = (int64_t)n * 4 * h * w;
  float *d10 = nullptr, *d20 = nullptr, *dout = nullptr;
  STC_CUDA(stc_dmalloc(&d10, px * 16)); STC_CUDA(stc_dmalloc(&d20, (int64_t)n * h * w * 24)); STC_CUDA(stc_dmalloc(&dout, px * 40));
  cudaMemcpyAsync(d10, s2_10_host, px * 16, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(d20, s2_20_host, (int64_t)n * h * w * 24, cudaMemcpyHostToDevice, ctx->stream);
  { TraceScope ts_(ctx, "k_build_sentinel2"); k_build_sentinel2<<<cdiv(px, 256), 256, 0, ctx->stream>>>(d10, d20, n, h, w, dout); }
  ctx->launches++;
  cudaMemcpyAsync(out_host, dout, px * 40, cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  stc_dfree(d10); stc_dfree(d20); stc_dfree(dout);
  STC_CUDA(e);
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}
