// Fused, device-resident versions of the per-tile steps of process_subtiles (src/download_and_predict_job.py:1125-1486):
// each entry uploads its inputs once, chains the existing kernels on the device and returns only what the caller
// keeps, instead of one host round trip per NumPy statement.  Same arithmetic as the unfused entry points.
//   stc_s2_medians_host        :1148-1159  NaN -> 0, np.median over the raw dates of the 10 bands and of the 4 indices
//   stc_smooth_quarterly_host  :1171-1174, 1274-1278 (+ deal_w_missing_px fill :1039-1047, make_indices, regrid/Whittaker)
//   stc_predict_postprocess_host :1398-1425, 1451-1483  normalise + forward + post-filters for a batch of subtiles
#include "stc_common.cuh"

int interp_missing_counts_dev(stc_ctx* ctx, const float* arr_dev, int n, int HW, int C, int* bad_px_dev, int* nan_vals_dev);
int interp_median_fill_dev(stc_ctx* ctx, float* arr_dev, int n, int64_t cols);
int post_subtiles_dev(stc_ctx* ctx, const float* preds_dev, const float* img_dev, const float* mc_dev, int nimg, int S, int F, int C,
                      unsigned char* a, unsigned char* b, int* d2, double* ramp, unsigned char* na, unsigned char* nb, unsigned char* vote,
                      float* out_dev);
int post_subtile_dev(stc_ctx* ctx, const float* preds_dev, const float* img_dev, const float* mc_dev, int S, int F, int C,
                     unsigned char* a, unsigned char* b, int* d2, double* ramp, unsigned char* na, unsigned char* nb, unsigned char* vote,
                     float* out_dev);

namespace {

struct FBuf { void* p = nullptr; ~FBuf() { if (p) stc_dfree(p); } template <typename T> T* as() { return (T*)p; } };

__global__ void __launch_bounds__(256) k_nan_to_zero(float* __restrict__ x, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && isnan(x[i])) x[i] = 0.f;
}
// out[f][p][0:ca] = a[f][p][:], out[f][p][ca:ca+cb] = b[f][p][:]
__global__ void __launch_bounds__(256) k_concat_channels(const float* __restrict__ a, int ca, const float* __restrict__ b, int cb, int64_t npx,
                                                         float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = ca + cb;
  if (i >= npx * c) return;
  const int64_t p = i / c; const int k = (int)(i % c);
  out[i] = k < ca ? a[p * ca + k] : b[p * cb + (k - ca)];
}
__global__ void __launch_bounds__(256) k_fill_value(float* __restrict__ x, int64_t n, float v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = v;
}

}  // namespace

#define TF_CHECK(call) do { int rc__ = (call); if (rc__) return rc__; } while (0)

// ---- device-resident cores (shared with stc_tile.cu) ----
// s2_dev [n,H,W,10] (NaN -> 0 in place), median14_dev [H,W,14]; bad_px_host[n], *nan_total_host on the host.
// The reference fills the NaNs first (interpolate_na_vals, :1148) and counts the missing pixels afterwards
// (smooth_large_tile -> deal_w_missing_px -> id_missing_px, :1171), so a NaN-filled zero counts as a missing value.
int tf_s2_medians_dev(stc_ctx* ctx, float* s2, int n, int H, int W, float* median14_dev, int32_t* bad_px_host, int64_t* nan_total_host) {
  const int HW = H * W; const int64_t px = (int64_t)n * HW;
  FBuf idx, m10, m4, cnt;
  STC_CUDA(stc_dmalloc(&idx.p, px * 16)); STC_CUDA(stc_dmalloc(&m10.p, (size_t)HW * 40));
  STC_CUDA(stc_dmalloc(&m4.p, (size_t)HW * 16)); STC_CUDA(stc_dmalloc(&cnt.p, 2 * n * 4));
  STC_CUDA(cudaMemsetAsync(cnt.p, 0, 2 * n * 4, ctx->stream));
  TF_CHECK(interp_missing_counts_dev(ctx, s2, n, HW, 10, cnt.as<int>(), cnt.as<int>() + n));
  std::vector<int> h(2 * n);
  STC_CUDA(cudaMemcpyAsync(h.data(), cnt.p, 2 * n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  int64_t nans = 0;
  for (int t = 0; t < n; ++t) nans += h[n + t];
  *nan_total_host = nans;
  if (nans) {            // interpolate_na_vals: NaN -> 0 (the reference fills in place), then the counts see the zeros
    { TraceScope ts_(ctx, "k_nan_to_zero"); k_nan_to_zero<<<cdiv(px * 10, 256), 256, 0, ctx->stream>>>(s2, px * 10); } ctx->launches++;
    STC_CUDA(cudaMemsetAsync(cnt.p, 0, 2 * n * 4, ctx->stream));
    TF_CHECK(interp_missing_counts_dev(ctx, s2, n, HW, 10, cnt.as<int>(), cnt.as<int>() + n));
    STC_CUDA(cudaMemcpyAsync(h.data(), cnt.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    STC_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  for (int t = 0; t < n; ++t) bad_px_host[t] = h[t];
  TF_CHECK(pre_temporal_median_dev(ctx, s2, n, (int64_t)HW * 10, m10.as<float>()));
  TF_CHECK(pre_indices_dev(ctx, s2, px, 10, idx.as<float>()));
  TF_CHECK(pre_temporal_median_dev(ctx, idx.as<float>(), n, (int64_t)HW * 4, m4.as<float>()));
  { TraceScope ts_(ctx, "k_concat_channels"); k_concat_channels<<<cdiv((int64_t)HW * 14, 256), 256, 0, ctx->stream>>>(m10.as<float>(), 10, m4.as<float>(), 4, HW, median14_dev); } ctx->launches++;
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

// s2 [n,H,W,10] device (median-filled in place), M_host [12,n]; s1_dev [12,H,W,2] optional.  Outputs (device, optional):
// s2_monthly [12,H,W,14], s2_quarterly [4,H,W,14], s1_quarterly [4,H,W,2], s1_median [H,W,2]; nan_after_host[n].
// Returns right after the NaN test when a date still holds NaNs (the caller drops it and calls again, :1048-1053).
int tf_smooth_quarterly_dev(stc_ctx* ctx, float* s2, int n, int H, int W, const float* M_host, const float* s1_dev,
                            float* s2_monthly_dev, float* s2_quarterly_dev, float* s1_quarterly_dev, float* s1_median_dev,
                            int32_t* nan_after_host, int skip_fill) {
  const int HW = H * W; const int64_t px = (int64_t)n * HW;
  FBuf idx, sm10, sm4, sm14, cnt;
  STC_CUDA(stc_dmalloc(&cnt.p, 2 * n * 4));
  if (!skip_fill) {
    TF_CHECK(interp_median_fill_dev(ctx, s2, n, (int64_t)HW * 10));                    // deal_w_missing_px :1039-1047
    STC_CUDA(cudaMemsetAsync(cnt.p, 0, 2 * n * 4, ctx->stream));
    TF_CHECK(interp_missing_counts_dev(ctx, s2, n, HW, 10, cnt.as<int>(), cnt.as<int>() + n));
    STC_CUDA(cudaMemcpyAsync(nan_after_host, cnt.as<int>() + n, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    STC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int t = 0; t < n; ++t) if (nan_after_host[t] > 0) return STC_OK;      // caller drops the NaN dates and retries (:1048-1053)
  }
  if ((((uintptr_t)s2) & 15) == 0 && ((int64_t)HW * 40) % 16 == 0) {
    // one fused pass (pre_smooth_fused_dev): indices, the 12 x n operator on all 14 channels, quarterly medians
    TF_CHECK(pre_smooth_fused_dev(ctx, s2, M_host, n, (int64_t)HW, s2_monthly_dev, s2_quarterly_dev));
  } else {
    // odd pixel count: the bulk copies of the fused kernel need 16-byte aligned date slabs -> separate kernels (same arithmetic)
    STC_CUDA(stc_dmalloc(&idx.p, px * 16)); STC_CUDA(stc_dmalloc(&sm10.p, (size_t)12 * HW * 40));
    STC_CUDA(stc_dmalloc(&sm4.p, (size_t)12 * HW * 16));
    float* sm14p = s2_monthly_dev;
    if (!sm14p) { STC_CUDA(stc_dmalloc(&sm14.p, (size_t)12 * HW * 56)); sm14p = sm14.as<float>(); }
    TF_CHECK(pre_indices_dev(ctx, s2, px, 10, idx.as<float>()));                         // make_indices :998
    TF_CHECK(pre_temporal_matmul_dev(ctx, s2, M_host, n, 12, (int64_t)HW * 10, sm10.as<float>()));
    TF_CHECK(pre_temporal_matmul_dev(ctx, idx.as<float>(), M_host, n, 12, (int64_t)HW * 4, sm4.as<float>()));
    { TraceScope ts_(ctx, "k_concat_channels"); k_concat_channels<<<cdiv((int64_t)12 * HW * 14, 256), 256, 0, ctx->stream>>>(sm10.as<float>(), 10, sm4.as<float>(), 4, (int64_t)12 * HW, sm14p); }
    ctx->launches++;
    if (s2_quarterly_dev)
      for (int k = 0; k < 4; ++k)
        TF_CHECK(pre_temporal_median_dev(ctx, sm14p + (size_t)3 * k * HW * 14, 3, (int64_t)HW * 14, s2_quarterly_dev + (size_t)k * HW * 14));
  }
  if (s1_dev) {
    if (s1_quarterly_dev)
      for (int k = 0; k < 4; ++k)
        TF_CHECK(pre_temporal_median_dev(ctx, s1_dev + (size_t)3 * k * HW * 2, 3, (int64_t)HW * 2, s1_quarterly_dev + (size_t)k * HW * 2));
    if (s1_median_dev) TF_CHECK(pre_temporal_median_dev(ctx, s1_dev, 12, (int64_t)HW * 2, s1_median_dev));
  }
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

extern "C" int stc_s2_medians_host(stc_ctx* ctx, float* s2_host, int n, int H, int W, float* median14_host, int32_t* bad_px_host,
                                   int64_t* nan_total_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!s2_host || !median14_host || !bad_px_host || !nan_total_host || n < 1 || H < 1 || W < 1) STC_FAIL(STC_ERR_ARG, "s2_medians: bad argument");
  const int HW = H * W; const int64_t px = (int64_t)n * HW;
  FBuf s2, out;
  STC_CUDA(stc_dmalloc(&s2.p, px * 40)); STC_CUDA(stc_dmalloc(&out.p, (size_t)HW * 56));
  STC_CUDA(cudaMemcpyAsync(s2.p, s2_host, px * 40, cudaMemcpyHostToDevice, ctx->stream));
  TF_CHECK(tf_s2_medians_dev(ctx, s2.as<float>(), n, H, W, out.as<float>(), bad_px_host, nan_total_host));
  if (*nan_total_host) STC_CUDA(cudaMemcpyAsync(s2_host, s2.p, px * 40, cudaMemcpyDeviceToHost, ctx->stream));   // filled in place, like the reference
  STC_CUDA(cudaMemcpyAsync(median14_host, out.p, (size_t)HW * 56, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

extern "C" int stc_smooth_quarterly_host(stc_ctx* ctx, const float* s2_host, int n, int H, int W, const float* M_host, const float* s1_host,
                                         float* s2_monthly_host, float* s2_quarterly_host, float* s1_quarterly_host, float* s1_median_host,
                                         int32_t* nan_after_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!s2_host || !M_host || !nan_after_host || n < 1 || H < 1 || W < 1) STC_FAIL(STC_ERR_ARG, "smooth_quarterly: bad argument");
  const int HW = H * W; const int64_t px = (int64_t)n * HW;
  FBuf s2, sm14, q, s1, s1q, s1m;
  STC_CUDA(stc_dmalloc(&s2.p, px * 40));
  STC_CUDA(cudaMemcpyAsync(s2.p, s2_host, px * 40, cudaMemcpyHostToDevice, ctx->stream));
  if (s2_monthly_host) STC_CUDA(stc_dmalloc(&sm14.p, (size_t)12 * HW * 56));
  if (s2_quarterly_host) STC_CUDA(stc_dmalloc(&q.p, (size_t)4 * HW * 56));
  if (s1_host) {
    STC_CUDA(stc_dmalloc(&s1.p, (size_t)12 * HW * 8)); STC_CUDA(stc_dmalloc(&s1q.p, (size_t)4 * HW * 8)); STC_CUDA(stc_dmalloc(&s1m.p, (size_t)HW * 8));
    STC_CUDA(cudaMemcpyAsync(s1.p, s1_host, (size_t)12 * HW * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  TF_CHECK(tf_smooth_quarterly_dev(ctx, s2.as<float>(), n, H, W, M_host, s1_host ? s1.as<float>() : nullptr, sm14.as<float>(), q.as<float>(),
                                   (s1_host && s1_quarterly_host) ? s1q.as<float>() : nullptr, (s1_host && s1_median_host) ? s1m.as<float>() : nullptr,
                                   nan_after_host, 0));
  for (int t = 0; t < n; ++t) if (nan_after_host[t] > 0) return STC_OK;
  if (s2_monthly_host) STC_CUDA(cudaMemcpyAsync(s2_monthly_host, sm14.p, (size_t)12 * HW * 56, cudaMemcpyDeviceToHost, ctx->stream));
  if (s2_quarterly_host) STC_CUDA(cudaMemcpyAsync(s2_quarterly_host, q.p, (size_t)4 * HW * 56, cudaMemcpyDeviceToHost, ctx->stream));
  if (s1_host && s1_quarterly_host) STC_CUDA(cudaMemcpyAsync(s1_quarterly_host, s1q.p, (size_t)4 * HW * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (s1_host && s1_median_host) STC_CUDA(cudaMemcpyAsync(s1_median_host, s1m.p, (size_t)HW * 8, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

extern "C" int stc_predict_postprocess_host(stc_ctx* ctx, const float* x_host, const float* min_clear_host, const int32_t* no_data_host, int B,
                                            int T, int H, int length, const double* min17, const double* max17, float* out_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!x_host || !min_clear_host || !no_data_host || !out_host || B < 1 || T < 1 || H < 28 || !min17 || !max17)
    STC_FAIL(STC_ERR_ARG, "predict_postprocess: bad argument");
  const int S = H - 14; const size_t per = (size_t)(T + 1) * H * H * 17;
  FBuf x, mc, preds, out, a, b, d2, ramp, na, nb, vote;
  STC_CUDA(stc_dmalloc(&x.p, per * B * 4)); STC_CUDA(stc_dmalloc(&mc.p, (size_t)B * H * H * 4)); STC_CUDA(stc_dmalloc(&preds.p, (size_t)B * S * S * 4));
  STC_CUDA(stc_dmalloc(&out.p, (size_t)B * S * S * 4)); STC_CUDA(stc_dmalloc(&a.p, H * H)); STC_CUDA(stc_dmalloc(&b.p, H * H));
  STC_CUDA(stc_dmalloc(&d2.p, (size_t)H * H * 4)); STC_CUDA(stc_dmalloc(&ramp.p, (size_t)S * S * 8)); STC_CUDA(stc_dmalloc(&na.p, (S + 2) * (S + 2)));
  STC_CUDA(stc_dmalloc(&nb.p, (S + 2) * (S + 2))); STC_CUDA(stc_dmalloc(&vote.p, 256));
  STC_CUDA(cudaMemcpyAsync(x.p, x_host, per * B * 4, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(mc.p, min_clear_host, (size_t)B * H * H * 4, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  // forward for the whole batch (normalize_subtile fused into the input packing; x stays un-normalised for the post-filters)
  TF_CHECK(model_predict_dev(ctx, x.as<float>(), B, T, H, H, length, 1, min17, max17, preds.as<float>()));
  for (int i = 0; i < B; ++i) {
    float* p = preds.as<float>() + (size_t)i * S * S;
    if (no_data_host[i]) { { TraceScope ts_(ctx, "k_fill_value"); k_fill_value<<<cdiv(S * S, 256), 256, 0, ctx->stream>>>(p, (int64_t)S * S, 255.f); } ctx->launches++; }   // np.full((SIZE, SIZE), 255)
    TF_CHECK(post_subtile_dev(ctx, p, x.as<float>() + per * i, mc.as<float>() + (size_t)i * H * H, S, T + 1, 17, a.as<unsigned char>(),
                              b.as<unsigned char>(), d2.as<int>(), ramp.as<double>(), na.as<unsigned char>(), nb.as<unsigned char>(),
                              vote.as<unsigned char>(), out.as<float>() + (size_t)i * S * S));
  }
  STC_CUDA(cudaMemcpyAsync(out_host, out.p, (size_t)B * S * S * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

// ---------------------------------------------------------------------------------------------
// stc_process_subtiles_host: the subtile loop of process_subtiles (:1345-1486) for one tile, on the device:
// window gather with the reference's reflect padding -> 17-channel stacks -> no-image test -> normalise + batched
// forward -> post-filters.  The host passes only the window table (integers) and gets the predictions back.
// ---------------------------------------------------------------------------------------------
namespace {

struct Win { int r0, c0, nr, nc, pr0, pr1, pc0, pc1, mr0, mr1, mc0, mc1; };   // data window, data pads (rows, cols), min_clear pads

__device__ __forceinline__ int reflect_idx(int i, int pad0, int len) {      // np.pad(..., 'reflect'): no edge repeat
  int j = i - pad0;
  if (j < 0) j = -j;
  if (j >= len) j = 2 * (len - 1) - j;
  return j;
}

// x [nt][T+1][P][P][17], P = S + 14
__global__ void __launch_bounds__(256) k_gather_subtiles(const float* __restrict__ s2q, const float* __restrict__ s1q,
                                                         const float* __restrict__ s2med, const float* __restrict__ s1med,
                                                         const float* __restrict__ dem, const Win* __restrict__ wins, int T, int H, int W,
                                                         int P, float* __restrict__ x) {
  // one thread per output float: the 68-byte pixel records are written coalesced (a thread per pixel scattered its 17 stores)
  const int t = blockIdx.z, f = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P * P * 17) return;
  const int i = e / 17, k = e - i * 17;
  const Win w = wins[t];
  const int r = w.r0 + reflect_idx(i / P, w.pr0, w.nr), c = w.c0 + reflect_idx(i % P, w.pc0, w.nc);
  const int64_t px = (int64_t)r * W + c;
  const float* b14 = (f < T) ? s2q + ((int64_t)f * H * W + px) * 14 : s2med + px * 14;
  const float* b2 = (f < T) ? s1q + ((int64_t)f * H * W + px) * 2 : s1med + px * 2;
  const float v = k < 10 ? b14[k] : k == 10 ? dem[px] : k < 13 ? b2[k - 11] : b14[k - 3];
  x[(((int64_t)t * (T + 1) + f) * P * P) * 17 + e] = v;
}
// min_clear maps [nt][P][P] (float32) with their own pads + the no-image test on the UNPADDED window (:1355-1357):
// np.percentile(min_clear, 50) < 1  <=>  both middle order statistics average below 1 (integers >= 0)
__global__ void __launch_bounds__(256) k_gather_clear(const int* __restrict__ clear, const Win* __restrict__ wins, int W, int P,
                                                      float* __restrict__ mc) {
  const int t = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * P) return;
  const Win w = wins[t];
  const int r = w.r0 + reflect_idx(i / P, w.mr0, w.nr), c = w.c0 + reflect_idx(i % P, w.mc0, w.nc);
  mc[(int64_t)t * P * P + i] = (float)clear[(int64_t)r * W + c];
}
__global__ void __launch_bounds__(256) k_no_image_test(const int* __restrict__ clear, const Win* __restrict__ wins, int W, int force,
                                                       int* __restrict__ flags) {
  const int t = blockIdx.x;
  const Win w = wins[t];
  __shared__ int c0, c01;
  if (threadIdx.x == 0) { c0 = 0; c01 = 0; }
  __syncthreads();
  int a = 0, b = 0;
  const int n = w.nr * w.nc;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int v = clear[(int64_t)(w.r0 + i / w.nc) * W + w.c0 + i % w.nc];
    a += (v == 0); b += (v <= 1);
  }
  atomicAdd(&c0, a); atomicAdd(&c01, b);
  __syncthreads();
  if (threadIdx.x == 0) {
    bool none;
    if (n & 1) none = c0 >= (n + 1) / 2;                       // the middle value is 0
    else none = (c0 >= n / 2) && (c01 >= n / 2 + 1);           // sorted[n/2-1] == 0 and sorted[n/2] <= 1: mean < 1
    flags[t] = (none || force) ? 1 : 0;
  }
}
__global__ void __launch_bounds__(256) k_fill_if(float* __restrict__ preds, const int* __restrict__ flags, int per, float v) {
  const int t = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < per && flags[t]) preds[(int64_t)t * per + i] = v;
}

}  // namespace

// Device-resident core of the subtile loop: all array pointers on the device, windows / min17 / max17 / flags on the host.
int tf_process_subtiles_dev(stc_ctx* ctx, const float* s2q, const float* s1q, const float* s2m, const float* s1m, const float* dem,
                            const int* clr, int H, int W, int nt, const int32_t* windows_host, int S, int T, int length,
                            int force_no_data, const double* min17, const double* max17, float* out_dev, int32_t* no_data_host,
                            float* early_dev, float* late_dev) {
  const int P = S + 14;
  for (int t = 0; t < nt; ++t) {
    const int32_t* w = windows_host + 12 * t;
    if (w[0] < 0 || w[1] < 0 || w[2] < 8 || w[3] < 8 || w[0] + w[2] > H || w[1] + w[3] > W) STC_FAIL(STC_ERR_ARG, "process_subtiles: window outside the tile");
    if (w[4] + w[5] + w[2] != P || w[6] + w[7] + w[3] != P || w[8] + w[9] + w[2] != P || w[10] + w[11] + w[3] != P)
      STC_FAIL(STC_ERR_ARG, "process_subtiles: window + padding does not give a (size+14)^2 subtile (the reference fails in its reshape here)");
    for (int k = 4; k < 12; ++k) if (w[k] < 0 || w[k] >= 8) STC_FAIL(STC_ERR_ARG, "process_subtiles: bad padding");
  }
  FBuf win, x, mc, flags, preds, a, b, d2, ramp, na, nb, vote;
  const size_t per = (size_t)(T + 1) * P * P * 17;
  STC_CUDA(stc_dmalloc(&win.p, (size_t)nt * sizeof(Win))); STC_CUDA(stc_dmalloc(&x.p, per * nt * 4)); STC_CUDA(stc_dmalloc(&mc.p, (size_t)nt * P * P * 4));
  STC_CUDA(stc_dmalloc(&flags.p, nt * 4)); STC_CUDA(stc_dmalloc(&preds.p, (size_t)nt * S * S * 4));
  STC_CUDA(stc_dmalloc(&a.p, (size_t)nt * P * P)); STC_CUDA(stc_dmalloc(&b.p, (size_t)nt * P * P)); STC_CUDA(stc_dmalloc(&d2.p, (size_t)nt * P * P * 4));
  STC_CUDA(stc_dmalloc(&ramp.p, (size_t)nt * S * S * 8)); STC_CUDA(stc_dmalloc(&na.p, (size_t)nt * (S + 2) * (S + 2)));
  STC_CUDA(stc_dmalloc(&nb.p, (size_t)nt * (S + 2) * (S + 2))); STC_CUDA(stc_dmalloc(&vote.p, (size_t)nt * 256));
  static_assert(sizeof(Win) == 48, "window table layout");
  STC_CUDA(cudaMemcpyAsync(win.p, windows_host, (size_t)nt * 48, cudaMemcpyHostToDevice, ctx->stream));
  { TraceScope ts_(ctx, "k_gather_subtiles"); k_gather_subtiles<<<dim3(cdiv(P * P * 17, 256), T + 1, nt), 256, 0, ctx->stream>>>(s2q, s1q, s2m, s1m, dem, win.as<Win>(), T, H, W, P, x.as<float>()); }
  { TraceScope ts_(ctx, "k_gather_clear"); k_gather_clear<<<dim3(cdiv(P * P, 256), nt), 256, 0, ctx->stream>>>(clr, win.as<Win>(), W, P, mc.as<float>()); }
  { TraceScope ts_(ctx, "k_no_image_test"); k_no_image_test<<<nt, 256, 0, ctx->stream>>>(clr, win.as<Win>(), W, force_no_data, flags.as<int>()); }
  ctx->launches += 3;
  STC_CUDA(cudaStreamSynchronize(ctx->stream));      // the caller's window table may go away
  ctx->feat_early_dev = early_dev; ctx->feat_late_dev = late_dev;       // --gen_feats taps [nt,S,S,64] (optional)
  const int rc_fwd = model_predict_dev(ctx, x.as<float>(), nt, T, P, P, length, 1, min17, max17, preds.as<float>());
  ctx->feat_early_dev = ctx->feat_late_dev = nullptr;
  if (rc_fwd) return rc_fwd;
  { TraceScope ts_(ctx, "k_fill_if"); k_fill_if<<<dim3(cdiv(S * S, 256), nt), 256, 0, ctx->stream>>>(preds.as<float>(), flags.as<int>(), S * S, 255.f); } ctx->launches++;
  TF_CHECK(post_subtiles_dev(ctx, preds.as<float>(), x.as<float>(), mc.as<float>(), nt, S, T + 1, 17, a.as<unsigned char>(), b.as<unsigned char>(),
                             d2.as<int>(), ramp.as<double>(), na.as<unsigned char>(), nb.as<unsigned char>(), vote.as<unsigned char>(), out_dev));
  STC_CUDA(cudaMemcpyAsync(no_data_host, flags.p, nt * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

static int process_subtiles_host_impl(stc_ctx* ctx, const float* s2q_host, const float* s1q_host, const float* s2med_host,
                                      const float* s1med_host, const float* dem_host, const int32_t* clear_host, int H, int W, int nt,
                                      const int32_t* windows_host /*[nt][12]*/, int S, int T, int length, int force_no_data,
                                      const double* min17, const double* max17, float* out_host, int32_t* no_data_host,
                                      float* early_host, float* late_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!s2q_host || !s1q_host || !s2med_host || !s1med_host || !dem_host || !clear_host || !windows_host || !out_host || !no_data_host ||
      nt < 1 || T < 1 || S < 14 || !min17 || !max17 || (!early_host != !late_host))
    STC_FAIL(STC_ERR_ARG, "process_subtiles: bad argument");
  const int64_t HW = (int64_t)H * W;
  FBuf s2q, s1q, s2m, s1m, dem, clr, out, fe, fl;
  STC_CUDA(stc_dmalloc(&s2q.p, (size_t)T * HW * 56)); STC_CUDA(stc_dmalloc(&s1q.p, (size_t)T * HW * 8)); STC_CUDA(stc_dmalloc(&s2m.p, HW * 56));
  STC_CUDA(stc_dmalloc(&s1m.p, HW * 8)); STC_CUDA(stc_dmalloc(&dem.p, HW * 4)); STC_CUDA(stc_dmalloc(&clr.p, HW * 4));
  STC_CUDA(stc_dmalloc(&out.p, (size_t)nt * S * S * 4));
  const size_t fbytes = (size_t)nt * S * S * 64 * 4;
  if (early_host) { STC_CUDA(stc_dmalloc(&fe.p, fbytes)); STC_CUDA(stc_dmalloc(&fl.p, fbytes)); }
  STC_CUDA(cudaMemcpyAsync(s2q.p, s2q_host, (size_t)T * HW * 56, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(s1q.p, s1q_host, (size_t)T * HW * 8, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(s2m.p, s2med_host, HW * 56, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(s1m.p, s1med_host, HW * 8, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(dem.p, dem_host, HW * 4, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(clr.p, clear_host, HW * 4, cudaMemcpyHostToDevice, ctx->stream));
  TF_CHECK(tf_process_subtiles_dev(ctx, s2q.as<float>(), s1q.as<float>(), s2m.as<float>(), s1m.as<float>(), dem.as<float>(), clr.as<int>(), H, W,
                                   nt, windows_host, S, T, length, force_no_data, min17, max17, out.as<float>(), no_data_host,
                                   fe.as<float>(), fl.as<float>()));
  STC_CUDA(cudaMemcpyAsync(out_host, out.p, (size_t)nt * S * S * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (early_host) {
    STC_CUDA(cudaMemcpyAsync(early_host, fe.p, fbytes, cudaMemcpyDeviceToHost, ctx->stream));
    STC_CUDA(cudaMemcpyAsync(late_host, fl.p, fbytes, cudaMemcpyDeviceToHost, ctx->stream));
  }
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

extern "C" int stc_process_subtiles_host(stc_ctx* ctx, const float* s2q_host, const float* s1q_host, const float* s2med_host,
                                         const float* s1med_host, const float* dem_host, const int32_t* clear_host, int H, int W, int nt,
                                         const int32_t* windows_host, int S, int T, int length, int force_no_data,
                                         const double* min17, const double* max17, float* out_host, int32_t* no_data_host) {
  return process_subtiles_host_impl(ctx, s2q_host, s1q_host, s2med_host, s1med_host, dem_host, clear_host, H, W, nt, windows_host, S, T, length,
                                    force_no_data, min17, max17, out_host, no_data_host, nullptr, nullptr);
}

extern "C" int stc_process_subtiles_feats_host(stc_ctx* ctx, const float* s2q_host, const float* s1q_host, const float* s2med_host,
                                               const float* s1med_host, const float* dem_host, const int32_t* clear_host, int H, int W, int nt,
                                               const int32_t* windows_host, int S, int T, int length, int force_no_data,
                                               const double* min17, const double* max17, float* out_host, int32_t* no_data_host,
                                               float* early_host, float* late_host) {
  if (!early_host || !late_host) { if (ctx) ctx->err = "process_subtiles_feats: feature outputs missing"; return STC_ERR_ARG; }
  return process_subtiles_host_impl(ctx, s2q_host, s1q_host, s2med_host, s1med_host, dem_host, clear_host, H, W, nt, windows_host, S, T, length,
                                    force_no_data, min17, max17, out_host, no_data_host, early_host, late_host);
}
