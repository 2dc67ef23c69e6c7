// Array work of process_tile's front half (src/download_and_predict_job.py:684-832, 995-997) that is not
// already covered by the codec / upsampling / cloud kernels: Sentinel-1 saturated-value fill, DEM 5x5
// median filter, the Sen2Cor consecutive-date rule, the snow mask, per-date threshold counts, clip / scale.
// All HBM-bound single-pass kernels.
#include "stc_common.cuh"
#include "stc_select.cuh"
#include <vector>

void maskop_dilate(stc_ctx* ctx, const unsigned char* in, unsigned char* out, int frames, int H, int W, int k, int conn, int inv_in,
                   int inv_out, int three_d);

namespace {

struct TBuf { void* p = nullptr; ~TBuf() { if (p) stc_dfree(p); } template <typename T> T* as() { return (T*)p; } };

// s1_i[s1_i == 1] = np.median(s1_i[s1_i < 65535]) (:702-705): the median runs over ALL values of the date (both
// polarisations; every value is < 65535 after the /65535 scaling), then replaces the saturated ones.  The per-date
// medians come from the GPU-wide radix select (stc_select.cu); round 1 ran one block per date (12 blocks, 1.5 ms).
__global__ void __launch_bounds__(256) k_s1_fill(float* __restrict__ s1, int len, const float* __restrict__ pairs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  float* a = s1 + (int64_t)blockIdx.y * len;
  if (a[i] != 1.0f) return;
  const float lo = pairs[blockIdx.y * SEL_MAX_COLS * 2], hi = pairs[blockIdx.y * SEL_MAX_COLS * 2 + 1];
  a[i] = (len & 1) ? lo : __fdiv_rn(__fadd_rn(lo, hi), 2.f);
}

// scipy.ndimage.median_filter(dem, size=5), mode='reflect' (d c b a | a b c d): rank 12 of the 25 window values
__global__ void __launch_bounds__(256) k_median5(const float* __restrict__ in, int H, int W, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int y = i / W, x = i % W;
  float v[25]; int n = 0;
  for (int dy = -2; dy <= 2; ++dy) {
    int yy = y + dy;
    while (yy < 0 || yy >= H) yy = yy < 0 ? -yy - 1 : 2 * H - 1 - yy;
    for (int dx = -2; dx <= 2; ++dx) {
      int xx = x + dx;
      while (xx < 0 || xx >= W) xx = xx < 0 ? -xx - 1 : 2 * W - 1 - xx;
      v[n++] = in[yy * W + xx];
    }
  }
  for (int a = 1; a < 25; ++a) { float t = v[a]; int b = a - 1; while (b >= 0 && v[b] > t) { v[b + 1] = v[b]; --b; } v[b + 1] = t; }
  out[i] = v[12];
}

// Sen2Cor mask rule (:688-695): walking the dates in order, a pixel flagged in two consecutive dates is cleared in both
__global__ void __launch_bounds__(256) k_clm_pairs(float* __restrict__ clm, int n, int HW) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  for (int i = 0; i < n; ++i) {
    const int lo = i - 1 > 0 ? i - 1 : 0, hi = i + 1 < n ? i + 1 : n;      // frames [lo, hi)
    float s = 0.f;
    for (int t = lo; t < hi; ++t) s = __fadd_rn(s, clm[(int64_t)t * HW + p]);
    if (s == 2.f) for (int t = lo; t < hi; ++t) clm[(int64_t)t * HW + p] = 0.f;
  }
}

// snow_filter(arr) > 0 (:808-825): per-date counts and the per-pixel count over dates
__device__ __forceinline__ bool snow_flag(const float* x) {
  float ndsi = __fdiv_rn(__fsub_rn(x[1], x[8]), __fadd_rn(x[1], x[8]));
  if (ndsi < 0.10f) ndsi = 0.f;
  if (ndsi > 0.42f) ndsi = 0.42f;
  float p = __fdiv_rn(__fsub_rn(ndsi, 0.1f), 0.32f);
  if (x[3] < 0.10f) p = 0.f;
  if (x[3] > 0.35f && p > 0.f) p = 1.f;
  if (x[0] < 0.10f) p = 0.f;
  if (x[0] > 0.22f && p > 0.f) p = 1.f;
  if (__fdiv_rn(x[0], x[2]) < 0.75f) p = 0.f;
  return p > 0.f;
}
__global__ void __launch_bounds__(256) k_snow(const float* __restrict__ s2, int n, int HW, int* __restrict__ per_date,
                                              unsigned char* __restrict__ low_snow /* mean_t < 0.7 */) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  int c = 0;
  for (int t = 0; t < n; ++t) {
    bool f = p < HW && snow_flag(s2 + ((int64_t)t * HW + p) * 10);
    c += f;
    unsigned bal = __ballot_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(per_date + t, __popc(bal));
  }
  if (p < HW) low_snow[p] = ((double)c / (double)n) < 0.7;
}

__global__ void __launch_bounds__(256) k_count_gt(const float* __restrict__ data, int len, float thresh, int* __restrict__ out) {
  const int sgm = blockIdx.y; int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool f = i < len && data[(int64_t)sgm * len + i] > thresh;
  unsigned bal = __ballot_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(out + sgm, __popc(bal));
}
// mode 0: np.clip(x, lo, hi) (NaN kept)     mode 1: x / lo     mode 2: NaN -> lo
__global__ void __launch_bounds__(256) k_elementwise(float* __restrict__ x, int64_t n, int mode, float lo, float hi) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i];
  if (mode == 0) { if (!isnan(v)) v = fminf(fmaxf(v, lo), hi); }
  else if (mode == 1) v = __fdiv_rn(v, lo);
  else if (isnan(v)) v = lo;
  x[i] = v;
}
// out[i] = #(data[t][i] < thresh over t)   (np.sum(interp < 0.33, axis=0), :1355)
__global__ void __launch_bounds__(256) k_count_lt_axis0(const float* __restrict__ data, int n, int64_t len, float thresh, int* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  int c = 0;
  for (int t = 0; t < n; ++t) c += data[(int64_t)t * len + i] < thresh;
  out[i] = c;
}
// a = max(a, b with fcps zeroed)  (clm[fcps] = 0; cloudshad = np.maximum(cloudshad, clm), :842-845)
__global__ void __launch_bounds__(256) k_max_masked(float* __restrict__ a, const float* __restrict__ b, const unsigned char* __restrict__ zero,
                                                    int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float bv = (zero && zero[i]) ? 0.f : b[i];
  a[i] = fmaxf(a[i], bv);
}

}  // namespace

// ---- device-level launchers (shared with stc_tile.cu: the device-resident tile chain) ----
int tp_s1_fill_dev(stc_ctx* ctx, float* s1_dev, int m, int len) {
  if (m < 1 || len < 1) return STC_OK;
  std::vector<SelJob> jobs(m);
  std::vector<int> ks((size_t)m * SEL_MAX_COLS, 0);
  for (int t = 0; t < m; ++t) { jobs[t] = SelJob{s1_dev + (int64_t)t * len, len, 1, 1}; ks[(size_t)t * SEL_MAX_COLS] = (len - 1) / 2; }
  PoolBuf d_ks, d_pairs;
  STC_CUDA(d_ks.alloc(ks.size() * 4)); STC_CUDA(d_pairs.alloc(ks.size() * 8));
  const void* hk = ctx_stage(ctx, ks.data(), ks.size() * 4);
  if (!hk) STC_FAIL(STC_ERR_NOMEM, "s1_fill: pinned staging");
  STC_CUDA(cudaMemcpyAsync(d_ks.p, hk, ks.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  int rc = select_ranks_dev(ctx, jobs.data(), m, d_ks.as<int>(), d_pairs.as<float>());
  if (rc) return rc;
  { TraceScope ts_(ctx, "k_s1_fill"); k_s1_fill<<<dim3(cdiv(len, 256), m), 256, 0, ctx->stream>>>(s1_dev, len, d_pairs.as<float>()); } ctx->launches++;
  return STC_OK;
}
int tp_median5_dev(stc_ctx* ctx, const float* in_dev, int H, int W, float* out_dev) {
  { TraceScope ts_(ctx, "k_median5"); k_median5<<<cdiv(H * W, 256), 256, 0, ctx->stream>>>(in_dev, H, W, out_dev); } ctx->launches++;
  return STC_OK;
}
int tp_clm_pairs_dev(stc_ctx* ctx, float* clm_dev, int n, int HW) {
  { TraceScope ts_(ctx, "k_clm_pairs"); k_clm_pairs<<<cdiv(HW, 256), 256, 0, ctx->stream>>>(clm_dev, n, HW); } ctx->launches++;
  return STC_OK;
}
// per_date_dev [n] int32 (zeroed here), low_tmp / snow_dev [HW] uint8
int tp_snow_dev(stc_ctx* ctx, const float* s2_dev, int n, int H, int W, int* per_date_dev, unsigned char* low_tmp, unsigned char* snow_dev) {
  const int HW = H * W;
  STC_CUDA(cudaMemsetAsync(per_date_dev, 0, n * 4, ctx->stream));
  { TraceScope ts_(ctx, "k_snow"); k_snow<<<cdiv(HW, 256), 256, 0, ctx->stream>>>(s2_dev, n, HW, per_date_dev, low_tmp); } ctx->launches++;
  maskop_dilate(ctx, low_tmp, snow_dev, 1, H, W, 2, 1, 0, 1, 0);     // 1 - binary_dilation(snow < 0.7, 2)
  return STC_OK;
}
int tp_count_gt_dev(stc_ctx* ctx, const float* data_dev, int nseg, int len, float thresh, int* counts_dev) {
  STC_CUDA(cudaMemsetAsync(counts_dev, 0, nseg * 4, ctx->stream));
  { TraceScope ts_(ctx, "k_count_gt"); k_count_gt<<<dim3(cdiv(len, 256), nseg), 256, 0, ctx->stream>>>(data_dev, len, thresh, counts_dev); } ctx->launches++;
  return STC_OK;
}
int tp_elementwise_dev(stc_ctx* ctx, float* x_dev, int64_t n, int mode, float a, float b) {
  { TraceScope ts_(ctx, "k_elementwise"); k_elementwise<<<cdiv(n, 256), 256, 0, ctx->stream>>>(x_dev, n, mode, a, b); } ctx->launches++;
  return STC_OK;
}
int tp_max_masked_dev(stc_ctx* ctx, float* a_dev, const float* b_dev, const unsigned char* zero_dev, int64_t n) {
  { TraceScope ts_(ctx, "k_max_masked"); k_max_masked<<<cdiv(n, 256), 256, 0, ctx->stream>>>(a_dev, b_dev, zero_dev, n); } ctx->launches++;
  return STC_OK;
}
int tp_count_lt_axis0_dev(stc_ctx* ctx, const float* data_dev, int n, int64_t len, float thresh, int* out_dev) {
  { TraceScope ts_(ctx, "k_count_lt_axis0"); k_count_lt_axis0<<<cdiv(len, 256), 256, 0, ctx->stream>>>(data_dev, n, len, thresh, out_dev); } ctx->launches++;
  return STC_OK;
}

#define TP_FINISH() do { STC_CUDA(cudaStreamSynchronize(ctx->stream)); STC_CUDA(cudaGetLastError()); return STC_OK; } while (0)

extern "C" int stc_s1_fill_host(stc_ctx* ctx, float* s1_host, int m, int H, int W, int C) {
  if (!ctx) return STC_ERR_ARG;
  if (!s1_host || m < 1 || H < 1 || W < 1 || C < 1) STC_FAIL(STC_ERR_ARG, "s1_fill: bad argument");
  const int len = H * W * C; TBuf d;
  STC_CUDA(stc_dmalloc(&d.p, (size_t)m * len * 4));
  STC_CUDA(cudaMemcpyAsync(d.p, s1_host, (size_t)m * len * 4, cudaMemcpyHostToDevice, ctx->stream));
  { int rc = tp_s1_fill_dev(ctx, d.as<float>(), m, len); if (rc) return rc; }
  STC_CUDA(cudaMemcpyAsync(s1_host, d.p, (size_t)m * len * 4, cudaMemcpyDeviceToHost, ctx->stream));
  TP_FINISH();
}

extern "C" int stc_median_filter5_host(stc_ctx* ctx, const float* in_host, int H, int W, float* out_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!in_host || !out_host || H < 1 || W < 1) STC_FAIL(STC_ERR_ARG, "median_filter5: bad argument");
  TBuf a, b;
  STC_CUDA(stc_dmalloc(&a.p, (size_t)H * W * 4)); STC_CUDA(stc_dmalloc(&b.p, (size_t)H * W * 4));
  STC_CUDA(cudaMemcpyAsync(a.p, in_host, (size_t)H * W * 4, cudaMemcpyHostToDevice, ctx->stream));
  { TraceScope ts_(ctx, "k_median5"); k_median5<<<cdiv(H * W, 256), 256, 0, ctx->stream>>>(a.as<float>(), H, W, b.as<float>()); } ctx->launches++;
  STC_CUDA(cudaMemcpyAsync(out_host, b.p, (size_t)H * W * 4, cudaMemcpyDeviceToHost, ctx->stream));
  TP_FINISH();
}

extern "C" int stc_clm_pairs_host(stc_ctx* ctx, float* clm_host, int n, int H, int W) {
  if (!ctx) return STC_ERR_ARG;
  if (!clm_host || n < 1 || H < 1 || W < 1) STC_FAIL(STC_ERR_ARG, "clm_pairs: bad argument");
  TBuf d; const size_t bytes = (size_t)n * H * W * 4;
  STC_CUDA(stc_dmalloc(&d.p, bytes));
  STC_CUDA(cudaMemcpyAsync(d.p, clm_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  { TraceScope ts_(ctx, "k_clm_pairs"); k_clm_pairs<<<cdiv(H * W, 256), 256, 0, ctx->stream>>>(d.as<float>(), n, H * W); } ctx->launches++;
  STC_CUDA(cudaMemcpyAsync(clm_host, d.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  TP_FINISH();
}

extern "C" int stc_snow_host(stc_ctx* ctx, const float* s2_host, int n, int H, int W, int32_t* per_date_host, uint8_t* snow_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!s2_host || !per_date_host || !snow_host || n < 1 || H < 1 || W < 1) STC_FAIL(STC_ERR_ARG, "snow: bad argument");
  const int HW = H * W; TBuf d, cnt, a, b;
  STC_CUDA(stc_dmalloc(&d.p, (size_t)n * HW * 40)); STC_CUDA(stc_dmalloc(&cnt.p, n * 4)); STC_CUDA(stc_dmalloc(&a.p, HW)); STC_CUDA(stc_dmalloc(&b.p, HW));
  STC_CUDA(cudaMemcpyAsync(d.p, s2_host, (size_t)n * HW * 40, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemsetAsync(cnt.p, 0, n * 4, ctx->stream));
  { TraceScope ts_(ctx, "k_snow"); k_snow<<<cdiv(HW, 256), 256, 0, ctx->stream>>>(d.as<float>(), n, HW, cnt.as<int>(), a.as<unsigned char>()); } ctx->launches++;
  maskop_dilate(ctx, a.as<unsigned char>(), b.as<unsigned char>(), 1, H, W, 2, 1, 0, 1, 0);     // 1 - binary_dilation(snow < 0.7, 2)
  STC_CUDA(cudaMemcpyAsync(per_date_host, cnt.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(snow_host, b.p, HW, cudaMemcpyDeviceToHost, ctx->stream));
  TP_FINISH();
}

extern "C" int stc_count_gt_host(stc_ctx* ctx, const float* data_host, int nseg, int len, float thresh, int32_t* counts_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!data_host || !counts_host || nseg < 1 || len < 1) STC_FAIL(STC_ERR_ARG, "count_gt: bad argument");
  TBuf d, cnt;
  STC_CUDA(stc_dmalloc(&d.p, (size_t)nseg * len * 4)); STC_CUDA(stc_dmalloc(&cnt.p, nseg * 4));
  STC_CUDA(cudaMemcpyAsync(d.p, data_host, (size_t)nseg * len * 4, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemsetAsync(cnt.p, 0, nseg * 4, ctx->stream));
  { TraceScope ts_(ctx, "k_count_gt"); k_count_gt<<<dim3(cdiv(len, 256), nseg), 256, 0, ctx->stream>>>(d.as<float>(), len, thresh, cnt.as<int>()); } ctx->launches++;
  STC_CUDA(cudaMemcpyAsync(counts_host, cnt.p, nseg * 4, cudaMemcpyDeviceToHost, ctx->stream));
  TP_FINISH();
}

extern "C" int stc_elementwise_host(stc_ctx* ctx, float* x_host, int64_t n, int mode, float a, float b) {
  if (!ctx) return STC_ERR_ARG;
  if (!x_host || n < 1 || mode < 0 || mode > 2) STC_FAIL(STC_ERR_ARG, "elementwise: bad argument");
  TBuf d;
  STC_CUDA(stc_dmalloc(&d.p, (size_t)n * 4));
  STC_CUDA(cudaMemcpyAsync(d.p, x_host, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  { TraceScope ts_(ctx, "k_elementwise"); k_elementwise<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d.as<float>(), n, mode, a, b); } ctx->launches++;
  STC_CUDA(cudaMemcpyAsync(x_host, d.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  TP_FINISH();
}

extern "C" int stc_max_masked_host(stc_ctx* ctx, float* a_host, const float* b_host, const uint8_t* zero_host, int64_t n) {
  if (!ctx) return STC_ERR_ARG;
  if (!a_host || !b_host || n < 1) STC_FAIL(STC_ERR_ARG, "max_masked: bad argument");
  TBuf a, b, z;
  STC_CUDA(stc_dmalloc(&a.p, (size_t)n * 4)); STC_CUDA(stc_dmalloc(&b.p, (size_t)n * 4));
  STC_CUDA(cudaMemcpyAsync(a.p, a_host, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(b.p, b_host, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (zero_host) { STC_CUDA(stc_dmalloc(&z.p, (size_t)n)); STC_CUDA(cudaMemcpyAsync(z.p, zero_host, (size_t)n, cudaMemcpyHostToDevice, ctx->stream)); }
  { TraceScope ts_(ctx, "k_max_masked"); k_max_masked<<<cdiv(n, 256), 256, 0, ctx->stream>>>(a.as<float>(), b.as<float>(), zero_host ? z.as<unsigned char>() : nullptr, n); } ctx->launches++;
  STC_CUDA(cudaMemcpyAsync(a_host, a.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  TP_FINISH();
}

extern "C" int stc_count_lt_axis0_host(stc_ctx* ctx, const float* data_host, int n, int64_t len, float thresh, int32_t* out_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!data_host || !out_host || n < 1 || len < 1) STC_FAIL(STC_ERR_ARG, "count_lt_axis0: bad argument");
  TBuf d, o;
  STC_CUDA(stc_dmalloc(&d.p, (size_t)n * len * 4)); STC_CUDA(stc_dmalloc(&o.p, (size_t)len * 4));
  STC_CUDA(cudaMemcpyAsync(d.p, data_host, (size_t)n * len * 4, cudaMemcpyHostToDevice, ctx->stream));
  { TraceScope ts_(ctx, "k_count_lt_axis0"); k_count_lt_axis0<<<cdiv(len, 256), 256, 0, ctx->stream>>>(d.as<float>(), n, len, thresh, o.as<int>()); } ctx->launches++;
  STC_CUDA(cudaMemcpyAsync(out_host, o.p, (size_t)len * 4, cudaMemcpyDeviceToHost, ctx->stream));
  TP_FINISH();
}
