// Fused, device-resident versions of the per-tile steps of process_subtiles (src/download_and_predict_job.py:1125-1486):
// each entry uploads its inputs once, chains the existing kernels on the device and returns only what the caller
// keeps, instead of one host round trip per NumPy statement.  Same arithmetic as the unfused entry points.
//   stc_s2_medians_host        :1148-1159  NaN -> 0, np.median over the raw dates of the 10 bands and of the 4 indices
//   stc_smooth_quarterly_host  :1171-1174, 1274-1278 (+ deal_w_missing_px fill :1039-1047, make_indices, regrid/Whittaker)
//   stc_predict_postprocess_host :1398-1425, 1451-1483  normalise + forward + post-filters for a batch of subtiles
#include "stc_common.cuh"

int interp_missing_counts_dev(stc_ctx* ctx, const float* arr_dev, int n, int HW, int C, int* bad_px_dev, int* nan_vals_dev);
int interp_median_fill_dev(stc_ctx* ctx, float* arr_dev, int n, int64_t cols);
int post_subtile_dev(stc_ctx* ctx, const float* preds_dev, const float* img_dev, const float* mc_dev, int S, int F, int C,
                     unsigned char* a, unsigned char* b, int* d2, double* ramp, unsigned char* na, unsigned char* nb, unsigned char* vote,
                     float* out_dev);

namespace {

struct FBuf { void* p = nullptr; ~FBuf() { if (p) cudaFree(p); } template <typename T> T* as() { return (T*)p; } };

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

extern "C" int stc_s2_medians_host(stc_ctx* ctx, float* s2_host, int n, int H, int W, float* median14_host, int32_t* bad_px_host,
                                   int64_t* nan_total_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!s2_host || !median14_host || !bad_px_host || !nan_total_host || n < 1 || H < 1 || W < 1) STC_FAIL(STC_ERR_ARG, "s2_medians: bad argument");
  const int HW = H * W; const int64_t px = (int64_t)n * HW;
  FBuf s2, idx, m10, m4, out, cnt;
  STC_CUDA(cudaMalloc(&s2.p, px * 40)); STC_CUDA(cudaMalloc(&idx.p, px * 16)); STC_CUDA(cudaMalloc(&m10.p, (size_t)HW * 40));
  STC_CUDA(cudaMalloc(&m4.p, (size_t)HW * 16)); STC_CUDA(cudaMalloc(&out.p, (size_t)HW * 56)); STC_CUDA(cudaMalloc(&cnt.p, 2 * n * 4));
  STC_CUDA(cudaMemcpyAsync(s2.p, s2_host, px * 40, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemsetAsync(cnt.p, 0, 2 * n * 4, ctx->stream));
  TF_CHECK(interp_missing_counts_dev(ctx, s2.as<float>(), n, HW, 10, cnt.as<int>(), cnt.as<int>() + n));
  std::vector<int> h(2 * n);
  STC_CUDA(cudaMemcpyAsync(h.data(), cnt.p, 2 * n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  int64_t nans = 0;
  for (int t = 0; t < n; ++t) { bad_px_host[t] = h[t]; nans += h[n + t]; }
  *nan_total_host = nans;
  if (nans) {            // interpolate_na_vals: NaN -> 0, visible to the caller (the reference fills in place)
    k_nan_to_zero<<<cdiv(px * 10, 256), 256, 0, ctx->stream>>>(s2.as<float>(), px * 10); ctx->launches++;
    STC_CUDA(cudaMemcpyAsync(s2_host, s2.p, px * 40, cudaMemcpyDeviceToHost, ctx->stream));
  }
  TF_CHECK(pre_temporal_median_dev(ctx, s2.as<float>(), n, (int64_t)HW * 10, m10.as<float>()));
  TF_CHECK(pre_indices_dev(ctx, s2.as<float>(), px, 10, idx.as<float>()));
  TF_CHECK(pre_temporal_median_dev(ctx, idx.as<float>(), n, (int64_t)HW * 4, m4.as<float>()));
  k_concat_channels<<<cdiv((int64_t)HW * 14, 256), 256, 0, ctx->stream>>>(m10.as<float>(), 10, m4.as<float>(), 4, HW, out.as<float>()); ctx->launches++;
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
  FBuf s2, idx, sm10, sm4, sm14, q, cnt, s1, s1q, s1m;
  STC_CUDA(cudaMalloc(&s2.p, px * 40)); STC_CUDA(cudaMalloc(&idx.p, px * 16)); STC_CUDA(cudaMalloc(&sm10.p, (size_t)12 * HW * 40));
  STC_CUDA(cudaMalloc(&sm4.p, (size_t)12 * HW * 16)); STC_CUDA(cudaMalloc(&sm14.p, (size_t)12 * HW * 56)); STC_CUDA(cudaMalloc(&q.p, (size_t)4 * HW * 56));
  STC_CUDA(cudaMalloc(&cnt.p, 2 * n * 4));
  STC_CUDA(cudaMemcpyAsync(s2.p, s2_host, px * 40, cudaMemcpyHostToDevice, ctx->stream));
  TF_CHECK(interp_median_fill_dev(ctx, s2.as<float>(), n, (int64_t)HW * 10));                    // deal_w_missing_px :1039-1047
  STC_CUDA(cudaMemsetAsync(cnt.p, 0, 2 * n * 4, ctx->stream));
  TF_CHECK(interp_missing_counts_dev(ctx, s2.as<float>(), n, HW, 10, cnt.as<int>(), cnt.as<int>() + n));
  STC_CUDA(cudaMemcpyAsync(nan_after_host, cnt.as<int>() + n, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int t = 0; t < n; ++t) if (nan_after_host[t] > 0) return STC_OK;      // caller drops the NaN dates and retries (:1048-1053)
  TF_CHECK(pre_indices_dev(ctx, s2.as<float>(), px, 10, idx.as<float>()));                         // make_indices :998
  TF_CHECK(pre_temporal_matmul_dev(ctx, s2.as<float>(), M_host, n, 12, (int64_t)HW * 10, sm10.as<float>()));
  TF_CHECK(pre_temporal_matmul_dev(ctx, idx.as<float>(), M_host, n, 12, (int64_t)HW * 4, sm4.as<float>()));
  k_concat_channels<<<cdiv((int64_t)12 * HW * 14, 256), 256, 0, ctx->stream>>>(sm10.as<float>(), 10, sm4.as<float>(), 4, (int64_t)12 * HW, sm14.as<float>());
  ctx->launches++;
  if (s2_monthly_host) STC_CUDA(cudaMemcpyAsync(s2_monthly_host, sm14.p, (size_t)12 * HW * 56, cudaMemcpyDeviceToHost, ctx->stream));
  if (s2_quarterly_host) {
    for (int k = 0; k < 4; ++k)
      TF_CHECK(pre_temporal_median_dev(ctx, sm14.as<float>() + (size_t)3 * k * HW * 14, 3, (int64_t)HW * 14, q.as<float>() + (size_t)k * HW * 14));
    STC_CUDA(cudaMemcpyAsync(s2_quarterly_host, q.p, (size_t)4 * HW * 56, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (s1_host) {
    STC_CUDA(cudaMalloc(&s1.p, (size_t)12 * HW * 8)); STC_CUDA(cudaMalloc(&s1q.p, (size_t)4 * HW * 8)); STC_CUDA(cudaMalloc(&s1m.p, (size_t)HW * 8));
    STC_CUDA(cudaMemcpyAsync(s1.p, s1_host, (size_t)12 * HW * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (s1_quarterly_host) {
      for (int k = 0; k < 4; ++k)
        TF_CHECK(pre_temporal_median_dev(ctx, s1.as<float>() + (size_t)3 * k * HW * 2, 3, (int64_t)HW * 2, s1q.as<float>() + (size_t)k * HW * 2));
      STC_CUDA(cudaMemcpyAsync(s1_quarterly_host, s1q.p, (size_t)4 * HW * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (s1_median_host) {
      TF_CHECK(pre_temporal_median_dev(ctx, s1.as<float>(), 12, (int64_t)HW * 2, s1m.as<float>()));
      STC_CUDA(cudaMemcpyAsync(s1_median_host, s1m.p, (size_t)HW * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
  }
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
  STC_CUDA(cudaMalloc(&x.p, per * B * 4)); STC_CUDA(cudaMalloc(&mc.p, (size_t)B * H * H * 4)); STC_CUDA(cudaMalloc(&preds.p, (size_t)B * S * S * 4));
  STC_CUDA(cudaMalloc(&out.p, (size_t)B * S * S * 4)); STC_CUDA(cudaMalloc(&a.p, H * H)); STC_CUDA(cudaMalloc(&b.p, H * H));
  STC_CUDA(cudaMalloc(&d2.p, (size_t)H * H * 4)); STC_CUDA(cudaMalloc(&ramp.p, (size_t)S * S * 8)); STC_CUDA(cudaMalloc(&na.p, (S + 2) * (S + 2)));
  STC_CUDA(cudaMalloc(&nb.p, (S + 2) * (S + 2))); STC_CUDA(cudaMalloc(&vote.p, 256));
  STC_CUDA(cudaMemcpyAsync(x.p, x_host, per * B * 4, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(mc.p, min_clear_host, (size_t)B * H * H * 4, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  // forward for the whole batch (normalize_subtile fused into the input packing; x stays un-normalised for the post-filters)
  TF_CHECK(model_predict_dev(ctx, x.as<float>(), B, T, H, H, length, 1, min17, max17, preds.as<float>()));
  for (int i = 0; i < B; ++i) {
    float* p = preds.as<float>() + (size_t)i * S * S;
    if (no_data_host[i]) { k_fill_value<<<cdiv(S * S, 256), 256, 0, ctx->stream>>>(p, (int64_t)S * S, 255.f); ctx->launches++; }   // np.full((SIZE, SIZE), 255)
    TF_CHECK(post_subtile_dev(ctx, p, x.as<float>() + per * i, mc.as<float>() + (size_t)i * H * H, S, T + 1, 17, a.as<unsigned char>(),
                              b.as<unsigned char>(), d2.as<int>(), ramp.as<double>(), na.as<unsigned char>(), nb.as<unsigned char>(),
                              vote.as<unsigned char>(), out.as<float>() + (size_t)i * S * S));
  }
  STC_CUDA(cudaMemcpyAsync(out_host, out.p, (size_t)B * S * S * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}
