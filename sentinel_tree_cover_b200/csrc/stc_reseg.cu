// Seam re-segmentation pass (src/resegment_tiles_wide.py): the array work that is new relative to the main job.
//   align_subtile_histograms (:284-345): a border subtile is made of the right edge of one tile and the left edge of its
//   neighbour (different acquisition dates).  Per time step the two halves are brought to a common per-band mean / standard
//   deviation (statistics over the non-water pixels of each half), and the change is kept only when it shrinks the jump
//   across the seam column.
// The statistics are accumulated in float64 (NumPy adds float32 rows sequentially here, np.nanmean / np.nanstd over axis 0
// of a boolean-indexed [npx, C] array); tests/test_resegment.py holds the result to rtol 1e-5 against the reference function.
#include "stc_common.cuh"
#include <vector>
#include <cmath>

namespace {

constexpr int RS_MAXC = 20;

// water[p] = NDWI(median_t arr[:, p, (1, 3)]) >= 0.1 (np.median over axis 0: NaN propagates -> False)
__global__ void __launch_bounds__(128) k_rs_water(const float* __restrict__ arr, int T, int HW, int C, unsigned char* __restrict__ water) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  float med[2];
  for (int k = 0; k < 2; ++k) {
    float v[32]; bool nan_ = false;
    for (int t = 0; t < T; ++t) { v[t] = arr[((int64_t)t * HW + p) * C + (k ? 3 : 1)]; nan_ = nan_ || isnan(v[t]); }
    for (int i = 1; i < T; ++i) { float x = v[i]; int j = i - 1; while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; } v[j + 1] = x; }
    med[k] = nan_ ? nanf("") : ((T & 1) ? v[T >> 1] : __fmul_rn(__fadd_rn(v[(T >> 1) - 1], v[T >> 1]), 0.5f));
  }
  water[p] = __fdiv_rn(__fsub_rn(med[0], med[1]), __fadd_rn(med[0], med[1])) >= 0.1f;
}

// per (t, side): nan-aware count / mean (pass 0) or sum of squared deviations (pass 1) per band over the non-water pixels
// side 0: columns >= half, side 1: columns < half.  out [T][2][C] doubles (+ counts [T][2][C] ints in pass 0)
__global__ void __launch_bounds__(256) k_rs_stats(const float* __restrict__ arr, const unsigned char* __restrict__ water, int H, int W, int C,
                                                  int half, int pass, const double* __restrict__ mean_in, double* __restrict__ out,
                                                  int* __restrict__ cnt) {
  const int t = blockIdx.x, side = blockIdx.y;
  const int c0 = side == 0 ? half : 0, c1 = side == 0 ? W : half, wd = c1 - c0;
  double acc[RS_MAXC]; int n[RS_MAXC];
  for (int b = 0; b < C; ++b) { acc[b] = 0.0; n[b] = 0; }
  const float* a = arr + (int64_t)t * H * W * C;
  const double* mu = mean_in + ((int64_t)t * 2 + side) * C;
  for (int i = threadIdx.x; i < H * wd; i += blockDim.x) {
    const int r = i / wd, c = c0 + i % wd;
    if (water[r * W + c]) continue;
    const float* x = a + ((int64_t)r * W + c) * C;
    for (int b = 0; b < C; ++b) {
      const float v = x[b];
      if (isnan(v)) continue;
      if (pass == 0) acc[b] += (double)v; else { const double d = (double)v - mu[b]; acc[b] += d * d; }
      n[b]++;
    }
  }
  __shared__ double s_acc[256]; __shared__ int s_n[256];
  for (int b = 0; b < C; ++b) {
    s_acc[threadIdx.x] = acc[b]; s_n[threadIdx.x] = n[b];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) { s_acc[threadIdx.x] += s_acc[threadIdx.x + o]; s_n[threadIdx.x] += s_n[threadIdx.x + o]; }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const int64_t o = ((int64_t)t * 2 + side) * C + b;
      const int nn = s_n[0];
      if (pass == 0) { out[o] = nn ? s_acc[0] / nn : nan(""); cnt[o] = nn; }
      else out[o] = nn ? sqrt(s_acc[0] / nn) : nan("");
    }
    __syncthreads();
  }
}

struct RsParams { float mult[2][RS_MAXC]; float add[2][RS_MAXC]; };      // [0]: applied to columns < half, [1]: to columns >= half

// jump across the seam column of time step t before and after the candidate transform: mean over rows and bands of
// |x[r, seam-1, b] - x[r, seam, b]| (np.roll(.., 1, axis=1) at column `seam`)
__global__ void __launch_bounds__(256) k_rs_seam(const float* __restrict__ arr, int H, int W, int C, int half, int seam,
                                                 const RsParams* __restrict__ prm, double* __restrict__ out /*[T][2]*/) {
  const int t = blockIdx.x;
  const float* a = arr + (int64_t)t * H * W * C;
  const RsParams& p = prm[t];
  const int cl = seam == 0 ? W - 1 : seam - 1;
  double before = 0.0, after = 0.0;
  for (int i = threadIdx.x; i < H * C; i += blockDim.x) {
    const int r = i / C, b = i % C;
    const float x0 = a[((int64_t)r * W + cl) * C + b], x1 = a[((int64_t)r * W + seam) * C + b];
    before += (double)fabsf(__fsub_rn(x0, x1));
    const int s0 = cl < half ? 0 : 1, s1 = seam < half ? 0 : 1;
    const float y0 = __fadd_rn(__fmul_rn(x0, p.mult[s0][b]), p.add[s0][b]), y1 = __fadd_rn(__fmul_rn(x1, p.mult[s1][b]), p.add[s1][b]);
    after += (double)fabsf(__fsub_rn(y0, y1));
  }
  __shared__ double s0_[256], s1_[256];
  s0_[threadIdx.x] = before; s1_[threadIdx.x] = after;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) { s0_[threadIdx.x] += s0_[threadIdx.x + o]; s1_[threadIdx.x] += s1_[threadIdx.x + o]; } __syncthreads(); }
  if (threadIdx.x == 0) { out[2 * t] = s0_[0] / ((double)H * C); out[2 * t + 1] = s1_[0] / ((double)H * C); }
}

__global__ void __launch_bounds__(256) k_rs_apply(float* __restrict__ arr, int HW, int W, int C, int half, const RsParams* __restrict__ prm,
                                                  const int* __restrict__ applied) {
  const int t = blockIdx.y;
  if (!applied[t]) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)HW * C) return;
  const int b = (int)(i % C); const int c = (int)((i / C) % W);
  const int s = c < half ? 0 : 1;
  float* x = arr + (int64_t)t * HW * C + i;
  *x = __fadd_rn(__fmul_rn(*x, prm[t].mult[s][b]), prm[t].add[s][b]);
}

}  // namespace

// arr [T,H,W,C] float32 on the device, transformed in place; applied_host[t] = 1 where the alignment was kept
int reseg_align_histograms_dev(stc_ctx* ctx, float* arr, int T, int H, int W, int C, int half, int seam, int32_t* applied_host) {
  if (T < 1 || T > 32 || C < 4 || C > RS_MAXC || half < 1 || half >= W || seam < 0 || seam >= W) STC_FAIL(STC_ERR_ARG, "align_histograms: bad argument");
  const int HW = H * W;
  PoolBuf water, mean, sd, cnt, prm, seamv, appl;
  STC_CUDA(water.alloc(HW)); STC_CUDA(mean.alloc((size_t)T * 2 * C * 8)); STC_CUDA(sd.alloc((size_t)T * 2 * C * 8)); STC_CUDA(cnt.alloc((size_t)T * 2 * C * 4));
  STC_CUDA(prm.alloc((size_t)T * sizeof(RsParams))); STC_CUDA(seamv.alloc((size_t)T * 16)); STC_CUDA(appl.alloc((size_t)T * 4));
  { TraceScope ts_(ctx, "k_rs_water"); k_rs_water<<<cdiv(HW, 128), 128, 0, ctx->stream>>>(arr, T, HW, C, water.as<unsigned char>()); }
  { TraceScope ts_(ctx, "k_rs_stats"); k_rs_stats<<<dim3(T, 2), 256, 0, ctx->stream>>>(arr, water.as<unsigned char>(), H, W, C, half, 0, mean.as<double>(), mean.as<double>(), cnt.as<int>()); }
  { TraceScope ts_(ctx, "k_rs_stats"); k_rs_stats<<<dim3(T, 2), 256, 0, ctx->stream>>>(arr, water.as<unsigned char>(), H, W, C, half, 1, mean.as<double>(), sd.as<double>(), cnt.as<int>()); }
  ctx->launches += 3;
  std::vector<double> h_mean((size_t)T * 2 * C), h_sd((size_t)T * 2 * C);
  STC_CUDA(cudaMemcpyAsync(h_mean.data(), mean.p, h_mean.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(h_sd.data(), sd.p, h_sd.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  // float32 vector arithmetic of :312-323.  The reference names the statistics of the columns >= half "left" and applies the
  // "left" transform to the columns < half (and vice versa); that is what the deployed code does, so it is kept.
  std::vector<RsParams> P(T);
  for (int t = 0; t < T; ++t)
    for (int b = 0; b < C; ++b) {
      const float std_left = (float)h_sd[((size_t)t * 2 + 0) * C + b], std_right = (float)h_sd[((size_t)t * 2 + 1) * C + b];
      const float mean_left = (float)h_mean[((size_t)t * 2 + 0) * C + b], mean_right = (float)h_mean[((size_t)t * 2 + 1) * C + b];
      volatile float std_ref = (std_right + std_left) / 2.f, mean_ref = (mean_right + mean_left) / 2.f;
      volatile float ml = std_left / std_ref, mr = std_right / std_ref;
      volatile float tl = mean_ref * ml, tr = mean_ref * mr;
      P[t].mult[0][b] = ml; P[t].add[0][b] = mean_left - tl;          // columns < half get the "left" transform
      P[t].mult[1][b] = mr; P[t].add[1][b] = mean_right - tr;
    }
  const void* hp = ctx_stage(ctx, P.data(), P.size() * sizeof(RsParams));
  if (!hp) STC_FAIL(STC_ERR_NOMEM, "align_histograms: pinned staging");
  STC_CUDA(cudaMemcpyAsync(prm.p, hp, P.size() * sizeof(RsParams), cudaMemcpyHostToDevice, ctx->stream));
  { TraceScope ts_(ctx, "k_rs_seam"); k_rs_seam<<<T, 256, 0, ctx->stream>>>(arr, H, W, C, half, seam, prm.as<RsParams>(), seamv.as<double>()); }
  std::vector<double> sv((size_t)T * 2);
  STC_CUDA(cudaMemcpyAsync(sv.data(), seamv.p, sv.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<int> ap(T);
  for (int t = 0; t < T; ++t) { ap[t] = ((float)sv[2 * t + 1] < (float)sv[2 * t]) ? 1 : 0; if (applied_host) applied_host[t] = ap[t]; }   // NaN: not applied
  const void* ha = ctx_stage(ctx, ap.data(), (size_t)T * 4);
  STC_CUDA(cudaMemcpyAsync(appl.p, ha, (size_t)T * 4, cudaMemcpyHostToDevice, ctx->stream));
  { TraceScope ts_(ctx, "k_rs_apply"); k_rs_apply<<<dim3(cdiv((int64_t)HW * C, 256), T), 256, 0, ctx->stream>>>(arr, HW, W, C, half, prm.as<RsParams>(), appl.as<int>()); }
  ctx->launches += 2;
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

extern "C" int stc_align_histograms_host(stc_ctx* ctx, float* arr_host, int T, int H, int W, int C, int half, int seam_col, int32_t* applied_out) {
  if (!ctx) return STC_ERR_ARG;
  if (!arr_host || H < 1 || W < 2) STC_FAIL(STC_ERR_ARG, "align_histograms: bad argument");
  const size_t bytes = (size_t)T * H * W * C * 4;
  PoolBuf d;
  STC_CUDA(d.alloc(bytes));
  STC_CUDA(cudaMemcpyAsync(d.p, arr_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  int rc = reseg_align_histograms_dev(ctx, d.as<float>(), T, H, W, C, half, seam_col, applied_out);
  if (rc) return rc;
  STC_CUDA(cudaMemcpyAsync(arr_host, d.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}
