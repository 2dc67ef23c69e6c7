// Spatial-neighbourhood kernels of the cloud-gap pipeline with SciPy's exact semantics.
//   feather: a = 1 - min(EDT(1 - mask), 12)/12 ; a < 0.2 -> 0 ; grey_closing(a, size)
//            id_areas_to_interp  (src/preprocessing/cloud_removal.py:774-798, size 15)
//            remove_cloud_and_shadows (:913-921, size 20)
//   binary dilation with the 4-connected cross or the full 3x3 element, k iterations,
//            border_value 0 (scipy.ndimage.binary_dilation; >= 15 call sites)
// Exactness notes (verified against SciPy 1.18 on the host): the EDT is capped at 12 px so an
// exact 25x25 windowed search suffices; grey_closing = erosion(dilation) with flat size x size
// windows, mode='reflect' (edge-repeating mirror); for even size the dilation window is
// [i-(s/2-1), i+s/2] and the erosion window [i-s/2, i+s/2-1].
#include "stc_common.cuh"

__device__ __forceinline__ int reflect_index(int i, int n) {
  int p = 2 * n;
  i %= p; if (i < 0) i += p;
  return (i >= n) ? (p - 1 - i) : i;
}

// one thread per pixel; skip[date] != 0 leaves the date untouched (np.sum(mask) == 0 guard)
__global__ void __launch_bounds__(256) edt_feather_kernel(const float* __restrict__ mask, const float* __restrict__ sums,
                                                          float* __restrict__ out, int n, int H, int W) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * H * W) return;
  int x = (int)(idx % W); int64_t r = idx / W; int y = (int)(r % H); int d = (int)(r / H);
  const float* m = mask + (int64_t)d * H * W;
  if (!(sums[d] > 0.f)) { out[idx] = m[(int64_t)y * W + x]; return; }
  int best = 1 << 30;
  for (int dy = -12; dy <= 12; ++dy) {
    int yy = y + dy; if (yy < 0 || yy >= H) continue;
    int rem = 144 - dy * dy;
    for (int dx = -12; dx <= 12; ++dx) {
      int xx = x + dx; if (xx < 0 || xx >= W) continue;
      int d2 = dx * dx;
      if (d2 > rem) continue;
      if (m[(int64_t)yy * W + xx] == 1.0f) { d2 += dy * dy; best = d2 < best ? d2 : best; }
    }
  }
  double dist = (best <= 144) ? sqrt((double)best) : 12.0;
  double v = 1.0 - (dist / 12.0);
  if (v < 0.2) v = 0.0;
  out[idx] = (float)v;
}

// per-date sum of the (clipped) mask: decides the `if np.sum(...) > 0` guard
__global__ void __launch_bounds__(256) date_sum_kernel(const float* __restrict__ mask, float* __restrict__ sums, int HW) {
  __shared__ float s[256];
  const float* m = mask + (int64_t)blockIdx.x * HW;
  float acc = 0.f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) acc += m[i];
  s[threadIdx.x] = acc; __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) sums[blockIdx.x] = s[0];
}

// separable flat max/min filter along one axis with SciPy 'reflect' boundary
__global__ void __launch_bounds__(256) window_reduce_kernel(const float* __restrict__ in, const float* __restrict__ sums,
                                                            float* __restrict__ out, int n, int H, int W, int lo, int hi,
                                                            int axis, int is_max) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * H * W) return;
  int x = (int)(idx % W); int64_t r = idx / W; int y = (int)(r % H); int d = (int)(r / H);
  const float* m = in + (int64_t)d * H * W;
  if (sums && !(sums[d] > 0.f)) { out[idx] = m[(int64_t)y * W + x]; return; }
  float v = is_max ? -INFINITY : INFINITY;
  for (int k = lo; k <= hi; ++k) {
    float t = axis == 0 ? m[(int64_t)reflect_index(y + k, H) * W + x] : m[(int64_t)y * W + reflect_index(x + k, W)];
    v = is_max ? fmaxf(v, t) : fminf(v, t);
  }
  out[idx] = v;
}

// binary dilation, k iterations of the cross (conn=1, L1 ball) or full 3x3 (conn=2, Linf ball)
__global__ void __launch_bounds__(256) binary_dilate_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out,
                                                            int n, int H, int W, int k, int conn) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * H * W) return;
  int x = (int)(idx % W); int64_t r = idx / W; int y = (int)(r % H); int d = (int)(r / H);
  const unsigned char* m = in + (int64_t)d * H * W;
  unsigned char o = 0;
  for (int dy = -k; dy <= k && !o; ++dy) {
    int yy = y + dy; if (yy < 0 || yy >= H) continue;
    int span = conn == 1 ? k - abs(dy) : k;
    for (int dx = -span; dx <= span; ++dx) {
      int xx = x + dx; if (xx < 0 || xx >= W) continue;
      if (m[(int64_t)yy * W + xx]) { o = 1; break; }
    }
  }
  out[idx] = o;
}

// squared Euclidean distance (exact, integer) from every pixel to the nearest non-zero pixel of
// `target` within `radius`; radius*radius + 1 where none is that close.  The callers cap
// distance_transform_edt at 3/5/12 px, so a windowed search is exact.
__global__ void __launch_bounds__(256) edt_sq_kernel(const unsigned char* __restrict__ target, int* __restrict__ out,
                                                     int n, int H, int W, int radius) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * H * W) return;
  int x = (int)(idx % W); int64_t r = idx / W; int y = (int)(r % H); int d = (int)(r / H);
  const unsigned char* m = target + (int64_t)d * H * W;
  const int r2 = radius * radius;
  int best = r2 + 1;
  for (int dy = -radius; dy <= radius; ++dy) {
    int yy = y + dy; if (yy < 0 || yy >= H) continue;
    for (int dx = -radius; dx <= radius; ++dx) {
      int xx = x + dx; if (xx < 0 || xx >= W) continue;
      int d2 = dx * dx + dy * dy;
      if (d2 < best && m[(int64_t)yy * W + xx]) best = d2;
    }
  }
  out[idx] = best;
}

int pre_edt_sq_dev(stc_ctx* ctx, const unsigned char* target_dev, int n, int H, int W, int radius, int* out_dev) {
  if (radius < 1 || radius > 64) STC_FAIL(STC_ERR_ARG, "edt_sq: radius must be in 1..64");
  { TraceScope ts_(ctx, "edt_sq_kernel"); edt_sq_kernel<<<cdiv((int64_t)n * H * W, 256), 256, 0, ctx->stream>>>(target_dev, out_dev, n, H, W, radius); }
  STC_CUDA(cudaGetLastError());
  ctx->launches++;
  return STC_OK;
}

int pre_feather_dev(stc_ctx* ctx, const float* mask_dev, int n, int H, int W, int size, float* tmp_a, float* tmp_b,
                    float* sums_dev, float* out_dev) {
  if (size < 1 || size > 64) STC_FAIL(STC_ERR_ARG, "feather: closing size must be in 1..64");
  int64_t tot = (int64_t)n * H * W;
  int grid = cdiv(tot, 256);
  { TraceScope ts_(ctx, "date_sum_kernel"); date_sum_kernel<<<n, 256, 0, ctx->stream>>>(mask_dev, sums_dev, H * W); }
  { TraceScope ts_(ctx, "edt_feather_kernel"); edt_feather_kernel<<<grid, 256, 0, ctx->stream>>>(mask_dev, sums_dev, tmp_a, n, H, W); }
  int dlo, dhi, elo, ehi;
  if (size & 1) { dlo = elo = -(size / 2); dhi = ehi = size / 2; }
  else { dlo = -(size / 2 - 1); dhi = size / 2; elo = -(size / 2); ehi = size / 2 - 1; }
  { TraceScope ts_(ctx, "window_reduce_kernel"); window_reduce_kernel<<<grid, 256, 0, ctx->stream>>>(tmp_a, sums_dev, tmp_b, n, H, W, dlo, dhi, 0, 1); }
  { TraceScope ts_(ctx, "window_reduce_kernel"); window_reduce_kernel<<<grid, 256, 0, ctx->stream>>>(tmp_b, sums_dev, tmp_a, n, H, W, dlo, dhi, 1, 1); }
  { TraceScope ts_(ctx, "window_reduce_kernel"); window_reduce_kernel<<<grid, 256, 0, ctx->stream>>>(tmp_a, sums_dev, tmp_b, n, H, W, elo, ehi, 0, 0); }
  { TraceScope ts_(ctx, "window_reduce_kernel"); window_reduce_kernel<<<grid, 256, 0, ctx->stream>>>(tmp_b, sums_dev, out_dev, n, H, W, elo, ehi, 1, 0); }
  STC_CUDA(cudaGetLastError());
  ctx->launches += 6;
  return STC_OK;
}

int pre_binary_dilate_dev(stc_ctx* ctx, const unsigned char* in_dev, int n, int H, int W, int iterations, int conn,
                          unsigned char* out_dev) {
  if (iterations < 1 || iterations > 64 || (conn != 1 && conn != 2)) STC_FAIL(STC_ERR_ARG, "binary_dilate: bad iterations/connectivity");
  { TraceScope ts_(ctx, "binary_dilate_kernel"); binary_dilate_kernel<<<cdiv((int64_t)n * H * W, 256), 256, 0, ctx->stream>>>(in_dev, out_dev, n, H, W, iterations, conn); }
  STC_CUDA(cudaGetLastError());
  ctx->launches++;
  return STC_OK;
}
