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

// ---------------------------------------------------------------------------------------------
// Row distance: g[f][y][x] = min |dx| over the set pixels of row (f, y), capped at cap + 1 (cap <= 32, out of image = not
// set).  One warp per 32-pixel segment: three coalesced loads (previous, own, next segment), three ballots, and the
// nearest set bit on either side by count-leading / find-first-set.  Every binary dilation (an L1 or Linf ball of radius
// k is exactly k iterations of SciPy's cross / 3x3 element with border_value 0) and every capped Euclidean distance
// transform of the pipeline is this pass plus ONE column pass over 2k + 1 entries: O(k) per pixel instead of the O(k^2)
// window search of round 1 (k_dilate: 233 launches at 18 GB/s, edt_feather_kernel: 625 probes per pixel).
// src_kind: 0 = uint8 != 0, 1 = uint8 == 0, 2 = float32 == 1.0f
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rowdist(const void* __restrict__ in, int src_kind, int64_t rows, int W, int segs, int cap,
                                                 unsigned char* __restrict__ g) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wid >= rows * segs) return;                      // whole warp
  const int64_t row = wid / segs; const int seg = (int)(wid - row * segs);
  const int x = seg * 32 + lane;
  auto is_set = [&](int xx) -> bool {
    if (xx < 0 || xx >= W) return false;
    const int64_t o = row * W + xx;
    if (src_kind == 2) return reinterpret_cast<const float*>(in)[o] == 1.0f;
    const unsigned char v = reinterpret_cast<const unsigned char*>(in)[o];
    return src_kind == 0 ? (v != 0) : (v == 0);
  };
  const unsigned prev = __ballot_sync(0xffffffffu, is_set(x - 32));
  const unsigned cur = __ballot_sync(0xffffffffu, is_set(x));
  const unsigned next = __ballot_sync(0xffffffffu, is_set(x + 32));
  const unsigned long long lo64 = (unsigned long long)prev | ((unsigned long long)cur << 32);
  const unsigned long long hi64 = (unsigned long long)cur | ((unsigned long long)next << 32);
  const int p = 32 + lane;
  const unsigned long long L = lo64 & (p == 63 ? ~0ull : ((1ull << (p + 1)) - 1ull));
  const int dl = L ? p - (63 - __clzll((long long)L)) : 255;
  const unsigned long long R = hi64 >> lane;
  const int dr = R ? (__ffsll((long long)R) - 1) : 255;
  int d = dl < dr ? dl : dr;
  if (d > cap) d = cap + 1;
  if (x < W) g[row * W + x] = (unsigned char)d;
}

int morph_rowdist_dev(stc_ctx* ctx, const void* in, int src_kind, int64_t rows, int W, int cap, unsigned char* g) {
  if (cap < 0 || cap > 32) STC_FAIL(STC_ERR_ARG, "rowdist: cap must be in 0..32");
  const int segs = (W + 31) / 32;
  { TraceScope ts_(ctx, "k_rowdist"); k_rowdist<<<cdiv(rows * segs, 8), 256, 0, ctx->stream>>>(in, src_kind, rows, W, segs, cap, g); }
  ctx->launches++;
  return STC_OK;
}

// column pass of a binary dilation: out = inv_out ^ (a set pixel within the L1 (conn 1) / Linf (conn 2) ball of radius k);
// three_d: the L1 ball also spans the frame axis (scipy binary_dilation of a 3-D array with the 3-D cross)
__global__ void __launch_bounds__(256) k_dilate_col(const unsigned char* __restrict__ g, unsigned char* __restrict__ out, int T, int H, int W,
                                                    int k, int conn, int inv_out, int three_d) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)T * H * W) return;
  const int x = (int)(idx % W); const int64_t r = idx / W; const int y = (int)(r % H); const int t = (int)(r / H);
  int hit = 0;
  const int dt0 = three_d ? -k : 0, dt1 = three_d ? k : 0;
  for (int dt = dt0; dt <= dt1 && !hit; ++dt) {
    const int tt = t + dt; if (tt < 0 || tt >= T) continue;
    const int kk = k - abs(dt);
    const unsigned char* gg = g + (int64_t)tt * H * W + x;
    const int y0 = max(y - kk, 0), y1 = min(y + kk, H - 1);
    for (int yy = y0; yy <= y1; ++yy) {
      const int lim = conn == 1 ? kk - abs(yy - y) : kk;
      if ((int)gg[(int64_t)yy * W] <= lim) { hit = 1; break; }
    }
  }
  out[idx] = (unsigned char)(hit ^ inv_out);
}

// The same column pass for larger radii (2-D balls only) as two one-sided running minima instead of a (2k + 1)-entry search:
// the L1 ball holds a set pixel iff min over yy of g[yy] + |yy - y| <= k (g = k + 1 where the row has none within k), and that
// minimum is min(forward scan, backward scan) of d <- min(g, d + 1); the Linf ball is the same with g' = (g <= k ? 0 : inf).
// One thread per (frame, 64-row segment, column), each scan started k rows outside the segment: 2 + k / 16 steps per pixel
// instead of 2k + 1 (the k = 32 and k = 50 dilations of the cloud stage took 0.09-0.24 ms each on a 24-date tile).
#define DCS_SEG 64
__global__ void __launch_bounds__(128) k_dilate_col_scan(const unsigned char* __restrict__ g, unsigned char* __restrict__ out, int T, int H, int W,
                                                         int k, int conn, int inv_out) {
  const int segs = (H + DCS_SEG - 1) / DCS_SEG;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)T * segs * W) return;
  const int x = (int)(idx % W); const int64_t r = idx / W; const int seg = (int)(r % segs); const int t = (int)(r / segs);
  const unsigned char* gg = g + (int64_t)t * H * W + x;
  unsigned char* oo = out + (int64_t)t * H * W + x;
  const int ys = seg * DCS_SEG, ye = min(ys + DCS_SEG, H);
  const bool linf = conn != 1;
  int d = 255;
  for (int yy = max(ys - k, 0); yy < ye; ++yy) {
    int gv = gg[(int64_t)yy * W];
    if (linf) gv = gv <= k ? 0 : 255;
    d = min(gv, d + 1);
    if (yy >= ys) oo[(int64_t)yy * W] = (unsigned char)min(d, 255);
  }
  d = 255;
  for (int yy = min(ye - 1 + k, H - 1); yy >= ys; --yy) {
    int gv = gg[(int64_t)yy * W];
    if (linf) gv = gv <= k ? 0 : 255;
    d = min(gv, d + 1);
    if (yy < ye) {
      const int best = min(d, (int)oo[(int64_t)yy * W]);
      oo[(int64_t)yy * W] = (unsigned char)((best <= k ? 1 : 0) ^ inv_out);
    }
  }
}
static void launch_dilate_col(stc_ctx* ctx, const unsigned char* g, unsigned char* out, int frames, int H, int W, int k, int conn, int inv_out,
                              int three_d) {
  const int64_t N = (int64_t)frames * H * W;
  if (!three_d && k >= 6) {
    const int64_t work = (int64_t)frames * ((H + DCS_SEG - 1) / DCS_SEG) * W;
    TraceScope ts_(ctx, "k_dilate_col_scan");
    k_dilate_col_scan<<<cdiv(work, 128), 128, 0, ctx->stream>>>(g, out, frames, H, W, k, conn, inv_out);
  } else {
    TraceScope ts_(ctx, "k_dilate_col");
    k_dilate_col<<<cdiv(N, 256), 256, 0, ctx->stream>>>(g, out, frames, H, W, k, conn, inv_out, three_d);
  }
  ctx->launches++;
}

// shared by stc_cloud.cu / stc_cloudfill.cu / the host wrapper below; radii above 32 are done as successive balls
// (an L1 / Linf ball of radius a + b is the ball of radius a dilated by the ball of radius b)
int morph_dilate_dev(stc_ctx* ctx, const unsigned char* in, unsigned char* out, int frames, int H, int W, int k, int conn, int inv_in,
                     int inv_out, int three_d) {
  PoolBuf g, tmp;
  const int64_t N = (int64_t)frames * H * W;
  STC_CUDA(g.alloc((size_t)N));
  const unsigned char* src = in; int inv = inv_in;
  while (k > 32) {
    if (!tmp.p) STC_CUDA(tmp.alloc((size_t)N));
    int rc = morph_rowdist_dev(ctx, src, inv ? 1 : 0, (int64_t)frames * H, W, 32, g.as<unsigned char>());
    if (rc) return rc;
    launch_dilate_col(ctx, g.as<unsigned char>(), tmp.as<unsigned char>(), frames, H, W, 32, conn, 0, three_d);
    src = tmp.as<unsigned char>(); inv = 0; k -= 32;
  }
  int rc = morph_rowdist_dev(ctx, src, inv ? 1 : 0, (int64_t)frames * H, W, k, g.as<unsigned char>());
  if (rc) return rc;
  launch_dilate_col(ctx, g.as<unsigned char>(), out, frames, H, W, k, conn, inv_out, three_d);
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

// column pass of the capped squared Euclidean distance: d2 = min over |dy| <= radius of dy^2 + g^2 (g <= radius), else
// radius^2 + 1
__device__ __forceinline__ int edt_col_min(const unsigned char* __restrict__ gcol, int y, int H, int W, int radius) {
  const int r2 = radius * radius;
  int best = r2 + 1;
  const int y0 = max(y - radius, 0), y1 = min(y + radius, H - 1);
  for (int yy = y0; yy <= y1; ++yy) {
    const int gx = gcol[(int64_t)yy * W];
    if (gx <= radius) { const int d2 = gx * gx + (yy - y) * (yy - y); best = d2 < best ? d2 : best; }
  }
  return best;
}

__global__ void __launch_bounds__(256) k_edt_sq_col(const unsigned char* __restrict__ g, int* __restrict__ out, int n, int H, int W, int radius) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * H * W) return;
  const int x = (int)(idx % W); const int64_t r = idx / W; const int y = (int)(r % H); const int d = (int)(r / H);
  const int best = edt_col_min(g + (int64_t)d * H * W + x, y, H, W, radius);
  out[idx] = best <= radius * radius ? best : radius * radius + 1;
}

// out = exists non-zero pixel within Euclidean distance <= radius  (1 - (edt(1 - in) > radius)).
// A frame with no non-zero pixel has no background for scipy's distance_transform_edt, which then measures from the
// virtual site (row -1, column 0): d^2 = (y+1)^2 + x^2 (scipy 1.x feature-transform initialisation; pinned by
// tests/test_cloud_masks.py against the reference run in this image).
__global__ void __launch_bounds__(256) k_edt_grow_col(const unsigned char* __restrict__ g, unsigned char* __restrict__ out, int T, int H, int W,
                                                      int radius, const int* __restrict__ frame_count) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)T * H * W) return;
  const int x = (int)(idx % W); const int64_t r = idx / W; const int y = (int)(r % H); const int t = (int)(r / H);
  const int r2 = radius * radius;
  if (frame_count[t] == 0) { out[idx] = ((y + 1) * (y + 1) + x * x) <= r2; return; }
  out[idx] = edt_col_min(g + (int64_t)t * H * W + x, y, H, W, radius) <= r2;
}

int morph_edt_grow_dev(stc_ctx* ctx, const unsigned char* in, unsigned char* out, int T, int H, int W, int radius, const int* frame_count_dev) {
  PoolBuf g;
  const int64_t N = (int64_t)T * H * W;
  STC_CUDA(g.alloc((size_t)N));
  int rc = morph_rowdist_dev(ctx, in, 0, (int64_t)T * H, W, radius, g.as<unsigned char>());
  if (rc) return rc;
  { TraceScope ts_(ctx, "k_edt_grow_col"); k_edt_grow_col<<<cdiv(N, 256), 256, 0, ctx->stream>>>(g.as<unsigned char>(), out, T, H, W, radius, frame_count_dev); }
  ctx->launches++;
  return STC_OK;
}

// feather value from the capped distance: a = 1 - min(EDT, 12) / 12, a < 0.2 -> 0 (float64 like SciPy / NumPy);
// flags[date] == 0 leaves the date untouched (np.sum(mask) == 0 guard)
__global__ void __launch_bounds__(256) k_feather_col(const float* __restrict__ mask, const unsigned char* __restrict__ g,
                                                     const int* __restrict__ flags, float* __restrict__ out, int n, int H, int W) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * H * W) return;
  const int x = (int)(idx % W); const int64_t r = idx / W; const int y = (int)(r % H); const int d = (int)(r / H);
  if (!flags[d]) { out[idx] = mask[idx]; return; }
  const int best = edt_col_min(g + (int64_t)d * H * W + x, y, H, W, 12);
  const double dist = (best <= 144) ? sqrt((double)best) : 12.0;
  double v = 1.0 - (dist / 12.0);
  if (v < 0.2) v = 0.0;
  out[idx] = (float)v;
}

// flags[date] = any(mask[date] > 0): the `if np.sum(mask[date]) > 0` guard (the masks are >= 0, so the sum is positive
// exactly when one element is)
__global__ void __launch_bounds__(256) k_any_positive(const float* __restrict__ mask, int* __restrict__ flags, int HW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool pos = i < HW && mask[(int64_t)blockIdx.y * HW + i] > 0.f;
  if (__any_sync(0xffffffffu, pos) && (threadIdx.x & 31) == 0) atomicOr(flags + blockIdx.y, 1);
}

// Flat 2-D max / min filter over the window [lo, hi] x [lo, hi] with SciPy 'reflect' boundary: one block per 64 x 16 output
// tile, the tile and its halo staged once in shared memory (the boundary folded in there), then the column pass and the row
// pass from shared memory.  The per-axis kernel of the first version read `hi - lo + 1` global values per pixel and pass and computed a
// reflected index for each of them: 0.1-0.2 ms per pass on a 12-24 date tile, 1.3-2.5 ms per tile for the three featherings.
#define W2_TX 64
#define W2_TY 16
template <bool IS_MAX>
__global__ void __launch_bounds__(256) k_window2d(const float* __restrict__ in, const int* __restrict__ flags, float* __restrict__ out,
                                                  int H, int W, int lo, int hi) {
  extern __shared__ float w2_sm[];
  const int d = blockIdx.z, x0 = blockIdx.x * W2_TX, y0 = blockIdx.y * W2_TY;
  const float* m = in + (int64_t)d * H * W;
  float* o = out + (int64_t)d * H * W;
  if (flags && !flags[d]) {                                         // date without any mask: the closing is skipped
    for (int e = threadIdx.x; e < W2_TX * W2_TY; e += blockDim.x) {
      const int x = x0 + e % W2_TX, y = y0 + e / W2_TX;
      if (x < W && y < H) o[(int64_t)y * W + x] = m[(int64_t)y * W + x];
    }
    return;
  }
  const int span = hi - lo, SW = W2_TX + span, SH = W2_TY + span;
  float* a = w2_sm;                 // [SH][SW] input tile + halo
  float* b = w2_sm + SH * SW;       // [W2_TY][SW] after the column pass
  for (int e = threadIdx.x; e < SH * SW; e += blockDim.x) {
    const int r = e / SW, c = e - r * SW;
    a[e] = m[(int64_t)reflect_index(y0 + lo + r, H) * W + reflect_index(x0 + lo + c, W)];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < W2_TY * SW; e += blockDim.x) {
    const int r = e / SW, c = e - r * SW;
    float v = a[r * SW + c];
    for (int k = 1; k <= span; ++k) { const float t = a[(r + k) * SW + c]; v = IS_MAX ? fmaxf(v, t) : fminf(v, t); }
    b[e] = v;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < W2_TY * W2_TX; e += blockDim.x) {
    const int r = e / W2_TX, c = e - r * W2_TX;
    if (x0 + c >= W || y0 + r >= H) continue;
    float v = b[r * SW + c];
    for (int k = 1; k <= span; ++k) { const float t = b[r * SW + c + k]; v = IS_MAX ? fmaxf(v, t) : fminf(v, t); }
    o[(int64_t)(y0 + r) * W + x0 + c] = v;
  }
}

int pre_edt_sq_dev(stc_ctx* ctx, const unsigned char* target_dev, int n, int H, int W, int radius, int* out_dev) {
  if (radius < 1 || radius > 64) STC_FAIL(STC_ERR_ARG, "edt_sq: radius must be in 1..64");
  if (radius > 32) STC_FAIL(STC_ERR_ARG, "edt_sq: radius must be in 1..32");
  PoolBuf g;
  STC_CUDA(g.alloc((size_t)n * H * W));
  int rc = morph_rowdist_dev(ctx, target_dev, 0, (int64_t)n * H, W, radius, g.as<unsigned char>());
  if (rc) return rc;
  { TraceScope ts_(ctx, "k_edt_sq_col"); k_edt_sq_col<<<cdiv((int64_t)n * H * W, 256), 256, 0, ctx->stream>>>(g.as<unsigned char>(), out_dev, n, H, W, radius); }
  STC_CUDA(cudaGetLastError());
  ctx->launches++;
  return STC_OK;
}

int pre_feather_dev(stc_ctx* ctx, const float* mask_dev, int n, int H, int W, int size, float* tmp_a, float* tmp_b,
                    float* sums_dev, float* out_dev) {
  if (size < 1 || size > 64) STC_FAIL(STC_ERR_ARG, "feather: closing size must be in 1..64");
  int64_t tot = (int64_t)n * H * W;
  int grid = cdiv(tot, 256);
  int* flags = reinterpret_cast<int*>(sums_dev);                 // [n] ints in the caller's scratch
  PoolBuf g;
  STC_CUDA(g.alloc((size_t)tot));
  STC_CUDA(cudaMemsetAsync(flags, 0, (size_t)n * 4, ctx->stream));
  { TraceScope ts_(ctx, "k_any_positive"); k_any_positive<<<dim3(cdiv(H * W, 256), n), 256, 0, ctx->stream>>>(mask_dev, flags, H * W); }
  int rc = morph_rowdist_dev(ctx, mask_dev, 2, (int64_t)n * H, W, 12, g.as<unsigned char>());
  if (rc) return rc;
  { TraceScope ts_(ctx, "k_feather_col"); k_feather_col<<<grid, 256, 0, ctx->stream>>>(mask_dev, g.as<unsigned char>(), flags, tmp_a, n, H, W); }
  int dlo, dhi, elo, ehi;
  if (size & 1) { dlo = elo = -(size / 2); dhi = ehi = size / 2; }
  else { dlo = -(size / 2 - 1); dhi = size / 2; elo = -(size / 2); ehi = size / 2 - 1; }
  {
    // grey closing = max filter, then min filter (window <= 64: <= 50 KB of shared memory)
    const dim3 g2(cdiv(W, W2_TX), cdiv(H, W2_TY), n);
    auto smem_of = [](int span) { return (size_t)((W2_TY + span) * (W2_TX + span) + W2_TY * (W2_TX + span)) * 4; };
    static bool configured = false;
    if (!configured) {
      STC_CUDA(cudaFuncSetAttribute(k_window2d<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of(63)));
      STC_CUDA(cudaFuncSetAttribute(k_window2d<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of(63)));
      configured = true;
    }
    { TraceScope ts_(ctx, "k_window2d"); k_window2d<true><<<g2, 256, smem_of(dhi - dlo), ctx->stream>>>(tmp_a, flags, tmp_b, H, W, dlo, dhi); }
    { TraceScope ts_(ctx, "k_window2d"); k_window2d<false><<<g2, 256, smem_of(ehi - elo), ctx->stream>>>(tmp_b, flags, out_dev, H, W, elo, ehi); }
  }
  STC_CUDA(cudaGetLastError());
  ctx->launches += 4;
  return STC_OK;
}

int pre_binary_dilate_dev(stc_ctx* ctx, const unsigned char* in_dev, int n, int H, int W, int iterations, int conn,
                          unsigned char* out_dev) {
  if (iterations < 1 || iterations > 64 || (conn != 1 && conn != 2)) STC_FAIL(STC_ERR_ARG, "binary_dilate: bad iterations/connectivity");
  return morph_dilate_dev(ctx, in_dev, out_dev, n, H, W, iterations, conn, 0, 0, 0);
}
