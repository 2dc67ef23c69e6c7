// Gaussian overlap-blend mosaic of subtile predictions -> uint8 tile
// (load_mosaic_predictions depth == 1, src/download_and_predict_job.py:1515-1641;
//  fspecial_gauss :1489-1501, calc_overlap :1503-1512).
//
// Layers keep the reference's order (the order of its os.listdir walk) because the float32
// reductions over the layer axis follow NumPy's pairwise-summation tree (8 strided
// accumulators, then the tail), which is order dependent.
#include "stc_common.cuh"

struct MosaicParams {
  const float* preds;      // [n][S][S] as saved (row = file y axis): probabilities 0..1 or 255 no-data
  const int* xs; const int* ys; const int* placed;   // [n] canvas offsets; placed=0 for all-255 subtiles
  const float* gauss;      // [S][S] float32 cast of the float64 kernel
  const float* mult;       // [n] per-layer multipliers (already capped)
  int n, S, Hc, Wc;
};

// value of layer i at canvas (X,Y) following :1565-1572: transpose, x100 where < 255
__device__ __forceinline__ bool layer_value(const MosaicParams& p, int i, int X, int Y, float& v, int& px, int& py) {
  if (!p.placed[i]) return false;
  px = X - p.xs[i]; py = Y - p.ys[i];
  if (px < 0 || py < 0 || px >= p.S || py >= p.S) return false;
  float raw = p.preds[((int64_t)i * p.S + py) * p.S + px];    // prediction.T[px, py] = prediction[py, px]
  v = (raw < 255.f) ? __fmul_rn(raw, 100.f) : raw;
  return true;
}

// NumPy float32 add.reduce over a contiguous axis of length n (pairwise_sum, n <= 128)
__device__ __forceinline__ float np_sum(const float* a, int n) {
  if (n < 8) {
    float r = 0.f;
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
    return r;
  }
  float r[8];
  for (int k = 0; k < 8; ++k) r[k] = a[k];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int k = 0; k < 8; ++k) r[k] = __fadd_rn(r[k], a[i + k]);
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, a[i]);
  return res;
}

// one block per layer i: diff[i][a][b] = | nanmean_{j != i}(v_j) - v_i | over layer i's footprint
// (calc_overlap :1503-1512 without the final nanmean, which the host takes with NumPy so the
// float32 pairwise summation order is NumPy's own).  The mean over the other layers follows
// np.nanmean(axis=-1) on the compacted layer axis: layer i deleted, layers that never touch the
// footprint deleted (:1508-1509), NaN -> 0, divided by the non-NaN count.
__global__ void __launch_bounds__(256) mosaic_ratio_kernel(MosaicParams p, float* diffs) {
  const int i = blockIdx.x;
  __shared__ int keep[64]; __shared__ int L;
  if (threadIdx.x == 0) {
    int l = 0;
    for (int j = 0; j < p.n; ++j) {
      if (j == i || !p.placed[j]) continue;
      int dx = p.xs[j] - p.xs[i], dy = p.ys[j] - p.ys[i];
      if (dx > -p.S && dx < p.S && dy > -p.S && dy < p.S) keep[l++] = j;
    }
    L = l;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < p.S * p.S; idx += blockDim.x) {
    float out = nanf("");
    if (p.placed[i]) {
      int px = idx / p.S, py = idx - px * p.S;
      int X = p.xs[i] + px, Y = p.ys[i] + py;
      float vi; int a, b;
      layer_value(p, i, X, Y, vi, a, b);
      // np.delete(...) hands nanmean an array whose layer axis has the LARGEST stride, so NumPy
      // reduces it as the outer loop: a plain sequential float32 sum in layer order (verified).
      float s = 0.f; int m = 0;
      for (int l = 0; l < L; ++l) {
        float vj;
        if (layer_value(p, keep[l], X, Y, vj, a, b)) { s = __fadd_rn(s, vj); ++m; }
      }
      if (m > 0) out = fabsf(__fsub_rn(__fdiv_rn(s, (float)m), vi));
    }
    diffs[(int64_t)i * p.S * p.S + idx] = out;
  }
}

// per canvas pixel: normalised Gaussian weights, weighted nansum, uint8 rules (:1609-1626)
__global__ void __launch_bounds__(128) mosaic_blend_kernel(MosaicParams p, unsigned char* out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.Hc * p.Wc) return;
  int X = idx / p.Wc, Y = idx - X * p.Wc;
  float w[64], v[64];
  int nan_count = 0;
  for (int i = 0; i < p.n; ++i) {
    float vi; int px, py;
    if (layer_value(p, i, X, Y, vi, px, py)) {
      float g = (vi > 100.f) ? 0.f : p.gauss[px * p.S + py];
      w[i] = __fmul_rn(g, p.mult[i]);
      if (vi > 100.f) { v[i] = 0.f; ++nan_count; } else v[i] = vi;
    } else { w[i] = 0.f; v[i] = 0.f; ++nan_count; }
  }
  float W = np_sum(w, p.n);
  for (int i = 0; i < p.n; ++i) v[i] = __fmul_rn(v[i], __fdiv_rn(w[i], W));   // NaN*x entries were zeroed (nansum)
  float r = np_sum(v, p.n);
  unsigned char o;
  if (nan_count == p.n || isnan(r)) o = 255;
  else {
    o = (unsigned char)r;                 // astype(np.uint8): truncation
    if (o <= 15) o = 0;
    if (o > 100) o = 255;
  }
  out[idx] = o;
}

// output[binary_dilation(output == 255, 3x3, iterations=10)] = 255   (:1636-1640)
__global__ void __launch_bounds__(256) mosaic_dilate_kernel(const unsigned char* in, unsigned char* out, int Hc, int Wc, int iters) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Hc * Wc) return;
  int X = idx / Wc, Y = idx - X * Wc;
  unsigned char o = in[idx];
  for (int dx = -iters; dx <= iters && o != 255; ++dx) {
    int xx = X + dx; if (xx < 0 || xx >= Hc) continue;
    for (int dy = -iters; dy <= iters; ++dy) {
      int yy = Y + dy; if (yy < 0 || yy >= Wc) continue;
      if (in[xx * Wc + yy] == 255) { o = 255; break; }
    }
  }
  out[idx] = o;
}

// Feature mosaic, load_mosaic_predictions with depth > 1 (src/download_and_predict_job.py:1540-1592,1628-1635):
// every saved int16 feature stack [S,S,D] is transposed, placed, weighted with the plain Gaussian (no no-data
// zeroing, no overlap reweighting), weights normalised over the layer axis, nansum, int16 truncation.  Float32
// reductions over the layer axis follow NumPy's pairwise order (np_sum).  out: [D][Hc][Wc] int16.
__global__ void __launch_bounds__(128) mosaic_feats_kernel(const short* feats, const int* xs, const int* ys, const float* gauss,
                                                           int n, int S, int D, int Hc, int Wc, short* out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Hc * Wc) return;
  int X = idx / Wc, Y = idx - X * Wc;
  float w[64], v[64];
  int64_t src[64];
  for (int i = 0; i < n; ++i) {
    int px = X - xs[i], py = Y - ys[i];
    if (px < 0 || py < 0 || px >= S || py >= S) { w[i] = 0.f; src[i] = -1; }
    else { w[i] = gauss[px * S + py]; src[i] = (((int64_t)i * S + py) * S + px) * D; }   // prediction.T[c, px, py] = prediction[py, px, c]
  }
  const float W = np_sum(w, n);
  for (int i = 0; i < n; ++i) w[i] = __fdiv_rn(w[i], W);          // 0/0 = NaN where nothing covers the pixel
  for (int c = 0; c < D; ++c) {
    for (int i = 0; i < n; ++i) {
      float prod = (src[i] >= 0) ? __fmul_rn((float)feats[src[i] + c], w[i]) : nanf("");
      v[i] = isnan(prod) ? 0.f : prod;                              // np.nansum
    }
    out[((int64_t)c * Hc + X) * Wc + Y] = (short)np_sum(v, n);     // np.int16(): truncation
  }
}

int pre_feature_mosaic_dev(stc_ctx* ctx, const short* feats_dev, const int* xs_dev, const int* ys_dev, const float* gauss_dev,
                           int n, int S, int D, int Hc, int Wc, short* out_dev) {
  if (n < 1 || n > 64) STC_FAIL(STC_ERR_ARG, "feature mosaic: 1..64 subtiles supported");
  { TraceScope ts_(ctx, "mosaic_feats_kernel"); mosaic_feats_kernel<<<cdiv((int64_t)Hc * Wc, 128), 128, 0, ctx->stream>>>(feats_dev, xs_dev, ys_dev, gauss_dev, n, S, D, Hc, Wc, out_dev); }
  STC_CUDA(cudaGetLastError()); ctx->launches++;
  return STC_OK;
}

int pre_gauss_mosaic_dev(stc_ctx* ctx, const float* preds_dev, const int* xs_dev, const int* ys_dev, const int* placed_dev,
                         const float* gauss_dev, float* mult_dev, float* diffs_dev, int stage,
                         int n, int S, int Hc, int Wc, unsigned char* tmp_dev, unsigned char* out_dev) {
  if (n < 1 || n > 64) STC_FAIL(STC_ERR_ARG, "mosaic: 1..64 subtiles supported");
  MosaicParams p{preds_dev, xs_dev, ys_dev, placed_dev, gauss_dev, mult_dev, n, S, Hc, Wc};
  if (stage == 0) {
    { TraceScope ts_(ctx, "mosaic_ratio_kernel"); mosaic_ratio_kernel<<<n, 256, 0, ctx->stream>>>(p, diffs_dev); }
    STC_CUDA(cudaGetLastError()); ctx->launches++;
  } else {
    { TraceScope ts_(ctx, "mosaic_blend_kernel"); mosaic_blend_kernel<<<cdiv((int64_t)Hc * Wc, 128), 128, 0, ctx->stream>>>(p, tmp_dev); }
    STC_CUDA(cudaGetLastError()); ctx->launches++;
    { TraceScope ts_(ctx, "mosaic_dilate_kernel"); mosaic_dilate_kernel<<<cdiv((int64_t)Hc * Wc, 256), 256, 0, ctx->stream>>>(tmp_dev, out_dev, Hc, Wc, 10); }
    STC_CUDA(cudaGetLastError()); ctx->launches++;
  }
  return STC_OK;
}
