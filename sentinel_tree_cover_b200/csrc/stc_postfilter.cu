// Per-subtile post-filters and the small reductions around the mosaic, so that no array arithmetic of
// the path is left to host NumPy:
//   normalize_subtile                 src/download_and_predict_job.py:316-325
//   identify_bright_bare_surfaces     :1099-1122
//   no-image block vote, attenuation, np.around(.., 3)   :1451-1483
//   np.sum / np.nanmean of float32 maps in NumPy's pairwise order (load_mosaic_predictions :1573, calc_overlap :1503-1512)
#include "stc_common.cuh"

void maskop_dilate(stc_ctx* ctx, const unsigned char* in, unsigned char* out, int frames, int H, int W, int k, int conn, int inv_in,
                   int inv_out, int three_d);
int pre_edt_sq_dev(stc_ctx* ctx, const unsigned char* target_dev, int n, int H, int W, int radius, int* out_dev);

namespace {

struct PBuf { void* p = nullptr; ~PBuf() { if (p) stc_dfree(p); } template <typename T> T* as() { return (T*)p; } };

// ---- np.sum of contiguous float32 segments, NumPy's pairwise order (see stc_cloud.cu k_np_tree / k_np_leaves) ----
// mode 0: x          mode 1: x < 255 ? x*100 : x  (the in-place scaling of :1570 before the sum of :1573)
// mode 2: NaN -> 0, valid[] counts the non-NaN values (np.nanmean)
__device__ __forceinline__ float seg_value(float v, int mode) {
  if (mode == 1) return (v < 255.f) ? __fmul_rn(v, 100.f) : v;
  if (mode == 2) return isnan(v) ? 0.f : v;
  return v;
}
__device__ float leaf_sum(const float* a, int n, int mode) {
  if (n < 8) { float r = 0.f; for (int i = 0; i < n; ++i) r = __fadd_rn(r, seg_value(a[i], mode)); return r; }
  float r[8];
  for (int k = 0; k < 8; ++k) r[k] = seg_value(a[k], mode);
  int i = 8;
  for (; i < n - (n % 8); i += 8) for (int k = 0; k < 8; ++k) r[k] = __fadd_rn(r[k], seg_value(a[i + k], mode));
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])), __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, seg_value(a[i], mode));
  return res;
}
__global__ void __launch_bounds__(1024) k_np_sum_seg(const float* __restrict__ data, int len, int mode, int leaf_cap, int2* __restrict__ leaves,
                                                     float* __restrict__ leafsum, float* __restrict__ sum_out, int* __restrict__ valid_out) {
  const int sgm = blockIdx.x;
  const float* a = data + (int64_t)sgm * len;
  int2* lv = leaves + (int64_t)sgm * leaf_cap; float* ls = leafsum + (int64_t)sgm * leaf_cap;
  __shared__ int L; __shared__ int nvalid;
  if (threadIdx.x == 0) {
    nvalid = 0;
    int2 st[40]; int sp = 0; st[0] = make_int2(0, len); int l = 0;
    while (sp >= 0) {
      int2 f = st[sp--];
      if (f.y <= 128) { lv[l++] = f; continue; }
      int n2 = f.y / 2; n2 -= n2 % 8;
      st[++sp] = make_int2(f.x + n2, f.y - n2);
      st[++sp] = make_int2(f.x, n2);
    }
    L = l;
  }
  __syncthreads();
  for (int l = threadIdx.x; l < L; l += blockDim.x) ls[l] = leaf_sum(a + lv[l].x, lv[l].y, mode);
  if (mode == 2) {
    int c = 0;
    for (int i = threadIdx.x; i < len; i += blockDim.x) c += !isnan(a[i]);
    atomicAdd(&nvalid, c);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    struct F { int n, stage; float a; };
    F st[40]; int sp = 0; st[0].n = len; st[0].stage = 0; st[0].a = 0.f;
    float ret = 0.f; int li = 0;
    while (sp >= 0) {
      F& f = st[sp];
      if (f.n <= 128) { ret = ls[li++]; --sp; continue; }
      int n2 = f.n / 2; n2 -= n2 % 8;
      if (f.stage == 0) { f.stage = 1; ++sp; st[sp].n = n2; st[sp].stage = 0; }
      else if (f.stage == 1) { f.a = ret; f.stage = 2; ++sp; st[sp].n = f.n - n2; st[sp].stage = 0; }
      else { ret = __fadd_rn(f.a, ret); --sp; }
    }
    sum_out[sgm] = ret;
    if (valid_out) valid_out[sgm] = (mode == 2) ? nvalid : len;
  }
}

// ---- normalize_subtile ----
struct NormParams { float lo[32], hi[32], mid[32], half[32]; };
__global__ void __launch_bounds__(256) k_normalize(float* __restrict__ x, int64_t n, int C, NormParams p) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  float v = x[i];
  // np.clip propagates NaN; fminf/fmaxf would not
  if (!isnan(v)) v = fminf(fmaxf(v, p.lo[c]), p.hi[c]);
  x[i] = __fdiv_rn(__fsub_rn(v, p.mid[c]), p.half[c]);
}

// ---- bright bare surfaces ----
// blockIdx.y = image of the batch: img [nimg][F][HW][C], out [nimg][HW]
__global__ void __launch_bounds__(256) k_bright_candidates(const float* __restrict__ img, int F, int HW, int C, unsigned char* __restrict__ out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  img += (int64_t)blockIdx.y * F * HW * C; out += (int64_t)blockIdx.y * HW;
  int cnt = 0;
  for (int f = 0; f < F; ++f) {
    const float* x = img + ((int64_t)f * HW + p) * C;
    auto clip01 = [](float v) { return isnan(v) ? v : fminf(fmaxf(v, 0.f), 1.f); };
    const float B = clip01(x[0]), R = clip01(x[2]), N = clip01(x[3]);
    float den = __fadd_rn(__fsub_rn(__fadd_rn(N, __fmul_rn(6.f, R)), __fmul_rn(7.5f, B)), 1.f);
    float evi = __fmul_rn(2.5f, __fdiv_rn(__fsub_rn(N, R), den));
    if (!isnan(evi)) evi = fminf(fmaxf(evi, -1.5f), 1.5f);
    bool c = __fdiv_rn(x[3], __fadd_rn(x[8], 0.01f)) < 0.9f;
    c = c && (__fdiv_rn(__fadd_rn(__fadd_rn(x[0], x[1]), x[2]), 3.f) > 0.2f);
    c = c && (evi < 0.3f);
    cnt += c;
  }
  out[p] = cnt > 1;
}
// ramp = min(sqrt(d2), 3) / 3 in float64, cropped by `crop` on every side
__global__ void __launch_bounds__(256) k_ramp_crop(const int* __restrict__ d2, int H, int W, int crop, double* __restrict__ out) {
  const int Ho = H - 2 * crop, Wo = W - 2 * crop;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Ho * Wo) return;
  d2 += (int64_t)blockIdx.y * H * W; out += (int64_t)blockIdx.y * Ho * Wo;
  int y = i / Wo + crop, x = i % Wo + crop;
  double d = sqrt((double)d2[y * W + x]);
  if (d > 3.0) d = 3.0;
  out[i] = __ddiv_rn(d, 3.0);
}

// ---- no-image block vote + attenuation + rounding ----
__global__ void __launch_bounds__(256) k_lt1(const float* __restrict__ m, int Hm, int Wm, int crop, unsigned char* __restrict__ out) {
  const int Ho = Hm - 2 * crop, Wo = Wm - 2 * crop;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Ho * Wo) return;
  m += (int64_t)blockIdx.y * Hm * Wm; out += (int64_t)blockIdx.y * Ho * Wo;
  out[i] = m[(i / Wo + crop) * Wm + (i % Wo + crop)] < 1.f;
}
__global__ void k_block_vote(const unsigned char* __restrict__ m, int64_t mstride, int blocks, int bs, int thresh, unsigned char* __restrict__ vote) {
  const int by = blockIdx.y, bx = blockIdx.x, side = blocks * bs;
  m += (int64_t)blockIdx.z * mstride; vote += (int64_t)blockIdx.z * 256;
  __shared__ int s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  int c = 0;
  for (int i = threadIdx.x; i < bs * bs; i += blockDim.x) c += m[(by * bs + i / bs) * side + bx * bs + i % bs];
  atomicAdd(&s, c);
  __syncthreads();
  if (threadIdx.x == 0) vote[by * blocks + bx] = s > thresh;
}
__global__ void __launch_bounds__(256) k_attenuate_round(const float* __restrict__ preds, const double* __restrict__ ramp,
                                                         const unsigned char* __restrict__ vote, int S, int blocks, int bs,
                                                         float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * S) return;
  preds += (int64_t)blockIdx.y * S * S; ramp += (int64_t)blockIdx.y * S * S; out += (int64_t)blockIdx.y * S * S;
  if (vote) vote += (int64_t)blockIdx.y * 256;
  const int y = i / S, x = i % S;
  float p = preds[i];
  if (vote && vote[((y + 1) / bs) * blocks + (x + 1) / bs]) p = 255.f;       // votes expanded to blocks, cropped by 1 (:1465-1466)
  double v = __dmul_rn((double)p, ramp[i]);
  v = __ddiv_rn(rint(__dmul_rn(v, 1000.0)), 1000.0);                          // np.around(x, 3)
  out[i] = (float)v;
}

// nimg images at once: img [nimg][F][H][W][C]; scratch a, b [nimg][H*W] u8, d2 [nimg][H*W] int; ramp [nimg][(H-14)*(W-14)] double
int bright_bare_dev(stc_ctx* ctx, const float* img_dev, int nimg, int F, int H, int W, int C, unsigned char* a, unsigned char* b, int* d2,
                    double* ramp_dev) {
  const int HW = H * W;
  { TraceScope ts_(ctx, "k_bright_candidates"); k_bright_candidates<<<dim3(cdiv(HW, 256), nimg), 256, 0, ctx->stream>>>(img_dev, F, HW, C, a); }
  maskop_dilate(ctx, a, b, nimg, H, W, 2, 1, 1, 0, 0);      // binary_dilation(1 - bright, 2)
  maskop_dilate(ctx, b, a, nimg, H, W, 1, 1, 1, 0, 0);      // binary_dilation(1 - that, 1)
  int rc = pre_edt_sq_dev(ctx, a, nimg, H, W, 3, d2);
  if (rc) return rc;
  { TraceScope ts_(ctx, "k_ramp_crop"); k_ramp_crop<<<dim3(cdiv((H - 14) * (W - 14), 256), nimg), 256, 0, ctx->stream>>>(d2, H, W, 7, ramp_dev); }
  ctx->launches += 2;
  return STC_OK;
}

}  // namespace

// device core of stc_np_sum_host: data_dev [nseg, len] -> sum_dev [nseg] float32, valid_dev [nseg] int32
int post_np_sum_dev(stc_ctx* ctx, const float* data_dev, int nseg, int len, int mode, float* sum_dev, int* valid_dev) {
  const int leaf_cap = len / 32 + 8;
  PBuf lv, ls;
  STC_CUDA(stc_dmalloc(&lv.p, (size_t)nseg * leaf_cap * 8)); STC_CUDA(stc_dmalloc(&ls.p, (size_t)nseg * leaf_cap * 4));
  { TraceScope ts_(ctx, "k_np_sum_seg"); k_np_sum_seg<<<nseg, 1024, 0, ctx->stream>>>(data_dev, len, mode, leaf_cap, lv.as<int2>(), ls.as<float>(), sum_dev, valid_dev); }
  STC_CUDA(cudaGetLastError()); ctx->launches++;
  return STC_OK;
}

extern "C" int stc_np_sum_host(stc_ctx* ctx, const float* data_host, int nseg, int len, int mode, float* sum_host, int32_t* valid_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!data_host || !sum_host || nseg < 1 || len < 1 || mode < 0 || mode > 2) STC_FAIL(STC_ERR_ARG, "np_sum: bad argument");
  const int leaf_cap = len / 32 + 8;
  PBuf d, lv, ls, so, vo;
  STC_CUDA(stc_dmalloc(&d.p, (size_t)nseg * len * 4)); STC_CUDA(stc_dmalloc(&lv.p, (size_t)nseg * leaf_cap * 8));
  STC_CUDA(stc_dmalloc(&ls.p, (size_t)nseg * leaf_cap * 4)); STC_CUDA(stc_dmalloc(&so.p, nseg * 4)); STC_CUDA(stc_dmalloc(&vo.p, nseg * 4));
  STC_CUDA(cudaMemcpyAsync(d.p, data_host, (size_t)nseg * len * 4, cudaMemcpyHostToDevice, ctx->stream));
  { TraceScope ts_(ctx, "k_np_sum_seg"); k_np_sum_seg<<<nseg, 1024, 0, ctx->stream>>>(d.as<float>(), len, mode, leaf_cap, lv.as<int2>(), ls.as<float>(), so.as<float>(), vo.as<int>()); }
  ctx->launches++;
  STC_CUDA(cudaMemcpyAsync(sum_host, so.p, nseg * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (valid_host) STC_CUDA(cudaMemcpyAsync(valid_host, vo.p, nseg * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

extern "C" int stc_normalize_host(stc_ctx* ctx, float* x_host, int64_t npx, int C, const double* mins, const double* maxs) {
  if (!ctx) return STC_ERR_ARG;
  if (!x_host || !mins || !maxs || npx < 1 || C < 1 || C > 32) STC_FAIL(STC_ERR_ARG, "normalize: bad argument (C <= 32)");
  NormParams p;
  for (int c = 0; c < C; ++c) {      // python-float (double) constants, float32 array arithmetic
    p.lo[c] = (float)mins[c]; p.hi[c] = (float)maxs[c];
    p.mid[c] = (float)((maxs[c] + mins[c]) / 2); p.half[c] = (float)((maxs[c] - mins[c]) / 2);
  }
  PBuf d;
  STC_CUDA(stc_dmalloc(&d.p, (size_t)npx * C * 4));
  STC_CUDA(cudaMemcpyAsync(d.p, x_host, (size_t)npx * C * 4, cudaMemcpyHostToDevice, ctx->stream));
  { TraceScope ts_(ctx, "k_normalize"); k_normalize<<<cdiv(npx * C, 256), 256, 0, ctx->stream>>>(d.as<float>(), npx * C, C, p); }
  ctx->launches++;
  STC_CUDA(cudaMemcpyAsync(x_host, d.p, (size_t)npx * C * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

extern "C" int stc_bright_bare_host(stc_ctx* ctx, const float* img_host, int F, int H, int W, int C, double* ramp_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!img_host || !ramp_host || F < 1 || H < 15 || W < 15 || C < 9) STC_FAIL(STC_ERR_ARG, "bright_bare: bad argument");
  PBuf img, a, b, d2, ramp;
  const size_t bytes = (size_t)F * H * W * C * 4;
  STC_CUDA(stc_dmalloc(&img.p, bytes)); STC_CUDA(stc_dmalloc(&a.p, H * W)); STC_CUDA(stc_dmalloc(&b.p, H * W));
  STC_CUDA(stc_dmalloc(&d2.p, (size_t)H * W * 4)); STC_CUDA(stc_dmalloc(&ramp.p, (size_t)(H - 14) * (W - 14) * 8));
  STC_CUDA(cudaMemcpyAsync(img.p, img_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  int rc = bright_bare_dev(ctx, img.as<float>(), 1, F, H, W, C, a.as<unsigned char>(), b.as<unsigned char>(), d2.as<int>(), ramp.as<double>());
  if (rc) return rc;
  STC_CUDA(cudaMemcpyAsync(ramp_host, ramp.p, (size_t)(H - 14) * (W - 14) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

// device-level post-filter of nimg subtiles in one pass per kernel (all pointers on the device): preds / out [nimg][S*S],
// img [nimg][F][H][H][C], mc [nimg][H*H]; scratch per image: a, b [H*H] u8, d2 [H*H] int, ramp [S*S] double, na, nb [(S+2)^2] u8,
// vote [256] u8.  (Round 1 looped over the 36 subtiles of a tile: 288 launches of 6-10 us.)
int post_subtiles_dev(stc_ctx* ctx, const float* preds_dev, const float* img_dev, const float* mc_dev, int nimg, int S, int F, int C,
                      unsigned char* a, unsigned char* b, int* d2, double* ramp, unsigned char* na, unsigned char* nb, unsigned char* vote,
                      float* out_dev) {
  const int H = S + 14, Hm = S + 2;
  int rc = bright_bare_dev(ctx, img_dev, nimg, F, H, H, C, a, b, d2, ramp);
  if (rc) return rc;
  int blocks = 0, bs = 0, thresh = 0;
  if (S == 158) { blocks = 4; bs = 40; thresh = 400; }         // sum > 40*40*0.25
  else if (S == 142) { blocks = 9; bs = 16; thresh = 192; }    // sum > 16*16*0.75
  { TraceScope ts_(ctx, "k_lt1"); k_lt1<<<dim3(cdiv(Hm * Hm, 256), nimg), 256, 0, ctx->stream>>>(mc_dev, H, H, 6, na); }
  maskop_dilate(ctx, na, nb, nimg, Hm, Hm, 6, 2, 1, 1, 0);   // 1 - dilate(1 - x, 3x3, 6)
  maskop_dilate(ctx, nb, na, nimg, Hm, Hm, 6, 2, 0, 0, 0);   // dilate(.., 3x3, 6)
  if (blocks) { TraceScope ts_(ctx, "k_block_vote"); k_block_vote<<<dim3(blocks, blocks, nimg), 256, 0, ctx->stream>>>(na, (int64_t)Hm * Hm, blocks, bs, thresh, vote); }
  { TraceScope ts_(ctx, "k_attenuate_round"); k_attenuate_round<<<dim3(cdiv(S * S, 256), nimg), 256, 0, ctx->stream>>>(preds_dev, ramp, blocks ? vote : nullptr, S, blocks, bs, out_dev); }
  ctx->launches += 3;
  return STC_OK;
}
int post_subtile_dev(stc_ctx* ctx, const float* preds_dev, const float* img_dev, const float* mc_dev, int S, int F, int C,
                     unsigned char* a, unsigned char* b, int* d2, double* ramp, unsigned char* na, unsigned char* nb, unsigned char* vote,
                     float* out_dev) {
  return post_subtiles_dev(ctx, preds_dev, img_dev, mc_dev, 1, S, F, C, a, b, d2, ramp, na, nb, vote, out_dev);
}

// preds [S,S] float32; img [F,S+14,S+14,C] (the subtile stack before normalisation); min_clear [S+14,S+14] float32
// (min_clear_images_per_date before its [6:-6] crop).  Block vote only for S == 158 (4x4 blocks of 40, > 25 %) and
// S == 142 (9x9 blocks of 16, > 75 %), as in the reference.
extern "C" int stc_postprocess_subtile_host(stc_ctx* ctx, const float* preds_host, const float* img_host, const float* min_clear_host,
                                            int S, int F, int C, float* out_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!preds_host || !img_host || !min_clear_host || !out_host || S < 8 || F < 1 || C < 9) STC_FAIL(STC_ERR_ARG, "postprocess: bad argument");
  const int H = S + 14, Hm = S + 2;
  PBuf img, a, b, d2, ramp, preds, mc, na, nb, vote, out;
  const size_t bytes = (size_t)F * H * H * C * 4;
  STC_CUDA(stc_dmalloc(&img.p, bytes)); STC_CUDA(stc_dmalloc(&a.p, H * H)); STC_CUDA(stc_dmalloc(&b.p, H * H));
  STC_CUDA(stc_dmalloc(&d2.p, (size_t)H * H * 4)); STC_CUDA(stc_dmalloc(&ramp.p, (size_t)S * S * 8));
  STC_CUDA(stc_dmalloc(&preds.p, (size_t)S * S * 4)); STC_CUDA(stc_dmalloc(&mc.p, (size_t)H * H * 4));
  STC_CUDA(stc_dmalloc(&na.p, Hm * Hm)); STC_CUDA(stc_dmalloc(&nb.p, Hm * Hm)); STC_CUDA(stc_dmalloc(&vote.p, 256)); STC_CUDA(stc_dmalloc(&out.p, (size_t)S * S * 4));
  STC_CUDA(cudaMemcpyAsync(img.p, img_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(preds.p, preds_host, (size_t)S * S * 4, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(mc.p, min_clear_host, (size_t)H * H * 4, cudaMemcpyHostToDevice, ctx->stream));
  int rc = post_subtile_dev(ctx, preds.as<float>(), img.as<float>(), mc.as<float>(), S, F, C, a.as<unsigned char>(), b.as<unsigned char>(),
                            d2.as<int>(), ramp.as<double>(), na.as<unsigned char>(), nb.as<unsigned char>(), vote.as<unsigned char>(), out.as<float>());
  if (rc) return rc;
  STC_CUDA(cudaMemcpyAsync(out_host, out.p, (size_t)S * S * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}
