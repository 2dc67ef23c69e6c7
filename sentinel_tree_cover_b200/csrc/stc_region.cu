// Region-scale driver pieces: a 1 x 1 degree area is a grid of R x C overlapping
// patches (168 px windows every 58 px over an 11,130 px canvas).  Rows of the patch grid shard across GPUs; what this
// file adds to the per-patch forward is the data movement on either side of it:
//   region_gather_kernel : patch windows out of a device-resident canvas band [T][Hc][Wc][Cc] (optionally periodic,
//                          which is how the synthetic region is generated from a small base cube) -> [B][T][P][P][Cc]
//   region_blend_kernel  : Gaussian overlap blend of the patch outputs [rows][C][S][S] -> uint8 canvas rows, the
//                          generalisation of load_mosaic_predictions (src/download_and_predict_job.py:1515-1641) to
//                          the region: weights fspecial_gauss(S, 36), value = sum(w * p*100) / sum(w), uint8
//                          truncation, <= 15 -> 0, uncovered -> 255.  A canvas pixel adds its (at most 3 x 3)
//                          covering patches in row-major grid order with float32 operations, so a band blended on
//                          any rank is bit-identical to the single-GPU result (oracle/region_ref.py).
// HBM-bound copies; no reference counterpart at this scale (the reference mosaics one 6 x 6 km tile at a time).
#include "stc_common.cuh"

__global__ void __launch_bounds__(256) region_gather_kernel(const float* __restrict__ canvas, int T, int Hc, int Wc, int Cc, int wrap,
                                                            const int* __restrict__ ys, const int* __restrict__ xs, int P,
                                                            float* __restrict__ out) {
  // grid (P, B*T): a block copies one patch row of one month, P*Cc contiguous floats on the patch side and -- unless the
  // periodic canvas wraps inside the row -- on the canvas side too
  const int b = blockIdx.y / T, t = blockIdx.y - b * T;
  const int py = blockIdx.x;
  int y = ys[b] + py;
  int x0 = xs[b];
  if (wrap) y %= Hc;
  else {                                    // origins are device data: keep a bad window inside the canvas instead of faulting
    y = y < 0 ? 0 : (y >= Hc ? Hc - 1 : y);
    x0 = x0 < 0 ? 0 : (x0 + P > Wc ? Wc - P : x0);
  }
  const float* src_row = canvas + ((int64_t)t * Hc + y) * Wc * Cc;
  float* dst = out + (((int64_t)b * T + t) * P + py) * P * Cc;
  const int n = P * Cc;
  const int xw = wrap ? x0 % Wc : x0;
  if (xw + P <= Wc) {                       // the row is contiguous on the canvas side
    const float* src = src_row + (int64_t)xw * Cc;
    for (int i = threadIdx.x; i < n; i += 256) dst[i] = src[i];
  } else {                                  // periodic canvas wrapping inside the row
    for (int i = threadIdx.x; i < n; i += 256) { const int px = i / Cc; dst[i] = src_row[(int64_t)((xw + px) % Wc) * Cc + (i - px * Cc)]; }
  }
}

struct BlendParams {
  const float* preds;      // [rows_have][C][S][S], first grid row = r_first
  int r_first, rows_have, R, C, S, stride, margin;
  const float* gauss;      // [S][S] float32
  int y0, y1, Wc;          // canvas rows [y0, y1) to produce, canvas width
};

__global__ void __launch_bounds__(256) region_blend_kernel(BlendParams p, unsigned char* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)(p.y1 - p.y0) * p.Wc;
  if (idx >= total) return;
  const int y = p.y0 + (int)(idx / p.Wc), x = (int)(idx % p.Wc);
  // covering patches: r*stride + margin <= y < r*stride + margin + S
  auto lo = [&](int v) { int a = v - p.margin - p.S; return (a < 0) ? -(((-a) + p.stride - 1) / p.stride) + 1 : a / p.stride + 1; };   // floor(a/stride) + 1
  auto hi = [&](int v) { int a = v - p.margin; return (a < 0) ? -1 : a / p.stride; };
  int r0 = lo(y), r1 = hi(y), c0 = lo(x), c1 = hi(x);
  if (r0 < 0) r0 = 0; if (c0 < 0) c0 = 0;
  if (r1 > p.R - 1) r1 = p.R - 1; if (c1 > p.C - 1) c1 = p.C - 1;
  float num = 0.f, den = 0.f;
  for (int r = r0; r <= r1; ++r) {
    const int rr = r - p.r_first;
    if (rr < 0 || rr >= p.rows_have) { out[idx] = 254; return; }      // caller did not provide a needed row (never in a correct call)
    const int py = y - r * p.stride - p.margin;
    for (int c = c0; c <= c1; ++c) {
      const int px = x - c * p.stride - p.margin;
      const float w = p.gauss[py * p.S + px];
      const float v = __fmul_rn(p.preds[(((int64_t)rr * p.C + c) * p.S + py) * p.S + px], 100.f);
      num = __fadd_rn(num, __fmul_rn(w, v));
      den = __fadd_rn(den, w);
    }
  }
  unsigned char o;
  if (!(den > 0.f)) o = 255;
  else {
    o = (unsigned char)__fdiv_rn(num, den);       // astype(np.uint8): truncation
    if (o <= 15) o = 0;
    if (o > 100) o = 255;
  }
  out[idx] = o;
}

extern "C" {

int stc_region_gather_dev(stc_ctx* ctx, const float* canvas_dev, int T, int Hc, int Wc, int Cc, int wrap,
                          const int32_t* ys_dev, const int32_t* xs_dev, int B, int P, float* out_dev) {
  if (!ctx) return STC_ERR_ARG;
  if (!canvas_dev || !ys_dev || !xs_dev || !out_dev || T < 1 || Hc < 1 || Wc < 1 || Cc < 1 || B < 1 || P < 1 || (!wrap && (P > Hc || P > Wc)))
    STC_FAIL(STC_ERR_ARG, "region_gather: bad argument");
  if ((int64_t)B * T > 65535) STC_FAIL(STC_ERR_ARG, "region_gather: at most 65535 patch-months per call");
  dim3 grid(P, B * T);
  region_gather_kernel<<<grid, 256, 0, ctx->stream>>>(canvas_dev, T, Hc, Wc, Cc, wrap, ys_dev, xs_dev, P, out_dev);
  STC_CUDA(cudaGetLastError()); ctx->launches++;
  return STC_OK;      // asynchronous: the window origins live on the device, nothing to wait for
}

// Forward over n patch windows of a canvas: gather -> fused front end + model, `batch` windows at a time.  The gather
// of batch k + 1 runs on its own stream into the second patch buffer while batch k is in the model (the gather alone
// costs ~10 % of a batch when it is serialised with the forward).
int stc_region_predict_dev(stc_ctx* ctx, const float* canvas_dev, int T, int Hc, int Wc, int Cc, int wrap,
                           const int32_t* ys_dev, const int32_t* xs_dev, int n, int batch, int P,
                           const double* min17, const double* max17, float* preds_dev) {
  if (!ctx) return STC_ERR_ARG;
  if (!canvas_dev || !ys_dev || !xs_dev || !preds_dev || !min17 || !max17 || n < 1 || batch < 1 || T != 12 || Cc != 13 || P < 28 || P % 4 ||
      (!wrap && (P > Hc || P > Wc)))
    STC_FAIL(STC_ERR_ARG, "region_predict: bad argument (the model front end takes 12 months x 13 bands, P a multiple of 4 >= 28)");
  if ((int64_t)batch * T > 65535) STC_FAIL(STC_ERR_ARG, "region_predict: at most 65535 patch-months per batch");
  if (batch > n) batch = n;
  const size_t patch_elems = (size_t)T * P * P * Cc;
  const int S = P - 14;
  float* buf[2] = {nullptr, nullptr};
  cudaStream_t gs = nullptr; cudaEvent_t eg[2] = {nullptr, nullptr}, ec[2] = {nullptr, nullptr};
  int rc = STC_OK;
  auto cleanup = [&]() {
    cudaStreamSynchronize(ctx->stream);
    if (gs) { cudaStreamSynchronize(gs); cudaStreamDestroy(gs); }
    for (int i = 0; i < 2; ++i) { if (buf[i]) cudaFree(buf[i]); if (eg[i]) cudaEventDestroy(eg[i]); if (ec[i]) cudaEventDestroy(ec[i]); }
  };
#define RG_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); cleanup(); return STC_ERR_CUDA; } } while (0)
  const int nbuf = n > batch ? 2 : 1;
  for (int i = 0; i < nbuf; ++i) RG_CUDA(cudaMalloc((void**)&buf[i], (size_t)batch * patch_elems * 4));
  RG_CUDA(cudaStreamCreateWithFlags(&gs, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    RG_CUDA(cudaEventCreateWithFlags(&eg[i], cudaEventDisableTiming));
    RG_CUDA(cudaEventCreateWithFlags(&ec[i], cudaEventDisableTiming));
  }
  // everything already enqueued on the caller's stream (e.g. the upload of the origins) precedes the first gather
  RG_CUDA(cudaEventRecord(ec[1], ctx->stream));
  RG_CUDA(cudaStreamWaitEvent(gs, ec[1], 0));
  int k = 0;
  for (int done = 0; done < n; done += batch, ++k) {
    const int b = (n - done) < batch ? (n - done) : batch;
    const int s = (nbuf == 2) ? (k & 1) : 0;
    if (k >= nbuf) RG_CUDA(cudaStreamWaitEvent(gs, ec[s], 0));            // the forward that last read this buffer is done
    dim3 grid(P, b * T);
    region_gather_kernel<<<grid, 256, 0, gs>>>(canvas_dev, T, Hc, Wc, Cc, wrap, ys_dev + done, xs_dev + done, P, buf[s]);
    RG_CUDA(cudaGetLastError()); ctx->launches++;
    RG_CUDA(cudaEventRecord(eg[s], gs));
    RG_CUDA(cudaStreamWaitEvent(ctx->stream, eg[s], 0));
    rc = model_predict_patches_dev(ctx, buf[s], b, P, P, min17, max17, preds_dev + (size_t)done * S * S);
    if (rc) { cleanup(); return rc; }
    RG_CUDA(cudaEventRecord(ec[s], ctx->stream));
  }
#undef RG_CUDA
  cleanup();
  return STC_OK;
}

int stc_region_blend_dev(stc_ctx* ctx, const float* preds_dev, int r_first, int rows_have, int R, int C, int S, int stride, int margin,
                         const float* gauss_host, int y0, int y1, int Wc, uint8_t* out_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!preds_dev || !gauss_host || !out_host || rows_have < 1 || R < 1 || C < 1 || S < 1 || stride < 1 || margin < 0 || y1 <= y0 || Wc < 1)
    STC_FAIL(STC_ERR_ARG, "region_blend: bad argument");
  float* dg = nullptr; unsigned char* dout = nullptr;
  const size_t n = (size_t)(y1 - y0) * Wc;
  STC_CUDA(cudaMalloc((void**)&dg, (size_t)S * S * 4));
  if (cudaMalloc((void**)&dout, n) != cudaSuccess) { cudaFree(dg); STC_FAIL(STC_ERR_NOMEM, "region_blend: out of device memory"); }
  cudaMemcpyAsync(dg, gauss_host, (size_t)S * S * 4, cudaMemcpyHostToDevice, ctx->stream);
  BlendParams p{preds_dev, r_first, rows_have, R, C, S, stride, margin, dg, y0, y1, Wc};
  region_blend_kernel<<<cdiv((int64_t)n, 256), 256, 0, ctx->stream>>>(p, dout);
  ctx->launches++;
  cudaMemcpyAsync(out_host, dout, n, cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFree(dg); cudaFree(dout);
  if (e != cudaSuccess) { ctx->err = std::string("region_blend: ") + cudaGetErrorString(e); return STC_ERR_CUDA; }
  return STC_OK;
}

}  // extern "C"
