// extern "C" entry points of the spatial-neighbourhood kernels (declared in include/stc.h).
#include "stc_common.cuh"
int pre_edt_sq_dev(stc_ctx* ctx, const unsigned char* target_dev, int n, int H, int W, int radius, int* out_dev);

namespace {
struct DevBuf2 {
  void* p = nullptr;
  ~DevBuf2() { if (p) stc_dfree(p); }
};
}

extern "C" {

int stc_feather_host(stc_ctx* ctx, const float* masks_host, int n, int H, int W, int closing_size, float* out_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!masks_host || !out_host || n < 1 || H < 1 || W < 1) STC_FAIL(STC_ERR_ARG, "feather: bad argument");
  size_t bytes = (size_t)n * H * W * 4;
  DevBuf2 din, da, db, ds, dout;
  STC_CUDA(stc_dmalloc(&din.p, bytes)); STC_CUDA(stc_dmalloc(&da.p, bytes)); STC_CUDA(stc_dmalloc(&db.p, bytes));
  STC_CUDA(stc_dmalloc(&ds.p, n * 4)); STC_CUDA(stc_dmalloc(&dout.p, bytes));
  STC_CUDA(cudaMemcpyAsync(din.p, masks_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  int rc = pre_feather_dev(ctx, (const float*)din.p, n, H, W, closing_size, (float*)da.p, (float*)db.p, (float*)ds.p, (float*)dout.p);
  if (rc) return rc;
  STC_CUDA(cudaMemcpyAsync(out_host, dout.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}

int stc_binary_dilate_host(stc_ctx* ctx, const uint8_t* in_host, int n, int H, int W, int iterations, int connectivity,
                           uint8_t* out_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!in_host || !out_host || n < 1 || H < 1 || W < 1) STC_FAIL(STC_ERR_ARG, "binary_dilate: bad argument");
  size_t bytes = (size_t)n * H * W;
  DevBuf2 din, dout;
  STC_CUDA(stc_dmalloc(&din.p, bytes)); STC_CUDA(stc_dmalloc(&dout.p, bytes));
  STC_CUDA(cudaMemcpyAsync(din.p, in_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  int rc = pre_binary_dilate_dev(ctx, (const unsigned char*)din.p, n, H, W, iterations, connectivity, (unsigned char*)dout.p);
  if (rc) return rc;
  STC_CUDA(cudaMemcpyAsync(out_host, dout.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}

int stc_edt_sq_host(stc_ctx* ctx, const uint8_t* target_host, int n, int H, int W, int radius, int32_t* out_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!target_host || !out_host || n < 1 || H < 1 || W < 1) STC_FAIL(STC_ERR_ARG, "edt_sq: bad argument");
  size_t px = (size_t)n * H * W;
  DevBuf2 din, dout;
  STC_CUDA(stc_dmalloc(&din.p, px)); STC_CUDA(stc_dmalloc(&dout.p, px * 4));
  STC_CUDA(cudaMemcpyAsync(din.p, target_host, px, cudaMemcpyHostToDevice, ctx->stream));
  int rc = pre_edt_sq_dev(ctx, (const unsigned char*)din.p, n, H, W, radius, (int*)dout.p);
  if (rc) return rc;
  STC_CUDA(cudaMemcpyAsync(out_host, dout.p, px * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}

}  // extern "C"
