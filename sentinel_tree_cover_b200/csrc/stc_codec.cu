// Storage codecs and the Sentinel-1 dB transform (HBM-bound elementwise kernels).
//   to_float32     : uint16 -> float32, x / 65535          (src/tof/tof_downloading.py:64-72)
//   to_int16       : float32 -> uint16, trunc(clip(x,0,1)*65535)   (:51-61)
//   convert_to_db  : 10*log10(x + 1/65535), floor at -min_db, rescale to [0,1]
//                    (src/download_and_predict_job.py:74-89)
#include "stc_common.cuh"

__global__ void __launch_bounds__(256) to_float32_kernel(const uint16_t* __restrict__ in, float* __restrict__ out, int64_t n) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    ushort4 u = *reinterpret_cast<const ushort4*>(in + i);
    *reinterpret_cast<float4*>(out + i) = make_float4(__fdiv_rn((float)u.x, 65535.f), __fdiv_rn((float)u.y, 65535.f),
                                                       __fdiv_rn((float)u.z, 65535.f), __fdiv_rn((float)u.w, 65535.f));
  } else {
    for (; i < n; ++i) out[i] = __fdiv_rn((float)in[i], 65535.f);
  }
}

__global__ void __launch_bounds__(256) to_uint16_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = fminf(fmaxf(in[i], 0.f), 1.f);
  out[i] = (uint16_t)truncf(__fmul_rn(v, 65535.f));
}

__global__ void __launch_bounds__(256) convert_to_db_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, float min_db) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = __fmul_rn(10.f, log10f(__fadd_rn(in[i], (float)(1.0 / 65535.0))));
  if (x < -min_db) x = -min_db;
  x = __fdiv_rn(__fadd_rn(x, min_db), min_db);
  out[i] = fminf(fmaxf(x, 0.f), 1.f);
}

// float_to_int16 (src/download_and_predict_job.py:174-180): NaN -> -32768, clip to [-32768/p, 32767/p] (float32 bounds),
// * p, np.int16() truncation.  Used for the --gen_feats feature stacks (p = 1000).
__global__ void __launch_bounds__(256) float_to_int16_kernel(const float* __restrict__ in, int16_t* __restrict__ out, int64_t n,
                                                              float lo, float hi, float precision) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = in[i];
  if (isnan(v)) v = -32768.f;
  v = fminf(fmaxf(v, lo), hi);
  out[i] = (int16_t)__fmul_rn(v, precision);
}

// ---- device-level launchers (shared with stc_tile.cu) ----
int codec_to_float32_dev(stc_ctx* ctx, const uint16_t* in_dev, int64_t n, float* out_dev) {
  { TraceScope ts_(ctx, "to_float32_kernel"); to_float32_kernel<<<cdiv((n + 3) / 4, 256), 256, 0, ctx->stream>>>(in_dev, out_dev, n); }
  STC_CUDA(cudaGetLastError()); ctx->launches++;
  return STC_OK;
}
int codec_convert_to_db_dev(stc_ctx* ctx, const float* in_dev, int64_t n, float min_db, float* out_dev) {
  { TraceScope ts_(ctx, "convert_to_db_kernel"); convert_to_db_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(in_dev, out_dev, n, min_db); }
  STC_CUDA(cudaGetLastError()); ctx->launches++;
  return STC_OK;
}

namespace {
struct Buf { void* p = nullptr; ~Buf() { if (p) stc_dfree(p); } };
}

extern "C" {

int stc_to_float32_host(stc_ctx* ctx, const uint16_t* in_host, int64_t n, float* out_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!in_host || !out_host || n < 1) STC_FAIL(STC_ERR_ARG, "to_float32: bad argument");
  Buf a, b;
  STC_CUDA(stc_dmalloc(&a.p, n * 2)); STC_CUDA(stc_dmalloc(&b.p, n * 4));
  STC_CUDA(cudaMemcpyAsync(a.p, in_host, n * 2, cudaMemcpyHostToDevice, ctx->stream));
  { TraceScope ts_(ctx, "to_float32_kernel"); to_float32_kernel<<<cdiv((n + 3) / 4, 256), 256, 0, ctx->stream>>>((const uint16_t*)a.p, (float*)b.p, n); }
  STC_CUDA(cudaGetLastError()); ctx->launches++;
  STC_CUDA(cudaMemcpyAsync(out_host, b.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}

int stc_to_uint16_host(stc_ctx* ctx, const float* in_host, int64_t n, uint16_t* out_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!in_host || !out_host || n < 1) STC_FAIL(STC_ERR_ARG, "to_uint16: bad argument");
  Buf a, b;
  STC_CUDA(stc_dmalloc(&a.p, n * 4)); STC_CUDA(stc_dmalloc(&b.p, n * 2));
  STC_CUDA(cudaMemcpyAsync(a.p, in_host, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  { TraceScope ts_(ctx, "to_uint16_kernel"); to_uint16_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>((const float*)a.p, (uint16_t*)b.p, n); }
  STC_CUDA(cudaGetLastError()); ctx->launches++;
  STC_CUDA(cudaMemcpyAsync(out_host, b.p, n * 2, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}

int stc_float_to_int16_host(stc_ctx* ctx, const float* in_host, int64_t n, int precision, int16_t* out_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!in_host || !out_host || n < 1 || precision < 1) STC_FAIL(STC_ERR_ARG, "float_to_int16: bad argument");
  Buf a, b;
  STC_CUDA(stc_dmalloc(&a.p, n * 4)); STC_CUDA(stc_dmalloc(&b.p, n * 2));
  STC_CUDA(cudaMemcpyAsync(a.p, in_host, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  // np.clip(float32 array, python float, python float): the bounds are float64 scalars cast to float32
  { TraceScope ts_(ctx, "float_to_int16_kernel"); float_to_int16_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>((const float*)a.p, (int16_t*)b.p, n, (float)(-32768.0 / precision),
                                                              (float)(32767.0 / precision), (float)precision); }
  STC_CUDA(cudaGetLastError()); ctx->launches++;
  STC_CUDA(cudaMemcpyAsync(out_host, b.p, n * 2, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}

int stc_convert_to_db_host(stc_ctx* ctx, const float* in_host, int64_t n, float min_db, float* out_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!in_host || !out_host || n < 1 || !(min_db > 0.f)) STC_FAIL(STC_ERR_ARG, "convert_to_db: bad argument");
  Buf a, b;
  STC_CUDA(stc_dmalloc(&a.p, n * 4)); STC_CUDA(stc_dmalloc(&b.p, n * 4));
  STC_CUDA(cudaMemcpyAsync(a.p, in_host, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  { TraceScope ts_(ctx, "convert_to_db_kernel"); convert_to_db_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>((const float*)a.p, (float*)b.p, n, min_db); }
  STC_CUDA(cudaGetLastError()); ctx->launches++;
  STC_CUDA(cudaMemcpyAsync(out_host, b.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}

}  // extern "C"
