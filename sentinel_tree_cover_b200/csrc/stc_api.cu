// extern "C" surface of libstc.so (declared in include/stc.h).
#include "stc_common.cuh"
#include <cstdio>
#include <cstring>

#define CTX_CHECK() do { if (!ctx) return STC_ERR_ARG; } while (0)

extern "C" {

const char* stc_version(void) { return "stc-b200 0.1 (sm_100a)"; }

int stc_create(int device, stc_ctx** out) {
  if (!out) return STC_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return STC_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return STC_ERR_CUDA;
  if (prop.major != 10) return STC_ERR_STATE;  // sm_100a binary only; no fallback path exists
  if (cudaSetDevice(device) != cudaSuccess) return STC_ERR_CUDA;
  stc_ctx* ctx = new stc_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return STC_ERR_CUDA; }
  cudaEventCreate(&ctx->t0); cudaEventCreate(&ctx->t1);
  *out = ctx;
  return STC_OK;
}

void stc_destroy(stc_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  model_destroy(ctx);
  sr_destroy(ctx);
  for (auto& e : ctx->conv_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  cudaEventDestroy(ctx->t0); cudaEventDestroy(ctx->t1);
  for (int i = 0; i < 2; ++i) {
    if (ctx->stage_in[i]) cudaFree(ctx->stage_in[i]);
    if (ctx->ev_ready[i]) cudaEventDestroy(ctx->ev_ready[i]);
    if (ctx->ev_free[i]) cudaEventDestroy(ctx->ev_free[i]);
  }
  if (ctx->stage_out) cudaFree(ctx->stage_out);
  for (unsigned char* p : {ctx->anc_forest, ctx->anc_urban_core, ctx->anc_urban_near}) if (p) cudaFree(p);
  if (ctx->pin_buf) cudaFreeHost(ctx->pin_buf);
  if (ctx->stage_ring) cudaFreeHost(ctx->stage_ring);
  if (ctx->aux_stream) { cudaStreamDestroy(ctx->aux_stream); cudaEventDestroy(ctx->aux_ev[0]); cudaEventDestroy(ctx->aux_ev[1]); }
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->d2h_stream) { cudaStreamDestroy(ctx->d2h_stream); cudaEventDestroy(ctx->d2h_fork); }
  for (cudaEvent_t e : ctx->d2h_events) cudaEventDestroy(e);
  if (ctx->smp_stream) cudaStreamDestroy(ctx->smp_stream);
  for (cudaEvent_t e : ctx->smp_events) cudaEventDestroy(e);
  for (int i = 0; i < stc_ctx::MAX_SLOTS; ++i) {
    if (ctx->hi_stream[i]) cudaStreamDestroy(ctx->hi_stream[i]);
    for (int j = 0; j < 2; ++j) if (ctx->ev_lane[i][j]) cudaEventDestroy(ctx->ev_lane[i][j]);
  }
  cudaStreamDestroy(ctx->stream);
  stc_pool_trim();
  delete ctx;
}

const char* stc_last_error(stc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int64_t stc_launch_count(stc_ctx* ctx) { return ctx ? ctx->launches : -1; }
int stc_set_conv_impl(stc_ctx* ctx, int impl) {
  CTX_CHECK();
  if (impl != 0 && impl != 1) STC_FAIL(STC_ERR_ARG, "conv impl must be 0 (tcgen05) or 1 (simt)");
  ctx->conv_impl = impl;
  return STC_OK;
}

int stc_set_weight(stc_ctx* ctx, const char* name, const float* data, int64_t n) {
  CTX_CHECK();
  if (!name || !data || n <= 0) STC_FAIL(STC_ERR_ARG, "set_weight: bad argument");
  ctx->host_w[name] = std::vector<float>(data, data + n);
  return STC_OK;
}

int stc_finalize_weights(stc_ctx* ctx, int which) {
  CTX_CHECK();
  cudaSetDevice(ctx->device);
  if (which == 0) return model_finalize_weights(ctx);
  if (which == 1) return sr_finalize_weights(ctx);
  STC_FAIL(STC_ERR_ARG, "finalize_weights: which must be 0 or 1");
}

int stc_malloc(stc_ctx* ctx, size_t bytes, void** dptr) { CTX_CHECK(); cudaSetDevice(ctx->device); STC_CUDA(cudaMalloc(dptr, bytes)); return STC_OK; }
int stc_free(stc_ctx* ctx, void* dptr) { CTX_CHECK(); STC_CUDA(cudaStreamSynchronize(ctx->stream)); STC_CUDA(cudaFree(dptr)); return STC_OK; }
int stc_malloc_host(stc_ctx* ctx, size_t bytes, void** hptr) { CTX_CHECK(); STC_CUDA(cudaMallocHost(hptr, bytes)); return STC_OK; }
// write_combined != 0: cudaHostAllocWriteCombined (upload-only buffers: the DMA engine reads them without snooping the CPU caches,
// the host must only write them sequentially and never read them back)
int stc_malloc_host_flags(stc_ctx* ctx, size_t bytes, int write_combined, void** hptr) {
  CTX_CHECK();
  STC_CUDA(cudaHostAlloc(hptr, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)));
  return STC_OK;
}
int stc_free_host(stc_ctx* ctx, void* hptr) { CTX_CHECK(); STC_CUDA(cudaFreeHost(hptr)); return STC_OK; }
int stc_h2d(stc_ctx* ctx, void* dst, const void* src, size_t bytes) {
  CTX_CHECK(); STC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream)); return STC_OK;
}
int stc_d2h(stc_ctx* ctx, void* dst, const void* src, size_t bytes) {
  CTX_CHECK(); STC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream)); return STC_OK;
}
int stc_sync(stc_ctx* ctx) { CTX_CHECK(); STC_CUDA(cudaStreamSynchronize(ctx->stream)); return STC_OK; }
int stc_timer_begin(stc_ctx* ctx) { CTX_CHECK(); STC_CUDA(cudaEventRecord(ctx->t0, ctx->stream)); return STC_OK; }
int stc_timer_end(stc_ctx* ctx, float* ms) {
  CTX_CHECK();
  STC_CUDA(cudaEventRecord(ctx->t1, ctx->stream));
  STC_CUDA(cudaEventSynchronize(ctx->t1));
  STC_CUDA(cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
  return STC_OK;
}
int stc_conv_timing(stc_ctx* ctx, int enable_reset, float* total_ms, int64_t* launches) {
  CTX_CHECK();
  float tot = 0.f;
  for (size_t i = 0; i < ctx->conv_events_used; ++i) {
    float ms = 0.f;
    STC_CUDA(cudaEventSynchronize(ctx->conv_events[i].second));
    STC_CUDA(cudaEventElapsedTime(&ms, ctx->conv_events[i].first, ctx->conv_events[i].second));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = (int64_t)ctx->conv_events_used;
  if (enable_reset >= 0) { ctx->time_convs = enable_reset != 0; ctx->conv_events_used = 0; }
  return STC_OK;
}

int stc_conv_timing_kind(stc_ctx* ctx, int N, int groups, int mode, float* total_ms, int64_t* launches) {
  CTX_CHECK();
  const int kind = (N * 100 + groups) * 10 + mode;
  float tot = 0.f; int64_t cnt = 0;
  for (size_t i = 0; i < ctx->conv_events_used; ++i) {
    if (ctx->conv_event_kind[i] != kind) continue;
    float ms = 0.f;
    STC_CUDA(cudaEventSynchronize(ctx->conv_events[i].second));
    STC_CUDA(cudaEventElapsedTime(&ms, ctx->conv_events[i].first, ctx->conv_events[i].second));
    tot += ms; ++cnt;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = cnt;
  return STC_OK;
}

int stc_trace(stc_ctx* ctx, int enable, const char* csv_path) {
  CTX_CHECK();
  if (enable) { ctx->trace.clear(); ctx->trace_on = true; return STC_OK; }
  ctx->trace_on = false;
  STC_CUDA(cudaDeviceSynchronize());
  if (!csv_path || ctx->trace.empty()) return STC_OK;
  FILE* f = fopen(csv_path, "w");
  if (!f) STC_FAIL(STC_ERR_ARG, "trace: cannot open csv path");
  fprintf(f, "label,slot,start_ms,end_ms\n");
  for (auto& r : ctx->trace) {
    float t0 = 0.f, t1 = 0.f;
    cudaEventElapsedTime(&t0, ctx->trace[0].a, r.a);
    cudaEventElapsedTime(&t1, ctx->trace[0].a, r.b);
    fprintf(f, "%s,%d,%.4f,%.4f\n", r.label, r.slot, t0, t1);
  }
  for (auto& r : ctx->trace) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  fclose(f);
  ctx->trace.clear();
  return STC_OK;
}

// ---- helpers for host-buffer variants ------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) stc_dfree(p); }
};

int stc_predict_dev(stc_ctx* ctx, const float* x_dev, int B, int T, int H, int W, int length,
                    int normalize, const double* min17, const double* max17, float* out_dev) {
  CTX_CHECK();
  return model_predict_dev(ctx, x_dev, B, T, H, W, length, normalize, min17, max17, out_dev);
}

int stc_predict_host(stc_ctx* ctx, const float* x_host, int B, int T, int H, int W, int length,
                     int normalize, const double* min17, const double* max17, float* out_host) {
  CTX_CHECK();
  if (!x_host || !out_host || B < 1) STC_FAIL(STC_ERR_ARG, "predict: bad argument");
  size_t nin = (size_t)B * (T + 1) * H * W * 17, nout = (size_t)B * (H - 14) * (W - 14);
  DevBuf din, dout;
  STC_CUDA(stc_dmalloc(&din.p, nin * 4)); STC_CUDA(stc_dmalloc(&dout.p, nout * 4));
  STC_CUDA(cudaMemcpyAsync(din.p, x_host, nin * 4, cudaMemcpyHostToDevice, ctx->stream));
  int rc = model_predict_dev(ctx, (const float*)din.p, B, T, H, W, length, normalize, min17, max17, (float*)dout.p);
  if (rc) return rc;
  STC_CUDA(cudaMemcpyAsync(out_host, dout.p, nout * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}

int stc_predict_feats_host(stc_ctx* ctx, const float* x_host, int B, int T, int H, int W, int length,
                           int normalize, const double* min17, const double* max17,
                           float* probs_host, float* early_host, float* late_host) {
  CTX_CHECK();
  if (!x_host || !early_host || !late_host || B < 1) STC_FAIL(STC_ERR_ARG, "predict_feats: bad argument");
  size_t nin = (size_t)B * (T + 1) * H * W * 17, nout = (size_t)B * (H - 14) * (W - 14);
  DevBuf din, dout, de, dl;
  STC_CUDA(stc_dmalloc(&din.p, nin * 4)); STC_CUDA(stc_dmalloc(&dout.p, nout * 4));
  STC_CUDA(stc_dmalloc(&de.p, nout * 64 * 4)); STC_CUDA(stc_dmalloc(&dl.p, nout * 64 * 4));
  STC_CUDA(cudaMemcpyAsync(din.p, x_host, nin * 4, cudaMemcpyHostToDevice, ctx->stream));
  ctx->feat_early_dev = (float*)de.p; ctx->feat_late_dev = (float*)dl.p;
  int rc = model_predict_dev(ctx, (const float*)din.p, B, T, H, W, length, normalize, min17, max17, (float*)dout.p);
  ctx->feat_early_dev = ctx->feat_late_dev = nullptr;
  if (rc) return rc;
  if (probs_host) STC_CUDA(cudaMemcpyAsync(probs_host, dout.p, nout * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(early_host, de.p, nout * 64 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(late_host, dl.p, nout * 64 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}

int stc_assemble_dev(stc_ctx* ctx, const float* monthly_dev, int B, int H, int W, float* out_dev) {
  CTX_CHECK();
  return pre_assemble_dev(ctx, monthly_dev, B, H, W, out_dev);
}

int stc_assemble_host(stc_ctx* ctx, const float* monthly_host, int B, int H, int W, float* out_host) {
  CTX_CHECK();
  size_t nin = (size_t)B * 12 * H * W * 13, nout = (size_t)B * 5 * H * W * 17;
  DevBuf din, dout;
  STC_CUDA(stc_dmalloc(&din.p, nin * 4)); STC_CUDA(stc_dmalloc(&dout.p, nout * 4));
  STC_CUDA(cudaMemcpyAsync(din.p, monthly_host, nin * 4, cudaMemcpyHostToDevice, ctx->stream));
  int rc = pre_assemble_dev(ctx, (const float*)din.p, B, H, W, (float*)dout.p);
  if (rc) return rc;
  STC_CUDA(cudaMemcpyAsync(out_host, dout.p, nout * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}

// monthly -> (fused assemble + normalize + pack) -> predict.  Host inputs are staged through
// two device buffers in sub-batches.
static int predict_patches_core(stc_ctx* ctx, const float* monthly, bool host_in, int B, int H, int W,
                                const double* min17, const double* max17, float* out, bool host_out) {
  if (B < 1 || !monthly || !out || !min17 || !max17) STC_FAIL(STC_ERR_ARG, "predict_patches: bad argument");
  if (H != W || H % 4 != 0 || H < 28) STC_FAIL(STC_ERR_ARG, "predict_patches: patches must be square, H a multiple of 4 and >= 28");
  const size_t esz = ctx->monthly_u16 ? 2 : 4;      // element size of the monthly patches
  size_t per_in = (size_t)12 * H * W * 13, per_out = (size_t)(H - 14) * (W - 14);
  if (!host_in && !host_out) return model_predict_patches_dev(ctx, monthly, B, H, W, min17, max17, out);
  const char* env = getenv("STC_CHUNK");
  int chunk = env ? atoi(env) : 32;
  if (chunk < 1) chunk = 1;
  int Bc = B < chunk ? B : chunk;
  if (!ctx->copy_stream) {
    STC_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      STC_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready[i], cudaEventDisableTiming));
      STC_CUDA(cudaEventCreateWithFlags(&ctx->ev_free[i], cudaEventDisableTiming));
    }
  }
  if (host_in && ctx->stage_in_bytes < Bc * per_in * esz) {
    STC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 2; ++i) { if (ctx->stage_in[i]) cudaFree(ctx->stage_in[i]); ctx->stage_in[i] = nullptr; }
    ctx->stage_in_bytes = Bc * per_in * esz;
    for (int i = 0; i < 2; ++i) STC_CUDA(cudaMalloc(&ctx->stage_in[i], ctx->stage_in_bytes));
  }
  if (host_out && ctx->stage_out_bytes < (size_t)B * per_out * 4) {
    STC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->stage_out) cudaFree(ctx->stage_out);
    ctx->stage_out_bytes = (size_t)B * per_out * 4;
    STC_CUDA(cudaMalloc(&ctx->stage_out, ctx->stage_out_bytes));
  }
  float* o_dev = host_out ? (float*)ctx->stage_out : out;
  // Sub-batch k: H2D on the copy stream into staging buffer k&1, compute on scratch slot k % ns (own stream), so
  // copy(k+1), conv(k) and the elementwise stages of the neighbouring chunks overlap.  A staging buffer is free again
  // as soon as the front-end kernel of its chunk has read it (ev_free).
  const int nchunks = (B + Bc - 1) / Bc;
  int ns = getenv("STC_SINGLE_STREAM") ? 1 : model_num_slots();
  if (ns > nchunks) ns = nchunks;
  if (!ctx->ev_fork) STC_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  if (ns > 1) {
    STC_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
    for (int i = 1; i < ns; ++i) {
      if (!ctx->ev_join[i]) STC_CUDA(cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming));
      STC_CUDA(cudaStreamWaitEvent(model_slot_stream(ctx, i), ctx->ev_fork, 0));
    }
  }
  int k = 0;
  for (int b0 = 0; b0 < B; b0 += Bc, ++k) {
    int nb = (B - b0) < Bc ? (B - b0) : Bc;
    const float* src = reinterpret_cast<const float*>(reinterpret_cast<const char*>(monthly) + (size_t)b0 * per_in * esz);
    const int sl = k & 1;
    const int slot = k % ns;
    cudaStream_t cs = model_slot_stream(ctx, slot);
    if (host_in) {
      // copy stream: wait until the front end that last read this staging buffer is done, then copy
      if (k >= 2) STC_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[sl], 0));
      STC_CUDA(cudaMemcpyAsync(ctx->stage_in[sl], src, nb * per_in * esz, cudaMemcpyHostToDevice, ctx->copy_stream));
      STC_CUDA(cudaEventRecord(ctx->ev_ready[sl], ctx->copy_stream));
      STC_CUDA(cudaStreamWaitEvent(cs, ctx->ev_ready[sl], 0));
      src = (const float*)ctx->stage_in[sl];
    }
    int rc = model_forward_slot(ctx, slot, src, nb, Bc, H, min17, max17, o_dev + (size_t)b0 * per_out, host_in ? ctx->ev_free[sl] : nullptr);
    if (rc) return rc;
  }
  for (int i = 1; i < ns; ++i) {
    STC_CUDA(cudaEventRecord(ctx->ev_join[i], ctx->slot_stream[i]));
    STC_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[i], 0));
  }
  if (host_out) STC_CUDA(cudaMemcpyAsync(out, o_dev, (size_t)B * per_out * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}

int stc_predict_patches_host(stc_ctx* ctx, const float* monthly_host, int B, int H, int W,
                             const double* min17, const double* max17, float* out_host) {
  CTX_CHECK();
  return predict_patches_core(ctx, monthly_host, true, B, H, W, min17, max17, out_host, true);
}
int stc_predict_patches_dev(stc_ctx* ctx, const float* monthly_dev, int B, int H, int W,
                            const double* min17, const double* max17, float* out_dev) {
  CTX_CHECK();
  return predict_patches_core(ctx, monthly_dev, false, B, H, W, min17, max17, out_dev, false);
}

// uint16 patches (the reference's integer storage convention, predict_subtile :345-347: x / 65535)
int stc_predict_patches_u16_host(stc_ctx* ctx, const uint16_t* monthly_host, int B, int H, int W,
                                 const double* min17, const double* max17, float* out_host) {
  CTX_CHECK();
  ctx->monthly_u16 = 1;
  int rc = predict_patches_core(ctx, reinterpret_cast<const float*>(monthly_host), true, B, H, W, min17, max17, out_host, true);
  ctx->monthly_u16 = 0;
  return rc;
}
int stc_predict_patches_u16_dev(stc_ctx* ctx, const uint16_t* monthly_dev, int B, int H, int W,
                                const double* min17, const double* max17, float* out_dev) {
  CTX_CHECK();
  ctx->monthly_u16 = 1;
  int rc = predict_patches_core(ctx, reinterpret_cast<const float*>(monthly_dev), false, B, H, W, min17, max17, out_dev, false);
  ctx->monthly_u16 = 0;
  return rc;
}

int stc_temporal_matmul_dev(stc_ctx* ctx, const float* in_dev, const float* M_host, int n_in, int n_out, int64_t inner,
                            float* out_dev) {
  CTX_CHECK();
  return pre_temporal_matmul_dev(ctx, in_dev, M_host, n_in, n_out, inner, out_dev);
}
int stc_temporal_matmul_host(stc_ctx* ctx, const float* in_host, const float* M_host, int n_in, int n_out, int64_t inner,
                             float* out_host) {
  CTX_CHECK();
  if (!in_host || !M_host || !out_host || inner < 1) STC_FAIL(STC_ERR_ARG, "temporal_matmul: bad argument");
  DevBuf din, dout;
  STC_CUDA(stc_dmalloc(&din.p, (size_t)n_in * inner * 4)); STC_CUDA(stc_dmalloc(&dout.p, (size_t)n_out * inner * 4));
  STC_CUDA(cudaMemcpyAsync(din.p, in_host, (size_t)n_in * inner * 4, cudaMemcpyHostToDevice, ctx->stream));
  int rc = pre_temporal_matmul_dev(ctx, (const float*)din.p, M_host, n_in, n_out, inner, (float*)dout.p);
  if (rc) return rc;
  STC_CUDA(cudaMemcpyAsync(out_host, dout.p, (size_t)n_out * inner * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}

int stc_indices_host(stc_ctx* ctx, const float* in_host, int64_t npix, int C, float* out_host) {
  CTX_CHECK();
  if (!in_host || !out_host || npix < 1) STC_FAIL(STC_ERR_ARG, "indices: bad argument");
  DevBuf din, dout;
  STC_CUDA(stc_dmalloc(&din.p, (size_t)npix * C * 4)); STC_CUDA(stc_dmalloc(&dout.p, (size_t)npix * 16));
  STC_CUDA(cudaMemcpyAsync(din.p, in_host, (size_t)npix * C * 4, cudaMemcpyHostToDevice, ctx->stream));
  int rc = pre_indices_dev(ctx, (const float*)din.p, npix, C, (float*)dout.p);
  if (rc) return rc;
  STC_CUDA(cudaMemcpyAsync(out_host, dout.p, (size_t)npix * 16, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}

int stc_temporal_median_host(stc_ctx* ctx, const float* in_host, int n, int64_t inner, float* out_host) {
  CTX_CHECK();
  if (!in_host || !out_host || inner < 1) STC_FAIL(STC_ERR_ARG, "temporal_median: bad argument");
  DevBuf din, dout;
  STC_CUDA(stc_dmalloc(&din.p, (size_t)n * inner * 4)); STC_CUDA(stc_dmalloc(&dout.p, (size_t)inner * 4));
  STC_CUDA(cudaMemcpyAsync(din.p, in_host, (size_t)n * inner * 4, cudaMemcpyHostToDevice, ctx->stream));
  int rc = pre_temporal_median_dev(ctx, (const float*)din.p, n, inner, (float*)dout.p);
  if (rc) return rc;
  STC_CUDA(cudaMemcpyAsync(out_host, dout.p, (size_t)inner * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}

int stc_superresolve_dev(stc_ctx* ctx, const float* x_dev, const float* bilinear_dev, int N, int H, int W, float* out_dev) {
  CTX_CHECK();
  if (!x_dev || !out_dev) STC_FAIL(STC_ERR_ARG, "superresolve: bad argument");
  return sr_forward_dev(ctx, x_dev, bilinear_dev, N, H, W, out_dev);     // bilinear_dev == NULL: bands 4..9 of x
}

int stc_superresolve_host(stc_ctx* ctx, const float* x_host, const float* bilinear_host, int N, int H, int W, float* out_host) {
  CTX_CHECK();
  if (!x_host || !out_host) STC_FAIL(STC_ERR_ARG, "superresolve: bad argument");
  size_t npx = (size_t)N * H * W;
  DevBuf dx, db, dout;
  STC_CUDA(stc_dmalloc(&dx.p, npx * 40)); STC_CUDA(stc_dmalloc(&dout.p, npx * 24));
  STC_CUDA(cudaMemcpyAsync(dx.p, x_host, npx * 40, cudaMemcpyHostToDevice, ctx->stream));
  if (bilinear_host) {          // NULL: the bilinear input is x[..., 4:] (superresolve_large_tile), no second upload
    STC_CUDA(stc_dmalloc(&db.p, npx * 24));
    STC_CUDA(cudaMemcpyAsync(db.p, bilinear_host, npx * 24, cudaMemcpyHostToDevice, ctx->stream));
  }
  int rc = sr_forward_dev(ctx, (const float*)dx.p, (const float*)db.p, N, H, W, (float*)dout.p);
  if (rc) return rc;
  STC_CUDA(cudaMemcpyAsync(out_host, dout.p, npx * 24, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}


static int mosaic_upload(stc_ctx* ctx, const float* preds, const int32_t* xs, const int32_t* ys, const int32_t* placed,
                         int n, int S, DevBuf& dp, DevBuf& dx, DevBuf& dy, DevBuf& dpl) {
  size_t np_ = (size_t)n * S * S * 4;
  STC_CUDA(stc_dmalloc(&dp.p, np_)); STC_CUDA(stc_dmalloc(&dx.p, n * 4)); STC_CUDA(stc_dmalloc(&dy.p, n * 4)); STC_CUDA(stc_dmalloc(&dpl.p, n * 4));
  STC_CUDA(cudaMemcpyAsync(dp.p, preds, np_, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(dx.p, xs, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(dy.p, ys, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(dpl.p, placed, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  return STC_OK;
}

int stc_mosaic_diffs_host(stc_ctx* ctx, const float* preds_host, const int32_t* xs, const int32_t* ys, const int32_t* placed,
                          int n, int S, float* diffs_host) {
  CTX_CHECK();
  if (!preds_host || !xs || !ys || !placed || !diffs_host || n < 1 || n > 64 || S < 1) STC_FAIL(STC_ERR_ARG, "mosaic_diffs: bad argument");
  DevBuf dp, dx, dy, dpl, dr;
  int rc = mosaic_upload(ctx, preds_host, xs, ys, placed, n, S, dp, dx, dy, dpl); if (rc) return rc;
  size_t bytes = (size_t)n * S * S * 4;
  STC_CUDA(stc_dmalloc(&dr.p, bytes));
  rc = pre_gauss_mosaic_dev(ctx, (const float*)dp.p, (const int*)dx.p, (const int*)dy.p, (const int*)dpl.p, nullptr, nullptr,
                            (float*)dr.p, 0, n, S, 0, 0, nullptr, nullptr);
  if (rc) return rc;
  STC_CUDA(cudaMemcpyAsync(diffs_host, dr.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}

int stc_gauss_mosaic_host(stc_ctx* ctx, const float* preds_host, const int32_t* xs, const int32_t* ys, const int32_t* placed,
                          const float* gauss_host, const float* mult_host, int n, int S, int out_h, int out_w, uint8_t* out_host) {
  CTX_CHECK();
  if (!preds_host || !xs || !ys || !placed || !gauss_host || !mult_host || !out_host || n < 1 || S < 1 || out_h < 1 || out_w < 1)
    STC_FAIL(STC_ERR_ARG, "gauss_mosaic: bad argument");
  DevBuf dp, dx, dy, dpl, dg, dm, dt, dout;
  int rc = mosaic_upload(ctx, preds_host, xs, ys, placed, n, S, dp, dx, dy, dpl); if (rc) return rc;
  STC_CUDA(stc_dmalloc(&dg.p, (size_t)S * S * 4)); STC_CUDA(stc_dmalloc(&dm.p, n * 4));
  STC_CUDA(stc_dmalloc(&dt.p, (size_t)out_h * out_w)); STC_CUDA(stc_dmalloc(&dout.p, (size_t)out_h * out_w));
  STC_CUDA(cudaMemcpyAsync(dg.p, gauss_host, (size_t)S * S * 4, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(dm.p, mult_host, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  rc = pre_gauss_mosaic_dev(ctx, (const float*)dp.p, (const int*)dx.p, (const int*)dy.p, (const int*)dpl.p, (const float*)dg.p,
                            (float*)dm.p, nullptr, 1, n, S, out_h, out_w, (unsigned char*)dt.p, (unsigned char*)dout.p);
  if (rc) return rc;
  STC_CUDA(cudaMemcpyAsync(out_host, dout.p, (size_t)out_h * out_w, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}

int stc_feature_mosaic_host(stc_ctx* ctx, const int16_t* feats_host, const int32_t* xs, const int32_t* ys, const float* gauss_host,
                            int n, int S, int D, int out_h, int out_w, int16_t* out_host) {
  CTX_CHECK();
  if (!feats_host || !xs || !ys || !gauss_host || !out_host || n < 1 || S < 1 || D < 1 || out_h < 1 || out_w < 1)
    STC_FAIL(STC_ERR_ARG, "feature_mosaic: bad argument");
  DevBuf df, dx, dy, dg, dout;
  const size_t nf = (size_t)n * S * S * D, no = (size_t)D * out_h * out_w;
  STC_CUDA(stc_dmalloc(&df.p, nf * 2)); STC_CUDA(stc_dmalloc(&dx.p, n * 4)); STC_CUDA(stc_dmalloc(&dy.p, n * 4));
  STC_CUDA(stc_dmalloc(&dg.p, (size_t)S * S * 4)); STC_CUDA(stc_dmalloc(&dout.p, no * 2));
  STC_CUDA(cudaMemcpyAsync(df.p, feats_host, nf * 2, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(dx.p, xs, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(dy.p, ys, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(dg.p, gauss_host, (size_t)S * S * 4, cudaMemcpyHostToDevice, ctx->stream));
  int rc = pre_feature_mosaic_dev(ctx, (const short*)df.p, (const int*)dx.p, (const int*)dy.p, (const float*)dg.p, n, S, D, out_h, out_w, (short*)dout.p);
  if (rc) return rc;
  STC_CUDA(cudaMemcpyAsync(out_host, dout.p, no * 2, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}

int stc_pool_info(stc_ctx* ctx, int64_t* hits, int64_t* misses, int64_t* cached_bytes, int64_t* total_bytes) {
  CTX_CHECK();
  size_t c = 0, t = 0;
  stc_pool_stats(hits, misses, &c, &t);
  if (cached_bytes) *cached_bytes = (int64_t)c;
  if (total_bytes) *total_bytes = (int64_t)t;
  return STC_OK;
}

int64_t stc_debug_read(stc_ctx* ctx, const char* name, float* out_host) {
  if (!ctx || !name) return STC_ERR_ARG;
  return model_debug_read(ctx, name, out_host);
}

}  // extern "C"
