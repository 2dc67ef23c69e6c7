// Shared declarations for libstc (sm_100a).  See DESIGN.md for the data layout.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string>
#include <cstring>
#include <map>
#include <vector>
#include "../../include/stc.h"

// ---------------------------------------------------------------------------------
// Activation layout ("chunk-major"): an fp16 activation tensor with C channels over a
// padded image batch [B][Hp][Wp] is stored as C/8 planes; plane c holds, for every
// flattened padded pixel p = (b*Hp + yp)*Wp + xp, the 8 channels 8c..8c+7 as one
// 16-byte unit (uint4).  Planes carry `guard` units of slack on both sides so that the
// shifted-window reads of the implicit-GEMM convolution (p + dy*Wp + dx) never leave
// the allocation.  A 3x3 tap is then a pure row shift of the K-major A operand: rows
// are 16 B apart, so any shift keeps the 16-byte alignment UMMA descriptors need.
// ---------------------------------------------------------------------------------
struct Act {
  uint4* base = nullptr;   // start of plane 0 (including guard)
  int64_t plane = 0;       // uint4 units per plane (guard + Ptot + guard)
  int guard = 0;
  int chunks = 0;          // C/8
  int B = 0, Hp = 0, Wp = 0;
  int64_t Ptot() const { return (int64_t)B * Hp * Wp; }
  __host__ __device__ uint4* at(int c) const { return base + (int64_t)c * plane + guard; }
};

// Raw (pre-normalisation) convolution output, fp32: N/4 planes of float4 per pixel.
struct Raw {
  float4* base = nullptr;
  int64_t plane = 0;       // float4 units per plane (>= Ptot rounded to tiles)
  int N = 0;
};

struct ConvParams {
  // A operand: up to two channel sources (K-steps of 16 channels = 2 chunks each)
  const uint4* a0[2]; int64_t a0_plane; int k0steps;
  const uint4* a1[2]; int64_t a1_plane; int k1steps;
  const uint4* w[2];            // packed weights [Ksteps][9 taps][2 chunks][N] uint4
  float4* out[2]; int64_t out_plane;
  double* stats[2];             // [B][G][2] (sum, sumsq) accumulated over valid outputs
  const float* sse_w[2];        // MODE_CAND: 1x1 squeeze weights [N]
  const float* bias;            // MODE_BIAS*: [N]
  int N, G;
  int B, Hp, Wp; int64_t Ptot;
  int vy0, vy1, vx0, vx1;       // valid output range in padded coordinates
  int mode;
  int out_fp16;                 // 1: write fp16 planes of uint4 (8 ch) instead of fp32 float4 planes
  int exp_flags;                // timing experiments (STC_EXP_FLAGS, results invalid): see stc_conv.cu
  // Fused DSen2 epilogues (MODE_BIAS / MODE_BIAS_RELU, tcgen05 kernels only; all null = plain raw output).  The network has
  // no normalisation between its convolutions, so the layer's next fp16 activation (with its reflect border), the fp32
  // residual and the final tanh + bilinear sum are produced straight from the accumulator instead of by a second pass over
  // an fp32 copy (stc_sr.cu).
  uint4* act16; int64_t act16_plane;       // fp16 activation of the valid outputs; border pixels mirrored from the interior
  float4* skip; int64_t skip_plane;        // fp32 residual planes
  int skip_mode;                           // 1: skip = x;  2: x = skip + 0.1 x, skip = x
  float* sr_out; const float* sr_bil; int sr_bil_stride, sr_bil_off;   // N = 16: out[px][k] = tanh(x_k) + bil[px][off + k], k < 6
};
enum { MODE_PLAIN = 0, MODE_PSCALE_SWISH = 1, MODE_SWISH = 2, MODE_CAND = 3,
       MODE_BIAS = 4, MODE_BIAS_RELU = 5 };

struct stc_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launches = 0;
  int conv_impl = 0;
  int num_sms = 148;
  std::map<std::string, std::vector<float>> host_w;   // canonical name -> f32 tensor
  // conv timing
  bool time_convs = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> conv_events;
  size_t conv_events_used = 0;
  std::vector<int> conv_event_kind;   // (N*100 + groups)*10 + mode of each timed conv launch
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  // host-buffer tile path: H2D copies run on their own stream, double-buffered against compute
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  void* stage_in[2] = {nullptr, nullptr}; size_t stage_in_bytes = 0;
  void* stage_out = nullptr; size_t stage_out_bytes = 0;
  // Chunk slots: consecutive sub-batches of a call run on separate scratch arenas / streams so that the HBM-bound
  // elementwise kernels of some overlap the tensor-bound convolutions of others (slot 0 uses `stream`).
  static constexpr int MAX_SLOTS = 4;
  void* slots[MAX_SLOTS] = {nullptr, nullptr, nullptr, nullptr};              // ModelState* per slot
  cudaStream_t slot_stream[MAX_SLOTS] = {nullptr, nullptr, nullptr, nullptr}; // [0] unused (= stream)
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  int cur_slot = 0;           // slot whose kernels are being enqueued (selects the conv priority lane)
  int last_slot = 0;
  cudaStream_t hi_stream[MAX_SLOTS] = {nullptr, nullptr, nullptr, nullptr};   // high-priority conv lanes (stc_conv.cu)
  cudaEvent_t ev_lane[MAX_SLOTS][2] = {};
  // kernel timeline (stc_trace)
  struct TraceRec { cudaEvent_t a, b; const char* label; int slot; };
  std::vector<TraceRec> trace; bool trace_on = false;
  // --gen_feats taps of the current call (device, [B,H-14,W-14,64] float32) and of the chunk being enqueued
  float* feat_early_dev = nullptr; float* feat_late_dev = nullptr;
  float* feat_early_chunk = nullptr; float* feat_late_chunk = nullptr;
  int monthly_u16 = 0;        // the monthly patches of the current call are uint16 (x/65535), not float32
  void* sr = nullptr;         // SuperresState*
  // grow-only pinned host scratch (index lists of the cloud-removal sampling stage, small result blocks)
  void* pin_buf = nullptr; size_t pin_bytes = 0;
  // pinned ring for small host -> device tables (job lists, window tables): cudaMemcpyAsync from it is truly asynchronous
  char* stage_ring = nullptr; size_t stage_pos = 0;
  // ancillary rasters of the current tile (stc_set_ancillary_masks_host): ESA WorldCover forest / urban masks at tile
  // resolution, [anc_H * anc_W] uint8 on the device, nullptr = absent (the reference's fallback when the .tif is missing)
  unsigned char* anc_forest = nullptr; unsigned char* anc_urban_core = nullptr; unsigned char* anc_urban_near = nullptr;
  int anc_H = 0, anc_W = 0;
  // side stream for latency-bound kernels that occupy a few SMs next to GPU-wide work (forked / joined with events)
  cudaStream_t aux_stream = nullptr; cudaEvent_t aux_ev[2] = {nullptr, nullptr};
  // remove_clouds: the per-date index lists go to the host on their own stream, one event per date (stc_cloudfill.cu)
  cudaStream_t d2h_stream = nullptr; cudaEvent_t d2h_fork = nullptr; std::vector<cudaEvent_t> d2h_events;
  // ... and the shuffled sample of a date comes back on a third one, issued by the worker thread that finished it
  cudaStream_t smp_stream = nullptr; std::vector<cudaEvent_t> smp_events;
};
static constexpr size_t STC_STAGE_RING = 8u << 20;
// copy `bytes` of host data into the pinned ring and return the pinned address (valid until the ring wraps: 8 MB of tables)
static inline const void* ctx_stage(stc_ctx* ctx, const void* src, size_t bytes) {
  if (bytes > STC_STAGE_RING / 4) return nullptr;
  if (!ctx->stage_ring && cudaMallocHost((void**)&ctx->stage_ring, STC_STAGE_RING) != cudaSuccess) { ctx->stage_ring = nullptr; return nullptr; }
  const size_t need = (bytes + 255) & ~(size_t)255;
  if (ctx->stage_pos + need > STC_STAGE_RING) ctx->stage_pos = 0;
  char* dst = ctx->stage_ring + ctx->stage_pos;
  ctx->stage_pos += need;
  memcpy(dst, src, bytes);
  return dst;
}
// pinned host scratch of at least `bytes` (contents are not preserved across calls that grow it)
static inline void* ctx_pinned(stc_ctx* ctx, size_t bytes) {
  if (ctx->pin_bytes < bytes) {
    if (ctx->pin_buf) { cudaStreamSynchronize(ctx->stream); cudaFreeHost(ctx->pin_buf); ctx->pin_buf = nullptr; ctx->pin_bytes = 0; }
    size_t want = bytes + bytes / 4;
    if (cudaMallocHost(&ctx->pin_buf, want) != cudaSuccess) { ctx->pin_buf = nullptr; return nullptr; }
    ctx->pin_bytes = want;
  }
  return ctx->pin_buf;
}

#define STC_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t _e = (call);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + std::to_string(__LINE__); \
      return STC_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define STC_FAIL(code, msg) do { ctx->err = (msg); return (code); } while (0)

// GroupNorm statistics slots (sum, sum of squares per sample and group) hold 64-bit FIXED-POINT integers (2^-24 units)
// inside the double-typed buffers: integer atomics are associative, so the totals do not depend on the order in which
// warps and CTAs arrive -- identical inputs give identical bytes, run to run.  Together with sample-aligned conv tiles
// (every partial sum covers the same pixels of a sample wherever the sample sits in the batch) the result is also
// independent of the batch position and of the number of CTAs.  Range: |sum| < 2^39 = 5e11.
#ifdef __CUDACC__
#define STC_STAT_SCALE 16777216.0
__device__ __forceinline__ void stat_add(double* slot, float v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(slot), (unsigned long long)__double2ll_rn((double)v * STC_STAT_SCALE));
}
__device__ __forceinline__ double stat_get(const double* slot) {
  return (double)(*reinterpret_cast<const long long*>(slot)) * (1.0 / STC_STAT_SCALE);
}
#endif

// timeline helpers: bracket one launch on ctx->stream
static inline void trace_begin(stc_ctx* ctx, const char* label) {
  if (!ctx->trace_on) return;
  stc_ctx::TraceRec r; r.label = label; r.slot = ctx->cur_slot;
  cudaEventCreate(&r.a); cudaEventCreate(&r.b);
  cudaEventRecord(r.a, ctx->stream);
  ctx->trace.push_back(r);
}
static inline void trace_end(stc_ctx* ctx) {
  if (!ctx->trace_on || ctx->trace.empty()) return;
  cudaEventRecord(ctx->trace.back().b, ctx->stream);
}

// RAII bracket around one launch statement: `{ TraceScope ts_(ctx, "kernel"); kernel<<<...>>>(...); }`
struct TraceScope {
  stc_ctx* c;
  TraceScope(stc_ctx* ctx, const char* label) : c(ctx) { trace_begin(ctx, label); }
  ~TraceScope() { trace_end(c); }
};

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- device-memory pool (stc_pool.cu): per-call scratch buffers are recycled instead of cudaMalloc'd / cudaFree'd ----
cudaError_t stc_dmalloc(void** p, size_t bytes);
cudaError_t stc_dfree(void* p);
template <typename T> static inline cudaError_t stc_dmalloc(T** p, size_t bytes) { return stc_dmalloc((void**)p, bytes); }
void stc_pool_trim();
void stc_pool_stats(int64_t* hits, int64_t* misses, size_t* cached_bytes, size_t* total_bytes);
// RAII scratch buffer from the pool
struct PoolBuf {
  void* p = nullptr;
  PoolBuf() = default;
  PoolBuf(const PoolBuf&) = delete; PoolBuf& operator=(const PoolBuf&) = delete;
  ~PoolBuf() { if (p) stc_dfree(p); }
  cudaError_t alloc(size_t bytes) { if (p) { stc_dfree(p); p = nullptr; } return stc_dmalloc(&p, bytes); }
  template <typename T> T* as() const { return (T*)p; }
};

// model entry points implemented in stc_model.cu
int model_finalize_weights(stc_ctx* ctx);
void model_destroy(stc_ctx* ctx);
int model_predict_dev(stc_ctx* ctx, const float* x_dev, int B, int T, int H, int W, int length,
                      int normalize, const double* min17, const double* max17, float* out_dev);
int model_predict_patches_dev(stc_ctx* ctx, const float* monthly_dev, int B, int H, int W,
                              const double* min17, const double* max17, float* out_dev);
int64_t model_debug_read(stc_ctx* ctx, const char* name, float* out_host);
int sr_finalize_weights(stc_ctx* ctx);
void sr_destroy(stc_ctx* ctx);
int sr_forward_dev(stc_ctx* ctx, const float* x_dev, const float* bil_dev, int N, int H, int W, float* out_dev);
// conv launchers (stc_conv.cu)
int launch_conv(stc_ctx* ctx, const ConvParams& p, int ndir);
// preprocessing (stc_preproc.cu)
int pre_assemble_dev(stc_ctx* ctx, const float* monthly_dev, int B, int H, int W, float* out_dev);
int pre_temporal_matmul_dev(stc_ctx* ctx, const float* in_dev, const float* M_host, int n_in, int n_out, int64_t inner, float* out_dev);
int pre_smooth_fused_dev(stc_ctx* ctx, const float* s2_dev, const float* M_host, int n, int64_t HW, float* monthly_dev, float* quarterly_dev);
int pre_indices_dev(stc_ctx* ctx, const float* in_dev, int64_t npix, int C, float* out_dev);
int pre_temporal_median_dev(stc_ctx* ctx, const float* in_dev, int n, int64_t inner, float* out_dev);
int pre_gauss_mosaic_dev(stc_ctx* ctx, const float* preds_dev, const int* xs_dev, const int* ys_dev, const int* placed_dev,
                         const float* gauss_dev, float* mult_dev, float* diffs_dev, int stage,
                         int n, int S, int Hc, int Wc, unsigned char* tmp_dev, unsigned char* out_dev);
int pre_feature_mosaic_dev(stc_ctx* ctx, const short* feats_dev, const int* xs_dev, const int* ys_dev, const float* gauss_dev,
                           int n, int S, int D, int Hc, int Wc, short* out_dev);
int pre_feather_dev(stc_ctx* ctx, const float* mask_dev, int n, int H, int W, int size, float* tmp_a, float* tmp_b,
                    float* sums_dev, float* out_dev);
// separable morphology (stc_morph.cu): row distance by warp ballots + one column pass
int morph_rowdist_dev(stc_ctx* ctx, const void* in, int src_kind, int64_t rows, int W, int cap, unsigned char* g);
int morph_dilate_dev(stc_ctx* ctx, const unsigned char* in, unsigned char* out, int frames, int H, int W, int k, int conn, int inv_in,
                     int inv_out, int three_d);
int morph_edt_grow_dev(stc_ctx* ctx, const unsigned char* in, unsigned char* out, int T, int H, int W, int radius, const int* frame_count_dev);
int pre_binary_dilate_dev(stc_ctx* ctx, const unsigned char* in_dev, int n, int H, int W, int iterations, int conn,
                          unsigned char* out_dev);
int model_forward_slot(stc_ctx* ctx, int slot, const float* monthly_dev, int nb, int Bc, int H,
                       const double* min17, const double* max17, float* out_dev, cudaEvent_t input_consumed);
int model_num_slots();
cudaStream_t model_slot_stream(stc_ctx* ctx, int slot);
