// HBM-bound temporal preprocessing kernels (one thread owns one pixel(-channel) column of
// n <= 32 time steps in registers; loads are coalesced across pixels).
//   assemble_kernel        : process_subtiles medians + 17-channel frame layout + indices
//                            (src/download_and_predict_job.py:1152-1160,1274-1283,1398-1407;
//                             src/preprocessing/indices.py:4-54)
//   temporal_matmul_kernel : calculate_and_save_best_images + Smoother.interpolate_array as
//                            one 12 x n operator (src/downloading/utils.py:176-347,
//                            src/preprocessing/whittaker_smoother.py:38-69)
//   indices_kernel         : make_indices (src/download_and_predict_job.py:998-1006)
//   temporal_median_kernel : np.median(axis=0) (:1152-1160)
#include "stc_common.cuh"
#include <algorithm>
#include <cstring>

#include "stc_indices.cuh"
#include "stc_sortnet.cuh"

// monthly [B,12,H,W,13] -> out [B,5,H,W,17]
__global__ void __launch_bounds__(128) assemble_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                       int B, int HW) {
  int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (int64_t)B * HW) return;
  int b = (int)(pix / HW); int r = (int)(pix - (int64_t)b * HW);
  const float* src = in + ((int64_t)b * 12 * HW + r) * 13;
  float* dst = out + ((int64_t)b * 5 * HW + r) * 17;
  const int64_t fs_in = (int64_t)HW * 13, fs_out = (int64_t)HW * 17;
  float bands[5][12];   // B2, B3, B4, B8, B11 (channels 0,1,2,3,8)
#pragma unroll
  for (int c = 0; c < 13; ++c) {
    float v[12];
#pragma unroll
    for (int t = 0; t < 12; ++t) v[t] = src[t * fs_in + c];
    if (c < 4) {
#pragma unroll
      for (int t = 0; t < 12; ++t) bands[c][t] = v[t];
    } else if (c == 8) {
#pragma unroll
      for (int t = 0; t < 12; ++t) bands[4][t] = v[t];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q * fs_out + c] = med3(v[3 * q], v[3 * q + 1], v[3 * q + 2]);
    dst[4 * fs_out + c] = median12_net(v);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float v[12];
#pragma unroll
    for (int t = 0; t < 12; ++t) {
      float b2 = bands[0][t], b3 = bands[1][t], b4 = bands[2][t], b8 = bands[3][t], b11 = bands[4][t];
      v[t] = (k == 0) ? idx_evi(b2, b3, b4, b8) : (k == 1) ? idx_bi(b2, b4, b8, b11)
           : (k == 2) ? idx_msavi2(b4, b8) : idx_grndvi(b3, b4, b8);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q * fs_out + 13 + k] = med3(v[3 * q], v[3 * q + 1], v[3 * q + 2]);
    dst[4 * fs_out + 13 + k] = median12_net(v);
  }
}

// Staged variants (default whenever H*W is a multiple of 4, which makes every 32-pixel run of the input 16-byte aligned):
// same arithmetic, different data movement.  The direct kernel above reads a pixel's 12 x 13 values with 4-byte loads 52 bytes
// apart and writes its 5 x 17 results 68 bytes apart: every 128-byte line passes through L1 thirteen (seventeen) times, and the
// kernel sits at 32 % of the HBM peak whatever the batch (profiles/r01_h_sweep.md).  Here a 32-pixel GROUP is the unit: one
// elected lane fetches the twelve contiguous 32 x 13-float runs with cp.async.bulk into the group's shared-memory tile
// (completion on the tile's mbarrier) and lane l reads pixel l's values at a 13-word stride (odd: bank-conflict free).
// Persistent grid, one CTA per SM; the warps drift apart so the copies of some overlap the sorting networks of others.
//   MODE 0: one warp per group, 7 tiles per CTA; results go through a per-warp output tile (17-word stride) and leave as five
//           contiguous 32 x 17-float runs, coalesced.  Measured 0.37 of the HBM peak (direct kernel 0.32): with 7 warps per SM
//           the kernel is bound by the latency of its own arithmetic (IEEE divisions of the exact index forms, 60-comparator
//           median networks), not by the memory system.
//   MODE 1: one warp per group, 11 tiles per CTA, results stored directly (68-byte stride; L2 merges the partial lines).
//   MODE 2: TWO warps per group, 11 tiles per CTA = 22 warps: warp 0 of a pair forms the medians of the 13 bands, warp 1
//           computes the four indices from the same tile and their medians; a named barrier per pair releases the tile.
//           MODES 1 and 2 measured 0.27 of the HBM peak, BELOW the direct kernel: the 68-byte-stride stores are what costs
//           (seventeen partial writes per 32-byte sector), not the number of warps -- 22 warps are no faster than 11.
//   MODE 3: two warps per group as in MODE 2 AND the output tile of MODE 0 (7 tiles, 14 warps); both warps store.  Measured
//           0.395 of the HBM peak at B = 256 (MODE 0 on the same box: 0.370): the default.  Twice the warps buy 7 %, so the
//           arithmetic latency is not the bound either; what is has not been profiled (profiles/r02_sweep.md).
template <int MODE> struct AsCfg {
  static constexpr bool HAS_OUT = (MODE == 0 || MODE == 3);         // results leave through an output tile, coalesced
  static constexpr int TILES = HAS_OUT ? 7 : 11;
  static constexpr int WPT = (MODE >= 2) ? 2 : 1;                    // warps per tile
  static constexpr int GE = 32 * 13, GO = 32 * 17;
  static constexpr int IN_BYTES = 12 * GE * 4;
  static constexpr int OUT_BYTES = HAS_OUT ? 5 * GO * 4 : 0;
  static constexpr int TILE_BYTES = IN_BYTES + OUT_BYTES;
  static constexpr int THREADS = TILES * WPT * 32;
  static constexpr int SMEM = TILES * TILE_BYTES + 8 * TILES;
};

template <int MODE>
__global__ void __launch_bounds__(AsCfg<MODE>::THREADS, 1) assemble_staged_kernel(const float* __restrict__ in, float* __restrict__ out, int HW,
                                                                                  int groups_per_sample, int total_groups) {
  using C = AsCfg<MODE>;
  constexpr int GE = C::GE, GO = C::GO;
  extern __shared__ __align__(128) uint8_t as_smem[];
  const int wid = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int tix = wid / C::WPT, role = wid - tix * C::WPT;
  float* tile = reinterpret_cast<float*>(as_smem + tix * C::TILE_BYTES);
  float* otile = reinterpret_cast<float*>(as_smem + tix * C::TILE_BYTES + C::IN_BYTES);      // MODE 0 only
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(as_smem + C::TILES * C::TILE_BYTES) + 8u * tix;
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
  if (role == 0 && lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t elected = 0;
  asm volatile("{\n.reg .pred px;\nelect.sync _|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}" : "=r"(elected));
  const bool leader = elected && role == 0;
  const int64_t fs_in = (int64_t)HW * 13, fs_out = (int64_t)HW * 17;
  uint32_t phase = 0;
  for (int g = blockIdx.x * C::TILES + tix; g < total_groups; g += gridDim.x * C::TILES) {
    const int b = g / groups_per_sample;
    const int r0 = (g - b * groups_per_sample) * 32;
    const int npx = (HW - r0 < 32) ? HW - r0 : 32;
    const uint32_t run_bytes = (uint32_t)(npx * 13 * 4);
    const float* src = in + ((int64_t)b * 12 * HW + r0) * 13;
    if (C::WPT == 2)      // both warps of the pair are done with the tile of the previous group
      asm volatile("bar.sync %0, 64;" ::"r"(1 + tix) : "memory");
    if (leader) {
      // the tile was last read through the generic proxy (previous iteration): order those reads before the async writes
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(12u * run_bytes) : "memory");
    }
#pragma unroll
    for (int t = 0; t < 12; ++t) {
      if (leader)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(tile_s + (uint32_t)(t * GE * 4)), "l"(src + (int64_t)t * fs_in), "r"(run_bytes), "r"(bar) : "memory");
    }
    {
      uint32_t ok = 0;
      long long t0 = clock64();
      while (!ok) {
        asm volatile("{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\nselp.u32 %0, 1, 0, q;\n}"
                     : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
        if (!ok && clock64() - t0 > 4000000000LL) __trap();          // never hang the box
      }
      phase ^= 1;
    }
    if (lane < npx) {
      const float* mine = tile + lane * 13;
      // where this lane's results go: its slot of the output tile (MODE 0) or its pixel of the five output frames
      float* mo = C::HAS_OUT ? otile + lane * 17 : out + ((int64_t)b * 5 * HW + r0 + lane) * 17;
      const int64_t fo = C::HAS_OUT ? (int64_t)GO : fs_out;           // frame stride of `mo`
      float bands[5][12];   // B2, B3, B4, B8, B11 (channels 0,1,2,3,8)
      if (C::WPT == 1 || role == 0) {
#pragma unroll
        for (int c = 0; c < 13; ++c) {
          float v[12];
#pragma unroll
          for (int t = 0; t < 12; ++t) v[t] = mine[t * GE + c];
          if (C::WPT == 1) {
            if (c < 4) {
#pragma unroll
              for (int t = 0; t < 12; ++t) bands[c][t] = v[t];
            } else if (c == 8) {
#pragma unroll
              for (int t = 0; t < 12; ++t) bands[4][t] = v[t];
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) mo[q * fo + c] = med3(v[3 * q], v[3 * q + 1], v[3 * q + 2]);
          mo[4 * fo + c] = median12_net(v);
        }
      }
      if (C::WPT == 1 || role == 1) {
        if (C::WPT == 2) {
#pragma unroll
          for (int t = 0; t < 12; ++t) {
#pragma unroll
            for (int k = 0; k < 4; ++k) bands[k][t] = mine[t * GE + k];
            bands[4][t] = mine[t * GE + 8];
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float v[12];
#pragma unroll
          for (int t = 0; t < 12; ++t) {
            float b2 = bands[0][t], b3 = bands[1][t], b4 = bands[2][t], b8 = bands[3][t], b11 = bands[4][t];
            v[t] = (k == 0) ? idx_evi(b2, b3, b4, b8) : (k == 1) ? idx_bi(b2, b4, b8, b11)
                 : (k == 2) ? idx_msavi2(b4, b8) : idx_grndvi(b3, b4, b8);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) mo[q * fo + 13 + k] = med3(v[3 * q], v[3 * q + 1], v[3 * q + 2]);
          mo[4 * fo + 13 + k] = median12_net(v);
        }
      }
    }
    __syncwarp();      // every lane is done with the input tile (and this warp's part of the output tile is written)
    if (C::HAS_OUT) {
      if (C::WPT == 2)   // the partner's part of the output tile is written too
        asm volatile("bar.sync %0, 64;" ::"r"(1 + tix) : "memory");
      float* dst = out + ((int64_t)b * 5 * HW + r0) * 17;
      const int nrun = npx * 17;
#pragma unroll
      for (int f = 0; f < 5; ++f)
        for (int i = lane + 32 * role; i < nrun; i += 32 * C::WPT) dst[f * fs_out + i] = otile[f * GO + i];
      if (C::WPT == 1) __syncwarp();    // the stores have read the output tile before the next group overwrites it
                                        // (two warps per tile: the pair barrier at the top of the next iteration)
    }
  }
}

template <int MODE>
static int launch_assemble_staged(stc_ctx* ctx, const float* monthly_dev, int B, int HW, float* out_dev) {
  using C = AsCfg<MODE>;
  static bool cfg = false;
  if (!cfg) { STC_CUDA(cudaFuncSetAttribute(assemble_staged_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM)); cfg = true; }
  const int gps = cdiv((int64_t)HW, 32);
  const int64_t total = (int64_t)gps * B;
  if (total >= (1ll << 31)) STC_FAIL(STC_ERR_ARG, "assemble: batch too large");
  const int grid = (int)std::min<int64_t>(cdiv(total, C::TILES), ctx->num_sms);
  TraceScope ts_(ctx, "assemble_staged_kernel");
  assemble_staged_kernel<MODE><<<grid, C::THREADS, C::SMEM, ctx->stream>>>(monthly_dev, out_dev, HW, gps, (int)total);
  return STC_OK;
}

int pre_assemble_dev(stc_ctx* ctx, const float* monthly_dev, int B, int H, int W, float* out_dev) {
  int64_t n = (int64_t)B * H * W;
  const int HW = H * W;
  // STC_ASSEMBLE_V: -1 = the direct-load kernel, 0 .. 3 = the staged variants above (A/B switch; default from the measurement)
  static const int variant = getenv("STC_ASSEMBLE_V") ? atoi(getenv("STC_ASSEMBLE_V")) : 3;
  const bool aligned = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(monthly_dev) & 15) == 0);
  if (aligned && variant >= 0) {
    int rc = variant == 3 ? launch_assemble_staged<3>(ctx, monthly_dev, B, HW, out_dev)
           : variant == 2 ? launch_assemble_staged<2>(ctx, monthly_dev, B, HW, out_dev)
           : variant == 1 ? launch_assemble_staged<1>(ctx, monthly_dev, B, HW, out_dev)
                          : launch_assemble_staged<0>(ctx, monthly_dev, B, HW, out_dev);
    if (rc) return rc;
  } else {
    TraceScope ts_(ctx, "assemble_kernel"); assemble_kernel<<<cdiv(n, 128), 128, 0, ctx->stream>>>(monthly_dev, out_dev, B, HW);
  }
  STC_CUDA(cudaGetLastError());
  ctx->launches++;
  return STC_OK;
}

// ---- out[o][i] = sum_n M[o][n] * in[n][i] ----------------------------------------------
// The operator travels as a kernel parameter (constant bank, 4 KB): no __constant__ upload, no stream synchronisation per call.
struct TMat { float m[32 * 32]; };

// NMAX = compile-time bound on n_in (8 / 16 / 24 / 32): the date loop is fully unrolled over NMAX with a uniform
// `n < n_in` guard, so the n_in x VEC inputs of a thread live in registers.  (A runtime-length loop indexes the array
// dynamically, which puts it in local memory: measured 1.65 TB/s = 25 % of the HBM peak for n = 24.)
template <int VEC, int NMAX>
__global__ void __launch_bounds__(256) temporal_matmul_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                              int n_in, int n_out, int64_t inner, const __grid_constant__ TMat Mk) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (i >= inner) return;
  float v[NMAX][VEC];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) {
    if (n < n_in) {
      if (VEC == 4) {
        float4 t = *reinterpret_cast<const float4*>(in + (int64_t)n * inner + i);
        v[n][0] = t.x; v[n][1 % VEC] = t.y; v[n][2 % VEC] = t.z; v[n][3 % VEC] = t.w;
      } else {
        v[n][0] = in[(int64_t)n * inner + i];
      }
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[n][k] = 0.f;
    }
  }
  for (int o = 0; o < n_out; ++o) {
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
      if (n < n_in) {                      // same sequential fma order over the dates as before
        const float m = Mk.m[o * 32 + n];
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = fmaf(m, v[n][k], acc[k]);
      }
    }
    if (VEC == 4) *reinterpret_cast<float4*>(out + (int64_t)o * inner + i) = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
    else out[(int64_t)o * inner + i] = acc[0];
  }
}

template <int VEC>
static void launch_temporal_matmul(const float* in_dev, float* out_dev, int n_in, int n_out, int64_t inner, const TMat& Mk, cudaStream_t st) {
  const int grid = cdiv(inner / VEC + (inner % VEC ? 1 : 0), 256);
  if (n_in <= 8) { temporal_matmul_kernel<VEC, 8><<<grid, 256, 0, st>>>(in_dev, out_dev, n_in, n_out, inner, Mk); }
  else if (n_in <= 16) { temporal_matmul_kernel<VEC, 16><<<grid, 256, 0, st>>>(in_dev, out_dev, n_in, n_out, inner, Mk); }
  else if (n_in <= 24) { temporal_matmul_kernel<VEC, 24><<<grid, 256, 0, st>>>(in_dev, out_dev, n_in, n_out, inner, Mk); }
  else { temporal_matmul_kernel<VEC, 32><<<grid, 256, 0, st>>>(in_dev, out_dev, n_in, n_out, inner, Mk); }
}

int pre_temporal_matmul_dev(stc_ctx* ctx, const float* in_dev, const float* M_host, int n_in, int n_out, int64_t inner,
                            float* out_dev) {
  if (n_in < 1 || n_in > 32 || n_out < 1 || n_out > 32) STC_FAIL(STC_ERR_ARG, "temporal_matmul: n_in/n_out must be in 1..32");
  TMat Mk; memset(&Mk, 0, sizeof(Mk));
  for (int o = 0; o < n_out; ++o)
    for (int n = 0; n < n_in; ++n) Mk.m[o * 32 + n] = M_host[o * n_in + n];
  bool vec = (inner % 4 == 0) && (((uintptr_t)in_dev & 15) == 0) && (((uintptr_t)out_dev & 15) == 0);
  { TraceScope ts_(ctx, "temporal_matmul_kernel");
    if (vec) launch_temporal_matmul<4>(in_dev, out_dev, n_in, n_out, inner, Mk, ctx->stream);
    else launch_temporal_matmul<1>(in_dev, out_dev, n_in, n_out, inner, Mk, ctx->stream); }
  STC_CUDA(cudaGetLastError());
  ctx->launches++;
  return STC_OK;
}

// ---- K1 fused: indices -> date regridding + Whittaker + monthly mean (one 12 x n operator) -> quarterly medians ----------
// smooth_large_tile (src/download_and_predict_job.py:1057-1096) + the quarterly composites of process_subtiles (:1274-1278) in
// ONE pass over the cube: s2 [n][HW][10] float32 is read exactly once, the 4 index channels are formed on the fly, the
// 14 x 12 monthly values of a pixel live in registers, and only what the caller needs is written (monthly [12][HW][14] and /
// or quarterly [4][HW][14]).  Algorithmic bytes per pixel: 40 n in, 56 x (12 and / or 4) out -- the unfused chain (indices,
// two products, a channel interleave, four medians: 9 launches) moved 2.6x that through HBM.
// Staging: a CTA owns tiles of FS_PX consecutive pixels; for a tile the n rows of FS_PX x 40 B (contiguous in the
// [n][HW][10] layout) are fetched by cp.async.bulk into one of two shared-memory stages, completion on an mbarrier
// (expect_tx = n x FS_PX x 40), so the copy of tile k + 1 runs under the arithmetic of tile k and no thread issues a
// global load.  FS_PX x 14 threads: thread (px, ch) reads its band (ch < 10) or recomputes its index (ch >= 10) per date.
// Arithmetic order = the unfused kernels (sequential fmaf over the dates from 0, exact index forms, insertion-sort median
// of 3): bit-identical results, which tests/test_tile_chain.py relies on (chain vs stage-by-stage mirrors).
constexpr int FS_PX = 32;
__device__ __forceinline__ uint32_t fs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fs_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
// NMAX = compile-time bound on n (8 / 16 / 24 / 32) so that the date loop unrolls: the operator is then read as immediate
// constant-bank operands of the FFMAs (a runtime-indexed M costs one LDC per FFMA) and a thread's n inputs sit in registers.
// Two phases per tile keep every warp convergent: (A) the 4 indices of the FS_PX x n (pixel, date) pairs, one pair per thread
// (32 consecutive threads = 32 pixels of one date: same code path), into shared memory; (B) thread (px, ch) gathers its n
// values (band from the staged rows, index from phase A) and forms the 12 months.  The first version let each (px, ch) thread
// recompute its own index inside the date loop: every warp ran the band path and all four index paths serially (IEEE divisions),
// 0.87 ms at n = 24, instruction-bound at 8 % of the HBM peak.
template <int NMAX>
__global__ void __launch_bounds__(FS_PX * 14) smooth_fused_kernel(const float* __restrict__ s2, int n, int64_t HW, const __grid_constant__ TMat Mk,
                                                                  float* __restrict__ monthly /*[12][HW][14] or null*/,
                                                                  float* __restrict__ quarterly /*[4][HW][14] or null*/, int stages) {
  extern __shared__ __align__(128) float fs_smem[];                  // `stages` x [n][FS_PX][10], then the index tile [n][FS_PX][4]
  __shared__ __align__(8) uint64_t bars[4];
  const int tid = threadIdx.x, px = tid / 14, ch = tid - px * 14;
  const int64_t ntiles = (HW + FS_PX - 1) / FS_PX;
  const uint32_t stage_floats = (uint32_t)n * FS_PX * 10;
  float4* idx_sm = reinterpret_cast<float4*>(fs_smem + (size_t)stages * stage_floats);
  if (tid == 0) {
    for (int q = 0; q < stages; ++q) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fs_smem_u32(&bars[q])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int64_t tile, int stage) {                        // thread 0: n bulk copies of one tile into `stage`
    const int64_t p0 = tile * FS_PX;
    const int npx = (int)((HW - p0) < FS_PX ? (HW - p0) : FS_PX);
    const uint32_t row_bytes = (uint32_t)npx * 40u;
    const uint32_t bar = fs_smem_u32(&bars[stage]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes * (uint32_t)n) : "memory");
    for (int t = 0; t < n; ++t) {
      const uint32_t dst = fs_smem_u32(fs_smem + (size_t)stage * stage_floats + (size_t)t * FS_PX * 10);
      const float* src = s2 + ((int64_t)t * HW + p0) * 10;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(src), "r"(row_bytes), "r"(bar) : "memory");
    }
  };
  int64_t tile = blockIdx.x;
  if (tid == 0)                                                      // prologue: stages - 1 tiles in flight (copy latency > one tile's arithmetic at small n)
    for (int q = 0; q < stages - 1; ++q)
      if (tile + (int64_t)q * gridDim.x < ntiles) issue(tile + (int64_t)q * gridDim.x, q);
  uint32_t phase_bits = 0;
  int stage = 0;
  for (; tile < ntiles; tile += gridDim.x) {
    const int64_t ahead = tile + (int64_t)(stages - 1) * gridDim.x;
    const int astage = stage == 0 ? stages - 1 : stage - 1;          // the stage consumed in the previous iteration (released by its barrier)
    if (tid == 0 && ahead < ntiles) issue(ahead, astage);
    fs_mbar_wait(fs_smem_u32(&bars[stage]), (phase_bits >> stage) & 1u);
    phase_bits ^= 1u << stage;
    const float* sm = fs_smem + (size_t)stage * stage_floats;
    const int64_t p0 = tile * FS_PX;
    const int npx = (int)((HW - p0) < FS_PX ? (HW - p0) : FS_PX);
    // ---- phase A: indices of every (date, pixel) of the tile ----
    for (int i = tid; i < n * FS_PX; i += FS_PX * 14) {
      const int q = i & (FS_PX - 1);
      if (q < npx) {
        const float* x = sm + (size_t)i * 10;
        idx_sm[i] = make_float4(idx_evi(x[0], x[1], x[2], x[3]), idx_bi(x[0], x[2], x[3], x[8]), idx_msavi2(x[2], x[3]), idx_grndvi(x[1], x[2], x[3]));
      }
    }
    __syncthreads();
    // ---- phase B: 12 months (and the quarterly medians) of channel ch of pixel px ----
    if (px < npx) {
      const int64_t p = p0 + px;
      float v[NMAX];
#pragma unroll
      for (int t = 0; t < NMAX; ++t)
        v[t] = (t < n) ? (ch < 10 ? sm[((size_t)t * FS_PX + px) * 10 + ch] : reinterpret_cast<const float*>(idx_sm)[((size_t)t * FS_PX + px) * 4 + (ch - 10)]) : 0.f;
      float acc[12];
#pragma unroll
      for (int o = 0; o < 12; ++o) {
        float a = 0.f;
#pragma unroll
        for (int t = 0; t < NMAX; ++t)
          if (t < n) a = fmaf(Mk.m[o * 32 + t], v[t], a);            // same sequential fma order over the dates as temporal_matmul_kernel
        acc[o] = a;
      }
      if (monthly) {
#pragma unroll
        for (int o = 0; o < 12; ++o) monthly[((int64_t)o * HW + p) * 14 + ch] = acc[o];
      }
      if (quarterly) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float a0 = acc[3 * k], a1 = acc[3 * k + 1], a2 = acc[3 * k + 2];
          float md;
          if (a0 == a0 && a1 == a1 && a2 == a2) md = med3(a0, a1, a2);          // min / max network: registers only
          else { float m3[3] = {a0, a1, a2}; md = median_n<3>(m3, 3); }          // NaN present: the insertion sort's order (same as the separate kernel)
          quarterly[((int64_t)k * HW + p) * 14 + ch] = md;
        }
      }
    }
    __syncthreads();                                                 // every thread is done with `stage` and the index tile
    stage = stage + 1 == stages ? 0 : stage + 1;
  }
}

int pre_smooth_fused_dev(stc_ctx* ctx, const float* s2_dev, const float* M_host, int n, int64_t HW, float* monthly_dev, float* quarterly_dev) {
  if (n < 1 || n > 32) STC_FAIL(STC_ERR_ARG, "smooth_fused: n must be in 1..32");
  if (((uintptr_t)s2_dev & 15) != 0 || (HW * 40) % 16 != 0) STC_FAIL(STC_ERR_ARG, "smooth_fused: the cube must be 16-byte aligned per date");
  TMat Mk; memset(&Mk, 0, sizeof(Mk));
  for (int o = 0; o < 12; ++o)
    for (int t = 0; t < n; ++t) Mk.m[o * 32 + t] = M_host[o * n + t];
  const size_t stage_bytes = (size_t)n * FS_PX * 40, idx_bytes = (size_t)n * FS_PX * 16;
  int stages = (int)((100 * 1024 - idx_bytes) / stage_bytes);                 // two CTAs per SM (registers) share ~227 KB
  stages = stages < 2 ? 2 : (stages > 4 ? 4 : stages);
  const size_t smem = stages * stage_bytes + idx_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    const int cap = 4 * 32 * FS_PX * 40 + 32 * FS_PX * 16;
    STC_CUDA(cudaFuncSetAttribute(smooth_fused_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    STC_CUDA(cudaFuncSetAttribute(smooth_fused_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    STC_CUDA(cudaFuncSetAttribute(smooth_fused_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    STC_CUDA(cudaFuncSetAttribute(smooth_fused_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ntiles = (HW + FS_PX - 1) / FS_PX;
  int per_sm = 2;                                             // resident CTAs per SM (registers: 61 x 448 -> 2; shared memory at small n allows more)
  {
    cudaError_t e = n <= 8 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smooth_fused_kernel<8>, FS_PX * 14, smem)
                  : n <= 16 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smooth_fused_kernel<16>, FS_PX * 14, smem)
                  : n <= 24 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smooth_fused_kernel<24>, FS_PX * 14, smem)
                            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smooth_fused_kernel<32>, FS_PX * 14, smem);
    if (e != cudaSuccess || per_sm < 1) per_sm = 1;
  }
  const int grid = (int)std::min<int64_t>(ntiles, (int64_t)sms * per_sm);          // one wave of persistent CTAs
  { TraceScope ts_(ctx, "smooth_fused_kernel");
    if (n <= 8) smooth_fused_kernel<8><<<grid, FS_PX * 14, smem, ctx->stream>>>(s2_dev, n, HW, Mk, monthly_dev, quarterly_dev, stages);
    else if (n <= 16) smooth_fused_kernel<16><<<grid, FS_PX * 14, smem, ctx->stream>>>(s2_dev, n, HW, Mk, monthly_dev, quarterly_dev, stages);
    else if (n <= 24) smooth_fused_kernel<24><<<grid, FS_PX * 14, smem, ctx->stream>>>(s2_dev, n, HW, Mk, monthly_dev, quarterly_dev, stages);
    else smooth_fused_kernel<32><<<grid, FS_PX * 14, smem, ctx->stream>>>(s2_dev, n, HW, Mk, monthly_dev, quarterly_dev, stages); }
  STC_CUDA(cudaGetLastError());
  ctx->launches++;
  return STC_OK;
}

// ---- indices [npix,C] -> [npix,4] ---------------------------------------------------------
__global__ void __launch_bounds__(256) indices_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t npix, int C) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const float* s = in + i * C;
  float b2 = s[0], b3 = s[1], b4 = s[2], b8 = s[3], b11 = s[8];
  float4 o = make_float4(idx_evi(b2, b3, b4, b8), idx_bi(b2, b4, b8, b11), idx_msavi2(b4, b8), idx_grndvi(b3, b4, b8));
  *reinterpret_cast<float4*>(out + i * 4) = o;
}

int pre_indices_dev(stc_ctx* ctx, const float* in_dev, int64_t npix, int C, float* out_dev) {
  if (C < 10) STC_FAIL(STC_ERR_ARG, "indices: need at least 10 bands");
  { TraceScope ts_(ctx, "indices_kernel"); indices_kernel<<<cdiv(npix, 256), 256, 0, ctx->stream>>>(in_dev, out_dev, npix, C); }
  STC_CUDA(cudaGetLastError());
  ctx->launches++;
  return STC_OK;
}

// ---- temporal median ----------------------------------------------------------------------
// N >= n slots in registers, sorted by a network (stc_sortnet.cuh); a column holding a NaN / infinity takes the insertion sort,
// whose placement of NaN the callers rely on (interpolate_na_vals goldens)
template <int N>
__global__ void __launch_bounds__(256) temporal_median_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int64_t inner) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= inner) return;
  float v[N];
  bool ok = true;
#pragma unroll
  for (int t = 0; t < N; ++t) {
    float x = INFINITY;
    if (t < n) { x = in[(int64_t)t * inner + i]; ok = ok && net_ok(x); }
    v[t] = x;
  }
  if (!ok) {
    float w[32];
    for (int t = 0; t < n; ++t) w[t] = in[(int64_t)t * inner + i];
    out[i] = median_n<32>(w, n);
    return;
  }
  sort_net<N>(v);
  out[i] = net_median<N>(v, n);
}

int pre_temporal_median_dev(stc_ctx* ctx, const float* in_dev, int n, int64_t inner, float* out_dev) {
  if (n < 1 || n > 32) STC_FAIL(STC_ERR_ARG, "temporal_median: n must be in 1..32");
  {
    TraceScope ts_(ctx, "temporal_median_kernel");
    const int grid = cdiv(inner, 256);
    if (n <= 4) temporal_median_kernel<4><<<grid, 256, 0, ctx->stream>>>(in_dev, out_dev, n, inner);
    else if (n <= 8) temporal_median_kernel<8><<<grid, 256, 0, ctx->stream>>>(in_dev, out_dev, n, inner);
    else if (n <= 16) temporal_median_kernel<16><<<grid, 256, 0, ctx->stream>>>(in_dev, out_dev, n, inner);
    else temporal_median_kernel<32><<<grid, 256, 0, ctx->stream>>>(in_dev, out_dev, n, inner);
  }
  STC_CUDA(cudaGetLastError());
  ctx->launches++;
  return STC_OK;
}
