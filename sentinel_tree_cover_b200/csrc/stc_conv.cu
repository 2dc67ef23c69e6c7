// 3x3 convolution as implicit GEMM over the chunk-major fp16 activation layout.
//
//   conv3x3_umma_kernel : tcgen05.mma (kind::f16, fp32 accumulate in TMEM), operands staged
//                         by cp.async.bulk (UBLKCP) into shared memory, persistent CTAs,
//                         warp-specialised (copy / MMA issue / 4 epilogue warps).
//   conv3x3_simt_kernel : CUDA-core verification kernel over the SAME buffers and packed
//                         weights (tests bisect the tensor path against it).
//
// Replaces the TF Conv2D nodes of the frozen graphs (pb:*/convolution, pb:*/WSConv2D,
// pb:*conv2d*/Conv2D; reference call site src/download_and_predict_job.py:353-357).
//
// GEMM view: M = flattened padded pixels (tiles of 128 consecutive pixels), N = Cout,
// K = 9 taps x Cin.  A tap (dy,dx) of the A operand is the activation plane shifted by
// dy*Wp+dx rows; rows are 16 B apart in the no-swizzle K-major core-matrix layout, so one
// staged row-segment per dy serves the three dx taps.  Border / garbage rows are computed
// and masked in the epilogue (2-3% extra MMA work, no im2col, no halo logic).
#include "stc_common.cuh"
#include <cstdio>

// --------------------------------------------------------------------------------------
// small device helpers
// --------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }
// epilogue variant: ex2.approx + rcp.approx (relative error ~1e-6, far below the fp16 operand rounding)
__device__ __forceinline__ float fast_sigmoid(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }

__device__ __forceinline__ void pixel_coords(int64_t P, int Hp, int Wp, int& b, int& yp, int& xp) {
  int hw = Hp * Wp;
  b = (int)(P / hw);
  int rem = (int)(P - (int64_t)b * hw);
  yp = rem / Wp;
  xp = rem - yp * Wp;
}

__device__ __forceinline__ float pscale(const ConvParams& p, int yp, int xp) {
  // pb:<blk>/mask/mul : 9/(cnt+1e-8)*clip(cnt,0,1) with cnt the 3x3 in-image tap count
  int cy = 3 - (yp == p.vy0) - (yp == p.vy1 - 1);
  int cx = 3 - (xp == p.vx0) - (xp == p.vx1 - 1);
  return 9.0f / ((float)(cy * cx) + 1e-8f);
}

// ======================================================================================
// SIMT verification kernel
// ======================================================================================
template <int NB>
__global__ void __launch_bounds__(256) conv3x3_simt_kernel(ConvParams p) {
  constexpr int CPT = NB / 8;  // output channels per thread
  __shared__ uint4 sA[2][3][132];
  __shared__ uint4 sB[9][2][NB];
  __shared__ float sStat[2][16][2];
  __shared__ float sDot[128][9];
  const int dir = blockIdx.z, tid = threadIdx.x;
  const int n0 = blockIdx.y * NB;
  const int64_t p0 = (int64_t)blockIdx.x * 128;
  const int pg = tid & 31, cg = tid >> 5;
  float acc[4][CPT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int n = 0; n < CPT; ++n) acc[i][n] = 0.f;
  if (tid < 64) ((float*)sStat)[tid] = 0.f;

  const int Ksteps = p.k0steps + p.k1steps;
  for (int ks = 0; ks < Ksteps; ++ks) {
    const uint4* src; int64_t plane; int c;
    if (ks < p.k0steps) { src = p.a0[dir]; plane = p.a0_plane; c = 2 * ks; }
    else { src = p.a1[dir]; plane = p.a1_plane; c = 2 * (ks - p.k0steps); }
    for (int idx = tid; idx < 2 * 3 * 130; idx += 256) {
      int ch = idx / 390, rem = idx - ch * 390, seg = rem / 130, r = rem - seg * 130;
      int64_t gp = p0 + (int64_t)(seg - 1) * p.Wp - 1 + r;
      sA[ch][seg][r] = src[(int64_t)(c + ch) * plane + gp];
    }
    const uint4* wk = p.w[dir] + (int64_t)ks * 9 * 2 * p.N;
    for (int idx = tid; idx < 9 * 2 * NB; idx += 256) {
      int n = idx % NB, tc = idx / NB;
      sB[tc >> 1][tc & 1][n] = wk[(int64_t)tc * p.N + n0 + n];
    }
    __syncthreads();
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3, dx = tap - dy * 3;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        float a[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 v = sA[ch][dy][4 * pg + i + dx];
          const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
          for (int k = 0; k < 4; ++k) { float2 f = __half22float2(h[k]); a[i][2 * k] = f.x; a[i][2 * k + 1] = f.y; }
        }
#pragma unroll
        for (int n = 0; n < CPT; ++n) {
          uint4 v = sB[tap][ch][cg * CPT + n];
          const __half2* h = reinterpret_cast<const __half2*>(&v);
          float bw[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) { float2 f = __half22float2(h[k]); bw[2 * k] = f.x; bw[2 * k + 1] = f.y; }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[i][n] = fmaf(a[i][k], bw[k], acc[i][n]);
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue ----
  const int hw = p.Hp * p.Wp;
  const int b0 = (int)(p0 / hw);
  if (p.mode == MODE_CAND) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float d = 0.f;
#pragma unroll
      for (int n = 0; n < CPT; ++n) d += acc[i][n] * p.sse_w[dir][n0 + cg * CPT + n];
      sDot[4 * pg + i][cg] = d;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float d = 0.f;
      for (int c = 0; c < 8; ++c) d += sDot[4 * pg + i][c];
      float s = sigmoidf_(d);
#pragma unroll
      for (int n = 0; n < CPT; ++n) acc[i][n] *= s;
    }
  }
  const int gs = p.N / p.G;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t P = p0 + 4 * pg + i;
    if (P >= p.Ptot) continue;
    int b, yp, xp;
    pixel_coords(P, p.Hp, p.Wp, b, yp, xp);
    bool valid = (yp >= p.vy0 && yp < p.vy1 && xp >= p.vx0 && xp < p.vx1);
    float sc = (p.mode == MODE_PSCALE_SWISH) ? pscale(p, yp, xp) : 1.f;
    float s = 0.f, ss = 0.f;
    int gcur = (n0 + cg * CPT) / gs;
#pragma unroll
    for (int n = 0; n < CPT; ++n) {
      float v = acc[i][n];
      int nc = n0 + cg * CPT + n;
      if (p.mode == MODE_PSCALE_SWISH) { v *= sc; v = v * sigmoidf_(v); }
      else if (p.mode == MODE_SWISH) { v = v * sigmoidf_(v); }
      else if (p.mode == MODE_BIAS) { v += p.bias[nc]; }
      else if (p.mode == MODE_BIAS_RELU) { v = fmaxf(v + p.bias[nc], 0.f); }
      acc[i][n] = v;
      if (valid && p.stats[dir]) {
        int g = nc / gs;
        if (g != gcur) {
          atomicAdd(&sStat[b - b0][gcur][0], s); atomicAdd(&sStat[b - b0][gcur][1], ss);
          s = 0.f; ss = 0.f; gcur = g;
        }
        s += v; ss += v * v;
      }
    }
    if (valid && p.stats[dir]) { atomicAdd(&sStat[b - b0][gcur][0], s); atomicAdd(&sStat[b - b0][gcur][1], ss); }
    if (p.out_fp16) {
      if constexpr (CPT >= 4) {
        __half2 hh[CPT / 2];
#pragma unroll
        for (int n = 0; n < CPT / 2; ++n) hh[n] = __floats2half2_rn(acc[i][2 * n], acc[i][2 * n + 1]);
        int nc = n0 + cg * CPT;
        uint2* o = reinterpret_cast<uint2*>(reinterpret_cast<uint4*>(p.out[dir]) + (int64_t)(nc >> 3) * p.out_plane + P);
#pragma unroll
        for (int q = 0; q < CPT / 4; ++q)
          o[((nc >> 2) & 1) + q] = make_uint2(*reinterpret_cast<uint32_t*>(&hh[2 * q]), *reinterpret_cast<uint32_t*>(&hh[2 * q + 1]));
      }
    } else if constexpr (CPT >= 4) {
#pragma unroll
      for (int q = 0; q < CPT / 4; ++q) {
        int nc = n0 + cg * CPT + 4 * q;
        p.out[dir][(int64_t)(nc >> 2) * p.out_plane + P] =
            make_float4(acc[i][4 * q], acc[i][4 * q + 1], acc[i][4 * q + 2], acc[i][4 * q + 3]);
      }
    } else {
#pragma unroll
      for (int n = 0; n < CPT; ++n) {
        int nc = n0 + cg * CPT + n;
        reinterpret_cast<float*>(&p.out[dir][(int64_t)(nc >> 2) * p.out_plane + P])[nc & 3] = acc[i][n];
      }
    }
  }
  __syncthreads();
  if (p.stats[dir] && tid < 2 * p.G * 2) {
    int slot = tid / (2 * p.G), rem = tid - slot * 2 * p.G;
    float v = ((float*)sStat)[slot * 32 + rem];
    if (b0 + slot < p.B && v != 0.f) stat_add(&p.stats[dir][(int64_t)(b0 + slot) * p.G * 2 + rem], v);
  }
}

// ======================================================================================
// tcgen05 / TMEM kernel
// ======================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = clock64();
  while (true) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) break;
    __nanosleep(32);                       // do not steal issue slots from the epilogue warps
    if (clock64() - t0 > 4000000000LL) {  // ~2 s: never hang the box, trap instead
      printf("stc: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// One elected lane of a converged warp.  The MMA / bulk-copy issue loops run on the WHOLE warp with warp-uniform values
// and predicate only the instruction with this flag: inside `if (lane == 0)` nvcc cannot prove the descriptors uniform
// and wraps every UTCHMMA in a convergence loop (R2UR, ELECT, PLOP3, BRA.U.ANY: 55-70 clk per MMA, the "issue floor"
// of profiles/r01_umma_microbench*.txt); with uniform operands the SASS is a straight line of UTCHMMA.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred px;\nelect.sync _|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// [0,14) start>>4, [16,30) LBO>>4 (K-direction core-matrix stride), [32,46) SBO>>4
// (stride between 8-row groups), [46,48) version=1, [61,64) layout_type=0.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

// Incremental form for the issue loops: shared-memory addresses are < 256 KB and 16-byte aligned, so the 14-bit start
// field never carries; a descriptor for "base + r rows" (r in 16-byte units) is (hi, lo + r): one uniform add per MMA.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo) { return (saddr >> 4) | ((lbo >> 4) << 16); }
__device__ __forceinline__ uint64_t desc_join(uint32_t lo) {
  constexpr uint32_t HI = (128u >> 4) | (1u << 14);           // SBO = 128 B, descriptor version 1
  return ((uint64_t)HI << 32) | (uint64_t)lo;
}

// OCC = CTAs resident per SM.  OCC 2 halves the shared-memory budget (2 or 3 shallower stages) and the TMEM budget
// (<= 256 columns per CTA) so that two MMA issuers feed the tensor pipe of one SM (profiles/r01_umma_microbench3.txt:
// the per-instruction issue floor is per issuer, two concurrent streams raise small-N throughput by 1.2-1.4x).
template <int N, int NT, int OCC = 1>
struct UmmaCfg {
  static constexpr int R = NT * 128 + 8;                 // rows per staged segment (multiple of 8: 128-B aligned blocks)
  static constexpr int A_BYTES = 3 * 2 * R * 16;
  static constexpr int B_BYTES = 9 * 2 * N * 16;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BUDGET = (OCC == 2) ? 110 * 1024 : 200 * 1024;
  static constexpr int STAGES = (3 * STAGE_BYTES <= BUDGET) ? 3 : 2;
  static_assert(2 * STAGE_BYTES + 256 <= ((OCC == 2) ? 113 : 227) * 1024, "stages do not fit the shared-memory budget");
  static_assert(OCC == 1 || 2 * NT * N <= 256, "two resident CTAs share the 512 TMEM columns");
  static constexpr int ACC_COLS = NT * N;                // per accumulator stage
  static constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128
                                   : (2 * ACC_COLS <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256;
};

__device__ __forceinline__ uint4 pack8h(const float* v) {
  uint4 r;
  __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
  r.x = *reinterpret_cast<uint32_t*>(&h0); r.y = *reinterpret_cast<uint32_t*>(&h1);
  r.z = *reinterpret_cast<uint32_t*>(&h2); r.w = *reinterpret_cast<uint32_t*>(&h3);
  return r;
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Epilogue shared by both tcgen05 kernels: warps 2-9, TMEM -> registers -> global, fused MODE arithmetic and
// GroupNorm partial sums.  `tile` is the global super-tile index (dir = tile / tiles_per_dir).
template <int N, int NT, int G, int MODE>
__device__ __forceinline__ void conv_epilogue(const ConvParams& p, int tiles_per_dir, int tps, int tile_begin, int tile_end,
                                              uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0, int warp, int lane) {
  constexpr int ACC_COLS = NT * N;
  auto TFULL = [&](int s) { return tfull0 + 8u * s; };
  auto TEMPTY = [&](int s) { return tempty0 + 8u * s; };
  {
    // ===================== epilogue: TMEM -> registers -> global =====================
    constexpr bool SPLIT = (N >= 32);             // two warps share a lane quarter, each takes half the columns
    constexpr int NH = SPLIT ? N / 2 : N;         // columns handled by a working warp
    constexpr int GA = (G > 0) ? G / 2 : 1;       // groups inside this warp's column half
    constexpr int GS = (G > 0) ? N / G : N;       // channels per group
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;             // 0: columns [0,N/2), 1: [N/2,N)
    const int cbase = SPLIT ? half * NH : 0;
    const bool working = SPLIT || half == 0;
    const int row = q * 32 + lane;
    const int hw = p.Hp * p.Wp;
    float acc_s[GA], acc_ss[GA];
#pragma unroll
    for (int g = 0; g < GA; ++g) { acc_s[g] = 0.f; acc_ss[g] = 0.f; }
    // one flush per super-tile: warp-shuffle reduction of the lane partials, then fixed-point integer atomics (stat_add).
    // Tiles are sample-aligned, so a partial covers the same pixels in the same order wherever the sample sits.
    auto flush_warp = [&](int b, int dir) {
      if (G == 0) return;
      double* st = p.stats[dir] + ((int64_t)b * G + half * GA) * 2;
#pragma unroll
      for (int g = 0; g < GA; ++g) {
        float s = acc_s[g], ss = acc_ss[g];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
        if (lane == 0) { stat_add(st + 2 * g, s); stat_add(st + 2 * g + 1, ss); }
        acc_s[g] = 0.f; acc_ss[g] = 0.f;
      }
    };
    int it = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
      const int as = it & 1;
      const int dir = tile / tiles_per_dir;
      const int tl = tile - dir * tiles_per_dir;
      const int tb = tl / tps;                                  // the sample this super-tile belongs to
      const int k0 = (tl - tb * tps) * (NT * 128);              // first row of the tile inside the sample
      const int p0 = tb * hw + k0;
      const bool has_stats = (G > 0) && (p.stats[dir] != nullptr);
      float* const outp = reinterpret_cast<float*>(p.out[dir]);
      // Fused DSen2 epilogues read global memory (the fp32 residual, the bilinear input).  Those loads are issued one
      // sub-tile AHEAD -- for sub-tile 0 before the wait on the accumulator -- so their latency hides behind the MMAs and the
      // previous sub-tile's stores; loaded where they are consumed they made the epilogue the critical path (0.78 ms per layer).
      constexpr bool DSEN = (MODE == MODE_BIAS || MODE == MODE_BIAS_RELU);
      float4 nx_skip[(DSEN && N == 32) ? 4 : 1], cu_skip[(DSEN && N == 32) ? 4 : 1];
      float nx_bil[(DSEN && N == 16) ? 6 : 1], cu_bil[(DSEN && N == 16) ? 6 : 1];
      auto prefetch = [&](int j) {
        if (!DSEN || !working) return;
        const int rem_raw = k0 + j * 128 + row;
        if (rem_raw >= hw) return;
        const int yp = rem_raw / p.Wp, xp = rem_raw - yp * p.Wp;
        if (!(yp >= p.vy0 && yp < p.vy1 && xp >= p.vx0 && xp < p.vx1)) return;
        const int P = p0 + j * 128 + row;
        if (N == 32 && p.act16 && p.skip_mode == 2) {
#pragma unroll
          for (int qd = 0; qd < ((DSEN && N == 32) ? 4 : 1); ++qd) nx_skip[qd] = p.skip[(int64_t)((cbase >> 2) + qd) * p.skip_plane + P];
        }
        if (N == 16 && p.sr_out) {
          const int64_t px = ((int64_t)tb * (p.vy1 - p.vy0) + (yp - p.vy0)) * (p.vx1 - p.vx0) + (xp - p.vx0);
#pragma unroll
          for (int k = 0; k < ((DSEN && N == 16) ? 6 : 1); ++k) nx_bil[k] = p.sr_bil[px * p.sr_bil_stride + p.sr_bil_off + k];
        }
      };
      prefetch(0);
      mbar_wait(TFULL(as), (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      // timing experiments (STC_EXP_FLAGS bits, results invalid): 2 = no epilogue work at all, 4 = no global stores
      for (int j = 0; j < ((working && !(p.exp_flags & 2)) ? NT : 0); ++j) {
        const int rem_raw = k0 + j * 128 + row;
        const bool inb = rem_raw < hw;                           // rows past the sample's end belong to the next sample's tiles
        const int P = p0 + j * 128 + row;
        const int rem = inb ? rem_raw : 0;
        const int yp = rem / p.Wp, xp = rem - yp * p.Wp;
        const bool valid = inb && (yp >= p.vy0 && yp < p.vy1 && xp >= p.vx0 && xp < p.vx1);
        if (DSEN) {
#pragma unroll
          for (int qd = 0; qd < ((DSEN && N == 32) ? 4 : 1); ++qd) cu_skip[qd] = nx_skip[qd];
#pragma unroll
          for (int k = 0; k < ((DSEN && N == 16) ? 6 : 1); ++k) cu_bil[k] = nx_bil[k];
          if (j + 1 < NT) prefetch(j + 1);
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * ACC_COLS + j * N);
        float scale = 1.f;
        if (MODE == MODE_PSCALE_SWISH) scale = pscale(p, yp, xp);
        uint32_t keep[(MODE == MODE_CAND) ? 16 : 1];     // MODE_CAND (N = 32): this warp's 16 columns, kept from the squeeze pass
        if (MODE == MODE_CAND) {           // 1x1 squeeze over all N channels of the pixel (pb:candidate/convolution_1)
          static_assert(MODE != MODE_CAND || N == 32, "the candidate epilogue keeps one 16-column half per warp");
          uint32_t ra[16], rb[16];
          tmem_ld16_nowait(taddr, ra);
          tmem_ld16_nowait(taddr + 16, rb);
          tmem_wait_ld();
          float d = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) d = fmaf(__uint_as_float(ra[i]), __ldg(&p.sse_w[dir][i]), d);
#pragma unroll
          for (int i = 0; i < 16; ++i) d = fmaf(__uint_as_float(rb[i]), __ldg(&p.sse_w[dir][16 + i]), d);
          scale = fast_sigmoid(d);
#pragma unroll
          for (int i = 0; i < 16; ++i) keep[i] = half ? rb[i] : ra[i];      // TMEM is read once (it is shared with the MMAs)
        }
#pragma unroll
        for (int c0 = 0; c0 < NH; c0 += 16) {
          uint32_t r[16];
          if (MODE == MODE_CAND) {
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = keep[i];
          } else {
            tmem_ld16_nowait(taddr + cbase + c0, r);
            tmem_wait_ld();
          }
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x = __uint_as_float(r[i]);
            if (MODE == MODE_PSCALE_SWISH) { x *= scale; x = x * fast_sigmoid(x); }
            else if (MODE == MODE_SWISH) { x = x * fast_sigmoid(x); }
            else if (MODE == MODE_CAND) { x *= scale; }
            else if (MODE == MODE_BIAS) { x += __ldg(&p.bias[cbase + c0 + i]); }
            else if (MODE == MODE_BIAS_RELU) { x = fmaxf(x + __ldg(&p.bias[cbase + c0 + i]), 0.f); }
            v[i] = x;
            if (G > 0) {
              const float xm = valid ? x : 0.f;
              acc_s[(c0 + i) / GS] += xm;
              acc_ss[(c0 + i) / GS] = fmaf(xm, xm, acc_ss[(c0 + i) / GS]);
            }
          }
          if ((MODE == MODE_BIAS || MODE == MODE_BIAS_RELU) && (p.act16 || p.sr_out)) {
            // fused DSen2 epilogues (see ConvParams): only valid outputs are stored
            if (valid && p.sr_out) {
              const int Ww = p.vx1 - p.vx0;
              const int64_t px = ((int64_t)tb * (p.vy1 - p.vy0) + (yp - p.vy0)) * Ww + (xp - p.vx0);
              if (cbase + c0 == 0) {
#pragma unroll
                for (int k = 0; k < 6; ++k) p.sr_out[px * 6 + k] = tanhf(v[k]) + cu_bil[(DSEN && N == 16) ? k : 0];
              }
            } else if (valid) {
              const int cq = (cbase + c0) >> 2;
              if (p.skip_mode == 2) {
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                  float4* sp = p.skip + (int64_t)(cq + qd) * p.skip_plane + P;
                  const float4 s = cu_skip[(DSEN && N == 32) ? qd : 0];
                  v[4 * qd] = fmaf(0.1f, v[4 * qd], s.x); v[4 * qd + 1] = fmaf(0.1f, v[4 * qd + 1], s.y);
                  v[4 * qd + 2] = fmaf(0.1f, v[4 * qd + 2], s.z); v[4 * qd + 3] = fmaf(0.1f, v[4 * qd + 3], s.w);
                  *sp = make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
                }
              } else if (p.skip_mode == 1) {
#pragma unroll
                for (int qd = 0; qd < 4; ++qd)
                  p.skip[(int64_t)(cq + qd) * p.skip_plane + P] = make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
              }
              const uint4 o0 = pack8h(v), o1 = pack8h(v + 8);
              uint4* const a16 = p.act16 + (int64_t)((cbase + c0) >> 3) * p.act16_plane;
              const int64_t sample0 = (int64_t)tb * hw;
              const int ym = (yp == 2) ? 0 : -1, yM = (yp == p.Hp - 3) ? p.Hp - 1 : -1;      // reflect pad 1: row 0 <- row 2, row Hp-1 <- row Hp-3
              const int xm = (xp == 2) ? 0 : -1, xM = (xp == p.Wp - 3) ? p.Wp - 1 : -1;
#pragma unroll
              for (int iy = 0; iy < 3; ++iy) {
                const int yy = iy == 0 ? yp : iy == 1 ? ym : yM;
                if (yy < 0) continue;
#pragma unroll
                for (int ix = 0; ix < 3; ++ix) {
                  const int xx = ix == 0 ? xp : ix == 1 ? xm : xM;
                  if (xx < 0) continue;
                  const int64_t Pm = sample0 + (int64_t)yy * p.Wp + xx;
                  a16[Pm] = o0; a16[p.act16_plane + Pm] = o1;
                }
              }
            }
          } else if (inb && !(p.exp_flags & 4)) {
            if (p.out_fp16) {
              uint4* o = reinterpret_cast<uint4*>(outp) + (int64_t)((cbase + c0) >> 3) * p.out_plane + P;
              o[0] = pack8h(v);
              o[p.out_plane] = pack8h(v + 8);
            } else {
              float4* o = reinterpret_cast<float4*>(outp) + (int64_t)((cbase + c0) >> 2) * p.out_plane + P;
#pragma unroll
              for (int qd = 0; qd < 4; ++qd) o[(int64_t)qd * p.out_plane] = make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(TEMPTY(as));                    // the accumulator stage is free again; the statistics follow
      if (has_stats && working) flush_warp(tb, dir);
    }
  }
}

// Template parameters: N output channels, NT accumulators (128-pixel sub-tiles) per super-tile,
// G GroupNorm groups whose (sum, sumsq) the epilogue accumulates (0 = none), MODE the fused
// epilogue.  320 threads: warp 0 bulk-copy producer, warp 1 TMEM allocator + MMA issuer,
// warps 2-9 epilogue (TMEM lane quarter = warp%4, column half = (warp-2)/4).
// Statistics live in per-lane registers for one (sample-aligned) super-tile and are flushed per tile: shuffle
// reduction + 64-bit fixed-point integer atomics (order independent, see stat_add in stc_common.cuh).
template <int N, int NT, int G, int MODE, int OCC = 1>
__global__ void __launch_bounds__(320, OCC) conv3x3_umma_kernel(ConvParams p, int tiles_per_dir, int total_tiles, int tiles_per_cta, int tps) {
  using C = UmmaCfg<N, NT, OCC>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
  const uint32_t bar_base = smem_u32(bars);
  auto FULL = [&](int s) { return bar_base + 8u * s; };
  auto EMPTY = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto TFULL = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
  auto TEMPTY = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 2 + s); };

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(TFULL(s), 1); mbar_init(TEMPTY(s), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"((uint32_t)C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  const int Ksteps = p.k0steps + p.k1steps;
  const uint32_t smem_base = smem_u32(smem);
  const int tile_begin = blockIdx.x * tiles_per_cta;
  const int tile_end = (tile_begin + tiles_per_cta < total_tiles) ? tile_begin + tiles_per_cta : total_tiles;

  if (warp == 0) {
    // ===================== producer: bulk copies global -> shared =====================
    {
      const uint32_t leader = elect_one();
      int stage = 0; uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int dir = tile / tiles_per_dir;
        const int tl = tile - dir * tiles_per_dir, tb = tl / tps;
        const int64_t p0 = (int64_t)tb * (p.Hp * p.Wp) + (int64_t)(tl - tb * tps) * (NT * 128);       // sample-aligned super-tiles
        for (int ks = 0; ks < Ksteps; ++ks) {
          const uint4* src; int64_t plane; int c;
          if (ks < p.k0steps) { src = p.a0[dir]; plane = p.a0_plane; c = 2 * ks; }
          else { src = p.a1[dir]; plane = p.a1_plane; c = 2 * (ks - p.k0steps); }
          mbar_wait(EMPTY(stage), phase ^ 1);
          if (leader) mbar_expect_tx(FULL(stage), (uint32_t)C::STAGE_BYTES);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
#pragma unroll
          for (int seg = 0; seg < 3; ++seg)
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
              const uint4* g = src + (int64_t)(c + ch) * plane + p0 + (int64_t)(seg - 1) * p.Wp - 1;
              if (leader) bulk_g2s(sa + (uint32_t)((seg * 2 + ch) * C::R * 16), g, (uint32_t)(C::R * 16), FULL(stage));
            }
          if (leader) bulk_g2s(sa + C::A_BYTES, p.w[dir] + (int64_t)ks * 9 * 2 * N, (uint32_t)C::B_BYTES, FULL(stage));
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp, instruction predicated on the elected lane) =====================
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t leader = elect_one();
    int stage = 0; uint32_t phase = 0; int it = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
      const int as = it & 1;
      mbar_wait(TEMPTY(as), ((uint32_t)(it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      for (int ks = 0; ks < Ksteps; ++ks) {
        mbar_wait(FULL(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
        const uint32_t a_lo = desc_lo(sa, C::R * 16), b_lo = desc_lo(sa + C::A_BYTES, N * 16);
        const uint32_t tacc = tmem_base + (uint32_t)(as * C::ACC_COLS);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3, dx = tap - dy * 3;
          const uint64_t bdesc = desc_join(b_lo + (uint32_t)(tap * 2 * N));
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            // STC_EXP_FLAGS bit 0 (timing experiment only): drop the dx row shift so every
            // A core matrix starts 128-B aligned -- results are wrong, the MMA rate is what is measured
            const int dxe = (p.exp_flags & 1) ? 0 : dx;
            const uint64_t adesc = desc_join(a_lo + (uint32_t)((dy * 2) * C::R + j * 128 + dxe));
            if (leader) tc_mma_f16(tacc + (uint32_t)(j * N), adesc, bdesc, IDESC, (ks > 0 || tap > 0) ? 1u : 0u);
          }
        }
        if (leader) tc_commit(EMPTY(stage));
        if (leader && ks == Ksteps - 1) tc_commit(TFULL(as));
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    conv_epilogue<N, NT, G, MODE>(p, tiles_per_dir, tps, tile_begin, tile_end, tmem_base, TFULL(0), TEMPTY(0), warp, lane);
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
}

// ======================================================================================
// tcgen05 kernel, second generation (default).
//
// Why: for M = 128, K = 16 a tcgen05.mma with both operands in shared memory reads 4 KB of A and 32*N B of B;
// measured (profiles/r01_umma_microbench3.txt) the tensor pipe retires one such instruction every
// max(N/2, 32 + N/4) clk once TWO threads issue concurrently (a single issuer saturates at 55-72 clk), i.e. small-N
// convolutions are bound by the shared-memory operand reads, and every byte the copy engine writes into shared
// memory competes with them.  So this kernel
//   * runs two MMA-issuing warps (1 and 10), each owning half of the super-tile's NT accumulators;
//   * stages ONE row range per 8-channel chunk, [p0 - Wp - 1, p0 + NT*128 + Wp + 1): all nine taps are row shifts
//     (dy*Wp + dx) into it.  The three per-dy segments of the first kernel overlapped by NT*128 - Wp rows each;
//     the union is 44 % fewer bytes through L2 and into shared memory at Wp = 174;
//   * keeps the packed weights resident in shared memory for the CTA's lifetime when they fit (WRES: both ConvGRU
//     convolutions, conv_median, DSen2) instead of re-fetching 9 taps x 2 chunks x N x 16 B per K-step;
//   * gives every CTA tiles of one direction only (resident weights are per direction).
// Geometry that depends on Wp is passed at launch: RU rows per chunk (multiple of 8), `stages` pipeline depth.
// ======================================================================================
template <int N, int NT, int G, int MODE, bool WRES>
__global__ void __launch_bounds__(352, 2) conv3x3_umma2_kernel(ConvParams p, int tiles_per_dir, int ndir, int tiles_per_cta,
                                                                int RU, int stages, int iss, int tps) {
  constexpr int B_BYTES = 9 * 2 * N * 16;
  constexpr int ACC_COLS = NT * N;
  constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128
                            : (2 * ACC_COLS <= 256) ? 256 : 512;
  constexpr int MAXS = 8;
  extern __shared__ __align__(128) uint8_t smem[];
  const int Ksteps = p.k0steps + p.k1steps;
  const uint32_t a_bytes = 2u * (uint32_t)RU * 16u;
  const uint32_t stage_bytes = a_bytes + (WRES ? 0u : (uint32_t)B_BYTES);
  const uint32_t w_bytes = WRES ? (uint32_t)(Ksteps * B_BYTES) : 0u;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t stage_base = smem_base + w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + w_bytes + (uint32_t)stages * stage_bytes);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * MAXS + 5);
  const uint32_t bar_base = smem_u32(bars);
  auto FULL = [&](int s) { return bar_base + 8u * s; };
  auto EMPTY = [&](int s) { return bar_base + 8u * (MAXS + s); };
  auto TFULL = [&](int s) { return bar_base + 8u * (2 * MAXS + s); };
  auto TEMPTY = [&](int s) { return bar_base + 8u * (2 * MAXS + 2 + s); };
  const uint32_t WFULL = bar_base + 8u * (2 * MAXS + 4);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), (uint32_t)iss); }
    for (int s = 0; s < 2; ++s) { mbar_init(TFULL(s), (uint32_t)iss); mbar_init(TEMPTY(s), 8); }
    mbar_init(WFULL, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  // this CTA's contiguous run of super-tiles, all of one direction
  const int dir = blockIdx.x % ndir;
  const int local0 = (blockIdx.x / ndir) * tiles_per_cta;
  const int local1 = (local0 + tiles_per_cta < tiles_per_dir) ? local0 + tiles_per_cta : tiles_per_dir;
  const int tile_begin = dir * tiles_per_dir + local0;
  const int tile_end = dir * tiles_per_dir + (local1 > local0 ? local1 : local0);

  if (warp == 0) {
    // ===================== producer: bulk copies global -> shared =====================
    if (tile_begin < tile_end) {
      const uint32_t leader = elect_one();
      if (WRES && leader) {
        mbar_expect_tx(WFULL, w_bytes);
        bulk_g2s(smem_base, p.w[dir], w_bytes, WFULL);
      }
      int stage = 0; uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int tl = tile - dir * tiles_per_dir, tb = tl / tps;
        const int64_t p0 = (int64_t)tb * (p.Hp * p.Wp) + (int64_t)(tl - tb * tps) * (NT * 128);       // sample-aligned super-tiles
        for (int ks = 0; ks < Ksteps; ++ks) {
          const uint4* src; int64_t plane; int c;
          if (ks < p.k0steps) { src = p.a0[dir]; plane = p.a0_plane; c = 2 * ks; }
          else { src = p.a1[dir]; plane = p.a1_plane; c = 2 * (ks - p.k0steps); }
          mbar_wait(EMPTY(stage), phase ^ 1);
          if (leader) mbar_expect_tx(FULL(stage), stage_bytes);
          const uint32_t sa = stage_base + (uint32_t)stage * stage_bytes;
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            const uint4* g = src + (int64_t)(c + ch) * plane + p0 - p.Wp - 1;
            if (leader) bulk_g2s(sa + (uint32_t)ch * (uint32_t)RU * 16u, g, (uint32_t)RU * 16u, FULL(stage));
          }
          if (!WRES && leader) bulk_g2s(sa + a_bytes, p.w[dir] + (int64_t)ks * 9 * 2 * N, (uint32_t)B_BYTES, FULL(stage));
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ===================== MMA issuers (whole warp, instruction predicated on the elected lane) =====================
    const int me = (warp == 1) ? 0 : 1;
    if (me < iss) {
      constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const int jper = NT / iss;                                      // accumulators of the super-tile per issuer
      const uint32_t lbo_a = (uint32_t)RU * 16u;
      const uint32_t leader = elect_one();
      const int jlo = me * jper, jhi = jlo + jper;                    // warp-uniform: this issuer's accumulators
      if (WRES && tile_begin < tile_end) { mbar_wait(WFULL, 0); tc_fence_after(); }
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
        const int as = it & 1;
        mbar_wait(TEMPTY(as), ((uint32_t)(it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        for (int ks = 0; ks < Ksteps; ++ks) {
          mbar_wait(FULL(stage), phase);
          tc_fence_after();
          const uint32_t sa = stage_base + (uint32_t)stage * stage_bytes;
          const uint32_t a_lo = desc_lo(sa, lbo_a);
          const uint32_t b_lo = desc_lo(WRES ? smem_base + (uint32_t)(ks * B_BYTES) : sa + a_bytes, N * 16);
          const uint32_t tacc = tmem_base + (uint32_t)(as * ACC_COLS);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap - dy * 3;
            const uint64_t bdesc = desc_join(b_lo + (uint32_t)(tap * 2 * N));
            const uint32_t a_tap = a_lo + (uint32_t)(dy * p.Wp + dx);       // tap = row shift of the staged range
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              if (j >= jlo && j < jhi)
                if (leader) tc_mma_f16(tacc + (uint32_t)(j * N), desc_join(a_tap + (uint32_t)(j * 128)), bdesc, IDESC, (ks > 0 || tap > 0) ? 1u : 0u);
            }
          }
          if (leader) tc_commit(EMPTY(stage));
          if (leader && ks == Ksteps - 1) tc_commit(TFULL(as));
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    conv_epilogue<N, NT, G, MODE>(p, tiles_per_dir, tps, tile_begin, tile_end, tmem_base, TFULL(0), TEMPTY(0), warp, lane);
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// --------------------------------------------------------------------------------------
// host launchers
// --------------------------------------------------------------------------------------
template <int N, int NT, int G, int MODE, int OCC = 1>
static int launch_umma(stc_ctx* ctx, const ConvParams& p, int ndir) {
  using C = UmmaCfg<N, NT, OCC>;
  static bool configured = false;
  auto kern = conv3x3_umma_kernel<N, NT, G, MODE, OCC>;
  if (!configured) {
    STC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  if (p.Ptot >= (1ll << 31) - 1024) STC_FAIL(STC_ERR_ARG, "conv: pixel space exceeds 2^31");
  const int tps = cdiv((int64_t)p.Hp * p.Wp, NT * 128);        // super-tiles per sample (sample-aligned, see conv_epilogue)
  int tiles_per_dir = tps * p.B;
  int total = tiles_per_dir * ndir;
  const int slots = ctx->num_sms * OCC;
  int grid = total < slots ? total : slots;
  int tiles_per_cta = cdiv(total, grid);
  grid = cdiv(total, tiles_per_cta);
  kern<<<grid, 320, C::SMEM_BYTES, ctx->stream>>>(p, tiles_per_dir, total, tiles_per_cta, tps);
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

// v2 launcher.  Returns STC_ERR_ARG + "fallback" (rc = 1) when the geometry does not fit shared memory.
template <int N, int NT, int G, int MODE, bool WRES>
static int launch_umma2(stc_ctx* ctx, const ConvParams& p, int ndir, int iss_req) {
  constexpr int B_BYTES = 9 * 2 * N * 16;
  // Shared-memory budget: leave room (default 64 KB + registers; measured best of 131/163/195/227) for blocks of the HBM-bound elementwise kernels of the
  // other chunk to be co-resident with the persistent conv CTA (STC_CONV_SMEM_KB, A/B switch).
  static const int SMEM_MAX_1 = (getenv("STC_CONV_SMEM_KB") ? atoi(getenv("STC_CONV_SMEM_KB")) : 163) * 1024;
  // The DSen2 convolutions (32 input channels: two K-steps, 18 small MMAs per sub-tile) are bound by their loads and stores,
  // not by the tensor pipe: two CTAs per SM (each with half the shared-memory budget and 256 of the 512 TMEM columns) keep
  // twice the copies and epilogue stores in flight.  STC_SR_OCC=1 restores one CTA per SM (A/B).
  static const int sr_occ = getenv("STC_SR_OCC") ? atoi(getenv("STC_SR_OCC")) : 2;
  // The GRU candidate convolution (N = 32, 256 TMEM columns) is neither tensor- nor HBM-bound with one CTA per SM (31 % tensor
  // pipe, 39 % of the HBM peak): STC_CAND_OCC=2 runs two, each with per-stage weights (resident weights + three stages do
  // not fit twice).  A/B switch, default from the measurement in DESIGN.md section 4.
  static const int cand_occ = getenv("STC_CAND_OCC") ? atoi(getenv("STC_CAND_OCC")) : 1;
  const bool two_ctas = ((MODE == MODE_BIAS || MODE == MODE_BIAS_RELU) && N <= 32 && sr_occ >= 2) || (MODE == MODE_CAND && cand_occ >= 2);
  const int SMEM_MAX = two_ctas ? std::min(SMEM_MAX_1, 104 * 1024) : SMEM_MAX_1;
  static bool configured = false;
  auto kern = conv3x3_umma2_kernel<N, NT, G, MODE, WRES>;
  if (!configured) {
    STC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  if (p.Ptot >= (1ll << 31) - 1024) STC_FAIL(STC_ERR_ARG, "conv: pixel space exceeds 2^31");
  const int Ksteps = p.k0steps + p.k1steps;
  const int RU = (NT * 128 + 2 * p.Wp + 2 + 7) / 8 * 8;
  const int stage_bytes = 2 * RU * 16 + (WRES ? 0 : B_BYTES);
  const int w_bytes = WRES ? Ksteps * B_BYTES : 0;
  int stages = (SMEM_MAX - 256 - w_bytes) / stage_bytes;
  if (stages > 8) stages = 8;
  if (WRES && stages < 3) return launch_umma2<N, NT, G, MODE, false>(ctx, p, ndir, iss_req);   // budget too small for resident weights
  if (stages < 2) return 1;                        // caller reports the geometry as unsupported
  const int smem_bytes = w_bytes + stages * stage_bytes + 256;
  const int tps = cdiv((int64_t)p.Hp * p.Wp, NT * 128);        // super-tiles per sample (sample-aligned, see conv_epilogue)
  const int tiles_per_dir = tps * p.B;
  int per_dir = (two_ctas ? 2 * ctx->num_sms : ctx->num_sms) / ndir;               // CTAs per direction
  if (per_dir > tiles_per_dir) per_dir = tiles_per_dir;
  const int tiles_per_cta = cdiv(tiles_per_dir, per_dir);
  per_dir = cdiv(tiles_per_dir, tiles_per_cta);
  const int iss = (NT >= 2 && iss_req >= 2) ? 2 : 1;
  kern<<<per_dir * ndir, 352, smem_bytes, ctx->stream>>>(p, tiles_per_dir, ndir, tiles_per_cta, RU, stages, iss, tps);
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

static int launch_simt(stc_ctx* ctx, const ConvParams& p, int ndir) {
  dim3 block(256);
  if (p.N % 64 == 0) {
    dim3 grid(cdiv(p.Ptot, 128), p.N / 64, ndir);
    conv3x3_simt_kernel<64><<<grid, block, 0, ctx->stream>>>(p);
  } else if (p.N == 32) {
    dim3 grid(cdiv(p.Ptot, 128), 1, ndir);
    conv3x3_simt_kernel<32><<<grid, block, 0, ctx->stream>>>(p);
  } else if (p.N == 16) {
    dim3 grid(cdiv(p.Ptot, 128), 1, ndir);
    conv3x3_simt_kernel<16><<<grid, block, 0, ctx->stream>>>(p);
  } else {
    STC_FAIL(STC_ERR_ARG, "conv: unsupported N");
  }
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

int launch_conv(stc_ctx* ctx, const ConvParams& p_in, int ndir) {
  static const int exp_flags = getenv("STC_EXP_FLAGS") ? atoi(getenv("STC_EXP_FLAGS")) : 0;
  // A/B switches (profiling): kernel generation, MMA issuers per CTA, resident weights
  static const int conv_v = getenv("STC_CONV_V") ? atoi(getenv("STC_CONV_V")) : 2;
  static const int conv_iss = getenv("STC_CONV_ISS") ? atoi(getenv("STC_CONV_ISS")) : 1;
  static const int conv_wres = getenv("STC_CONV_WRES") ? atoi(getenv("STC_CONV_WRES")) : 1;
  static const int conv_prio = getenv("STC_CONV_PRIO") ? atoi(getenv("STC_CONV_PRIO")) : 1;
  ConvParams p = p_in;
  p.exp_flags = exp_flags;
  if (p.mode == MODE_CAND && p.N != 32) STC_FAIL(STC_ERR_ARG, "conv: MODE_CAND requires N == 32");
  if (p.G > 16 || (p.G > 0 && p.N % p.G)) STC_FAIL(STC_ERR_ARG, "conv: bad group count");
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->time_convs) {
    if (ctx->conv_events_used == ctx->conv_events.size()) {
      cudaEvent_t a, b;
      STC_CUDA(cudaEventCreate(&a)); STC_CUDA(cudaEventCreate(&b));
      ctx->conv_events.push_back({a, b});
    }
    e0 = ctx->conv_events[ctx->conv_events_used].first;
    e1 = ctx->conv_events[ctx->conv_events_used].second;
    if (ctx->conv_event_kind.size() < ctx->conv_events.size()) ctx->conv_event_kind.resize(ctx->conv_events.size());
    ctx->conv_event_kind[ctx->conv_events_used] = (p.N * 100 + (p.stats[0] ? p.G : 0)) * 10 + p.mode;
    ctx->conv_events_used++;
  }
  // Priority lane: the tensor-bound conv kernels of a chunk are enqueued on a high-priority side stream (forked from and
  // joined back into the chunk's own stream), so that while the HBM-bound elementwise kernels of the OTHER chunk occupy
  // the SMs the block scheduler places the persistent conv CTAs first and lets elementwise blocks fill the leftover
  // threads / registers, instead of running the two kernel classes back to back (scheduling).
  cudaStream_t base_stream = ctx->stream;
  int lane_slot = -1;
  if (conv_prio && ctx->conv_impl != 1) {
    lane_slot = ctx->cur_slot;
    if (!ctx->hi_stream[lane_slot]) {
      int lo = 0, hi = 0;
      STC_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      STC_CUDA(cudaStreamCreateWithPriority(&ctx->hi_stream[lane_slot], cudaStreamNonBlocking, hi));
      STC_CUDA(cudaEventCreateWithFlags(&ctx->ev_lane[lane_slot][0], cudaEventDisableTiming));
      STC_CUDA(cudaEventCreateWithFlags(&ctx->ev_lane[lane_slot][1], cudaEventDisableTiming));
    }
    STC_CUDA(cudaEventRecord(ctx->ev_lane[lane_slot][0], base_stream));
    STC_CUDA(cudaStreamWaitEvent(ctx->hi_stream[lane_slot], ctx->ev_lane[lane_slot][0], 0));
    ctx->stream = ctx->hi_stream[lane_slot];
  }
  struct Restore { stc_ctx* c; cudaStream_t s; ~Restore() { c->stream = s; } } restore{ctx, base_stream};
  if (e0) STC_CUDA(cudaEventRecord(e0, ctx->stream));
  trace_begin(ctx, p.N == 64 && p.mode == MODE_PLAIN ? "conv_gates" : p.mode == MODE_CAND ? "conv_cand" : p.N == 64 ? "conv_n64" : p.N == 128 ? "conv_n128" : p.N == 256 ? "conv_n256" : "conv_other");
  int rc;
  const int g = p.stats[0] ? p.G : 0;
  const int key = (p.N * 100 + g) * 10 + p.mode;
  rc = STC_OK;
  if (ctx->conv_impl == 1) {
    rc = launch_simt(ctx, p, ndir);
  } else if (conv_v >= 2) {
    // second-generation kernel; weights resident in shared memory when all K-steps fit next to >= 3 stages
    // (the launcher itself falls back to per-stage weights when the resident copy would leave fewer than 3 stages)
    const bool wres = conv_wres && (p.k0steps + p.k1steps) * 9 * 2 * p.N * 16 <= 80 * 1024;
#define STC_V2(NN, NTT, GG, MM) (wres ? launch_umma2<NN, NTT, GG, MM, true>(ctx, p, ndir, conv_iss) \
                                      : launch_umma2<NN, NTT, GG, MM, false>(ctx, p, ndir, conv_iss))
    switch (key) {
      case (6400 + 16) * 10 + MODE_PLAIN:        rc = STC_V2(64, 4, 16, MODE_PLAIN); break;          // GRU gates
      case (3200 + 8) * 10 + MODE_CAND:          rc = STC_V2(32, 4, 8, MODE_CAND); break;            // GRU candidate
      case (6400 + 8) * 10 + MODE_PSCALE_SWISH:  rc = STC_V2(64, 4, 8, MODE_PSCALE_SWISH); break;    // conv_median, conv_concat, up3
      case (6400 + 8) * 10 + MODE_SWISH:         rc = launch_umma2<64, 4, 8, MODE_SWISH, false>(ctx, p, ndir, conv_iss); break;         // out
      case (12800 + 8) * 10 + MODE_PSCALE_SWISH: rc = launch_umma2<128, 2, 8, MODE_PSCALE_SWISH, false>(ctx, p, ndir, conv_iss); break; // up2, up2_out
      case (12800 + 8) * 10 + MODE_SWISH:        rc = launch_umma2<128, 2, 8, MODE_SWISH, false>(ctx, p, ndir, conv_iss); break;        // conv1
      case (25600 + 8) * 10 + MODE_SWISH:        rc = launch_umma2<256, 1, 8, MODE_SWISH, false>(ctx, p, ndir, conv_iss); break;        // conv2
      case (3200 + 0) * 10 + MODE_BIAS_RELU:     rc = launch_umma2<32, 4, 0, MODE_BIAS_RELU, true>(ctx, p, ndir, conv_iss); break;      // DSen2
      case (3200 + 0) * 10 + MODE_BIAS:          rc = launch_umma2<32, 4, 0, MODE_BIAS, true>(ctx, p, ndir, conv_iss); break;
      case (1600 + 0) * 10 + MODE_BIAS:          rc = launch_umma2<16, 4, 0, MODE_BIAS, true>(ctx, p, ndir, conv_iss); break;
      default: STC_FAIL(STC_ERR_ARG, "conv: unsupported (N, groups, mode) combination");
    }
#undef STC_V2
  }
  if (ctx->conv_impl != 1 && (conv_v < 2 || rc == 1)) {
    // first-generation kernel: A/B reference (STC_CONV_V=1) and the fallback for image rows too wide for the one-range staging
    switch (key) {
      case (6400 + 16) * 10 + MODE_PLAIN:                                                                               // GRU gates
        rc = launch_umma<64, 4, 16, MODE_PLAIN>(ctx, p, ndir); break;
      case (3200 + 8) * 10 + MODE_CAND:                                                                                 // GRU candidate
        rc = launch_umma<32, 4, 8, MODE_CAND>(ctx, p, ndir); break;
      case (6400 + 8) * 10 + MODE_PSCALE_SWISH:                                                                         // conv_median, conv_concat, up3
        rc = launch_umma<64, 4, 8, MODE_PSCALE_SWISH>(ctx, p, ndir); break;
      case (6400 + 8) * 10 + MODE_SWISH:                                                                                // out
        rc = launch_umma<64, 4, 8, MODE_SWISH>(ctx, p, ndir); break;
      case (12800 + 8) * 10 + MODE_PSCALE_SWISH: rc = launch_umma<128, 2, 8, MODE_PSCALE_SWISH>(ctx, p, ndir); break; // up2, up2_out
      case (12800 + 8) * 10 + MODE_SWISH:        rc = launch_umma<128, 2, 8, MODE_SWISH>(ctx, p, ndir); break;        // conv1
      case (25600 + 8) * 10 + MODE_SWISH:        rc = launch_umma<256, 1, 8, MODE_SWISH>(ctx, p, ndir); break;        // conv2
      case (3200 + 0) * 10 + MODE_BIAS_RELU:     rc = launch_umma<32, 4, 0, MODE_BIAS_RELU>(ctx, p, ndir); break;     // DSen2
      case (3200 + 0) * 10 + MODE_BIAS:          rc = launch_umma<32, 4, 0, MODE_BIAS>(ctx, p, ndir); break;
      case (1600 + 0) * 10 + MODE_BIAS:          rc = launch_umma<16, 4, 0, MODE_BIAS>(ctx, p, ndir); break;
      default: STC_FAIL(STC_ERR_ARG, "conv: unsupported (N, groups, mode) combination");
    }
  }
  if (rc != STC_OK) return rc;
  ctx->launches++;
  trace_end(ctx);
  if (e1) STC_CUDA(cudaEventRecord(e1, ctx->stream));
  if (lane_slot >= 0) {
    STC_CUDA(cudaEventRecord(ctx->ev_lane[lane_slot][1], ctx->stream));
    STC_CUDA(cudaStreamWaitEvent(base_stream, ctx->ev_lane[lane_slot][1], 0));
  }
  return STC_OK;
}
