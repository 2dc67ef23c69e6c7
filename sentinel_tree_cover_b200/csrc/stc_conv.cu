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
    if (b0 + slot < p.B && v != 0.f) atomicAdd(&p.stats[dir][(int64_t)(b0 + slot) * p.G * 2 + rem], (double)v);
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

template <int N, int NT>
struct UmmaCfg {
  static constexpr int R = NT * 128 + 2;                 // rows per staged segment
  static constexpr int A_BYTES = 3 * 2 * R * 16;
  static constexpr int B_BYTES = 9 * 2 * N * 16;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (3 * STAGE_BYTES <= 200 * 1024) ? 3 : 2;
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

// G = GroupNorm groups whose (sum, sumsq) the epilogue accumulates (0 = no statistics).
// Statistics live in per-lane registers across the CTA's contiguous run of tiles and are
// flushed (warp shuffle reduction + one fp64 atomic per value) only when the sample changes:
// a few hundred atomics per launch instead of one per warp-row (the first version spent
// >90% of the kernel serialised on same-address L2 reductions; profiles/r01_*).
template <int N, int NT, int G>
__global__ void __launch_bounds__(192, 1) conv3x3_umma_kernel(ConvParams p, int tiles_per_dir, int total_tiles, int tiles_per_cta) {
  using C = UmmaCfg<N, NT>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
  const uint32_t bar_base = smem_u32(bars);
  auto FULL = [&](int s) { return bar_base + 8u * s; };
  auto EMPTY = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto TFULL = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
  auto TEMPTY = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 2 + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(TFULL(s), 1); mbar_init(TEMPTY(s), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"((uint32_t)C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int Ksteps = p.k0steps + p.k1steps;
  const uint32_t smem_base = smem_u32(smem);
  const int tile_begin = blockIdx.x * tiles_per_cta;
  const int tile_end = (tile_begin + tiles_per_cta < total_tiles) ? tile_begin + tiles_per_cta : total_tiles;

  if (warp == 0) {
    // ===================== producer: bulk copies global -> shared =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int dir = tile / tiles_per_dir;
        const int64_t p0 = (int64_t)(tile - dir * tiles_per_dir) * (NT * 128);
        for (int ks = 0; ks < Ksteps; ++ks) {
          const uint4* src; int64_t plane; int c;
          if (ks < p.k0steps) { src = p.a0[dir]; plane = p.a0_plane; c = 2 * ks; }
          else { src = p.a1[dir]; plane = p.a1_plane; c = 2 * (ks - p.k0steps); }
          mbar_wait(EMPTY(stage), phase ^ 1);
          mbar_expect_tx(FULL(stage), (uint32_t)C::STAGE_BYTES);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
#pragma unroll
          for (int seg = 0; seg < 3; ++seg)
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
              const uint4* g = src + (int64_t)(c + ch) * plane + p0 + (int64_t)(seg - 1) * p.Wp - 1;
              bulk_g2s(sa + (uint32_t)((seg * 2 + ch) * C::R * 16), g, (uint32_t)(C::R * 16), FULL(stage));
            }
          bulk_g2s(sa + C::A_BYTES, p.w[dir] + (int64_t)ks * 9 * 2 * N, (uint32_t)C::B_BYTES, FULL(stage));
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected lane) =====================
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int stage = 0; uint32_t phase = 0; int it = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
      const int as = it & 1;
      mbar_wait(TEMPTY(as), ((uint32_t)(it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      for (int ks = 0; ks < Ksteps; ++ks) {
        mbar_wait(FULL(stage), phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap - dy * 3;
            const uint64_t bdesc = make_desc(sb + (uint32_t)(tap * 2 * N * 16), N * 16, 128);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              const uint64_t adesc = make_desc(sa + (uint32_t)(((dy * 2) * C::R + j * 128 + dx) * 16), C::R * 16, 128);
              tc_mma_f16(tmem_base + (uint32_t)(as * C::ACC_COLS + j * N), adesc, bdesc, IDESC, (ks > 0 || tap > 0) ? 1u : 0u);
            }
          }
          tc_commit(EMPTY(stage));
          if (ks == Ksteps - 1) tc_commit(TFULL(as));
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    constexpr int GA = (G > 0) ? G : 1;
    constexpr int GS = N / GA;            // channels per group
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    float acc_s[GA], acc_ss[GA];
#pragma unroll
    for (int g = 0; g < GA; ++g) { acc_s[g] = 0.f; acc_ss[g] = 0.f; }
    int cur_b = -1, cur_dir = 0;
    auto flush_warp = [&]() {   // warp-uniform: reduce over lanes, lanes 0..2G-1 each add one value
      if (G == 0 || cur_b < 0) return;
#pragma unroll
      for (int g = 0; g < GA; ++g) {
        float s = acc_s[g], ss = acc_ss[g];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
        if (lane == 0) {
          atomicAdd(&p.stats[cur_dir][((int64_t)cur_b * GA + g) * 2 + 0], (double)s);
          atomicAdd(&p.stats[cur_dir][((int64_t)cur_b * GA + g) * 2 + 1], (double)ss);
        }
        acc_s[g] = 0.f; acc_ss[g] = 0.f;
      }
    };
    int it = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
      const int as = it & 1;
      const int dir = tile / tiles_per_dir;
      const int64_t p0 = (int64_t)(tile - dir * tiles_per_dir) * (NT * 128);
      mbar_wait(TFULL(as), (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < NT; ++j) {
        const int64_t P = p0 + j * 128 + row;
        const bool inb = P < p.Ptot;
        int b, yp, xp;
        pixel_coords(inb ? P : 0, p.Hp, p.Wp, b, yp, xp);
        const bool valid = inb && (yp >= p.vy0 && yp < p.vy1 && xp >= p.vx0 && xp < p.vx1);
        const float sc = (p.mode == MODE_PSCALE_SWISH) ? pscale(p, yp, xp) : 1.f;
        const float vm = valid ? 1.f : 0.f;
        bool per_lane_flush = false;
        if (G > 0 && p.stats[dir]) {
          const unsigned mk = __ballot_sync(0xffffffffu, inb);
          if (mk) {
            const int bw = __shfl_sync(0xffffffffu, b, __ffs(mk) - 1);
            const bool uniform = __all_sync(0xffffffffu, !inb || b == bw);
            if (uniform) {
              if (bw != cur_b || dir != cur_dir) { flush_warp(); cur_b = bw; cur_dir = dir; }
            } else {              // the warp's 32 pixels straddle two samples (once per sample boundary)
              flush_warp(); cur_b = -1; per_lane_flush = true;
            }
          }
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * C::ACC_COLS + j * N);
        float sse = 1.f;
        if (p.mode == MODE_CAND) {
          float d = 0.f;
#pragma unroll
          for (int c0 = 0; c0 < N; c0 += 16) {
            float v[16];
            tmem_ld16(taddr + c0, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) d += v[i] * __ldg(&p.sse_w[dir][c0 + i]);
          }
          sse = sigmoidf_(d);
        }
#pragma unroll
        for (int c0 = 0; c0 < N; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + c0, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x = v[i];
            if (p.mode == MODE_PSCALE_SWISH) { x *= sc; x = x * sigmoidf_(x); }
            else if (p.mode == MODE_SWISH) { x = x * sigmoidf_(x); }
            else if (p.mode == MODE_CAND) { x *= sse; }
            else if (p.mode == MODE_BIAS) { x += __ldg(&p.bias[c0 + i]); }
            else if (p.mode == MODE_BIAS_RELU) { x = fmaxf(x + __ldg(&p.bias[c0 + i]), 0.f); }
            v[i] = x;
            if (G > 0) {
              const float xm = x * vm;
              acc_s[(c0 + i) / GS] += xm;
              acc_ss[(c0 + i) / GS] += xm * xm;
            }
          }
          if (inb) {
            if (p.out_fp16) {
              uint4* o = reinterpret_cast<uint4*>(p.out[dir]);
              o[(int64_t)((c0 >> 3) + 0) * p.out_plane + P] = pack8h(v);
              o[(int64_t)((c0 >> 3) + 1) * p.out_plane + P] = pack8h(v + 8);
            } else {
#pragma unroll
              for (int qd = 0; qd < 4; ++qd)
                p.out[dir][(int64_t)((c0 >> 2) + qd) * p.out_plane + P] = make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
            }
          }
        }
        if (G > 0 && per_lane_flush) {
#pragma unroll
          for (int g = 0; g < GA; ++g) {
            if (valid) {
              atomicAdd(&p.stats[dir][((int64_t)b * GA + g) * 2 + 0], (double)acc_s[g]);
              atomicAdd(&p.stats[dir][((int64_t)b * GA + g) * 2 + 1], (double)acc_ss[g]);
            }
            acc_s[g] = 0.f; acc_ss[g] = 0.f;
          }
        } else if (G > 0 && !p.stats[dir]) {
#pragma unroll
          for (int g = 0; g < GA; ++g) { acc_s[g] = 0.f; acc_ss[g] = 0.f; }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(TEMPTY(as));
    }
    flush_warp();
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
}

// --------------------------------------------------------------------------------------
// host launchers
// --------------------------------------------------------------------------------------
template <int N, int NT, int G>
static int launch_umma(stc_ctx* ctx, const ConvParams& p, int ndir) {
  using C = UmmaCfg<N, NT>;
  static bool configured = false;
  auto kern = conv3x3_umma_kernel<N, NT, G>;
  if (!configured) {
    STC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  int tiles_per_dir = cdiv(p.Ptot, NT * 128);
  int total = tiles_per_dir * ndir;
  int grid = total < ctx->num_sms ? total : ctx->num_sms;
  int tiles_per_cta = cdiv(total, grid);
  grid = cdiv(total, tiles_per_cta);
  kern<<<grid, 192, C::SMEM_BYTES, ctx->stream>>>(p, tiles_per_dir, total, tiles_per_cta);
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

int launch_conv(stc_ctx* ctx, const ConvParams& p, int ndir) {
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
    ctx->conv_events_used++;
    STC_CUDA(cudaEventRecord(e0, ctx->stream));
  }
  int rc;
  if (ctx->conv_impl == 1) {
    rc = launch_simt(ctx, p, ndir);
  } else {
    const int key = p.N * 100 + (p.stats[0] ? p.G : 0);
    switch (key) {
      case 1600: rc = launch_umma<16, 4, 0>(ctx, p, ndir); break;
      case 3200: rc = launch_umma<32, 4, 0>(ctx, p, ndir); break;
      case 3208: rc = launch_umma<32, 4, 8>(ctx, p, ndir); break;
      case 6408: rc = launch_umma<64, 4, 8>(ctx, p, ndir); break;
      case 6416: rc = launch_umma<64, 4, 16>(ctx, p, ndir); break;
      case 12808: rc = launch_umma<128, 2, 8>(ctx, p, ndir); break;
      case 25608: rc = launch_umma<256, 1, 8>(ctx, p, ndir); break;
      default: STC_FAIL(STC_ERR_ARG, "conv: unsupported (N, groups) combination");
    }
  }
  if (rc != STC_OK) return rc;
  ctx->launches++;
  if (e1) STC_CUDA(cudaEventRecord(e1, ctx->stream));
  return STC_OK;
}
