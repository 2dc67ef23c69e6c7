// Micro-benchmark 5: tcgen05.mma rate with a warp-uniform issue path.
// Benchmarks 1-4 issued the MMA from inside `if (lane == 0)`: nvcc then cannot prove the descriptors warp-uniform and
// wraps every UTCHMMA in a convergence loop (R2UR + ELECT + UTCHMMA + 2 PLOP3 + BRA.U.ANY, ~55-70 clk per MMA) -- the
// "single-issuer floor" those benchmarks reported.  Here the whole warp runs the loop on uniform values and only the
// instruction is predicated by elect.sync; the SASS is a straight line of UTCHMMA.  Operands: zero, timing only.
// Pattern: tap (dy,dx) reads A rows shifted by dy*wp + dx*dxs (16-byte rows), K chunks `lbo` bytes apart.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred px;\nelect.sync _|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ bool wait_bar(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (clock64() - t0 > 400000000LL) return false;
  }
  return true;
}

template <int N, int NT>
__global__ void __launch_bounds__(160) bench(int n_round, int issuers, int wp, int dxs, int lbo, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4]; __shared__ uint32_t tptr;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int who = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t tm = __shfl_sync(0xffffffffu, tptr, 0);
  if (who < issuers) {
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sa = smem_u32(smem), sb = sa + 160u * 1024u;   // A region 160 KB, B region 40 KB
    const uint32_t leader = elect_one();
    const int jbase = who * NT;
    long long t0 = clock64();
    for (int r = 0; r < n_round; ++r) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3, dx = tap % 3;
        const uint64_t bd = desc(sb + (uint32_t)(tap * 2 * N * 16), N * 16, 128);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const uint64_t ad = desc(sa + (uint32_t)((dy * wp + dx * dxs + (jbase + j) * 128) * 16), (uint32_t)lbo, 128);
          const uint32_t acc = (r > 0 || tap > 0);
          if (leader)
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                         ::"r"(tm + (uint32_t)((jbase + j) * N)), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
      }
    }
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[who])) : "memory");
    bool ok = wait_bar(smem_u32(&bar[who]), 0);
    if (leader) out[blockIdx.x * 4 + who] = ok ? (clock64() - t0) : -1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

template <int N, int NT>
static void run(int issuers, int wp, int dxs, int lbo, const char* what, long long* d) {
  auto kern = bench<N, NT>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int rounds = 200;
  cudaMemset(d, 0, 148 * 4 * 8);
  kern<<<148, 160, 200 * 1024>>>(rounds, issuers, wp, dxs, lbo, d);
  cudaError_t e = cudaDeviceSynchronize();
  static long long h[148 * 4];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; bool bad = false;
  for (int i = 0; i < 148; ++i) for (int w = 0; w < issuers; ++w) { if (h[4 * i + w] <= 0) bad = true; if (h[4 * i + w] > mx) mx = h[4 * i + w]; }
  const double n_mma_sm = (double)rounds * 9 * NT * issuers;
  printf("N%-3d issuers %d x NT %d  %-34s: %s%s  %6.1f clk per MMA (SM aggregate)  %5.0f MAC/clk/SM\n", N, issuers, NT, what, cudaGetErrorString(e),
         bad ? " TIMEOUT" : "", (double)mx / n_mma_sm, 128.0 * N * 16 * n_mma_sm / (double)mx);
  if (e != cudaSuccess) exit(1);
}

template <int N>
static void sweep(long long* d) {
  run<N, 4>(1, 0, 0, 8320, "aligned, same rows", d);
  run<N, 4>(1, 1040, 0, 8320, "aligned, 3 segments", d);
  run<N, 4>(1, 1040, 1, 8320, "v1: 3 segments + dx", d);
  run<N, 4>(1, 170, 1, 13696, "v2: union, dy*170 + dx", d);
  run<N, 4>(1, 176, 1, 13696, "union, dy*176 + dx", d);
  run<N, 4>(1, 176, 0, 13696, "union, dy*176 (aligned)", d);
  run<N, 4>(1, 4, 0, 8320, "all taps 64 B off", d);
  run<N, 2>(2, 0, 0, 8320, "2 issuers, aligned", d);
  run<N, 2>(2, 1040, 1, 8320, "2 issuers, v1", d);
  run<N, 2>(2, 170, 1, 13696, "2 issuers, v2", d);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 4 * 8);
  sweep<32>(d); sweep<64>(d); sweep<128>(d);
  // N = 256 rows removed: this harness sizes one shared-memory B tile for N <= 128 (32 * N bytes per K-step overflows the
  // staging area at N = 256 -> illegal address, the last line of the round-1 log); the production kernel's N = 256 instantiation
  // (conv2) has its own geometry and is measured in profiles/r02_launches_bench.md instead.
  run<16, 4>(1, 0, 0, 8320, "aligned", d);
  run<16, 4>(1, 170, 1, 13696, "v2", d);
  return 0;
}
