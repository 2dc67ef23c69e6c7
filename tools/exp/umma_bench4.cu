// Micro-benchmark 4: cost of a tcgen05.mma (M128, K16, kind::f16, both operands in shared memory, no-swizzle K-major)
// as a function of the A operand's start-row shift (the 3x3 taps of the implicit-GEMM convolution are row shifts of
// dy*Wp + dx 16-byte rows), the chunk stride (LBO) and the number of issuing threads.  Timing only (zero operands).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
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
struct Pat { int shift[9]; int lbo; int nt; };   // row shifts of the 9 taps, chunk stride in bytes, sub-tiles per issuer

__global__ void __launch_bounds__(160) bench(int N, int n_round, int issuers, Pat pat, long long* out) {
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
  const uint32_t tm = tptr;
  const int who = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && who < issuers) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sa = smem_u32(smem), sb = sa + 160u * 1024u;   // A region 160 KB, B region 40 KB
    long long t0 = clock64();
    for (int r = 0; r < n_round; ++r) {
#pragma unroll 1
      for (int tap = 0; tap < 9; ++tap) {
        const uint64_t bd = desc(sb + (uint32_t)(tap * 2 * N * 16), N * 16, 128);
        for (int j = 0; j < pat.nt; ++j) {
          const int jj = who * pat.nt + j;
          const uint64_t ad = desc(sa + (uint32_t)((pat.shift[tap] + jj * 128) * 16), (uint32_t)pat.lbo, 128);
          const uint32_t acc = (r > 0 || tap > 0);
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                       ::"r"(tm + (uint32_t)(jj * N)), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[who])) : "memory");
    bool ok = wait_bar(smem_u32(&bar[who]), 0);
    out[blockIdx.x * 4 + who] = ok ? (clock64() - t0) : -1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 4 * 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int rounds = 200;
  struct Cfg { int N, issuers, nt; Pat p; const char* what; };
  auto mk = [](int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7, int s8, int lbo) {
    Pat p; int s[9] = {s0, s1, s2, s3, s4, s5, s6, s7, s8}; for (int i = 0; i < 9; ++i) p.shift[i] = s[i]; p.lbo = lbo; p.nt = 0; return p; };
  const int W = 170, R = 520;
  Pat aligned = mk(0, 0, 0, 0, 0, 0, 0, 0, 0, 8320);
  Pat v1 = mk(0, 1, 2, 2 * R, 2 * R + 1, 2 * R + 2, 4 * R, 4 * R + 1, 4 * R + 2, 8320);          // three aligned segments (chunk pairs), dx shifts
  Pat v2 = mk(0, 1, 2, W, W + 1, W + 2, 2 * W, 2 * W + 1, 2 * W + 2, 13696);                      // one union range, dy*Wp + dx shifts
  Pat v2a = mk(0, 1, 2, 176, 177, 178, 352, 353, 354, 13696);                                      // same with Wp a multiple of 8
  std::vector<Cfg> cfgs;
  for (int N : {64, 32}) for (int iss : {1, 2, 4}) {
    int nt = 4 / iss;
    cfgs.push_back({N, iss, nt, aligned, "aligned"});
    cfgs.push_back({N, iss, nt, v1, "v1 (3 segments, dx)"});
    cfgs.push_back({N, iss, nt, v2, "v2 (union, dy*170+dx)"});
    cfgs.push_back({N, iss, nt, v2a, "v2a (union, dy*176+dx)"});
  }
  for (int s = 1; s < 8; ++s) {
    Pat p = mk(s, s, s, s, s, s, s, s, s, 8320);
    static char names[8][32]; snprintf(names[s], 32, "all taps shifted %d rows", s);
    cfgs.push_back({64, 2, 2, p, names[s]});
  }
  for (int lbo : {8192, 8320, 13696, 13824, 16384}) {
    Pat p = aligned; p.lbo = lbo;
    static char names[5][32]; static int k = 0; snprintf(names[k], 32, "aligned, LBO %d", lbo);
    cfgs.push_back({64, 2, 2, p, names[k++]});
  }
  for (auto& c : cfgs) {
    c.p.nt = c.nt;
    cudaMemset(d, 0, 148 * 4 * 8);
    bench<<<148, 160, 200 * 1024>>>(c.N, rounds, c.issuers, c.p, d);
    cudaError_t e = cudaDeviceSynchronize();
    static long long h[148 * 4];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; bool bad = false;
    for (int i = 0; i < 148; ++i) for (int w = 0; w < c.issuers; ++w) { if (h[4 * i + w] < 0) bad = true; if (h[4 * i + w] > mx) mx = h[4 * i + w]; }
    const double n_mma_sm = (double)rounds * 9 * c.nt * c.issuers;
    printf("N%-3d issuers %d  %-28s: %s%s  %.1f clk per MMA (SM aggregate)\n", c.N, c.issuers, c.what, cudaGetErrorString(e), bad ? " TIMEOUT" : "", (double)mx / n_mma_sm);
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
