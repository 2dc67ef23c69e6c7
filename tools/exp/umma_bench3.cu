// Micro-benchmark 3: is the ~72-clk tcgen05.mma issue floor per SM or per issuer?
//  (a) two CTAs co-resident on one SM (100 KB smem, 256 TMEM columns each), each issuing its own MMA stream;
//  (b) one CTA with two issuing threads (different warps), each with its own TMEM half and barrier.
// Timing only (zero operands); all waits are clock-bounded.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

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

// issuers = 1 or 2 issuing threads per CTA (threads 0 and 32); cols = TMEM columns allocated by this CTA
__global__ void __launch_bounds__(128) bench(int N, int n_mma, int issuers, int cols, int smem_kb, long long* out, int* smid) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2]; __shared__ uint32_t tptr;
  for (int i = threadIdx.x; i < smem_kb * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); smid[blockIdx.x] = (int)s;
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"((uint32_t)cols) : "memory");
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
    const uint32_t sa = smem_u32(smem), sb = sa + (uint32_t)(smem_kb / 2) * 1024;
    const uint32_t tmem = tm + (uint32_t)who * (uint32_t)(cols / 2);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      uint64_t ad = desc(sa + (uint32_t)((i % 4) * 4096), 4096 * 2, 128), bd = desc(sb + (uint32_t)((i % 4) * 4096), 256 * 16, 128);
      uint32_t acc = i > 0;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                   ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[who])) : "memory");
    bool ok = wait_bar(smem_u32(&bar[who]), 0);
    out[blockIdx.x * 2 + who] = ok ? (clock64() - t0) : -1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"((uint32_t)cols) : "memory");
}

int main() {
  long long* d; int* s; cudaMalloc(&d, 1024 * 8); cudaMalloc(&s, 1024 * 4);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024);
  const int n = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct Cfg { int N, grid, issuers, cols, kb; const char* what; };
  Cfg cfgs[] = {{64, 148, 1, 512, 160, "1 CTA/SM, 1 issuer"}, {64, 296, 1, 256, 96, "2 CTAs/SM, 1 issuer each"},
                {64, 148, 2, 512, 160, "1 CTA/SM, 2 issuers"}, {32, 148, 1, 512, 160, "1 CTA/SM, 1 issuer"},
                {32, 296, 1, 256, 96, "2 CTAs/SM, 1 issuer each"}, {32, 148, 2, 512, 160, "1 CTA/SM, 2 issuers"},
                {32, 592, 1, 128, 48, "4 CTAs/SM, 1 issuer each"}, {128, 296, 1, 256, 96, "2 CTAs/SM, 1 issuer each"}};
  for (auto& c : cfgs) {
    cudaMemset(d, 0, 1024 * 8);
    cudaEventRecord(e0);
    bench<<<c.grid, 128, c.kb * 1024>>>(c.N, n, c.issuers, c.cols, c.kb, d, s);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    static long long h[1200]; static int sm[600];
    cudaMemcpy(h, d, (size_t)c.grid * 16, cudaMemcpyDeviceToHost); cudaMemcpy(sm, s, (size_t)c.grid * 4, cudaMemcpyDeviceToHost);
    long long mx = 0; bool bad = false; int per_sm[256] = {0}, maxco = 0;
    for (int i = 0; i < c.grid; ++i) {
      for (int w = 0; w < c.issuers; ++w) { if (h[2 * i + w] < 0) bad = true; if (h[2 * i + w] > mx) mx = h[2 * i + w]; }
      if (sm[i] >= 0 && sm[i] < 256) { per_sm[sm[i]]++; if (per_sm[sm[i]] > maxco) maxco = per_sm[sm[i]]; }
    }
    double streams_per_sm = (double)c.grid * c.issuers / 148.0;
    printf("N%-3d %-28s: %s%s  %.1f clk per MMA per issuer (max), CTAs on one SM: %d, kernel %.3f ms -> %.0f MAC/clk per SM aggregate\n", c.N, c.what,
           cudaGetErrorString(e), bad ? " TIMEOUT" : "", (double)mx / n, maxco, ms, mx > 0 ? 128.0 * c.N * 16 * n * streams_per_sm / mx : 0.0);
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
