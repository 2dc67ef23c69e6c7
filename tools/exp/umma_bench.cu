// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, M=128) from shared-memory operands for
// several N / layout / alignment settings.  Timing only (operands are zeros).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp/umma_bench tools/exp/umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         ((uint64_t)(base_off & 7) << 49) | ((uint64_t)layout << 61);
}

struct Cfg { int N; int layout; int a_off; int lbo_a, sbo_a, lbo_b, sbo_b; int n_mma; int distinct; };

__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
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
  uint32_t tm = tptr;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    uint32_t idesc = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 96 * 1024;
    t0 = clock64();
    for (int i = 0; i < c.n_mma; ++i) {
      uint32_t ao = c.a_off + (c.distinct ? (uint32_t)((i % 9) * 8320) : 0u);
      uint64_t ad = desc(sa + ao, c.lbo_a, c.sbo_a, c.layout, c.layout ? ((sa + ao) >> 7) : 0);
      uint64_t bd = desc(sb + (c.distinct ? (uint32_t)((i % 9) * 2 * c.N * 16) : 0u), c.lbo_b, c.sbo_b, c.layout, 0);
      mma(tm + (uint32_t)((i & 3) * (c.N <= 64 ? c.N : 0)), ad, bd, idesc, i > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Named { const char* name; Cfg c; int grid; };
  const int R16 = 8320;
  Named tests[] = {
      {"none N64  aligned  1CTA", {64, 0, 0, R16, 128, 64 * 16, 128, 4096, 0}, 1},
      {"none N64  +16B     1CTA", {64, 0, 16, R16, 128, 64 * 16, 128, 4096, 0}, 1},
      {"none N64  distinct 1CTA", {64, 0, 16, R16, 128, 64 * 16, 128, 4096, 1}, 1},
      {"none N64  distinct 148 ", {64, 0, 16, R16, 128, 64 * 16, 128, 4096, 1}, 148},
      {"none N32  aligned  1CTA", {32, 0, 0, R16, 128, 32 * 16, 128, 4096, 0}, 1},
      {"none N128 aligned  1CTA", {128, 0, 0, R16, 128, 128 * 16, 128, 4096, 0}, 1},
      {"none N256 aligned  1CTA", {256, 0, 0, R16, 128, 256 * 16, 128, 4096, 0}, 1},
      {"none N64 LBO=128 SBO=256 (contiguous K)", {64, 0, 0, 128, 256, 128, 256, 4096, 0}, 1},
      {"sw128 N64  aligned 1CTA", {64, 2, 0, 16, 1024, 16, 1024, 4096, 0}, 1},
      {"sw128 N64  +128B   1CTA", {64, 2, 128, 16, 1024, 16, 1024, 4096, 0}, 1},
      {"sw128 N128 aligned 1CTA", {128, 2, 0, 16, 1024, 16, 1024, 4096, 0}, 1},
      {"sw128 N256 aligned 1CTA", {256, 2, 0, 16, 1024, 16, 1024, 4096, 0}, 1},
      {"sw128 N64  aligned 148 ", {64, 2, 0, 16, 1024, 16, 1024, 4096, 0}, 148},
  };
  for (auto& t : tests) {
    bench<<<t.grid, 128, 200 * 1024>>>(t.c, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, t.grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < t.grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-45s : %s  %.1f clk/MMA (M128 N%d K16)\n", t.name, cudaGetErrorString(e), (double)mx / t.c.n_mma, t.c.N);
  }
  return 0;
}
