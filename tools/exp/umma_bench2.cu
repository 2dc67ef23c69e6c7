// Micro-benchmark 2: tcgen05.mma issue rate for M=64 (cta_group::1) and M=256 (cta_group::2, CTA pair).
// Timing only (zero operands).  All waits are clock-bounded so a protocol mistake cannot hang the GPU.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

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

// ---------------- cta_group::1, runtime M (64 or 128) and N ----------------
__global__ void __launch_bounds__(128, 1) bench1(int M, int N, int n_mma, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar; __shared__ uint32_t tptr;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tm = tptr;
  if (threadIdx.x == 0) {
    uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    uint32_t sa = smem_u32(smem), sb = sa + 64 * 1024;
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      uint64_t ad = desc(sa + (uint32_t)((i % 9) * 4096), 4096 * 2, 128), bd = desc(sb + (uint32_t)((i % 9) * 8192), 256 * 16, 128);
      uint32_t acc = i > 0;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                   ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    bool ok = wait_bar(smem_u32(&bar), 0);
    out[blockIdx.x] = ok ? (clock64() - t0) : -1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

// ---------------- cta_group::2: CTA pair, M=256 (128 rows per CTA), N columns ----------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench2(int N, int n_mma, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar; __shared__ uint32_t tptr;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  cluster.sync();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster.sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tm = tptr;
  if (rank == 0 && threadIdx.x == 0) {
    uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    uint32_t sa = smem_u32(smem), sb = sa + 64 * 1024;
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      uint64_t ad = desc(sa + (uint32_t)((i % 9) * 4096), 4096 * 2, 128), bd = desc(sb + (uint32_t)((i % 9) * 4096), 128 * 16, 128);
      uint32_t acc = i > 0;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                   ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
    bool ok = wait_bar(smem_u32(&bar), 0);
    out[blockIdx.x / 2] = ok ? (clock64() - t0) : -1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster.sync();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 256 * 8);
  cudaFuncSetAttribute(bench1, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024);
  cudaFuncSetAttribute(bench2, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024);
  const int n = 4096;
  int cfg1[][2] = {{64, 64}, {64, 128}, {64, 256}, {128, 64}, {128, 128}, {128, 256}, {128, 192}, {128, 224}, {128, 240}};
  for (auto& c : cfg1) {
    bench1<<<1, 128, 180 * 1024>>>(c[0], c[1], n, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("cta_group::1 M%-3d N%-3d : %s  %.1f clk/MMA  (%.0f MAC/clk)\n", c[0], c[1], cudaGetErrorString(e), (double)h / n,
           h > 0 ? (double)c[0] * c[1] * 16 * n / h : 0.0);
    if (e != cudaSuccess) return 1;
  }
  for (int N : {64, 128, 256}) {
    for (int grid : {2, 148}) {
      bench2<<<grid, 128, 180 * 1024>>>(N, n, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[74] = {0}; cudaMemcpy(h, d, (grid / 2) * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < grid / 2; ++i) mx = (h[i] < 0) ? -1 : (h[i] > mx && mx >= 0 ? h[i] : mx);
      printf("cta_group::2 M256 N%-3d grid %3d : %s  %.1f clk/MMA  (%.0f MAC/clk per SM)\n", N, grid, cudaGetErrorString(e), (double)mx / n,
             mx > 0 ? 256.0 * N * 16 * n / mx / 2 : 0.0);
      if (e != cudaSuccess) return 1;
    }
  }
  return 0;
}
