// Device-memory pool of libstc: every per-call scratch buffer of the host-buffer entry points and of the tile chain
// comes from here instead of cudaMalloc / cudaFree (both synchronise the device and cost 0.1-1 ms each; a whole-tile
// run used to issue ~400 of them).  Blocks are cached per size class (8 classes per octave, <= 12.5 % slack) and
// handed out again without touching the driver.  Reuse is stream-ordered: all scratch work of a context runs on
// ctx->stream, so a block released by one stage and taken by the next is only written after the kernels that still read
// it have been enqueued ahead.  The pool is per process and device (one context per device, include/stc.h).
#include "stc_common.cuh"
#include <mutex>
#include <unordered_map>

namespace {
struct Pool {
  std::mutex mu;
  std::multimap<size_t, void*> free_blocks;          // class size -> block
  std::unordered_map<void*, size_t> live;            // block -> class size (handed out)
  size_t cached_bytes = 0, total_bytes = 0;
  int64_t hits = 0, misses = 0;
};
Pool g_pool[16];

size_t size_class(size_t bytes) {
  if (bytes < 512) return 512;
  size_t p = 512;
  while (p * 2 <= bytes) p *= 2;                      // p <= bytes < 2p
  const size_t step = p / 8;
  return (bytes + step - 1) / step * step;
}
void trim_locked(Pool& P) {
  for (auto& kv : P.free_blocks) { cudaFree(kv.second); P.total_bytes -= kv.first; }
  P.free_blocks.clear();
  P.cached_bytes = 0;
}
Pool& cur_pool() { int d = 0; cudaGetDevice(&d); return g_pool[d & 15]; }
}  // namespace

cudaError_t stc_dmalloc(void** p, size_t bytes) {
  Pool& P = cur_pool();
  const size_t cls = size_class(bytes ? bytes : 1);
  std::lock_guard<std::mutex> lk(P.mu);
  auto it = P.free_blocks.find(cls);
  if (it != P.free_blocks.end()) {
    *p = it->second; P.free_blocks.erase(it); P.cached_bytes -= cls; P.live[*p] = cls; ++P.hits;
    return cudaSuccess;
  }
  static const size_t cap = [] { const char* e = getenv("STC_POOL_CAP_GB"); return (size_t)(e ? atof(e) : 48.0) * (size_t)(1u << 30); }();
  if (P.cached_bytes + cls > cap) trim_locked(P);
  cudaError_t e = cudaMalloc(p, cls);
  if (e != cudaSuccess) {                              // out of memory: give the cached blocks back and retry once
    cudaGetLastError();
    trim_locked(P);
    e = cudaMalloc(p, cls);
    if (e != cudaSuccess) return e;
  }
  P.live[*p] = cls; P.total_bytes += cls; ++P.misses;
  return cudaSuccess;
}

cudaError_t stc_dfree(void* p) {
  if (!p) return cudaSuccess;
  Pool& P = cur_pool();
  std::lock_guard<std::mutex> lk(P.mu);
  auto it = P.live.find(p);
  if (it == P.live.end()) return cudaFree(p);          // not ours (allocated before the pool existed)
  P.free_blocks.emplace(it->second, p); P.cached_bytes += it->second; P.live.erase(it);
  return cudaSuccess;
}

void stc_pool_trim() {
  Pool& P = cur_pool();
  std::lock_guard<std::mutex> lk(P.mu);
  trim_locked(P);
}

void stc_pool_stats(int64_t* hits, int64_t* misses, size_t* cached_bytes, size_t* total_bytes) {
  Pool& P = cur_pool();
  std::lock_guard<std::mutex> lk(P.mu);
  if (hits) *hits = P.hits; if (misses) *misses = P.misses;
  if (cached_bytes) *cached_bytes = P.cached_bytes; if (total_bytes) *total_bytes = P.total_bytes;
}
