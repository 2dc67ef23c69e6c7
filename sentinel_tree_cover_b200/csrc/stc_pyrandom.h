// Python's `random` generator replayed on the host (CPython Modules/_randommodule.c MT19937, Lib/random.py shuffle ->
// _randbelow_with_getrandbits -> getrandbits(k) = genrand_uint32() >> (32 - k)).  remove_cloud_and_shadows samples its fit
// pixels with the interpreter's global generator (cloud_removal.py:447-497), so the library takes the 624-word state,
// draws exactly what Python would draw and hands the advanced state back.  Host-only code (stc_pyrandom.cpp).
#pragma once
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>

// The generator's output stream does not depend on how it is consumed, so a second host thread can run AHEAD of the
// consumer: it regenerates and tempers blocks into a ring of chunks of OUTPUTS, and the consumer only scans.  The state
// behind a block is not stored: tempering is a bijection, so export_state recovers the 624 state words of the block it
// stands in from that block's outputs.  Used by the data-less walk of remove_clouds (PyRandom::attach).
struct PyRandomProducer {
  static constexpr int CB = 16;                       // 624-word blocks per chunk (39 KB)
  static constexpr int R = 8;                         // chunks in the ring
  struct Chunk { alignas(64) uint32_t out[CB * 624]; };
  ~PyRandomProducer() { stop(); }
  void start(const uint32_t* state624);               // chunk 0 begins with the regeneration that follows `state624`
  void stop();
  const Chunk* acquire(long c);                       // blocks until chunk c is complete
  void release(long c) { released.store(c + 1, std::memory_order_release); }     // chunks <= c may be overwritten
 private:
  std::unique_ptr<Chunk[]> ring;
  alignas(64) uint32_t seed[624];
  std::atomic<long> produced{0}, released{0};
  std::atomic<bool> quit{false};
  std::thread th;
};

struct PyRandom {
  static constexpr int NB = 4;                        // 624-word blocks generated per refill
  // load / store the state Python exports with random.getstate()[1]: 624 untempered words + position (0..624)
  void import_state(const uint32_t* mt624, int idx);
  void export_state(uint32_t* mt624, int* idx) const;
  // random.shuffle(v): for i = len-1 .. 1: j = _randbelow(i + 1); v[i], v[j] = v[j], v[i]
  void shuffle(int* v, size_t len);
  // the generator after shuffle() of `len` elements, without touching any data
  void skip_shuffle(size_t len);
  // from the next refill on, outputs come from `p` (started here, from this generator's last state).  `p` must outlive
  // every later call, export_state included.
  void attach(PyRandomProducer* p);

 private:
  void refill();
  alignas(64) uint32_t hist_own[NB][624];              // untempered state after each regeneration of the buffer
  alignas(64) uint32_t out_own[NB * 624];              // tempered outputs
  const uint32_t (*hist)[624] = hist_own;              // states of the own buffer; null on a ring chunk (recovered from the outputs)
  const uint32_t* out = out_own;
  int pos = 0, len_ = 0;                               // next output, outputs in the buffer
  PyRandomProducer* prod = nullptr;
  long chunk = -1;                                     // ring chunk in use (-1: still on the own buffer)
};
