// Python's `random` generator replayed on the host (CPython Modules/_randommodule.c MT19937, Lib/random.py shuffle ->
// _randbelow_with_getrandbits -> getrandbits(k) = genrand_uint32() >> (32 - k)).  remove_cloud_and_shadows samples its fit
// pixels with the interpreter's global generator (cloud_removal.py:447-497), so the library takes the 624-word state,
// draws exactly what Python would draw and hands the advanced state back.  Host-only code (stc_pyrandom.cpp).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

struct PyRandom {
  static constexpr int NB = 4;                        // 624-word blocks generated per refill
  // load / store the state Python exports with random.getstate()[1]: 624 untempered words + position (0..624)
  void import_state(const uint32_t* mt624, int idx);
  void export_state(uint32_t* mt624, int* idx) const;
  // random.shuffle(v): for i = len-1 .. 1: j = _randbelow(i + 1); v[i], v[j] = v[j], v[i]
  void shuffle(int* v, size_t len);
  // the generator after shuffle() of `len` elements, without touching any data
  void skip_shuffle(size_t len);

 private:
  void refill();
  alignas(64) uint32_t hist[NB][624];                  // untempered state after each regeneration of the buffer
  alignas(64) uint32_t out[NB * 624];                  // tempered outputs
  int pos = 0, len_ = 0;                               // next output, outputs in the buffer
};
