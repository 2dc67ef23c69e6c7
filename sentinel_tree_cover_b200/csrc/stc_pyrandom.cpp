// See stc_pyrandom.h.  The replay of random.shuffle over ~0.7 M list elements per date is the one long sequential host
// loop of the cloud-removal stage, so:
//  * the generator regenerates and tempers 4 x 624 outputs at a time with loops the compiler vectorises (function
//    multi-versioning: AVX-512 / AVX2 / baseline picked at load time): 0.25-0.5 ns per output;
//  * every draw consumes one output whether it is accepted or not (_randbelow draws k = (i+1).bit_length() bits until the
//    value is <= i), so shuffle() advances one output per iteration and turns the accept test into arithmetic -- a
//    rejected draw swaps v[i] with itself and leaves i alone: no unpredictable branch;
//  * skip_shuffle() walks the generator through a shuffle WITHOUT data, which is what lets the caller farm the shuffles
//    themselves (dependent random memory accesses, ~3.5 ns per element) out to worker threads that each start from a
//    recorded state.  i only drops by the number of accepts, so inside a block of B outputs every r <= i - B is
//    accepted and every r > i rejected whatever the order: blocks without a value in between are counted with one
//    branch-free pass (B grows with i: the chance of an in-between value is ~B^2 / i), the others are halved down
//    to the scalar walk.
#include "stc_pyrandom.h"

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define STC_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define STC_CLONES
#endif

namespace {
inline uint32_t twist(uint32_t u, uint32_t v, uint32_t m) {
  const uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
  return m ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
}
// next 624-word state from `prev` (may alias `mt`), and its tempered outputs
STC_CLONES void regen_block(const uint32_t* prev, uint32_t* __restrict__ mt, uint32_t* __restrict__ out) {
  if (prev != mt) memcpy(mt, prev, 624 * 4);
  for (int k = 0; k < 227; ++k) mt[k] = twist(mt[k], mt[k + 1], mt[k + 397]);
  for (int k = 227; k < 454; ++k) mt[k] = twist(mt[k], mt[k + 1], mt[k - 227]);
  for (int k = 454; k < 623; ++k) mt[k] = twist(mt[k], mt[k + 1], mt[k - 227]);
  mt[623] = twist(mt[623], mt[0], mt[396]);
  for (int k = 0; k < 624; ++k) {
    uint32_t y = mt[k];
    y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
    out[k] = y;
  }
}
STC_CLONES void temper_block(const uint32_t* __restrict__ mt, uint32_t* __restrict__ out) {
  for (int k = 0; k < 624; ++k) {
    uint32_t y = mt[k];
    y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
    out[k] = y;
  }
}
STC_CLONES void block_counts(const uint32_t* __restrict__ o, int nb, int sh, uint32_t safe, uint32_t hi, int* sure, int* upto_hi) {
  int s = 0, a = 0;
  for (int q = 0; q < nb; ++q) { const uint32_t r = o[q] >> sh; s += (r <= safe); a += (r <= hi); }
  *sure = s; *upto_hi = a;
}
}  // namespace

void PyRandom::import_state(const uint32_t* mt624, int idx) {
  memcpy(hist[0], mt624, 624 * 4);
  temper_block(hist[0], out);
  pos = idx; len_ = 624;
}

void PyRandom::export_state(uint32_t* mt624, int* idx) const {
  const int b = pos == 0 ? 0 : (pos - 1) / 624;
  memcpy(mt624, hist[b], 624 * 4);
  *idx = pos - b * 624;
}

void PyRandom::refill() {                              // only called when an output is needed and the buffer is spent
  const int last = len_ / 624 - 1;
  alignas(64) uint32_t prev[624];
  memcpy(prev, hist[last], sizeof(prev));
  regen_block(prev, hist[0], out);
  for (int b = 1; b < NB; ++b) regen_block(hist[b - 1], hist[b], out + b * 624);
  pos = 0; len_ = NB * 624;
}

void PyRandom::shuffle(int* v, size_t len) {
  size_t i = len;
  if (i < 2) return;
  --i;                                                 // i = len - 1
  while (i >= 1) {
    const int sh = __builtin_clz((uint32_t)i + 1u);    // all i of one power-of-two band of i + 1 share the shift
    const size_t band_lo = (sh == 31) ? 1 : ((size_t)1 << (31 - sh));
    const size_t lo = band_lo > 1 ? band_lo - 1 : 1;   // i + 1 >= band_lo  <=>  i >= band_lo - 1
    while (i >= lo) {
      if (pos >= len_) refill();
      const int avail = len_ - pos;
      const uint32_t* o = out + pos;
      int used = 0;
      while (used < avail && i >= lo) {
        const uint32_t r = o[used++] >> sh;
        const bool ok = r <= (uint32_t)i;
        const size_t j = ok ? (size_t)r : i;
        const int t = v[i]; v[i] = v[j]; v[j] = t;
        i -= ok;
      }
      pos += used;
    }
  }
}

void PyRandom::skip_shuffle(size_t len) {
  size_t i = len;
  if (i < 2) return;
  --i;
  while (i >= 1) {
    const int sh = __builtin_clz((uint32_t)i + 1u);
    const size_t band_lo = (sh == 31) ? 1 : ((size_t)1 << (31 - sh));
    const size_t lo = band_lo > 1 ? band_lo - 1 : 1;
    while (i >= lo) {
      if (pos >= len_) refill();
      const int avail = len_ - pos;
      const uint32_t* o = out + pos;
      int used = 0;
      while (used < avail && i >= lo) {
        int B = i >= (1u << 22) ? 512 : i >= (1u << 20) ? 256 : i >= (1u << 18) ? 128 : i >= (1u << 16) ? 64 : i >= (1u << 13) ? 32 : 0;
        bool done = false;
        for (; B >= 32; B >>= 1) {
          if (avail - used < B || i < lo + (size_t)B) continue;
          int sure, upto;
          block_counts(o + used, B, sh, (uint32_t)(i - (size_t)B), (uint32_t)i, &sure, &upto);
          if (upto == sure) { i -= (size_t)sure; used += B; done = true; break; }
        }
        if (done) continue;
        const uint32_t r = o[used++] >> sh;
        i -= (r <= (uint32_t)i);
      }
      pos += used;
    }
  }
}
