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
#include <cstdlib>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif

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
#if defined(__x86_64__) && defined(__GNUC__)
// The same regeneration + tempering with explicit AVX-512: the state is 39 aligned vectors; mt[k + 1] and mt[k + 397] are
// built from two aligned loads each (valignd by 1 and by 13 lanes: 397 = 24 * 16 + 13), so no load crosses a cache line,
// and updating in place with wrapped vector indices IS the reference recurrence (a wrapped index reads a vector this pass
// already rewrote, which is exactly mt[k - 227] / the new mt[0] of the scalar loops).  ~1.5x the auto-vectorised loops.
__attribute__((target("avx512f"))) void regen_block_avx512(uint32_t* __restrict__ mt /*64-byte aligned, in place*/, uint32_t* __restrict__ out) {
  const __m512i UPPER = _mm512_set1_epi32((int)0x80000000u), MATRIX = _mm512_set1_epi32((int)0x9908b0dfu), ONE = _mm512_set1_epi32(1);
  const __m512i TB = _mm512_set1_epi32((int)0x9d2c5680u), TC = _mm512_set1_epi32((int)0xefc60000u);
  for (int j = 0; j < 39; ++j) {
    const int j1 = j + 1 == 39 ? 0 : j + 1, ja = j + 24 >= 39 ? j + 24 - 39 : j + 24, jb = j + 25 >= 39 ? j + 25 - 39 : j + 25;
    const __m512i cur = _mm512_load_si512(mt + 16 * j), nxt = _mm512_load_si512(mt + 16 * j1);
    const __m512i mlo = _mm512_load_si512(mt + 16 * ja), mhi = _mm512_load_si512(mt + 16 * jb);
    const __m512i v = _mm512_alignr_epi32(nxt, cur, 1), m = _mm512_alignr_epi32(mhi, mlo, 13);
    const __m512i y = _mm512_ternarylogic_epi32(UPPER, cur, v, 0xCA);                 // (cur & UPPER) | (v & ~UPPER)
    __m512i t = _mm512_xor_si512(m, _mm512_srli_epi32(y, 1));
    t = _mm512_mask_xor_epi32(t, _mm512_test_epi32_mask(y, ONE), t, MATRIX);
    _mm512_store_si512(mt + 16 * j, t);
    __m512i z = _mm512_xor_si512(t, _mm512_srli_epi32(t, 11));
    z = _mm512_xor_si512(z, _mm512_and_si512(_mm512_slli_epi32(z, 7), TB));
    z = _mm512_xor_si512(z, _mm512_and_si512(_mm512_slli_epi32(z, 15), TC));
    z = _mm512_xor_si512(z, _mm512_srli_epi32(z, 18));
    _mm512_storeu_si512(out + 16 * j, z);
  }
}
const bool have_avx512 = __builtin_cpu_supports("avx512f") && !(getenv("STC_PYRANDOM_ISA") && atoi(getenv("STC_PYRANDOM_ISA")) < 2);
#else
const bool have_avx512 = false;
inline void regen_block_avx512(uint32_t*, uint32_t*) {}
#endif
// in-place regeneration of a 64-byte aligned state with the widest instruction set at hand
inline void regen_inplace(uint32_t* mt, uint32_t* out) {
  if (have_avx512) regen_block_avx512(mt, out); else regen_block(mt, mt, out);
}
STC_CLONES void temper_block(const uint32_t* __restrict__ mt, uint32_t* __restrict__ out) {
  for (int k = 0; k < 624; ++k) {
    uint32_t y = mt[k];
    y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
    out[k] = y;
  }
}
// inverse of the tempering: y ^= y >> 11; y ^= (y << 7) & B; y ^= (y << 15) & C; y ^= y >> 18, undone last to first
// (x = y ^ f(x) solved by iterating from x = y: every round fixes another `shift` bits)
STC_CLONES void untemper_block(const uint32_t* __restrict__ out, uint32_t* __restrict__ mt) {
  for (int k = 0; k < 624; ++k) {
    uint32_t y = out[k], x;
    y ^= (y >> 18);
    x = y; x = y ^ ((x << 15) & 0xefc60000u); x = y ^ ((x << 15) & 0xefc60000u); y = x;
    x = y; for (int q = 0; q < 5; ++q) x = y ^ ((x << 7) & 0x9d2c5680u); y = x;
    x = y; x = y ^ (x >> 11); x = y ^ (x >> 11); y = x;
    mt[k] = y;
  }
}
STC_CLONES void block_counts(const uint32_t* __restrict__ o, int nb, int sh, uint32_t safe, uint32_t hi, int* sure, int* upto_hi) {
  int s = 0, a = 0;
  for (int q = 0; q < nb; ++q) { const uint32_t r = o[q] >> sh; s += (r <= safe); a += (r <= hi); }
  *sure = s; *upto_hi = a;
}
}  // namespace

void PyRandom::import_state(const uint32_t* mt624, int idx) {
  hist = hist_own; out = out_own; prod = nullptr; chunk = -1;
  memcpy(hist_own[0], mt624, 624 * 4);
  temper_block(hist_own[0], out_own);
  pos = idx; len_ = 624;
}

void PyRandom::attach(PyRandomProducer* p) {
  prod = p; chunk = -1;
  p->start(hist[len_ / 624 - 1]);
}

void PyRandomProducer::start(const uint32_t* state624) {
  stop();
  if (!ring) ring.reset(new Chunk[R]);
  produced.store(0); released.store(0); quit.store(false);
  memcpy(seed, state624, sizeof(seed));
  th = std::thread([this]() {
    alignas(64) uint32_t mt[624];
    memcpy(mt, seed, sizeof(mt));
    for (long c = 0;; ++c) {
      while (c - released.load(std::memory_order_acquire) >= R) {          // the consumer still reads the slot
        if (quit.load(std::memory_order_relaxed)) return;
        std::this_thread::yield();
      }
      if (quit.load(std::memory_order_relaxed)) return;
      Chunk& k = ring[c % R];
      for (int b = 0; b < CB; ++b) regen_inplace(mt, k.out + b * 624);
      produced.store(c + 1, std::memory_order_release);
    }
  });
}

void PyRandomProducer::stop() {
  if (th.joinable()) { quit.store(true); th.join(); }
}

const PyRandomProducer::Chunk* PyRandomProducer::acquire(long c) {
  while (produced.load(std::memory_order_acquire) <= c) std::this_thread::yield();
  return &ring[c % R];
}

void PyRandom::export_state(uint32_t* mt624, int* idx) const {
  const int b = pos == 0 ? 0 : (pos - 1) / 624;
  if (hist) memcpy(mt624, hist[b], 624 * 4);
  else untemper_block(out + b * 624, mt624);            // ring chunk: the state words are the untempered outputs
  *idx = pos - b * 624;
}

void PyRandom::refill() {                              // only called when an output is needed and the buffer is spent
  if (prod) {
    if (chunk >= 0) prod->release(chunk);
    const PyRandomProducer::Chunk* k = prod->acquire(++chunk);
    hist = nullptr; out = k->out;
    pos = 0; len_ = PyRandomProducer::CB * 624;
    return;
  }
  const int last = len_ / 624 - 1;
  alignas(64) uint32_t prev[624];
  memcpy(prev, hist[last], sizeof(prev));
  memcpy(hist_own[0], prev, sizeof(prev));
  regen_inplace(hist_own[0], out_own);
  for (int b = 1; b < NB; ++b) { memcpy(hist_own[b], hist_own[b - 1], 624 * 4); regen_inplace(hist_own[b], out_own + b * 624); }
  hist = hist_own; out = out_own;
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

// ---- data-less walk ----------------------------------------------------------------------------------------------------
// scan_*(o, avail, sh, i, lo): consume outputs o[0..) while i >= lo, at most `avail`; returns the number consumed and
// updates i (every accepted draw, r = o >> sh <= i, lowers i by one).  Inside a group of B outputs i only drops by the
// number of accepts, so every r <= i - B is accepted and every r > i rejected whatever the order: a group without a
// value in between is counted from two compare masks (64 outputs per step while 64^2 is small against i, then 16), a
// group with one is walked output by output.  The loop-carried chain is one popcount and one subtraction per group.
namespace {
inline int scan_scalar(const uint32_t* o, int avail, int sh, size_t& i, size_t lo) {
  int used = 0;
  size_t ii = i;
  while (used < avail && ii >= lo) { const uint32_t r = o[used++] >> sh; ii -= (r <= (uint32_t)ii); }
  i = ii;
  return used;
}
#if defined(__x86_64__) && defined(__GNUC__)
#define STC_T512 __attribute__((target("avx512f,popcnt"), always_inline)) inline
// one group of 16 / 64 / 256 outputs; a group with a value in (i - W, i] is redone with four narrower groups (scalar below 16)
STC_T512 void step16(const uint32_t* o, __m128i vsh, int sh, size_t& i) {
  const __m512i r0 = _mm512_srl_epi32(_mm512_loadu_si512(o), vsh);
  const unsigned mh = _mm512_cmple_epu32_mask(r0, _mm512_set1_epi32((int)(uint32_t)i));
  const unsigned ml = _mm512_cmple_epu32_mask(r0, _mm512_set1_epi32((int)(uint32_t)(i - 16)));
  if (mh == ml) i -= (size_t)__builtin_popcount(mh);
  else for (int q = 0; q < 16; ++q) { const uint32_t r = o[q] >> sh; i -= (r <= (uint32_t)i); }
}
STC_T512 bool masks64(const uint32_t* o, __m128i vsh, __m512i hi, __m512i lw, int& accepted) {
  const __m512i r0 = _mm512_srl_epi32(_mm512_loadu_si512(o), vsh), r1 = _mm512_srl_epi32(_mm512_loadu_si512(o + 16), vsh);
  const __m512i r2 = _mm512_srl_epi32(_mm512_loadu_si512(o + 32), vsh), r3 = _mm512_srl_epi32(_mm512_loadu_si512(o + 48), vsh);
  const uint64_t mh = (uint64_t)_mm512_cmple_epu32_mask(r0, hi) | ((uint64_t)_mm512_cmple_epu32_mask(r1, hi) << 16) |
                      ((uint64_t)_mm512_cmple_epu32_mask(r2, hi) << 32) | ((uint64_t)_mm512_cmple_epu32_mask(r3, hi) << 48);
  const uint64_t ml = (uint64_t)_mm512_cmple_epu32_mask(r0, lw) | ((uint64_t)_mm512_cmple_epu32_mask(r1, lw) << 16) |
                      ((uint64_t)_mm512_cmple_epu32_mask(r2, lw) << 32) | ((uint64_t)_mm512_cmple_epu32_mask(r3, lw) << 48);
  accepted = __builtin_popcountll(mh);
  return mh == ml;
}
STC_T512 void step64(const uint32_t* o, __m128i vsh, int sh, size_t& i) {
  int acc;
  if (masks64(o, vsh, _mm512_set1_epi32((int)(uint32_t)i), _mm512_set1_epi32((int)(uint32_t)(i - 64)), acc)) i -= (size_t)acc;
  else for (int q = 0; q < 64; q += 16) step16(o + q, vsh, sh, i);
}
__attribute__((target("avx512f,popcnt"))) int scan_avx512(const uint32_t* o, int avail, int sh, size_t& i_io, size_t lo) {
  int used = 0;
  size_t i = i_io;
  const __m128i vsh = _mm_cvtsi32_si128(sh);
  while (avail - used >= 256 && i >= lo + 256 && i >= (1u << 19)) {          // 256^2 / 2^20: a redo every ~16 groups at worst
    const __m512i hi = _mm512_set1_epi32((int)(uint32_t)i), lw = _mm512_set1_epi32((int)(uint32_t)(i - 256));
    int a0, a1, a2, a3;
    const bool ok = masks64(o + used, vsh, hi, lw, a0) & masks64(o + used + 64, vsh, hi, lw, a1) &
                    masks64(o + used + 128, vsh, hi, lw, a2) & masks64(o + used + 192, vsh, hi, lw, a3);
    if (ok) i -= (size_t)(a0 + a1 + a2 + a3);
    else for (int q = 0; q < 256; q += 64) step64(o + used + q, vsh, sh, i);
    used += 256;
  }
  while (avail - used >= 64 && i >= lo + 64 && i >= (1u << 12)) { step64(o + used, vsh, sh, i); used += 64; }
  while (avail - used >= 16 && i >= lo + 16 && i >= 256) { step16(o + used, vsh, sh, i); used += 16; }
  i_io = i;
  return used;
}
__attribute__((target("avx2,popcnt"))) int scan_avx2(const uint32_t* o, int avail, int sh, size_t& i_io, size_t lo) {
  int used = 0;
  size_t i = i_io;
  const __m128i vsh = _mm_cvtsi32_si128(sh);
#define le_mask(r, t) ((unsigned)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpeq_epi32(_mm256_min_epu32((r), (t)), (r)))))   /* unsigned r <= t */
  while (avail - used >= 32 && i >= lo + 32 && i >= (1u << 14)) {
    const __m256i hi = _mm256_set1_epi32((int)(uint32_t)i), lw = _mm256_set1_epi32((int)(uint32_t)(i - 32));
    unsigned mh = 0, ml = 0;
    for (int v = 0; v < 4; ++v) {
      const __m256i r = _mm256_srl_epi32(_mm256_loadu_si256((const __m256i*)(o + used + 8 * v)), vsh);
      mh |= le_mask(r, hi) << (8 * v); ml |= le_mask(r, lw) << (8 * v);
    }
    if (mh == ml) { i -= (size_t)__builtin_popcount(mh); used += 32; }
    else { for (int q = 0; q < 32; ++q) { const uint32_t r = o[used + q] >> sh; i -= (r <= (uint32_t)i); } used += 32; }
  }
  while (avail - used >= 8 && i >= lo + 8 && i >= 128) {
    const __m256i r = _mm256_srl_epi32(_mm256_loadu_si256((const __m256i*)(o + used)), vsh);
    const unsigned mh = le_mask(r, _mm256_set1_epi32((int)(uint32_t)i)), ml = le_mask(r, _mm256_set1_epi32((int)(uint32_t)(i - 8)));
    if (mh == ml) { i -= (size_t)__builtin_popcount(mh); used += 8; }
    else { for (int q = 0; q < 8; ++q) { const uint32_t r = o[used + q] >> sh; i -= (r <= (uint32_t)i); } used += 8; }
  }
#undef le_mask
  i_io = i;
  return used;
}
int pick_scan_isa() {                       // STC_PYRANDOM_ISA=0|1|2 caps the instruction set (tests walk all three paths)
  const int have = __builtin_cpu_supports("avx512f") ? 2 : __builtin_cpu_supports("avx2") ? 1 : 0;
  const char* e = getenv("STC_PYRANDOM_ISA");
  const int cap = e ? atoi(e) : 2;
  return have < cap ? have : cap;
}
const int scan_isa = pick_scan_isa();
#else
const int scan_isa = 0;
#endif
}  // namespace

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
#if defined(__x86_64__) && defined(__GNUC__)
      if (scan_isa == 2) used = scan_avx512(o, avail, sh, i, lo);
      else if (scan_isa == 1) used = scan_avx2(o, avail, sh, i, lo);
#endif
      // the vector scans stop a group short of the band edge / the end of the buffer
      used += scan_scalar(o + used, (avail - used < 64 ? avail - used : 64), sh, i, lo);
      pos += used;
    }
  }
}
