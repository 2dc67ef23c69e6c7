// Register-resident sorting networks for the per-pixel order statistics over the date axis (n <= 32 values).
//
// The first versions sorted a per-thread array by insertion (dynamic indices -> local memory, O(n^2) dependent loads and
// stores): k_static_refs / k_shadow_candidates / temporal_median_kernel spent 2-3 ms of a 24-date tile there.  A bitonic
// network of fixed size N (power of two >= n, unused slots = +inf) has compile-time indices only, so the N values live in
// registers and a sort is N/2 * log2(N) * (log2(N) + 1) / 2 compare-exchanges of two FMNMX each (N = 32: 240).
// fminf / fmaxf drop a NaN operand and cannot tell +inf padding from a real +inf, so callers keep the insertion sort for
// pixels that hold a NaN or an infinity (never the case for reflectances; the branch keeps the semantics exact).
// Equal keys: -0.0 and +0.0 may come out in either order; every consumer compares or averages them, so no result changes.
#pragma once

template <int N>
__device__ __forceinline__ void sort_net(float (&v)[N]) {
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int l = i ^ j;
        if (l > i) {
          const bool up = ((i & k) == 0);
          const float a = v[i], b = v[l];
          const float lo = fminf(a, b), hi = fmaxf(a, b);
          v[i] = up ? lo : hi;
          v[l] = up ? hi : lo;
        }
      }
    }
  }
}

// v[k] for a run-time k without turning the array into local memory
template <int N>
__device__ __forceinline__ float net_pick(const float (&v)[N], int k) {
  float r = v[0];
#pragma unroll
  for (int i = 1; i < N; ++i) r = (i == k) ? v[i] : r;
  return r;
}

// np.median of the first n entries of a sorted array (n = 0: NaN)
template <int N>
__device__ __forceinline__ float net_median(const float (&v)[N], int n) {
  if (n == 0) return nanf("");
  if (n & 1) return net_pick(v, n >> 1);
  return __fmul_rn(__fadd_rn(net_pick(v, (n >> 1) - 1), net_pick(v, n >> 1)), 0.5f);
}

__device__ __forceinline__ bool net_ok(float x) { return fabsf(x) < INFINITY; }      // false for NaN and +-inf
