// Cloud / shadow removal: remove_cloud_and_shadows (src/preprocessing/cloud_removal.py:888-973) with
// make_aligned_mosaic (:578-699), align_interp_array_randomforest (:316-575) and
// calculate_clouds_in_mosaic (:703-732).
//
// Division of labour.  Device: every array operation -- feathering, the cloud-free mosaic with its
// per-date median / std matching, snow and EVI features, order statistics (radix select), the
// float64 Gram matrices of the sampled clear pixels, the 11-variable non-negative least squares
// (Lawson-Hanson on the Gram form, the algorithm scipy.optimize.nnls itself runs), prediction and
// blending, the residual-cloud mask.  Host (this file, C++): control flow on scalar counts and the
// reference's sampling bookkeeping -- index lists per EVI stratum shuffled with Python's Mersenne
// Twister (random.shuffle), whose 625-word state the caller passes in and gets back, so a pinned
// random.seed reproduces the reference's sample exactly.
//
// Exactness.  Everything up to the regression is bit-identical to NumPy (float32 reductions follow
// NumPy's order: axis-0 sums are sequential, medians / percentiles are exact order statistics with
// NumPy's float32 lerp).  The regression is float64 with a different summation order than
// LAPACK/BLAS: filled pixels agree to ~1e-6 relative (tests/test_cloud_fill.py: rtol 1e-4).
#include "stc_common.cuh"
#include "stc_select.cuh"
#include "stc_pyrandom.h"
#include <chrono>
#include <atomic>
#include <thread>
#include <memory>
#include <sched.h>
#include <cstdio>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#define CF_MAX_DATES 32
#define NF 11          // regression features: 10 mosaic bands + snow

void maskop_dilate(stc_ctx* ctx, const unsigned char* in, unsigned char* out, int frames, int H, int W, int k, int conn, int inv_in,
                   int inv_out, int three_d);

namespace {

struct DBuf {
  void* p = nullptr;
  ~DBuf() { if (p) stc_dfree(p); }
  template <typename T> T* as() { return (T*)p; }
};

__device__ __forceinline__ void isortf(float* v, int n) {
  for (int i = 1; i < n; ++i) { float x = v[i]; int j = i - 1; while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; } v[j + 1] = x; }
}
// np.median over <= 32 values (NaN propagates)
__device__ __forceinline__ float np_median_small(float* v, int n) {
  for (int i = 0; i < n; ++i) if (isnan(v[i])) return v[i];
  isortf(v, n);
  return (n & 1) ? v[n >> 1] : __fdiv_rn(__fadd_rn(v[(n >> 1) - 1], v[n >> 1]), 2.f);
}
// NumPy's float32 linear-interpolation quantile between two neighbouring order statistics
// Position of the `pct`-th percentile among n sorted float32 values exactly as np.percentile(a, pct) computes it for a
// float32 array (NumPy 2.x): q = float32(pct) / float32(100) (percentile divides by a.dtype.type(100)), virtual index =
// (n - 1) * q in FLOAT32 (method 'linear': lambda n, q: (n - 1) * q on a 0-d float32 array), gamma = index - floor(index).
// For n in the 10^4..10^5 range the float32 index carries only 7-9 fractional bits, so gamma differs from the exact
// fraction in the 3rd-5th digit; the float64 formula used before this was 1 ulp off on ~25 % of the EVI percentiles
// (found with the T = 8, 96 x 104 case of tests/test_resegment.py: a boundary pixel changed stratum).
__device__ __host__ inline void np_quantile_pos(int n, double pct, int& lo, float& g) {
  volatile float q = (float)pct / 100.f;
  volatile float vi = (float)(n - 1) * q;
  float v = vi;
  if (v < 0.f) v = 0.f;
  const float fl = floorf(v);
  lo = (int)fl; g = v - fl;
  if (lo >= n - 1) { lo = n - 1; g = 0.f; }
}
__device__ __forceinline__ float np_lerp(float a, float b, float g) {
  float d = __fsub_rn(b, a);
  return (g >= 0.5f) ? __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.f, g))) : __fadd_rn(a, __fmul_rn(d, g));
}

// ---------------------------------------------------------------------------------------------
// mosaic stage
// ---------------------------------------------------------------------------------------------
// water0[p] = np.median_t NDWI > 0 ; divisor[p] = sum_t (1 - a_t) (sequential float32)
__global__ void __launch_bounds__(128) k_mosaic_prep(const float* __restrict__ tiles, const float* __restrict__ areas, int n, int HW,
                                                     unsigned char* __restrict__ water0, float* __restrict__ divisor) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  float v[CF_MAX_DATES];
  float d = 0.f;
  for (int t = 0; t < n; ++t) {
    const float* x = tiles + ((int64_t)t * HW + p) * 10;
    v[t] = __fdiv_rn(__fsub_rn(x[1], x[3]), __fadd_rn(x[1], x[3]));
    float om = __fsub_rn(1.f, areas[(int64_t)t * HW + p]);
    d = t ? __fadd_rn(d, om) : om;
  }
  water0[p] = np_median_small(v, n) > 0.f;
  divisor[p] = d;
}

// Reference image of every date i >= i0: mean of the OTHER dates over the pixels usable for both (:598-616), summed in date
// order.  One thread per (pixel, band) keeps the band's values of all dates in registers and forms the n ordered sums from
// them, so the cube is read once (the per-date version re-read it for every date: n^2 * HW * 40 bytes, 1.7 ms at n = 24).
// ref / flag are per-date slabs ([n][HW][10], [n][HW]).
template <int NMAX>
__global__ void __launch_bounds__(256) k_mosaic_ref(const float* __restrict__ tiles, const float* __restrict__ areas,
                                                    const unsigned char* __restrict__ water, int n, int HW, int i0,
                                                    float* __restrict__ ref_all, unsigned char* __restrict__ flag_all) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)HW * 10) return;
  const int p = (int)(e / 10), c = (int)(e - (int64_t)p * 10);
  float x[NMAX];
  unsigned use = 0, lowa = 0;                       // bit b: areas[b] < 1 (contributes), areas[b] < 0.25 (gets a reference)
#pragma unroll
  for (int b = 0; b < NMAX; ++b) {
    x[b] = 0.f;
    if (b < n) {
      x[b] = tiles[(int64_t)b * HW * 10 + e];
      const float a = areas[(int64_t)b * HW + p];
      use |= (a < 1.f) ? (1u << b) : 0u;
      lowa |= (a < 0.25f) ? (1u << b) : 0u;
    }
  }
  const bool land = !water[p];
  // the ordered sum over the other dates = (sum of the usable dates before i, shared by all later i) continued behind i
  float prefix = 0.f;
#pragma unroll
  for (int i = 0; i < NMAX; ++i) {
    if (i < n) {
      if (i >= i0) {
        const unsigned others = use & ~(1u << i);
        const bool ok = land && ((lowa >> i) & 1u) && others;
        if (ok) {
          float s = prefix;
#pragma unroll
          for (int b = i + 1; b < NMAX; ++b) if ((use >> b) & 1u) s = __fadd_rn(s, x[b]);
          ref_all[(int64_t)i * HW * 10 + e] = __fdiv_rn(s, (float)__popc(others));
        }
        if (c == 0) flag_all[(int64_t)i * HW + p] = ok;
      }
      if ((use >> i) & 1u) prefix = __fadd_rn(prefix, x[i]);
    }
  }
}

// order-preserving compaction positions of a flag image: pos[p] = rank of p among flagged pixels; total -> *count
// (single block; HW <= a few 100k)
// Three GPU-wide passes: flags per 1024-pixel chunk counted, chunk counts scanned per image, positions written (round 1
// walked every image with one block: 0.33 ms per call on 12-24 of 148 SMs).
__global__ void __launch_bounds__(1024) k_flag_chunk_count(const unsigned char* __restrict__ flag_all, int HW, int* __restrict__ chunk_cnt) {
  const int p = blockIdx.x * 1024 + threadIdx.x;
  const bool f = p < HW && flag_all[(int64_t)blockIdx.y * HW + p];
  __shared__ int tot;
  if (threadIdx.x == 0) tot = 0;
  __syncthreads();
  const unsigned bal = __ballot_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&tot, __popc(bal));
  __syncthreads();
  if (threadIdx.x == 0) chunk_cnt[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
}
__global__ void __launch_bounds__(1024) k_flag_chunk_scan(int* __restrict__ chunk_cnt /* in: counts, out: exclusive bases */, int chunks,
                                                          int* __restrict__ count) {
  int* c = chunk_cnt + (int64_t)blockIdx.x * chunks;
  __shared__ int wtot[32]; __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int c0 = 0; c0 < chunks; c0 += 1024) {
    const int i = c0 + threadIdx.x;
    const int v = i < chunks ? c[i] : 0;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) wtot[wid] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < wid; ++w) woff += wtot[w];
    if (i < chunks) c[i] = carry + woff + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) { int sum = 0; for (int w = 0; w < 32; ++w) sum += wtot[w]; carry += sum; }
    __syncthreads();
  }
  if (threadIdx.x == 0) count[blockIdx.x] = carry;
}
__global__ void __launch_bounds__(1024) k_flag_positions(const unsigned char* __restrict__ flag_all, int HW, const int* __restrict__ chunk_base,
                                                         int* __restrict__ pos_all) {
  const int p = blockIdx.x * 1024 + threadIdx.x;
  const bool f = p < HW && flag_all[(int64_t)blockIdx.y * HW + p];
  __shared__ int wtot[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, f);
  if (lane == 0) wtot[wid] = __popc(bal);
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < wid; ++w) woff += wtot[w];
  if (p < HW)
    pos_all[(int64_t)blockIdx.y * HW + p] = f ? chunk_base[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] + woff + __popc(bal & ((1u << lane) - 1u)) : -1;
}

// gather the flagged rows of date i = i0 + blockIdx.y and of its reference image into per-date slabs:
// src_rows[i] = [K_i][10], ref_rows[i] = [K_i][10]
__global__ void __launch_bounds__(256) k_gather_rows(const float* __restrict__ tiles, const float* __restrict__ ref_all,
                                                     const int* __restrict__ pos_all, int HW, int i0, float* __restrict__ src_rows_all,
                                                     float* __restrict__ ref_rows_all) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)HW * 10) return;
  const int i = i0 + blockIdx.y;
  const int64_t slab = (int64_t)i * HW * 10;
  int p = (int)(idx / 10), c = (int)(idx % 10);
  int r = pos_all[(int64_t)i * HW + p];
  if (r < 0) return;
  src_rows_all[slab + (int64_t)r * 10 + c] = tiles[slab + idx];
  ref_rows_all[slab + (int64_t)r * 10 + c] = ref_all[slab + idx];
}

// np.median / np.percentile from the (x_(k), x_(k+1)) pairs of select_ranks_dev: NumPy's even-length median is (a + b) / 2,
// its percentile the float32 lerp between the two neighbouring order statistics.
struct QSpec { int slot; int n; double q /* percentile, 0..100 */; int median; };
__global__ void __launch_bounds__(128) k_quantile_finish(const float* __restrict__ pairs, const QSpec* __restrict__ specs, int nspec,
                                                         float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nspec) return;
  const QSpec s = specs[i];
  if (s.n <= 0) { out[i] = nanf(""); return; }
  const float a = pairs[2 * s.slot];
  float b = pairs[2 * s.slot + 1];
  if (s.median) { out[i] = (s.n & 1) ? a : __fdiv_rn(__fadd_rn(a, b), 2.f); return; }
  int lo; float g;
  np_quantile_pos(s.n, s.q, lo, g);
  if (!(lo + 1 < s.n && g != 0.f)) b = a;
  out[i] = np_lerp(a, b, g);
}

// np.nanstd(rows, axis=0) of a [K][10] float32 matrix: NumPy reduces axis 0 row by row, i.e. each column is a
// SEQUENTIAL float32 sum (a dependent chain of K adds, 4 clk each at best: 2 passes x 3.8e5 rows = 1.6 ms whatever the
// machine).  One warp per column: the 32 lanes prefetch a tile of 512 rows into shared memory (double-buffered through
// registers), then every lane replays the same chain from broadcast shared-memory reads, so the loads are off the
// dependency chain.  The chain is the longest single item of the mosaic phase, so nothing else may slow it: the values come
// back as 128-bit shared-memory loads (one per four additions -- with one 32-bit load per addition the ten warps of a block
// asked the shared-memory pipe for 2.5 loads per clock and ran at 8.7 clk per row), and the 20 chains of a date are spread
// over ten blocks of two warps instead of two blocks of ten.
#define CS_TILE 512
__global__ void __launch_bounds__(64) k_col_std(const float* __restrict__ m0_all, const float* __restrict__ m1_all, int64_t slab,
                                                const int* __restrict__ Ks, int i0, float* __restrict__ out_all) {
  __shared__ __align__(16) float tile[2][CS_TILE];
  const int date = i0 + blockIdx.y;
  const int K = Ks[date];
  if (K <= 1000) return;                                   // date not aligned (:617)
  const int mat = blockIdx.x / 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = (blockIdx.x % 5) * 2 + w;
  const float* m = (mat ? m1_all : m0_all) + (int64_t)date * slab;
  float* out = out_all + (int64_t)date * 20;
  float* tl = tile[w];
  float avg = 0.f;
  for (int pass = 0; pass < 2; ++pass) {
    float s = 0.f;
    float pre[CS_TILE / 32];
#pragma unroll
    for (int j = 0; j < CS_TILE / 32; ++j) { int r = j * 32 + lane; pre[j] = (r < K) ? m[(int64_t)r * 10 + c] : 0.f; }
    for (int r0 = 0; r0 < K; r0 += CS_TILE) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < CS_TILE / 32; ++j) {
        float v = pre[j];
        if (pass) { float d = __fsub_rn(v, avg); v = __fmul_rn(d, d); }
        tl[j * 32 + lane] = v;
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < CS_TILE / 32; ++j) { int r = r0 + CS_TILE + j * 32 + lane; pre[j] = (r < K) ? m[(int64_t)r * 10 + c] : 0.f; }
      const int cnt = (K - r0) < CS_TILE ? (K - r0) : CS_TILE;
      if (cnt == CS_TILE) {
        const float4* t4 = reinterpret_cast<const float4*>(tl);
#pragma unroll 8
        for (int l = 0; l < CS_TILE / 4; ++l) {
          const float4 v = t4[l];
          s = __fadd_rn(s, v.x); s = __fadd_rn(s, v.y); s = __fadd_rn(s, v.z); s = __fadd_rn(s, v.w);
        }
      } else {
        for (int l = 0; l < cnt; ++l) s = __fadd_rn(s, tl[l]);
      }
    }
    float r = (float)((double)s / (double)K);
    if (pass == 0) avg = r; else if (lane == 0) out[mat * 10 + c] = __fsqrt_rn(r);
  }
}

// params[c] = std_ref/std_src, params[10+c] = med_ref - med_src*mult   (stats: med[0..9]=src, [10..19]=ref; std same)
__global__ void k_scale_params(const float* __restrict__ med_all, const float* __restrict__ sd_all, float* __restrict__ params_all) {
  int c = threadIdx.x;
  if (c >= 10) return;
  const float* med = med_all + blockIdx.x * 20; const float* sd = sd_all + blockIdx.x * 20; float* params = params_all + blockIdx.x * 20;
  float mult = __fdiv_rn(sd[10 + c], sd[c]);
  params[c] = mult;
  params[10 + c] = __fsub_rn(med[10 + c], __fmul_rn(med[c], mult));
}

// mosaic += (1 - area_i) * aligned(tiles_i) for the dates [i0, i1), in date order (float32 sum order) -- one pass: the
// running sum stays in a register (a launch per date re-read and re-wrote the 15 MB mosaic every time)
__global__ void __launch_bounds__(256) k_mosaic_accum(const float* __restrict__ tiles, const float* __restrict__ areas,
                                                      const unsigned char* __restrict__ water, const float* __restrict__ params_all,
                                                      int HW, int i0, int i1, float* __restrict__ mosaic) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)HW * 10) return;
  int p = (int)(idx / 10), c = (int)(idx % 10);
  const bool land = !water[p];
  float m = mosaic[idx];
  for (int i = i0; i < i1; ++i) {
    float x = tiles[(int64_t)i * HW * 10 + idx];
    const float* params = params_all + i * 20;
    if (land) x = __fadd_rn(__fmul_rn(x, params[c]), params[10 + c]);
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(1.f, areas[(int64_t)i * HW + p]), x));
  }
  mosaic[idx] = m;
}

__global__ void __launch_bounds__(128) k_mosaic_final(const float* __restrict__ tiles, const float* __restrict__ divisor, int n, int HW,
                                                      float* __restrict__ mosaic) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)HW * 10) return;
  int p = (int)(idx / 10);
  float d = divisor[p]; if (d < 0.f) d = 0.f;
  float m = __fdiv_rn(mosaic[idx], d);
  float v[CF_MAX_DATES]; float mn = INFINITY, mx = -INFINITY;
  for (int t = 0; t < n; ++t) { v[t] = tiles[(int64_t)t * HW * 10 + idx]; mn = fminf(mn, v[t]); mx = fmaxf(mx, v[t]); }
  if (isnan(m)) {
    isortf(v, n);
    int lo; float g; np_quantile_pos(n, 10.0, lo, g);
    m = np_lerp(v[lo], v[lo + 1 < n ? lo + 1 : n - 1], g);
  }
  m = fmaxf(m, mn); m = fminf(m, mx);
  mosaic[idx] = m;
}

__global__ void __launch_bounds__(256) k_fill_f(float* p, int64_t n, float v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------------------------
// per-date alignment stage
// ---------------------------------------------------------------------------------------------
// water1[p] = NDWI(np.median_t tiles) > 0 (:939)
__global__ void __launch_bounds__(128) k_water_of_median(const float* __restrict__ tiles, int n, int HW, unsigned char* __restrict__ water) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  float v[CF_MAX_DATES];
  for (int t = 0; t < n; ++t) v[t] = tiles[((int64_t)t * HW + p) * 10 + 1];
  float g = np_median_small(v, n);
  for (int t = 0; t < n; ++t) v[t] = tiles[((int64_t)t * HW + p) * 10 + 3];
  float nir = np_median_small(v, n);
  water[p] = __fdiv_rn(__fsub_rn(g, nir), __fadd_rn(g, nir)) > 0.f;
}

// counts[t] = {#a>0, #a==0, #a<1, #a==1, #(a==0 && !water)}
__global__ void __launch_bounds__(1024) k_area_counts(const float* __restrict__ areas, const unsigned char* __restrict__ water, int HW,
                                                     int* __restrict__ counts) {
  const int t = blockIdx.y; int p = blockIdx.x * blockDim.x + threadIdx.x;
  float a = p < HW ? areas[(int64_t)t * HW + p] : -1.f;
  bool in = p < HW;
  unsigned b0 = __ballot_sync(0xffffffffu, in && a > 0.f), b1 = __ballot_sync(0xffffffffu, in && a == 0.f),
           b2 = __ballot_sync(0xffffffffu, in && a < 1.f), b3 = __ballot_sync(0xffffffffu, in && a == 1.f),
           b4 = __ballot_sync(0xffffffffu, in && a == 0.f && !water[p < HW ? p : 0]);
  // block totals in shared memory, then five global atomics per block (one set per warp serialised on 5 n addresses: 0.28-0.52 ms)
  __shared__ int tot[5];
  if (threadIdx.x < 5) tot[threadIdx.x] = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    if (b0) atomicAdd(&tot[0], __popc(b0));
    if (b1) atomicAdd(&tot[1], __popc(b1));
    if (b2) atomicAdd(&tot[2], __popc(b2));
    if (b3) atomicAdd(&tot[3], __popc(b3));
    if (b4) atomicAdd(&tot[4], __popc(b4));
  }
  __syncthreads();
  if (threadIdx.x < 5 && tot[threadIdx.x]) atomicAdd(counts + t * 5 + threadIdx.x, tot[threadIdx.x]);
}

// snow probability of one pixel-date (:348-370), float32 as NumPy evaluates it
__device__ __forceinline__ float snow_prob(const float* x) {
  float ndsi = __fdiv_rn(__fsub_rn(x[1], x[8]), __fadd_rn(x[1], x[8]));
  if (ndsi < 0.10f) ndsi = 0.f;
  if (ndsi > 0.42f) ndsi = 0.42f;
  float p = __fdiv_rn(__fsub_rn(ndsi, 0.1f), 0.32f);
  if (x[3] < 0.10f) p = 0.f;
  if (x[3] > 0.35f && p > 0.f) p = 1.f;
  if (x[0] < 0.10f) p = 0.f;
  if (x[0] > 0.22f && p > 0.f) p = 1.f;
  if (__fdiv_rn(x[0], x[2]) < 0.75f) p = 0.f;
  return p;
}
// Mean snow probability over the dates (:505-511), recomputed for every fitted date because the blend of the previous date
// changed its tile.  The per-date probabilities are kept as planes sp[t][p]: k_snow_planes fills them once, the blend
// refreshes the pixels it rewrites, and the per-date mean adds n floats per pixel in date order instead of re-reading the
// whole cube (same float32 sum order, identical results).
__global__ void __launch_bounds__(256) k_snow_planes(const float* __restrict__ tiles, int n, int HW, float* __restrict__ sp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * HW) return;
  sp[i] = snow_prob(tiles + i * 10);
}
__global__ void __launch_bounds__(256) k_snow_mean(const float* __restrict__ sp, int n, int HW, float* __restrict__ snow) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  float s = 0.f;
  for (int t = 0; t < n; ++t) { float v = sp[(int64_t)t * HW + p]; s = t ? __fadd_rn(s, v) : v; }
  snow[p] = (float)((double)s / (double)n);
}

__device__ __forceinline__ float evi_of(const float* x) {
  float den = __fadd_rn(__fsub_rn(__fadd_rn(x[3], __fmul_rn(6.f, x[2])), __fmul_rn(7.5f, x[0])), 1.f);
  float e = __fmul_rn(2.5f, __fdiv_rn(__fsub_rn(x[3], x[2]), den));
  return fminf(fmaxf(e, -1.5f), 1.5f);
}
// rows of date t usable for the fit (a_t == 0 and not water), appended at row_base in pixel order:
// rowsrc[r] = t*HW + p, evi[r] = EVI(tiles[t][p])
// Clear-land pixels of date t in pixel order (:421-433).  The selection (weights == 0, not water) does not change while
// the dates are blended, so the order-preserving positions of ALL dates are computed once (k_flag_clear_land +
// k_scan_flags, one block per date) and the per-date step is a fully parallel gather of the row sources and of the EVI
// of the (already partly blended) tiles.
__global__ void __launch_bounds__(256) k_flag_clear_land(const float* __restrict__ areas, const unsigned char* __restrict__ water, int HW,
                                                         unsigned char* __restrict__ flag_all) {
  const int t = blockIdx.y; const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  flag_all[(int64_t)t * HW + p] = areas[(int64_t)t * HW + p] == 0.f && !water[p];
}
__global__ void __launch_bounds__(256) k_collect_rows(const float* __restrict__ tiles, const int* __restrict__ pos_t, int HW, int t, int row_base,
                                                      int* __restrict__ rowsrc, float* __restrict__ evi) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const int r = pos_t[p];
  if (r < 0) return;
  rowsrc[row_base + r] = t * HW + p;
  evi[row_base + r] = evi_of(tiles + ((int64_t)t * HW + p) * 10);
}

// stratum bits of every row against the six EVI percentiles b = {2,20,40,60,80,98} (:455-467)
__global__ void __launch_bounds__(256) k_strata(const float* __restrict__ evi, const float* __restrict__ b, int K, unsigned char* __restrict__ lab,
                                                int* __restrict__ counts /*[7], zeroed*/) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned char m = 0;
  if (r < K) {
    float e = evi[r];
    if (e < b[0]) m |= 1;                       // p2
    if (e < b[1]) m |= 2;                       // p20
    if (e >= b[1] && e < b[2]) m |= 4;          // p40
    if (e >= b[2] && e < b[3]) m |= 8;          // p60
    if (e >= b[3] && e < b[4]) m |= 16;         // p80
    if (e >= b[4]) m |= 32;                     // p100
    if (e >= b[5]) m |= 64;                     // p98
    lab[r] = m;
  }
  // block totals first: one global atomic per stratum and block (one per warp made 150,000 atomics on seven addresses: 33 us)
  __shared__ int tot[7];
  if (threadIdx.x < 7) tot[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    unsigned bal = __ballot_sync(0xffffffffu, (m >> k) & 1);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&tot[k], __popc(bal));
  }
  __syncthreads();
  if (threadIdx.x < 7 && tot[threadIdx.x]) atomicAdd(counts + threadIdx.x, tot[threadIdx.x]);
}
// Index lists of the seven strata of one date, rows in ascending order (np.argwhere), the two 2 % tails repeated ten
// times per row (np.repeat(.., 10), :468-471).  Order-preserving compaction in three GPU-wide steps (the first version
// walked each list with one block: 84-168 blocks of serial 1024-row rounds, 0.5-0.9 ms): per-block stratum counts, a scan
// of the block counts per (date, stratum), then every block places its rows.
struct BucketJob { const unsigned char* lab; int K; int* out[7]; int* chunk; /* [blocks][7]: counts, then exclusive bases */ };
#define BL_ROWS 1024                     // rows per block: 256 threads x 4 consecutive rows
__device__ __forceinline__ void bucket_thread_counts(const BucketJob& j, int r0, unsigned char (&lab)[4], int (&c)[7]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) lab[q] = (r0 + q < j.K) ? j.lab[r0 + q] : (unsigned char)0;
#pragma unroll
  for (int k = 0; k < 7; ++k) c[k] = ((lab[0] >> k) & 1) + ((lab[1] >> k) & 1) + ((lab[2] >> k) & 1) + ((lab[3] >> k) & 1);
}
__global__ void __launch_bounds__(256) k_bucket_count(const BucketJob* __restrict__ jobs) {
  const BucketJob j = jobs[blockIdx.y];
  if ((int64_t)blockIdx.x * BL_ROWS >= j.K) return;
  unsigned char lab[4]; int c[7];
  bucket_thread_counts(j, blockIdx.x * BL_ROWS + threadIdx.x * 4, lab, c);
  __shared__ int tot[7];
  if (threadIdx.x < 7) tot[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    int v = c[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&tot[k], v);
  }
  __syncthreads();
  if (threadIdx.x < 7) j.chunk[blockIdx.x * 7 + threadIdx.x] = tot[threadIdx.x];
}
__global__ void __launch_bounds__(224) k_bucket_scan(const BucketJob* __restrict__ jobs) {      // one warp per (date, stratum)
  const BucketJob j = jobs[blockIdx.x];
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blocks = (j.K + BL_ROWS - 1) / BL_ROWS;
  int running = 0;
  for (int c0 = 0; c0 < blocks; c0 += 32) {
    const int v = (c0 + lane < blocks) ? j.chunk[(c0 + lane) * 7 + k] : 0;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (c0 + lane < blocks) j.chunk[(c0 + lane) * 7 + k] = running + incl - v;
    running += __shfl_sync(0xffffffffu, incl, 31);
  }
}
__global__ void __launch_bounds__(256) k_bucket_scatter(const BucketJob* __restrict__ jobs) {
  const BucketJob j = jobs[blockIdx.y];
  if ((int64_t)blockIdx.x * BL_ROWS >= j.K) return;
  const int r0 = blockIdx.x * BL_ROWS + threadIdx.x * 4;
  unsigned char lab[4]; int c[7];
  bucket_thread_counts(j, r0, lab, c);
  __shared__ int wtot[8][7];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int excl[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    int incl = c[k];
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    excl[k] = incl - c[k];
    if (lane == 31) wtot[wid][k] = incl;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    if (!c[k]) continue;
    int pos = j.chunk[blockIdx.x * 7 + k] + excl[k];
    for (int w = 0; w < wid; ++w) pos += wtot[w][k];
    const int rep = (k == 0 || k == 6) ? 10 : 1;
    int* out = j.out[k];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if ((lab[q] >> k) & 1) { for (int t = 0; t < rep; ++t) out[pos * rep + t] = r0 + q; ++pos; }
  }
}

// float64 Gram sums over the sampled rows.  Features u_j = [mosaic bands, snow], c_j = clip(u_j, 0.005, 1)
// for j < 10 (c_10 = u_10), targets y_b = tile bands of the row's (date, pixel).
// gram layout (583 doubles): UU[11][11], CU[11][11] (c_j * u_k), CC[11][11], UY[11][10], CY[11][10].
//
// k_gram_rows follows the sample -> row -> (date, pixel) indirection once, with the whole GPU, and writes the row
// z = [u0..u10 | c0..c9 | y0..y9 | 0] (32 floats) of every sampled row.  k_gram then forms z_a * z_b for a < 24, all b, from
// register tiles (5 shared-memory loads per 12 FMAs; the round-1 kernel spent its time on two loads and two float->double
// conversions per FMA: 206 us per date).  Every entry is still accumulated in row order over its partial's 64-row groups
// (group g of partial b = rows [64 (b + g P), +64)), and the P partials are added in order, so the sums do not depend on
// the schedule and are the ones the first version produced.
#define GRAM_N (3 * NF * NF + 2 * NF * 10)
#define GRAM_ROWS 64
#define GRAM_ZW 32                      // floats per staged row
#define GRAM_PA 24                      // a-side entries computed per row (21 needed)
#define GRAM_P (GRAM_PA * GRAM_ZW)      // products per partial
// Everything of a row except the snow feature is known before the per-date loop starts (the rows are clear land, weights == 0,
// which no blend touches; the mosaic is final), so this gather runs AHEAD on the sample stream; k_gram_snow then drops the one
// value that depends on the blends of the earlier dates into slot 10.
__global__ void __launch_bounds__(256) k_gram_rows(const float* __restrict__ tiles, const float* __restrict__ mosaic,
                                                   const int* __restrict__ rowsrc, const int* __restrict__ sample, int S, int HW,
                                                   float* __restrict__ Z /*[S][32]*/, int* __restrict__ pix /*[S]*/) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t s = e >> 5;
  const int f = (int)(e & 31);
  if (s >= S) return;
  const int src = rowsrc[sample[s]];
  const int p = src % HW;
  float v = 0.f;
  if (f < 10) v = mosaic[(int64_t)p * 10 + f];
  else if (f == 10) pix[s] = p;
  else if (f < 21) v = fminf(fmaxf(mosaic[(int64_t)p * 10 + (f - 11)], 0.005f), 1.f);
  else if (f < 31) v = tiles[(int64_t)src * 10 + (f - 21)];
  Z[e] = v;
}
__global__ void __launch_bounds__(256) k_gram_snow(const float* __restrict__ snow, const int* __restrict__ pix, int S, float* __restrict__ Z) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < S) Z[(int64_t)s * GRAM_ZW + 10] = snow[pix[s]];
}
// One partial sum = two warps (a-rows 0..11 and 12..23 of the 24 x 32 products, a 3 x 4 register tile per lane) sharing the
// staged rows; two partials per block.  The pair synchronises on a named barrier, rows are double-buffered in shared memory and
// the next sub-batch is already in registers while this one is multiplied.
__global__ void __launch_bounds__(128) k_gram(const float* __restrict__ Z, int S, int nparts, double* __restrict__ partial /*[nparts][GRAM_P]*/) {
  __shared__ __align__(16) double zs[2][2][32][GRAM_ZW];
  const int q = threadIdx.x >> 6, t64 = threadIdx.x & 63, h = t64 >> 5, lane = threadIdx.x & 31;
  const int part = blockIdx.x * 2 + q;
  if (part >= nparts) return;                       // both warps of the pair leave together: the pair's barrier is never half-used
  const int a0 = 12 * h + (lane >> 3) * 3, b0 = (lane & 7) * 4;
  double acc[3][4];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  float4 pre[4];
  auto rows_of = [&](int t, int64_t& s) { s = ((int64_t)part + (int64_t)(t >> 1) * nparts) * GRAM_ROWS + (t & 1) * 32; const int64_t left = (int64_t)S - s; return (int)(left < 32 ? (left < 0 ? 0 : left) : 32); };
  auto fetch = [&](int64_t s, int rr) {
    const float4* src = reinterpret_cast<const float4*>(Z + s * GRAM_ZW);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int i = t64 + 64 * k; pre[k] = (i < rr * 8) ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f); }
  };
  int t = 0, buf = 0; int64_t s; int rr = rows_of(0, s);
  if (rr > 0) fetch(s, rr);
  while (rr > 0) {
    double (*zz)[GRAM_ZW] = zs[q][buf];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = t64 + 64 * k, r = i >> 3, c = (i & 7) * 4;
      *reinterpret_cast<double2*>(&zz[r][c]) = make_double2((double)pre[k].x, (double)pre[k].y);
      *reinterpret_cast<double2*>(&zz[r][c + 2]) = make_double2((double)pre[k].z, (double)pre[k].w);
    }
    asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory");
    const int cur = rr;
    ++t; rr = rows_of(t, s);
    if (rr == 0 && (t & 1)) { ++t; rr = rows_of(t, s); }
    if (rr > 0) fetch(s, rr);
#pragma unroll 4
    for (int r = 0; r < cur; ++r) {
      const double a[3] = {zz[r][a0], zz[r][a0 + 1], zz[r][a0 + 2]};
      const double2 b01 = *reinterpret_cast<const double2*>(&zz[r][b0]), b23 = *reinterpret_cast<const double2*>(&zz[r][b0 + 2]);
      const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);       // the product of two floats is exact in double
    }
    buf ^= 1;
  }
  double* out = partial + (int64_t)part * GRAM_P;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[(a0 + i) * GRAM_ZW + b0 + j] = acc[i][j];
}
// gram[q] = sum over the partials, in order; q -> (a, b) of the staged row (c_10 is u_10).  Eight lanes per entry: each
// loads its 37 consecutive partials at once (all loads in flight), then the running sum walks through the eight lanes in
// order -- the same chain of additions as a serial loop, without waiting for 296 loads one batch after the other.
#define GRAM_PARTS 296
#define GRAM_PER ((GRAM_PARTS + 7) / 8)
__global__ void __launch_bounds__(256) k_gram_reduce(const double* __restrict__ partial, int nparts, double* __restrict__ gram) {
  const int q = blockIdx.x * 32 + (threadIdx.x >> 3), sub = threadIdx.x & 7, lane = threadIdx.x & 31;
  const bool valid = q < GRAM_N;
  const int qq = valid ? q : 0;
  int kind, j, k;
  if (qq < 3 * NF * NF) { kind = qq / (NF * NF); j = (qq % (NF * NF)) / NF; k = qq % NF; }
  else { const int r = qq - 3 * NF * NF; kind = 3 + r / (NF * 10); j = (r % (NF * 10)) / 10; k = r % 10; }
  const int cj = j < 10 ? 11 + j : 10, ck = k < 10 ? 11 + k : 10;      // position of c_j / c_k in the staged row
  int a, b;
  switch (kind) {
    case 0: a = j; b = k; break;                 // u_j u_k
    case 1: a = cj; b = k; break;                // c_j u_k
    case 2: a = cj; b = ck; break;               // c_j c_k
    case 3: a = j; b = 21 + k; break;            // u_j y_k
    default: a = cj; b = 21 + k; break;          // c_j y_k
  }
  const double* src = partial + a * GRAM_ZW + b;
  const int p0 = sub * GRAM_PER, p1 = min(p0 + GRAM_PER, nparts);
  double v[GRAM_PER];
#pragma unroll
  for (int i = 0; i < GRAM_PER; ++i) v[i] = (p0 + i < p1) ? src[(int64_t)(p0 + i) * GRAM_P] : 0.0;
  double s = 0.0;
#pragma unroll
  for (int turn = 0; turn < 8; ++turn) {
    if (sub == turn) {
#pragma unroll
      for (int i = 0; i < GRAM_PER; ++i) if (p0 + i < p1) s += v[i];
    }
    s = __shfl_sync(0xffffffffu, s, (lane & ~7) + turn);
  }
  if (sub == 0 && valid) gram[q] = s;
}

// Lawson-Hanson NNLS on the normal equations, as scipy.optimize.nnls does (AtA, Atb in float64,
// tol = 10*max(m,n)*eps, maxiter = 3n).  Band b's design matrix has bands < b clipped (the reference
// clips column b in place AFTER copying train_x for band b, :527-540).  One thread per band.
__device__ bool solve_sym(const double* G, const double* r, const bool* P, double* s) {
  int id[NF], m = 0;
  for (int i = 0; i < NF; ++i) { if (P[i]) id[m++] = i; s[i] = 0.0; }
  double A[NF][NF + 1];
  for (int a = 0; a < m; ++a) { for (int b = 0; b < m; ++b) A[a][b] = G[id[a] * NF + id[b]]; A[a][m] = r[id[a]]; }
  for (int col = 0; col < m; ++col) {
    int piv = col; double best = fabs(A[col][col]);
    for (int a = col + 1; a < m; ++a) if (fabs(A[a][col]) > best) { best = fabs(A[a][col]); piv = a; }
    if (best == 0.0) return false;
    if (piv != col) for (int b = col; b <= m; ++b) { double t = A[col][b]; A[col][b] = A[piv][b]; A[piv][b] = t; }
    for (int a = col + 1; a < m; ++a) {
      double f = A[a][col] / A[col][col];
      for (int b = col; b <= m; ++b) A[a][b] -= f * A[col][b];
    }
  }
  for (int a = m - 1; a >= 0; --a) {
    double t = A[a][m];
    for (int b = a + 1; b < m; ++b) t -= A[a][b] * s[id[b]];
    s[id[a]] = t / A[a][a];
  }
  return true;
}
__global__ void k_nnls_serial(const double* __restrict__ gram, int S, double* __restrict__ coef /*[10][NF]*/, int* __restrict__ status) {
  const int band = threadIdx.x;
  if (band >= 10) return;
  const double *UU = gram, *CU = gram + NF * NF, *CC = gram + 2 * NF * NF, *UY = gram + 3 * NF * NF, *CY = UY + NF * 10;
  double G[NF * NF], r[NF];
  for (int j = 0; j < NF; ++j) {
    const bool cj = j < band;
    r[j] = cj ? CY[j * 10 + band] : UY[j * 10 + band];
    for (int k = 0; k < NF; ++k) {
      const bool ck = k < band;
      G[j * NF + k] = (cj && ck) ? CC[j * NF + k] : cj ? CU[j * NF + k] : ck ? CU[k * NF + j] : UU[j * NF + k];
    }
  }
  const double tol = 10.0 * (double)(S > NF ? S : NF) * 2.220446049250313e-16;
  const int maxiter = 3 * NF;
  double x[NF], s[NF], w[NF]; bool P[NF];
  for (int j = 0; j < NF; ++j) { x[j] = 0.0; s[j] = 0.0; w[j] = r[j]; P[j] = false; }
  int iter = 0, st = 1;
  while (true) {
    bool allP = true, any = false;
    for (int j = 0; j < NF; ++j) { if (!P[j]) { allP = false; if (w[j] > tol) any = true; } }
    if (allP || !any) break;
    int kbest = 0; double best = -INFINITY;
    for (int j = 0; j < NF; ++j) { double v = P[j] ? 0.0 : w[j]; if (v > best) { best = v; kbest = j; } }   // argmax(w * ~P)
    P[kbest] = true;
    if (!solve_sym(G, r, P, s)) { st = -2; break; }
    while (iter < maxiter) {
      double mn = INFINITY;
      for (int j = 0; j < NF; ++j) if (P[j] && s[j] < mn) mn = s[j];
      if (!(mn < 0)) break;
      ++iter;
      double alpha = INFINITY;
      for (int j = 0; j < NF; ++j) if (P[j] && s[j] < 0) { double a = x[j] / (x[j] - s[j]); if (a < alpha) alpha = a; }
      for (int j = 0; j < NF; ++j) { x[j] *= (1 - alpha); x[j] += alpha * s[j]; }
      for (int j = 0; j < NF; ++j) if (x[j] <= tol) P[j] = false;
      if (!solve_sym(G, r, P, s)) { st = -2; break; }
    }
    if (st < 0) break;
    for (int j = 0; j < NF; ++j) { x[j] = s[j]; }
    for (int j = 0; j < NF; ++j) { double t = r[j]; for (int k = 0; k < NF; ++k) t -= G[j * NF + k] * x[k]; w[j] = t; }
    if (iter == maxiter) { st = -1; break; }
  }
  for (int j = 0; j < NF; ++j) coef[band * NF + j] = x[j];
  status[band] = st;
}

// The same algorithm, one WARP per band: the O(m^3) elimination of every solve runs over the lanes (each element update is
// the same fused multiply-add with the same operands, so the bits do not change), the O(m) bookkeeping is done by lane j
// for entry j with exact warp reductions (min / max / ballot), and only the back substitution -- a sum whose order
// matters -- stays on one lane.  107 -> ~30 us per date; k_nnls_serial above is kept as the A/B reference
// (STC_NNLS_SERIAL=1, tests/test_cloud_fill.py).
struct NnlsShared { double G[NF * NF], A[NF][NF + 1], r[NF], x[NF], s[NF], f[NF]; int id[NF]; };
__device__ bool solve_sym_warp(NnlsShared& sh, unsigned P, int lane) {
  const int m = __popc(P);
  if (lane < NF) sh.s[lane] = 0.0;
  if (lane < m) sh.id[lane] = __fns(P, 0, lane + 1);
  __syncwarp();
  for (int e = lane; e < m * (m + 1); e += 32) {
    const int a = e / (m + 1), b = e - a * (m + 1);
    sh.A[a][b] = (b < m) ? sh.G[sh.id[a] * NF + sh.id[b]] : sh.r[sh.id[a]];
  }
  __syncwarp();
  for (int col = 0; col < m; ++col) {
    int piv = col; double best = fabs(sh.A[col][col]);                  // every lane: the same <= 11 broadcast reads
    for (int a = col + 1; a < m; ++a) { const double v = fabs(sh.A[a][col]); if (v > best) { best = v; piv = a; } }
    if (best == 0.0) return false;
    if (piv != col && lane >= col && lane <= m) { const double t = sh.A[col][lane]; sh.A[col][lane] = sh.A[piv][lane]; sh.A[piv][lane] = t; }
    __syncwarp();
    if (lane > col && lane < m) sh.f[lane] = sh.A[lane][col] / sh.A[col][col];
    __syncwarp();
    const int wdt = m - col;                                            // columns col+1 .. m (column col is never read again)
    for (int e = lane; e < (m - col - 1) * wdt; e += 32) {
      const int a = col + 1 + e / wdt, b = col + 1 + e % wdt;
      sh.A[a][b] = fma(-sh.f[a], sh.A[col][b], sh.A[a][b]);
    }
    __syncwarp();
  }
  if (lane == 0) {
    for (int a = m - 1; a >= 0; --a) {
      double t = sh.A[a][m];
      for (int b = a + 1; b < m; ++b) t = fma(-sh.A[a][b], sh.s[sh.id[b]], t);
      sh.s[sh.id[a]] = t / sh.A[a][a];
    }
  }
  __syncwarp();
  return true;
}
__global__ void __launch_bounds__(320) k_nnls(const double* __restrict__ gram, int S, double* __restrict__ coef /*[10][NF]*/, int* __restrict__ status) {
  __shared__ NnlsShared shs[10];
  const int band = threadIdx.x >> 5, lane = threadIdx.x & 31;
  NnlsShared& sh = shs[band];
  const double *UU = gram, *CU = gram + NF * NF, *CC = gram + 2 * NF * NF, *UY = gram + 3 * NF * NF, *CY = UY + NF * 10;
  for (int e = lane; e < NF * NF; e += 32) {
    const int j = e / NF, k = e - j * NF;
    const bool cj = j < band, ck = k < band;
    sh.G[e] = (cj && ck) ? CC[j * NF + k] : cj ? CU[j * NF + k] : ck ? CU[k * NF + j] : UU[j * NF + k];
  }
  const bool mine = lane < NF;
  double r = 0.0, x = 0.0, w = 0.0;
  if (mine) { r = (lane < band) ? CY[lane * 10 + band] : UY[lane * 10 + band]; sh.r[lane] = r; sh.x[lane] = 0.0; w = r; }
  __syncwarp();
  const double tol = 10.0 * (double)(S > NF ? S : NF) * 2.220446049250313e-16;
  const int maxiter = 3 * NF;
  const unsigned ALL = (1u << NF) - 1u, FULL = 0xffffffffu;
  unsigned P = 0;
  int iter = 0, st = 1;
  while (true) {
    if (P == ALL) break;
    const bool free_j = mine && !((P >> lane) & 1u);
    if (!__any_sync(FULL, free_j && w > tol)) break;
    // argmax over j of (P[j] ? 0 : w[j]), first maximum
    const double v = mine ? (free_j ? w : 0.0) : -INFINITY;
    double mx = v;
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
    const unsigned hit = __ballot_sync(FULL, mine && v == mx);
    const int kbest = hit ? (__ffs(hit) - 1) : 0;
    P |= 1u << kbest;
    if (!solve_sym_warp(sh, P, lane)) { st = -2; break; }
    while (iter < maxiter) {
      const bool in_p = mine && ((P >> lane) & 1u);
      const double sj = mine ? sh.s[lane] : 0.0;
      if (!__any_sync(FULL, in_p && sj < 0)) break;                     // min over P of s is not negative
      ++iter;
      double al = (in_p && sj < 0) ? x / (x - sj) : INFINITY;
      for (int o = 16; o > 0; o >>= 1) al = fmin(al, __shfl_xor_sync(FULL, al, o));
      if (mine) { x *= (1 - al); x = fma(al, sj, x); }
      P &= ~__ballot_sync(FULL, mine && x <= tol);
      if (!solve_sym_warp(sh, P, lane)) { st = -2; break; }
    }
    if (st < 0) break;
    if (mine) { x = sh.s[lane]; sh.x[lane] = x; }
    __syncwarp();
    if (mine) { double t = r; for (int k = 0; k < NF; ++k) t = fma(-sh.G[lane * NF + k], sh.x[k], t); w = t; }
    if (iter == maxiter) { st = -1; break; }
  }
  if (mine) coef[band * NF + lane] = x;
  if (lane == 0) status[band] = st;
}

// tiles[d] = tiles[d]*(1-a) + fill*a with fill = regression prediction from [mosaic, snow] (use_coef) or the mosaic itself
__global__ void __launch_bounds__(256) k_predict_blend(float* __restrict__ tiles_d, const float* __restrict__ area_d,
                                                       const float* __restrict__ mosaic, const float* __restrict__ snow,
                                                       const double* __restrict__ coef, int use_coef, int HW, float* __restrict__ sp_d) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const float a = area_d[p];
  if (!(a > 0.f)) return;            // a == 0: tiles*1 + 0*0 leaves the pixel unchanged
  float f[NF];
#pragma unroll
  for (int j = 0; j < 10; ++j) f[j] = mosaic[(int64_t)p * 10 + j];
  f[10] = snow ? snow[p] : 0.f;
  const float om = __fsub_rn(1.f, a);
  for (int b = 0; b < 10; ++b) {
    float fill;
    if (use_coef) {
      double t = 0.0;
      for (int j = 0; j < NF; ++j) t += (double)f[j] * coef[b * NF + j];
      fill = (float)t;
    } else fill = f[b];
    float* x = tiles_d + (int64_t)p * 10 + b;
    *x = __fadd_rn(__fmul_rn(*x, om), __fmul_rn(fill, a));
  }
  if (sp_d) sp_d[p] = snow_prob(tiles_d + (int64_t)p * 10);
}

// ---------------------------------------------------------------------------------------------
// residual clouds in the mosaic (:703-732)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_only_one(const float* __restrict__ areas, const unsigned char* __restrict__ pf, int n, int HW,
                                                  unsigned char* __restrict__ only1, int* __restrict__ count) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned char v = 0;
  if (p < HW) {
    int clear = 0;
    for (int t = 0; t < n; ++t) clear += !(areas[(int64_t)t * HW + p] > 0.f);
    v = (clear < 2) || pf[p];
    only1[p] = v;
  }
  unsigned bal = __ballot_sync(0xffffffffu, v != 0);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(count, __popc(bal));
}
// compact mosaic blue / red of the multi-image pixels (~only1) in pixel order
__global__ void __launch_bounds__(256) k_gather_br(const float* __restrict__ mosaic, const int* __restrict__ pos, int HW,
                                                   float* __restrict__ blue, float* __restrict__ red) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  int r = pos[p];
  if (r < 0) return;
  blue[r] = mosaic[(int64_t)p * 10]; red[r] = mosaic[(int64_t)p * 10 + 2];
}
__global__ void __launch_bounds__(256) k_not(const unsigned char* in, unsigned char* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = !in[i];
}
__global__ void __launch_bounds__(256) k_mosaic_clouds(const float* __restrict__ mosaic, const unsigned char* __restrict__ only1,
                                                       const unsigned char* __restrict__ pf, const float* __restrict__ refs, int HW,
                                                       unsigned char* __restrict__ out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const float* m = mosaic + (int64_t)p * 10;
  bool c = (m[0] > refs[0]) && (m[2] > refs[1]) && only1[p] && (__fadd_rn(__fadd_rn(m[0], m[1]), m[2]) < 1.f);
  out[p] = c && !pf[p];
}
__global__ void __launch_bounds__(256) k_add_clip(float* __restrict__ areas, const unsigned char* __restrict__ c, int n, int HW) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * HW) return;
  float v = __fadd_rn(areas[i], c[i % HW] ? 1.f : 0.f);
  areas[i] = v > 1.f ? 1.f : v;
}

}  // namespace

namespace {
__global__ void __launch_bounds__(256) k_count_zero(const unsigned char* __restrict__ m, int n, int* __restrict__ count) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned bal = __ballot_sync(0xffffffffu, p < n && m[p] == 0);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(count, __popc(bal));
}
}  // namespace

#define CF_LAUNCH(kern, grid, block, ...) do { TraceScope ts_(ctx, #kern); kern<<<(grid), (block), 0, ctx->stream>>>(__VA_ARGS__); ctx->launches++; } while (0)
#define CF_SYNC() STC_CUDA(cudaStreamSynchronize(ctx->stream))

// pos[img][p] = rank of pixel p among the flagged pixels of image img (row-major), -1 when not flagged; count[img] = flagged pixels
static int scan_flags_dev(stc_ctx* ctx, const unsigned char* flags, int nimg, int HW, int* pos, int* count) {
  const int chunks = cdiv(HW, 1024);
  PoolBuf cc;
  STC_CUDA(cc.alloc((size_t)nimg * chunks * 4));
  CF_LAUNCH(k_flag_chunk_count, dim3(chunks, nimg), 1024, flags, HW, cc.as<int>());
  CF_LAUNCH(k_flag_chunk_scan, nimg, 1024, cc.as<int>(), chunks, count);
  CF_LAUNCH(k_flag_positions, dim3(chunks, nimg), 1024, flags, HW, cc.as<int>(), pos);
  return STC_OK;
}

namespace {
__global__ void __launch_bounds__(256) k_clip01(float* __restrict__ x, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float v = x[i]; x[i] = isnan(v) ? v : fminf(fmaxf(v, 0.f), 1.f); }     // np.clip keeps NaN
}
}  // namespace

// Worker threads for the shuffle replay: STC_HOST_THREADS, else this process' share of the cores it may run on (the ranks
// of one node divide them: LOCAL_WORLD_SIZE, set by torchrun) minus the three threads that are busy anyway (the caller, the
// generator walk and its producer); 2..12.  The shuffles are ~60 ms of CPU time for a 24-date tile, and a date's device
// work cannot be enqueued before its sample is shuffled, so too few workers put the host on the critical path.
static int host_threads() {
  if (const char* e = getenv("STC_HOST_THREADS")) { const int v = atoi(e); if (v >= 1) return std::min(v, 64); }
  cpu_set_t set; CPU_ZERO(&set);
  int cores = 8;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
  int ranks = 1;
  if (const char* e = getenv("LOCAL_WORLD_SIZE")) { const int v = atoi(e); if (v >= 1) ranks = v; }
  return std::max(2, std::min(cores / ranks - 3, 12));
}

// Device-resident core: tiles [n,H,W,10] float32 (blended in place), probs [n,H,W] float32, pfcps [>= H*W] uint8 (date 0 is read),
// areas [n,H,W] float32 out, mosaic_out [H,W,10] optional -- all device pointers; mt_state / to_remove / clipped_out on the host.
int remove_clouds_dev(stc_ctx* ctx, float* tiles, const float* probs_dev, const unsigned char* pfcps_dev, int n, int H,
                      int W, uint32_t* mt_state, float* areas, int32_t* to_remove_host, float* mosaic_out_dev,
                      int clip_when_all_kept, int32_t* clipped_out) {
  if (!ctx) return STC_ERR_ARG;
  if (clipped_out) *clipped_out = 0;
  if (!tiles || !probs_dev || !pfcps_dev || !mt_state || !areas || !to_remove_host || n < 1 || n > CF_MAX_DATES || H < 3 ||
      W < 3 || mt_state[624] > 624)
    STC_FAIL(STC_ERR_ARG, "remove_clouds: bad argument (1 <= n <= 32, MT19937 state of 624 words + position)");
  const int HW = H * W; const int64_t N = (int64_t)n * HW;
  // STC_CF_TIMING=1: wall time of each phase on stderr (the marks synchronise the stream)
  static const bool cf_timing = getenv("STC_CF_TIMING") != nullptr;
  auto cf_now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double cf_t = cf_now();
  auto cf_mark = [&](const char* what) {
    if (!cf_timing) return;
    cudaStreamSynchronize(ctx->stream);
    const double t = cf_now();
    fprintf(stderr, "[remove_clouds] %-34s %8.1f ms\n", what, t - cf_t);
    cf_t = t;
  };
  DBuf d_ta, d_tb, d_sums, d_water0, d_water1, d_flag, d_u8a, d_u8b, d_pf, d_ref, d_pos, d_src_rows,
      d_ref_rows, d_mosaic, d_div, d_snow, d_partial, d_gramz, d_pix, d_gram, d_coef, d_status, d_qout,
      d_sd, d_params, d_cnt, d_counts;
  const int gram_blocks = GRAM_PARTS;
  STC_CUDA(stc_dmalloc(&d_ta.p, N * 4)); STC_CUDA(stc_dmalloc(&d_tb.p, N * 4)); STC_CUDA(stc_dmalloc(&d_sums.p, CF_MAX_DATES * 4));
  for (DBuf* b : {&d_water0, &d_water1, &d_flag, &d_u8a, &d_u8b, &d_pf}) STC_CUDA(stc_dmalloc(&b->p, HW));
  STC_CUDA(stc_dmalloc(&d_ref.p, (int64_t)HW * 40)); STC_CUDA(stc_dmalloc(&d_pos.p, (int64_t)HW * 4));
  STC_CUDA(stc_dmalloc(&d_src_rows.p, (int64_t)HW * 40)); STC_CUDA(stc_dmalloc(&d_ref_rows.p, (int64_t)HW * 40));
  STC_CUDA(stc_dmalloc(&d_mosaic.p, (int64_t)HW * 40)); STC_CUDA(stc_dmalloc(&d_div.p, (int64_t)HW * 4)); STC_CUDA(stc_dmalloc(&d_snow.p, (int64_t)HW * 4));
  STC_CUDA(stc_dmalloc(&d_partial.p, (size_t)gram_blocks * GRAM_P * 8)); STC_CUDA(stc_dmalloc(&d_gram.p, GRAM_N * 8));
  STC_CUDA(stc_dmalloc(&d_coef.p, 10 * NF * 8)); STC_CUDA(stc_dmalloc(&d_status.p, 64));
  STC_CUDA(stc_dmalloc(&d_qout.p, 32 * 4));
  STC_CUDA(stc_dmalloc(&d_sd.p, 32 * 4)); STC_CUDA(stc_dmalloc(&d_params.p, 32 * 4)); STC_CUDA(stc_dmalloc(&d_cnt.p, 64));
  STC_CUDA(stc_dmalloc(&d_counts.p, CF_MAX_DATES * 5 * 4));
  float* mosaic = d_mosaic.as<float>();
  unsigned char *water0 = d_water0.as<unsigned char>(), *water1 = d_water1.as<unsigned char>(), *flag = d_flag.as<unsigned char>(),
                *u8a = d_u8a.as<unsigned char>(), *u8b = d_u8b.as<unsigned char>(), *pf = d_pf.as<unsigned char>();
  int* cnt = d_cnt.as<int>();

  // out_dev[i] = the median / percentile described by specs[i]; every (job, column) slot serves one spec.  Asynchronous.
  auto run_select = [&](const std::vector<SelJob>& jobs, const std::vector<QSpec>& specs, float* out_dev) -> int {
    const int nj = (int)jobs.size(), ns = (int)specs.size();
    std::vector<int> ks((size_t)nj * SEL_MAX_COLS, 0);
    for (const QSpec& sp : specs) {
      if (sp.n <= 0) continue;
      if (sp.median) ks[sp.slot] = (sp.n - 1) / 2;
      else { int lo; float g; np_quantile_pos(sp.n, sp.q, lo, g); ks[sp.slot] = lo; }
    }
    DBuf d_ks, d_pairs, d_specs;
    STC_CUDA(stc_dmalloc(&d_ks.p, ks.size() * 4)); STC_CUDA(stc_dmalloc(&d_pairs.p, ks.size() * 8)); STC_CUDA(stc_dmalloc(&d_specs.p, (size_t)ns * sizeof(QSpec)));
    const void* hk = ctx_stage(ctx, ks.data(), ks.size() * 4);
    const void* hs = ctx_stage(ctx, specs.data(), (size_t)ns * sizeof(QSpec));
    if (!hk || !hs) STC_FAIL(STC_ERR_NOMEM, "remove_clouds: pinned staging");
    STC_CUDA(cudaMemcpyAsync(d_ks.p, hk, ks.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    STC_CUDA(cudaMemcpyAsync(d_specs.p, hs, (size_t)ns * sizeof(QSpec), cudaMemcpyHostToDevice, ctx->stream));
    int rcs = select_ranks_dev(ctx, jobs.data(), nj, d_ks.as<int>(), d_pairs.as<float>());
    if (rcs) return rcs;
    CF_LAUNCH(k_quantile_finish, cdiv(ns, 128), 128, d_pairs.as<float>(), d_specs.as<QSpec>(), ns, out_dev);
    return STC_OK;
  };
  int rc;

  cf_mark("alloc + upload");
  // ---- 1. feather the masks (:908-921, closing size 20) ----
  if ((rc = pre_feather_dev(ctx, probs_dev, n, H, W, 20, d_ta.as<float>(), d_tb.as<float>(), d_sums.as<float>(), areas))) return rc;

  cf_mark("feather");
  // ---- state of the per-date fits (phase 3), declared here because phase 3a may run inside the mosaic phase (see below) ----
  struct FitJob { int d, lo, hi, K; int64_t row0; int cnt[7]; int64_t list0[7]; };
  struct ShufTask { int j; int k; uint32_t mt[624]; int idx; };            // k in 0..6: list k; k == 7: the sample
  std::vector<int> counts(n * 5);
  PyRandomProducer rng_ahead;                     // generates the walk's outputs on its own thread (stc_pyrandom.h); outlives rng's last use
  PyRandom rng; rng.import_state(mt_state, (int)mt_state[624]);
  std::vector<FitJob> fits;
  std::vector<int> fit_of(n, -1);
  int64_t total_rows = 0;
  int nf = 0;
  DBuf d_rows_all, d_evi_all, d_lab_all, d_fqout, d_fcnt, d_lists, d_bjobs, d_bchunk, d_sample_all, d_status_all, d_flag3, d_pos3, d_K3;
  int* pin_lists = nullptr; int* pin_samples = nullptr;
  std::vector<int64_t> sample0;
  std::vector<ShufTask> tasks;
  std::atomic<size_t> published{0};
  std::vector<size_t> S_of, zoff;
  double tt0 = 0, t_skip = 0;
  long long n_draw = 0;
  std::unique_ptr<std::atomic<int>[]> lists_done, sample_done;
  std::atomic<size_t> next_task{0};
  std::vector<std::thread> pool;
  const int nthreads = host_threads();
  struct Joiner { std::vector<std::thread>& p; ~Joiner() { for (auto& t : p) if (t.joinable()) t.join(); } } joiner{pool};
  bool fits_ready = false;
  // Phase 3a + the host-side shuffle replay.  Nothing here depends on the mosaic VALUES: the fit rows are the clear land
  // pixels (weights == 0, never blended), their EVI comes from the untouched cube.  It does depend on the weights, which
  // the mosaic phase changes only when a date has <= 1000 usable pixels (:679-680).  So when the first pass over the
  // dates shows that no date is in that case, this runs INSIDE the mosaic phase and the generator walk + the shuffles
  // (10-20 ms of host work) overlap the mosaic kernels; otherwise it runs after the mosaic, as the reference orders it.
  auto prepare_fits = [&]() -> int {
    STC_CUDA(stc_dmalloc(&d_flag3.p, (size_t)n * HW)); STC_CUDA(stc_dmalloc(&d_pos3.p, (size_t)n * HW * 4)); STC_CUDA(stc_dmalloc(&d_K3.p, CF_MAX_DATES * 4));
    CF_LAUNCH(k_water_of_median, cdiv(HW, 128), 128, tiles, n, HW, water1);
    STC_CUDA(cudaMemsetAsync(d_counts.p, 0, CF_MAX_DATES * 5 * 4, ctx->stream));
    CF_LAUNCH(k_area_counts, dim3(cdiv(HW, 1024), n), 1024, areas, water1, HW, d_counts.as<int>());
    CF_LAUNCH(k_flag_clear_land, dim3(cdiv(HW, 256), n), 256, areas, water1, HW, d_flag3.as<unsigned char>());
    if ((rc = scan_flags_dev(ctx, d_flag3.as<unsigned char>(), n, HW, d_pos3.as<int>(), d_K3.as<int>()))) return rc;
    STC_CUDA(cudaMemcpyAsync(counts.data(), d_counts.p, n * 20, cudaMemcpyDeviceToHost, ctx->stream));
    CF_SYNC();
    // ---- 3a. everything about the per-date fits that does NOT depend on the blending of earlier dates, for all dates at once:
    //      the clear-land rows of a date's window [lo, hi) (weights == 0: the blend never touches them), their EVI, the six EVI
    //      percentiles, the stratum bits and the seven ordered index lists (:421-471).  One synchronisation for the stratum
    //      sizes, one for the lists; the lists land in pinned host memory.
    for (int d = 0; d < n; ++d) {
      const int c_pos = counts[d * 5], c_zero = counts[d * 5 + 1], c_lt1 = counts[d * 5 + 2], c_one = counts[d * 5 + 3];
      to_remove_host[d] = (c_one == HW);
      if (!(c_pos > 0 && c_zero > 0)) continue;
      if (!((double)c_lt1 / (double)HW > 0.01))
        STC_FAIL(STC_ERR_STATE, "remove_clouds: date with <= 1% non-saturated pixels -- the reference raises UnboundLocalError here (cloud_removal.py:575)");
      FitJob f; f.d = d;
      if (c_zero > 40000) { f.lo = d; f.hi = d + 1; }
      else { f.lo = (d == n - 1) ? std::max(d - 2, 0) : std::max(d - 1, 0); f.hi = std::min(d + 2, n); }
      f.K = 0;
      for (int t = f.lo; t < f.hi; ++t) f.K += counts[t * 5 + 4];
      if (f.K < 1) STC_FAIL(STC_ERR_STATE, "remove_clouds: no clear land pixel to fit on -- the reference fails in np.percentile here");
      f.row0 = total_rows; total_rows += f.K;
      fit_of[d] = (int)fits.size(); fits.push_back(f);
    }
    nf = (int)fits.size();
    if (nf > 0) {
      STC_CUDA(stc_dmalloc(&d_rows_all.p, (size_t)total_rows * 4)); STC_CUDA(stc_dmalloc(&d_evi_all.p, (size_t)total_rows * 4));
      STC_CUDA(stc_dmalloc(&d_lab_all.p, (size_t)total_rows));
      STC_CUDA(stc_dmalloc(&d_fqout.p, (size_t)nf * 6 * 4)); STC_CUDA(stc_dmalloc(&d_fcnt.p, (size_t)nf * 7 * 4));
      STC_CUDA(stc_dmalloc(&d_status_all.p, (size_t)nf * 10 * 4));
      std::vector<SelJob> jobs; std::vector<QSpec> specs;
      const double qs[6] = {2, 20, 40, 60, 80, 98};           // percentiles (np_quantile_pos divides by 100 in float32)
      for (const FitJob& f : fits) {
        int K = 0;
        for (int t = f.lo; t < f.hi; ++t) {
          CF_LAUNCH(k_collect_rows, cdiv(HW, 256), 256, tiles, d_pos3.as<int>() + (int64_t)t * HW, HW, t, (int)(f.row0 + K), d_rows_all.as<int>(), d_evi_all.as<float>());
          K += counts[t * 5 + 4];
        }
        for (int k = 0; k < 6; ++k) {
          jobs.push_back(SelJob{d_evi_all.as<float>() + f.row0, f.K, 1, 1});
          specs.push_back(QSpec{(int)(jobs.size() - 1) * SEL_MAX_COLS, f.K, qs[k], 0});
        }
      }
      if ((rc = run_select(jobs, specs, d_fqout.as<float>()))) return rc;
      STC_CUDA(cudaMemsetAsync(d_fcnt.p, 0, (size_t)nf * 7 * 4, ctx->stream));
      for (int j = 0; j < nf; ++j)
        CF_LAUNCH(k_strata, cdiv(fits[j].K, 256), 256, d_evi_all.as<float>() + fits[j].row0, d_fqout.as<float>() + 6 * j, fits[j].K,
                  d_lab_all.as<unsigned char>() + fits[j].row0, d_fcnt.as<int>() + 7 * j);
      std::vector<int> h_cnt((size_t)nf * 7);
      STC_CUDA(cudaMemcpyAsync(h_cnt.data(), d_fcnt.p, (size_t)nf * 7 * 4, cudaMemcpyDeviceToHost, ctx->stream));
      CF_SYNC();
      int64_t total_list = 0;
      const int rep[7] = {10, 1, 1, 1, 1, 1, 10};
      for (int j = 0; j < nf; ++j)
        for (int k = 0; k < 7; ++k) { fits[j].cnt[k] = h_cnt[(size_t)j * 7 + k] * rep[k]; fits[j].list0[k] = total_list; total_list += fits[j].cnt[k]; }
      for (const FitJob& f : fits)
        for (int k = 1; k <= 5; ++k)                          // p2 / p98 go through np.repeat first, which accepts a 0-d array
          if (f.cnt[k] == 1)
            STC_FAIL(STC_ERR_STATE, "remove_clouds: single-element EVI stratum -- the reference raises TypeError (shuffle of a 0-d array)");
      STC_CUDA(stc_dmalloc(&d_lists.p, (size_t)std::max<int64_t>(total_list, 1) * 4)); STC_CUDA(stc_dmalloc(&d_bjobs.p, (size_t)nf * sizeof(BucketJob)));
      int max_blocks = 1;
      for (int j = 0; j < nf; ++j) max_blocks = std::max(max_blocks, cdiv(fits[j].K, BL_ROWS));
      STC_CUDA(stc_dmalloc(&d_bchunk.p, (size_t)nf * max_blocks * 7 * 4));
      std::vector<BucketJob> bj(nf);
      for (int j = 0; j < nf; ++j) {
        bj[j].lab = d_lab_all.as<unsigned char>() + fits[j].row0; bj[j].K = fits[j].K;
        for (int k = 0; k < 7; ++k) bj[j].out[k] = d_lists.as<int>() + fits[j].list0[k];
        bj[j].chunk = d_bchunk.as<int>() + (size_t)j * max_blocks * 7;
      }
      {
        const void* staged = ctx_stage(ctx, bj.data(), (size_t)nf * sizeof(BucketJob));
        if (!staged) STC_FAIL(STC_ERR_NOMEM, "remove_clouds: pinned staging");
        STC_CUDA(cudaMemcpyAsync(d_bjobs.p, staged, (size_t)nf * sizeof(BucketJob), cudaMemcpyHostToDevice, ctx->stream));
      }
      CF_LAUNCH(k_bucket_count, dim3(max_blocks, nf), 256, d_bjobs.as<BucketJob>());
      CF_LAUNCH(k_bucket_scan, nf, 224, d_bjobs.as<BucketJob>());
      CF_LAUNCH(k_bucket_scatter, dim3(max_blocks, nf), 256, d_bjobs.as<BucketJob>());
      // pinned host scratch: the lists, then (behind them) one sample slot per date
      int64_t sample_cap = 0;
      sample0.assign(nf, 0);
      for (int j = 0; j < nf; ++j) {
        const int64_t n_i = std::min(90000, fits[j].K) / 5;
        int64_t cap = (int64_t)fits[j].cnt[0] + fits[j].cnt[6];
        for (int k = 1; k <= 5; ++k) cap += std::min<int64_t>(n_i, fits[j].cnt[k]);
        sample0[j] = sample_cap; sample_cap += cap;
      }
      pin_lists = (int*)ctx_pinned(ctx, (size_t)(total_list + sample_cap + 16) * 4);
      if (!pin_lists) STC_FAIL(STC_ERR_NOMEM, "remove_clouds: pinned host scratch");
      STC_CUDA(stc_dmalloc(&d_sample_all.p, (size_t)std::max<int64_t>(sample_cap, 1) * 4));
      // The lists (4 MB per date) travel on their own stream, one event per date: the compute stream goes on with the mosaic,
      // the generator walk below needs only the list LENGTHS and starts now, and a worker waits for its date's event.
      if (!ctx->d2h_stream) {
        STC_CUDA(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
        STC_CUDA(cudaEventCreateWithFlags(&ctx->d2h_fork, cudaEventDisableTiming));
      }
      if (!ctx->smp_stream) STC_CUDA(cudaStreamCreateWithFlags(&ctx->smp_stream, cudaStreamNonBlocking));
      while ((int)ctx->d2h_events.size() < nf) {
        cudaEvent_t e; STC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->d2h_events.push_back(e);
        STC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->smp_events.push_back(e);
      }
      STC_CUDA(cudaEventRecord(ctx->d2h_fork, ctx->stream));
      STC_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, ctx->d2h_fork, 0));
      STC_CUDA(cudaStreamWaitEvent(ctx->smp_stream, ctx->d2h_fork, 0));      // d_sample_all comes from the stream-ordered pool
      for (int j = 0; j < nf; ++j) {
        const int64_t a = fits[j].list0[0], b = fits[j].list0[6] + fits[j].cnt[6];
        if (b > a) STC_CUDA(cudaMemcpyAsync(pin_lists + a, d_lists.as<int>() + a, (size_t)(b - a) * 4, cudaMemcpyDeviceToHost, ctx->d2h_stream));
        STC_CUDA(cudaEventRecord(ctx->d2h_events[j], ctx->d2h_stream));
      }
      cf_mark("fit rows, strata, index lists (all dates)");
      // ---- 3b. in date order: Python's random.shuffle replayed on the host for date d while the device still works on date
      //      d - 1 (launches are asynchronous; nothing below synchronises), then Gram sums, NNLS and the blend of date d.
      pin_samples = pin_lists + total_list;
      // Python's random.shuffle, replayed on the host: 8 shuffles per date (:472-496) whose swaps are dependent random
      // memory accesses (~3.5 ns per element, ~0.5 M elements per date).  The generator is first walked through all of
      // them WITHOUT data (skip_shuffle) to record the state each one starts from; the shuffles themselves then run on
      // worker threads (they touch disjoint lists; a date's sample shuffle waits for its seven list shuffles), and the
      // device work of date d is enqueued as soon as that date's sample is ready.
      tasks.resize((size_t)nf * 8);
      S_of.assign(nf, 0);
      for (int j = 0; j < nf; ++j) {
        const FitJob& f = fits[j];
        const size_t n_i = (size_t)(std::min(90000, f.K) / 5);
        size_t S = (size_t)f.cnt[0] + (size_t)f.cnt[6];
        for (int k = 1; k <= 5; ++k) S += std::min(n_i, (size_t)f.cnt[k]);
        S_of[j] = S;
      }
      {
        // staged rows of every date (k_gram_rows runs ahead of the per-date chain, so each date keeps its own)
        zoff.assign(nf, 0);
        size_t s_tot = 0;
        for (int j = 0; j < nf; ++j) { zoff[j] = s_tot; s_tot += std::min(S_of[j], (size_t)fits[j].K); }
        STC_CUDA(stc_dmalloc(&d_gramz.p, std::max<size_t>(s_tot, 1) * GRAM_ZW * 4));
        STC_CUDA(stc_dmalloc(&d_pix.p, std::max<size_t>(s_tot, 1) * 4));
      }
      tt0 = cf_timing ? cf_now() : 0;
      auto walker = [&, this_n = n]() {
        size_t ti = 0;
        rng.attach(&rng_ahead);
        for (int d = 0; d < n; ++d) {
          const int j = fit_of[d];
          if (j < 0) continue;
          const FitJob& f = fits[j];
          const int order[8] = {0, 6, 1, 2, 3, 4, 5, 7};                        // :472-478 shuffle order p2, p98, p20 ... p100, then the sample
          for (int q = 0; q < 8; ++q) {
            ShufTask& t = tasks[ti];
            t.j = j; t.k = order[q]; rng.export_state(t.mt, &t.idx);
            published.store(++ti, std::memory_order_release);
            const size_t len = order[q] < 7 ? (size_t)f.cnt[order[q]] : S_of[j];
            rng.skip_shuffle(len);
            n_draw += (long long)len;
          }
        }
        if (cf_timing) t_skip = cf_now() - tt0;
      };
      lists_done.reset(new std::atomic<int>[nf]); sample_done.reset(new std::atomic<int>[nf]);
      for (int j = 0; j < nf; ++j) { lists_done[j].store(0); sample_done[j].store(0); }
      auto worker = [&]() {
        PyRandom r;
        cudaSetDevice(ctx->device);                          // event queries and the sample upload below are CUDA calls of this thread
        for (;;) {
          const size_t ti = next_task.fetch_add(1);
          if (ti >= tasks.size()) return;
          while (published.load(std::memory_order_acquire) <= ti) std::this_thread::yield();
          const ShufTask& t = tasks[ti];
          const FitJob& f = fits[t.j];
          while (cudaEventQuery(ctx->d2h_events[t.j]) == cudaErrorNotReady) std::this_thread::yield();      // the date's lists are on the host
          r.import_state(t.mt, t.idx);
          if (t.k < 7) {
            r.shuffle(pin_lists + f.list0[t.k], (size_t)f.cnt[t.k]);
            lists_done[t.j].fetch_add(1, std::memory_order_release);
          } else {
            while (lists_done[t.j].load(std::memory_order_acquire) < 7) std::this_thread::yield();   // earlier tasks: already taken by a worker
            const size_t n_i = (size_t)(std::min(90000, f.K) / 5);
            int* smp = pin_samples + sample0[t.j];
            size_t S = 0;
            auto append = [&](int k, size_t limit) { const size_t c = std::min(limit, (size_t)f.cnt[k]); memcpy(smp + S, pin_lists + f.list0[k], c * 4); S += c; };
            append(0, (size_t)f.cnt[0]); for (int k = 1; k <= 5; ++k) append(k, n_i); append(6, (size_t)f.cnt[6]);   // [p2, p20, p40, p60, p80, p100, p98]
            r.shuffle(smp, S);
            // the upload starts here, next to the device work of the earlier dates; the compute stream waits on the event
            const size_t up = std::min(S, (size_t)f.K);
            const cudaError_t ce = cudaMemcpyAsync(d_sample_all.as<int>() + sample0[t.j], smp, up * 4, cudaMemcpyHostToDevice, ctx->smp_stream);
            sample_done[t.j].store(ce == cudaSuccess ? 1 : -1, std::memory_order_release);
          }
        }
      };
      pool.emplace_back(walker);
      for (int q = 0; q < nthreads; ++q) pool.emplace_back(worker);
    }
    fits_ready = true;
    return STC_OK;
    };
  // ---- 2. cloud-free mosaic (:578-699) ----
  CF_LAUNCH(k_mosaic_prep, cdiv(HW, 128), 128, tiles, areas, n, HW, u8a, d_div.as<float>());
  maskop_dilate(ctx, u8a, u8b, 1, H, W, 2, 1, 1, 0, 0);          // dilate(1 - water, 2)
  maskop_dilate(ctx, u8b, water0, 1, H, W, 5, 1, 1, 0, 0);       // dilate(1 - that, 5)
  STC_CUDA(cudaMemsetAsync(mosaic, 0, (int64_t)HW * 40, ctx->stream));
  STC_CUDA(cudaMemsetAsync(cnt, 0, 64, ctx->stream));
  CF_LAUNCH(k_count_zero, cdiv(HW, 256), 256, water0, HW, cnt + 1);
  // All dates at once: reference images, compaction positions, row gathers, 20 medians + 20 standard deviations and the
  // scale parameters of every date run in ONE launch each (per-date slabs), because a single date offers two blocks of
  // work to 148 SMs (the column chains of np.nanstd are sequential by definition).  Only the accumulation into the mosaic
  // stays in date order (float32 sum order).  A date that cannot be aligned (<= 1000 usable pixels) sets its weights to 1
  // (:679-680), which changes the reference images of the LATER dates: the batch is then cut at that date and restarted
  // behind it, exactly reproducing the sequential loop.
  DBuf d_refall, d_flagall, d_posall, d_srcall, d_refrowsall, d_Ks, d_medall, d_sdall, d_paramsall;
  const int64_t slab = (int64_t)HW * 10;
  STC_CUDA(stc_dmalloc(&d_refall.p, (size_t)n * slab * 4)); STC_CUDA(stc_dmalloc(&d_srcall.p, (size_t)n * slab * 4));
  STC_CUDA(stc_dmalloc(&d_refrowsall.p, (size_t)n * slab * 4));
  STC_CUDA(stc_dmalloc(&d_flagall.p, (size_t)n * HW)); STC_CUDA(stc_dmalloc(&d_posall.p, (size_t)n * HW * 4));
  STC_CUDA(stc_dmalloc(&d_Ks.p, CF_MAX_DATES * 4)); STC_CUDA(stc_dmalloc(&d_medall.p, CF_MAX_DATES * 20 * 4));
  STC_CUDA(stc_dmalloc(&d_sdall.p, CF_MAX_DATES * 20 * 4)); STC_CUDA(stc_dmalloc(&d_paramsall.p, CF_MAX_DATES * 20 * 4));
  int land_px = 0;
  STC_CUDA(cudaMemcpyAsync(&land_px, cnt + 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<int> Ks(n, 0);
  int start = 0;
  while (start < n) {
    const int m = n - start;
    if (n <= 8) CF_LAUNCH(k_mosaic_ref<8>, cdiv((int64_t)HW * 10, 256), 256, tiles, areas, water0, n, HW, start, d_refall.as<float>(), d_flagall.as<unsigned char>());
    else if (n <= 16) CF_LAUNCH(k_mosaic_ref<16>, cdiv((int64_t)HW * 10, 256), 256, tiles, areas, water0, n, HW, start, d_refall.as<float>(), d_flagall.as<unsigned char>());
    else CF_LAUNCH(k_mosaic_ref<32>, cdiv((int64_t)HW * 10, 256), 256, tiles, areas, water0, n, HW, start, d_refall.as<float>(), d_flagall.as<unsigned char>());
    if ((rc = scan_flags_dev(ctx, d_flagall.as<unsigned char>() + (int64_t)start * HW, m, HW, d_posall.as<int>() + (int64_t)start * HW,
                             d_Ks.as<int>() + start))) return rc;
    STC_CUDA(cudaMemcpyAsync(Ks.data() + start, d_Ks.as<int>() + start, m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CF_SYNC();
    int f = start;
    while (f < n && Ks[f] > 1000) ++f;                       // dates [start, f) are final
    if (start == 0 && f == n && !fits_ready && (rc = prepare_fits())) return rc;      // no weight changes ahead: overlap (see prepare_fits)
    if (f > start) {
      const int cv = f - start;
      CF_LAUNCH(k_gather_rows, dim3(cdiv(slab, 256), cv), 256, tiles, d_refall.as<float>(), d_posall.as<int>(), HW, start,
                d_srcall.as<float>(), d_refrowsall.as<float>());
      // np.nanstd's column chains are sequential by definition (k_col_std: 20 warps per date, ~3 ms): they run on a side
      // stream next to the medians, which use the whole GPU
      if (!ctx->aux_stream) {
        STC_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
        STC_CUDA(cudaEventCreateWithFlags(&ctx->aux_ev[0], cudaEventDisableTiming)); STC_CUDA(cudaEventCreateWithFlags(&ctx->aux_ev[1], cudaEventDisableTiming));
      }
      STC_CUDA(cudaEventRecord(ctx->aux_ev[0], ctx->stream));
      STC_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_ev[0], 0));
      { k_col_std<<<dim3(10, cv), 64, 0, ctx->aux_stream>>>(d_srcall.as<float>(), d_refrowsall.as<float>(), slab, d_Ks.as<int>(), start, d_sdall.as<float>()); }
      ctx->launches++;
      STC_CUDA(cudaEventRecord(ctx->aux_ev[1], ctx->aux_stream));
      std::vector<SelJob> jobs; std::vector<QSpec> specs;
      for (int i = start; i < f; ++i)
        for (int mm = 0; mm < 2; ++mm) {
          jobs.push_back(SelJob{(mm ? d_refrowsall.as<float>() : d_srcall.as<float>()) + (int64_t)i * slab, Ks[i], 10, 10});
          for (int c = 0; c < 10; ++c) specs.push_back(QSpec{(int)(jobs.size() - 1) * SEL_MAX_COLS + c, Ks[i], 0.0, 1});
        }
      if ((rc = run_select(jobs, specs, d_medall.as<float>() + start * 20))) return rc;
      STC_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->aux_ev[1], 0));
      CF_LAUNCH(k_scale_params, cv, 32, d_medall.as<float>() + start * 20, d_sdall.as<float>() + start * 20, d_paramsall.as<float>() + start * 20);
      CF_LAUNCH(k_mosaic_accum, cdiv(slab, 256), 256, tiles, areas, water0, d_paramsall.as<float>(), HW, start, f, mosaic);
    }
    if (f < n && land_px > 0)
      CF_LAUNCH(k_fill_f, cdiv(HW, 256), 256, areas + (int64_t)f * HW, (int64_t)HW, 1.f);      // interp[i] = 1. (:679-680)
    start = f + 1;
  }
  CF_LAUNCH(k_mosaic_final, cdiv((int64_t)HW * 10, 128), 128, tiles, d_div.as<float>(), n, HW, mosaic);
  if (mosaic_out_dev) STC_CUDA(cudaMemcpyAsync(mosaic_out_dev, mosaic, (int64_t)HW * 40, cudaMemcpyDeviceToDevice, ctx->stream));

  cf_mark("mosaic");
  // ---- 3. per-date alignment and blending (:939-959, :316-575) ----
  if (!fits_ready && (rc = prepare_fits())) return rc;
  if (nf > 0) {
    const char* nnls_env = getenv("STC_NNLS_SERIAL");                  // A/B reference, read per call
    const bool nnls_serial = nnls_env && atoi(nnls_env) == 1;
    DBuf d_sp;
    STC_CUDA(stc_dmalloc(&d_sp.p, (size_t)n * HW * 4));
    float* sp = d_sp.as<float>();
    CF_LAUNCH(k_snow_planes, cdiv((int64_t)n * HW, 256), 256, tiles, n, HW, sp);
    STC_CUDA(cudaEventRecord(ctx->d2h_fork, ctx->stream));                 // the mosaic is final: the sample stream may gather rows
    STC_CUDA(cudaStreamWaitEvent(ctx->smp_stream, ctx->d2h_fork, 0));
    for (int d = 0; d < n; ++d) {
      const int j = fit_of[d];
      if (j < 0) {
        if (counts[d * 5] > 0 && !(counts[d * 5 + 1] > 0))    // no clear pixel at all: the interpolated array stays the raw mosaic
          CF_LAUNCH(k_predict_blend, cdiv(HW, 256), 256, tiles + (int64_t)d * HW * 10, areas + (int64_t)d * HW, mosaic, (const float*)nullptr,
                    d_coef.as<double>(), 0, HW, sp + (int64_t)d * HW);
        continue;
      }
      const FitJob& f = fits[j];
      int sdone;
      while (!(sdone = sample_done[j].load(std::memory_order_acquire))) std::this_thread::yield();
      if (sdone < 0) STC_FAIL(STC_ERR_CUDA, "remove_clouds: sample upload failed");
      int* smp = pin_samples + sample0[j];
      size_t S = S_of[j];
      if (S > (size_t)f.K) S = (size_t)f.K;
      int* d_smp = d_sample_all.as<int>() + sample0[j];
      float* Zj = d_gramz.as<float>() + zoff[j] * GRAM_ZW;
      int* pixj = d_pix.as<int>() + zoff[j];
      {
        // the snow-independent part of the rows: on the sample stream, behind the upload the worker enqueued there, next to the
        // chain of the earlier dates on the compute stream
        cudaStream_t keep = ctx->stream; ctx->stream = ctx->smp_stream;
        CF_LAUNCH(k_gram_rows, cdiv((int64_t)S * GRAM_ZW, 256), 256, tiles, mosaic, d_rows_all.as<int>() + f.row0, d_smp, (int)S, HW, Zj, pixj);
        ctx->stream = keep;
        STC_CUDA(cudaEventRecord(ctx->smp_events[j], ctx->smp_stream));
        STC_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->smp_events[j], 0));
      }
      CF_LAUNCH(k_snow_mean, cdiv(HW, 256), 256, sp, n, HW, d_snow.as<float>());
      CF_LAUNCH(k_gram_snow, cdiv((int64_t)S, 256), 256, d_snow.as<float>(), pixj, (int)S, Zj);
      const int gb = std::min(gram_blocks, cdiv((int64_t)S, GRAM_ROWS));
      CF_LAUNCH(k_gram, cdiv(gb, 2), 128, Zj, (int)S, gb, d_partial.as<double>());
      CF_LAUNCH(k_gram_reduce, cdiv(GRAM_N, 32), 256, d_partial.as<double>(), gb, d_gram.as<double>());
      if (nnls_serial) CF_LAUNCH(k_nnls_serial, 1, 32, d_gram.as<double>(), (int)S, d_coef.as<double>(), d_status_all.as<int>() + 10 * j);
      else CF_LAUNCH(k_nnls, 1, 320, d_gram.as<double>(), (int)S, d_coef.as<double>(), d_status_all.as<int>() + 10 * j);
      if (getenv("STC_CF_DEBUG")) {                        // test aid: fit inputs / coefficients of every date on stderr
        double hc[10 * NF]; int hs[6] = {0}; int hr[6] = {0}; float hq[6] = {0};
        cudaMemcpyAsync(hq, d_fqout.as<float>() + 6 * j, sizeof(hq), cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(hc, d_coef.p, sizeof(hc), cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(hr, d_rows_all.as<int>() + f.row0, sizeof(int) * (f.K < 6 ? f.K : 6), cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        for (int q = 0; q < 6 && q < (int)S; ++q) hs[q] = smp[q];
        fprintf(stderr, "[cf_debug] date %d window [%d,%d) K %d S %zu sample %d %d %d %d %d %d rows %d %d %d %d %d %d cnt %d %d %d %d %d %d %d\n", d, f.lo, f.hi, f.K, S,
                hs[0], hs[1], hs[2], hs[3], hs[4], hs[5], hr[0], hr[1], hr[2], hr[3], hr[4], hr[5], f.cnt[0], f.cnt[1], f.cnt[2], f.cnt[3], f.cnt[4], f.cnt[5], f.cnt[6]);
        fprintf(stderr, "[cf_debug]   evi percentiles %.9g %.9g %.9g %.9g %.9g %.9g\n", hq[0], hq[1], hq[2], hq[3], hq[4], hq[5]);
        for (int b = 0; b < 10; b += 8) { fprintf(stderr, "[cf_debug]   band %d coef", b); for (int q = 0; q < NF; ++q) fprintf(stderr, " %.6g", hc[b * NF + q]); fprintf(stderr, "\n"); }
      }
      CF_LAUNCH(k_predict_blend, cdiv(HW, 256), 256, tiles + (int64_t)d * HW * 10, areas + (int64_t)d * HW, mosaic, d_snow.as<float>(),
                d_coef.as<double>(), 1, HW, sp + (int64_t)d * HW);
    }
    for (auto& t : pool) t.join();
    if (cf_timing) fprintf(stderr, "[remove_clouds]   generator walk %.1f ms (%lld elements), shuffles on %d host threads, all enqueued after %.1f ms\n",
                           t_skip, n_draw, nthreads, cf_now() - tt0);
    std::vector<int> status((size_t)nf * 10);
    STC_CUDA(cudaMemcpyAsync(status.data(), d_status_all.p, (size_t)nf * 40, cudaMemcpyDeviceToHost, ctx->stream));
    CF_SYNC();
    for (int v : status)
      if (v != 1) STC_FAIL(STC_ERR_STATE, "remove_clouds: NNLS did not converge (scipy.optimize.nnls raises RuntimeError)");
  } else {
    for (int d = 0; d < n; ++d)
      if (counts[d * 5] > 0 && !(counts[d * 5 + 1] > 0))
        CF_LAUNCH(k_predict_blend, cdiv(HW, 256), 256, tiles + (int64_t)d * HW * 10, areas + (int64_t)d * HW, mosaic, (const float*)nullptr,
                  d_coef.as<double>(), 0, HW, (float*)nullptr);
  }
  { int idx_out = 0; rng.export_state(mt_state, &idx_out); mt_state[624] = (uint32_t)idx_out; rng_ahead.stop(); }

  cf_mark("per-date alignment + blend");
  // ---- 4. residual clouds in the mosaic (:703-732) ----
  maskop_dilate(ctx, pfcps_dev, pf, 1, H, W, 10, 1, 0, 0, 0);      // pfcps[0] (single frame when n == 1)
  STC_CUDA(cudaMemsetAsync(cnt, 0, 8, ctx->stream));
  CF_LAUNCH(k_only_one, cdiv(HW, 256), 256, areas, pf, n, HW, u8a, cnt);
  int only_cnt = 0;
  STC_CUDA(cudaMemcpyAsync(&only_cnt, cnt, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CF_SYNC();
  if (only_cnt != HW) {
    CF_LAUNCH(k_not, cdiv(HW, 256), 256, u8a, flag, HW);
    if ((rc = scan_flags_dev(ctx, flag, 1, HW, d_pos.as<int>(), cnt))) return rc;
    float* blue = d_src_rows.as<float>(); float* red = d_ref_rows.as<float>();
    CF_LAUNCH(k_gather_br, cdiv(HW, 256), 256, mosaic, d_pos.as<int>(), HW, blue, red);
    const int K2 = HW - only_cnt;
    std::vector<SelJob> jobs = {SelJob{blue, K2, 1, 1}, SelJob{red, K2, 1, 1}};
    std::vector<QSpec> specs = {QSpec{0, K2, 99.0, 0}, QSpec{SEL_MAX_COLS, K2, 99.0, 0}};
    if ((rc = run_select(jobs, specs, d_qout.as<float>()))) return rc;
    CF_LAUNCH(k_mosaic_clouds, cdiv(HW, 256), 256, mosaic, u8a, pf, d_qout.as<float>(), HW, u8b);
    maskop_dilate(ctx, u8b, flag, 1, H, W, 3, 1, 1, 0, 0);
    maskop_dilate(ctx, flag, u8b, 1, H, W, 8, 1, 1, 0, 0);
    CF_LAUNCH(k_add_clip, cdiv(N, 256), 256, areas, u8b, n, HW);
  }
  STC_CUDA(cudaGetLastError());
  cf_mark("residual clouds");
  if (clip_when_all_kept) {
    // process_tile clips the cube right after this call (src/download_and_predict_job.py:996) unless a fully
    // interpolated date has to be dropped first (:972-990, the masks are then recomputed on the unclipped cube): doing it
    // here saves a round trip of the whole cube
    bool all_kept = true;
    for (int d = 0; d < n; ++d) all_kept = all_kept && !to_remove_host[d];
    if (all_kept) {
      CF_LAUNCH(k_clip01, cdiv(N * 10, 256), 256, tiles, N * 10);
      if (clipped_out) *clipped_out = 1;
    }
  }
  return STC_OK;
}

static int remove_clouds_impl(stc_ctx* ctx, float* tiles_host, const float* probs_host, const uint8_t* pfcps_host, int n, int H,
                              int W, uint32_t* mt_state, float* areas_host, int32_t* to_remove_host, float* mosaic_host,
                              int clip_when_all_kept, int32_t* clipped_out) {
  if (!ctx) return STC_ERR_ARG;
  if (!tiles_host || !probs_host || !pfcps_host || !areas_host || n < 1 || n > CF_MAX_DATES || H < 3 || W < 3)
    STC_FAIL(STC_ERR_ARG, "remove_clouds: bad argument (1 <= n <= 32, MT19937 state of 624 words + position)");
  const int64_t N = (int64_t)n * H * W;
  DBuf d_tiles, d_areas, d_probs, d_pfall, d_mosaic;
  STC_CUDA(stc_dmalloc(&d_tiles.p, N * 40)); STC_CUDA(stc_dmalloc(&d_areas.p, N * 4)); STC_CUDA(stc_dmalloc(&d_probs.p, N * 4));
  STC_CUDA(stc_dmalloc(&d_pfall.p, N));
  if (mosaic_host) STC_CUDA(stc_dmalloc(&d_mosaic.p, (size_t)H * W * 40));
  STC_CUDA(cudaMemcpyAsync(d_tiles.p, tiles_host, N * 40, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(d_probs.p, probs_host, N * 4, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(d_pfall.p, pfcps_host, N, cudaMemcpyHostToDevice, ctx->stream));
  int rc = remove_clouds_dev(ctx, d_tiles.as<float>(), d_probs.as<float>(), d_pfall.as<unsigned char>(), n, H, W, mt_state,
                             d_areas.as<float>(), to_remove_host, mosaic_host ? d_mosaic.as<float>() : nullptr, clip_when_all_kept, clipped_out);
  if (rc) return rc;
  if (mosaic_host) STC_CUDA(cudaMemcpyAsync(mosaic_host, d_mosaic.p, (size_t)H * W * 40, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(tiles_host, d_tiles.p, N * 40, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(areas_host, d_areas.p, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CF_SYNC();
  return STC_OK;
}

extern "C" int stc_remove_clouds_host(stc_ctx* ctx, float* tiles_host, const float* probs_host, const uint8_t* pfcps_host, int n, int H,
                                      int W, uint32_t* mt_state, float* areas_host, int32_t* to_remove_host, float* mosaic_host) {
  return remove_clouds_impl(ctx, tiles_host, probs_host, pfcps_host, n, H, W, mt_state, areas_host, to_remove_host, mosaic_host, 0, nullptr);
}

extern "C" int stc_remove_clouds_clip_host(stc_ctx* ctx, float* tiles_host, const float* probs_host, const uint8_t* pfcps_host, int n,
                                           int H, int W, uint32_t* mt_state, float* areas_host, int32_t* to_remove_host,
                                           int32_t* clipped_out) {
  return remove_clouds_impl(ctx, tiles_host, probs_host, pfcps_host, n, H, W, mt_state, areas_host, to_remove_host, nullptr, 1, clipped_out);
}

// Test hook for the host-side generator replay (no device, no context): shuffles data[0..n) exactly like Python's
// random.shuffle would with the generator state mt_state (624 words + position), and writes the advanced state back.
extern "C" int stc_py_shuffle(uint32_t* mt_state, int32_t* data, int64_t n) {
  if (!mt_state || n < 0 || mt_state[624] > 624) return STC_ERR_ARG;
  PyRandomProducer ahead;
  PyRandom rng; rng.import_state(mt_state, (int)mt_state[624]);
  if (!data) { rng.attach(&ahead); rng.skip_shuffle((size_t)n); }     // generator walk only, fed by the producer thread: what remove_clouds does before it farms the shuffles out
  else rng.shuffle(data, (size_t)n);
  int idx_out = 0; rng.export_state(mt_state, &idx_out); mt_state[624] = (uint32_t)idx_out;
  return STC_OK;
}
