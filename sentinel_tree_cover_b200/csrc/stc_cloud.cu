// Multi-temporal cloud / shadow mask, identify_clouds_shadows
// (src/preprocessing/cloud_removal.py:1215-1677) for the configuration the reference tree runs in
// (no urbanmask.tif / forestmask.tif: forest mask and potential-false-positive masks are zero).
// All array arithmetic runs on the device; the host only drives data-dependent control flow
// (the adaptive threshold loop :1425-1440, the per-date plausibility tests) from scalars the
// device reduces.  One thread owns one pixel column of T <= 32 dates in registers for the
// temporal stages; spatial stages are windowed searches with SciPy's border semantics.
// Stage-by-stage parity is checked against oracle/cloud_ref.py (tests/test_cloud_masks.py).
#include "stc_common.cuh"
#include "stc_select.cuh"
#include "stc_sortnet.cuh"
#include <cmath>
int pfcp_detect_dev(stc_ctx* ctx, const float* img, const float* dem, const unsigned char* urban_core, const unsigned char* urban_near, int T,
                    int H, int W, unsigned char* fcps_out, unsigned char* pfps_out);
#include <cstring>
#include <algorithm>
#include <vector>

#define CT_MAX 32

namespace {

struct Buf { void* p = nullptr; ~Buf() { if (p) stc_dfree(p); } };

// ---------------------------------------------------------------------------------------------
// generic spatial primitives on [T][H][W] uint8 masks
// ---------------------------------------------------------------------------------------------
// binary dilation / erosion-by-dilation and the capped Euclidean grow are the separable passes of stc_morph.cu
// (morph_dilate_dev, morph_edt_grow_dev)

// 3x3 window sum with np.pad(mode='reflect') borders (cloud_removal.py:1244-1249, windowsize 3)
__device__ __forceinline__ int winsum3(const unsigned char* m, int y, int x, int H, int W) {
  int s = 0;
  for (int dy = -1; dy <= 1; ++dy) {
    int yy = y + dy; if (yy < 0) yy = -yy; if (yy >= H) yy = 2 * H - 2 - yy;
    for (int dx = -1; dx <= 1; ++dx) {
      int xx = x + dx; if (xx < 0) xx = -xx; if (xx >= W) xx = 2 * W - 2 - xx;
      s += m[(int64_t)yy * W + xx];
    }
  }
  return s;
}

// ---------------------------------------------------------------------------------------------
// small register sorts / order statistics
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void isort(float* v, int n) {
  for (int i = 1; i < n; ++i) { float x = v[i]; int j = i - 1; while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; } v[j + 1] = x; }
}
__device__ __forceinline__ float median_sorted(const float* v, int n) {   // np.median / nanmedian on n valid values
  if (n == 0) return nanf("");
  return (n & 1) ? v[n >> 1] : __fmul_rn(__fadd_rn(v[(n >> 1) - 1], v[n >> 1]), 0.5f);
}
// np.percentile(a, 25, axis=0) for float32: linear interpolation with NumPy's _lerp in float32
__device__ __forceinline__ float percentile25_sorted(const float* v, int n) {
  double vi = 0.25 * (double)(n - 1);
  int lo = (int)floor(vi);
  float g = (float)(vi - (double)lo);
  float a = v[lo], b = v[lo + 1 < n ? lo + 1 : n - 1];
  float d = __fsub_rn(b, a);
  return (g >= 0.5f) ? __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.f, g))) : __fadd_rn(a, __fmul_rn(d, g));
}

// ---------------------------------------------------------------------------------------------
// stage A: water mask, Hollstein mask, all-date shadow reference, 25th percentiles, minima
// ---------------------------------------------------------------------------------------------
struct StaticRefs {
  float* water;       // [HW]   nanmedian_t NDWI
  float* allref;      // [HW][4] nanmedian over dates of bands (0,1,7,8) with Hollstein-flagged dates removed
  float* minb4;       // [HW][4] min over dates of bands (0,1,7,8)
  float* p25;         // [HW][3] 25th percentile over dates of bands 0,1,2
  float* minrgb;      // [HW][3] min over dates of bands 0,1,2
};

__global__ void __launch_bounds__(256) k_hollstein(const float* __restrict__ img, unsigned char* __restrict__ clm, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = img + i * 10;
  clm[i] = (p[7] > 0.166f) && (p[1] > 0.28f) && (__fdiv_rn(p[5], p[8]) < 4.292f);
}

// one pixel by insertion sorts (the reference transcription): the route for pixels holding a NaN / infinity
__device__ void static_refs_px(const float* __restrict__ img, const unsigned char* __restrict__ clm, int T, int HW, int p, const StaticRefs& o) {
  float v[CT_MAX];
  // water = nanmedian_t (B3 - B8)/(B3 + B8)
  int n = 0;
  for (int t = 0; t < T; ++t) {
    const float* q = img + ((int64_t)t * HW + p) * 10;
    float x = __fdiv_rn(__fsub_rn(q[1], q[3]), __fadd_rn(q[1], q[3]));
    if (!isnan(x)) v[n++] = x;
  }
  isort(v, n);
  o.water[p] = median_sorted(v, n);
  const int bsel[4] = {0, 1, 7, 8};
  for (int k = 0; k < 4; ++k) {
    int nn = 0; float mn = INFINITY; float all[CT_MAX];
    for (int t = 0; t < T; ++t) {
      float x = img[((int64_t)t * HW + p) * 10 + bsel[k]];
      all[t] = x; mn = fminf(mn, x);
      if (!clm[(int64_t)t * HW + p]) v[nn++] = x;
    }
    isort(v, nn);
    float r = median_sorted(v, nn);
    if (isnan(r)) { isort(all, T); r = median_sorted(all, T); }   // np.median over all dates (:1303-1304)
    o.allref[p * 4 + k] = r;
    o.minb4[p * 4 + k] = mn;
  }
  for (int k = 0; k < 3; ++k) {
    float mn = INFINITY;
    for (int t = 0; t < T; ++t) { float x = img[((int64_t)t * HW + p) * 10 + k]; v[t] = x; mn = fminf(mn, x); }
    isort(v, T);
    o.p25[p * 3 + k] = percentile25_sorted(v, T);
    o.minrgb[p * 3 + k] = mn;
  }
}
// The same statistics with the dates of a pixel in registers and sorting networks of N >= T slots (stc_sortnet.cuh): the
// seven bands the stage uses are loaded once per date, the eight sorts per pixel cost 8 x 240 compare-exchanges at N = 32.
template <int N>
__global__ void __launch_bounds__(128) k_static_refs(const float* __restrict__ img, const unsigned char* __restrict__ clm, int T, int HW,
                                                     StaticRefs o) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  // bands 0, 1, 2, 3, 7, 8 of every date (b[.][t]), Hollstein flags as a bit mask
  float b0[N], b1[N], b2[N], b3[N], b7[N], b8[N];
  unsigned cloudy = 0; bool ok = true;
#pragma unroll
  for (int t = 0; t < N; ++t) {
    b0[t] = b1[t] = b2[t] = b3[t] = b7[t] = b8[t] = INFINITY;
    if (t < T) {
      const float* q = img + ((int64_t)t * HW + p) * 10;
      b0[t] = q[0]; b1[t] = q[1]; b2[t] = q[2]; b3[t] = q[3]; b7[t] = q[7]; b8[t] = q[8];
      ok = ok && net_ok(b0[t]) && net_ok(b1[t]) && net_ok(b2[t]) && net_ok(b3[t]) && net_ok(b7[t]) && net_ok(b8[t]);
      if (clm[(int64_t)t * HW + p]) cloudy |= 1u << t;
    }
  }
  float v[N];
  int n = 0;
#pragma unroll
  for (int t = 0; t < N; ++t) {
    float x = INFINITY;
    if (t < T) {
      x = __fdiv_rn(__fsub_rn(b1[t], b3[t]), __fadd_rn(b1[t], b3[t]));
      if (isnan(x)) x = INFINITY; else { ++n; ok = ok && net_ok(x); }
    }
    v[t] = x;
  }
  if (!ok) { static_refs_px(img, clm, T, HW, p, o); return; }
  sort_net<N>(v);
  o.water[p] = net_median<N>(v, n);
  const int nn = T - __popc(cloudy);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float (&bk)[N] = k == 0 ? b0 : k == 1 ? b1 : k == 2 ? b7 : b8;
    float mn = INFINITY;
#pragma unroll
    for (int t = 0; t < N; ++t) { if (t < T) mn = fminf(mn, bk[t]); v[t] = (t < T && !((cloudy >> t) & 1u)) ? bk[t] : INFINITY; }
    float r;
    if (nn > 0) { sort_net<N>(v); r = net_median<N>(v, nn); }
    else {                                                      // every date flagged: np.median over all dates (:1303-1304)
#pragma unroll
      for (int t = 0; t < N; ++t) v[t] = bk[t];
      sort_net<N>(v); r = net_median<N>(v, T);
    }
    o.allref[p * 4 + k] = r;
    o.minb4[p * 4 + k] = mn;
  }
  // np.percentile(a, 25, axis=0): positions as in percentile25_sorted
  const double vi = 0.25 * (double)(T - 1);
  const int lo = (int)floor(vi);
  const float g = (float)(vi - (double)lo);
  const int hi = lo + 1 < T ? lo + 1 : T - 1;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float (&bk)[N] = k == 0 ? b0 : k == 1 ? b1 : b2;
    float mn = INFINITY;
#pragma unroll
    for (int t = 0; t < N; ++t) { if (t < T) mn = fminf(mn, bk[t]); v[t] = bk[t]; }
    sort_net<N>(v);
    const float a = net_pick<N>(v, lo), bb = net_pick<N>(v, hi);
    const float d = __fsub_rn(bb, a);
    o.p25[p * 3 + k] = (g >= 0.5f) ? __fsub_rn(bb, __fmul_rn(d, __fsub_rn(1.f, g))) : __fadd_rn(a, __fmul_rn(d, g));
    o.minrgb[p * 3 + k] = mn;
  }
}

// ---------------------------------------------------------------------------------------------
// stage B: per-date shadow candidates (:1265-1324)
// ---------------------------------------------------------------------------------------------
// N >= the longest date window: the window's values sit in registers and are sorted by a network (stc_sortnet.cuh)
template <int N>
__global__ void __launch_bounds__(128) k_shadow_candidates(const float* __restrict__ img, const unsigned char* __restrict__ clm,
                                                           const float* __restrict__ dem, StaticRefs s, int T, int HW,
                                                           const int* __restrict__ win_lo, const int* __restrict__ win_hi,
                                                           unsigned char* __restrict__ shadows) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  if (p >= HW) return;
  const int lo = win_lo[t], hi = win_hi[t];
  float rmed[4], rmax[4];
  unsigned use = 0;                                   // window dates that are not Hollstein-flagged
#pragma unroll
  for (int i = 0; i < N; ++i) if (lo + i < hi && !clm[(int64_t)(lo + i) * HW + p]) use |= 1u << i;
  const int n = __popc(use);
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int band = k == 0 ? 0 : k == 1 ? 1 : k == 2 ? 7 : 8;
    float v[N]; float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float x = INFINITY;
      if ((use >> i) & 1u) { x = img[((int64_t)(lo + i) * HW + p) * 10 + band]; ok = ok && net_ok(x); mx = fmaxf(mx, x); }
      v[i] = x;
    }
    sort_net<N>(v);
    rmax[k] = n ? mx : nanf("");
    rmed[k] = n ? net_median<N>(v, n) : s.minb4[p * 4 + k];
  }
  if (!ok) {                                          // a NaN / infinity in the window: the insertion-sort transcription
    const int bsel[4] = {0, 1, 7, 8};
    for (int k = 0; k < 4; ++k) {
      float v[CT_MAX]; int nn = 0; float mx = -INFINITY;
      for (int tt = lo; tt < hi; ++tt) {
        if (clm[(int64_t)tt * HW + p]) continue;
        float x = img[((int64_t)tt * HW + p) * 10 + bsel[k]];
        v[nn++] = x; mx = fmaxf(mx, x);
      }
      isort(v, nn);
      float m = median_sorted(v, nn);
      rmax[k] = nn ? mx : nanf("");
      rmed[k] = isnan(m) ? s.minb4[p * 4 + k] : m;
    }
  }
  const float* x = img + ((int64_t)t * HW + p) * 10;
  const float water = s.water[p];
  const bool wpos = water > 0.f;
  bool sh = (__fsub_rn(x[8], rmed[3]) < -0.04f) && (__fsub_rn(x[7], rmed[2]) < -0.04f) && (x[0] < 0.09f) &&
            (__fsub_rn(x[0], rmed[0]) < -0.02f) && (x[7] < 0.17f);
  const bool d8a = __fsub_rn(x[7], rmax[2]) < -0.04f, d11 = __fsub_rn(x[8], rmax[3]) < -0.04f;   // NaN compares false
  bool dark = d11 && d8a && (x[0] < 0.03f) && (x[7] < 0.18f) && !wpos;
  sh = (sh || dark) && !wpos;
  float rgbsum = __fadd_rn(__fadd_rn(x[0], x[1]), x[2]);
  bool slope = d8a && d11 && (x[0] < 0.07f) && (x[7] < 0.18f) && (rgbsum < 0.28f) && !wpos && (dem[p] >= 25.f);
  sh = sh || slope;
  const float* ar = s.allref + p * 4;
  bool wsh = (__fsub_rn(x[0], ar[0]) < -0.05f) && (__fsub_rn(x[1], ar[1]) < -0.05f) && (x[7] < 0.03f) &&
             (__fsub_rn(ar[1], x[1]) > 0.02f) && wpos;
  shadows[(int64_t)t * HW + p] = sh || wsh;
}

// ---------------------------------------------------------------------------------------------
// stage C: cloud references and candidates (:1342-1447)
// ---------------------------------------------------------------------------------------------
struct CloudWin { int others_lo, others_hi; int close[3]; int nclose; };

// Per pixel, ALL dates in one pass: the T x 3 visible-band references are loaded once into registers and every date's
// ri_upper (3), ri_close (3), close_thresh and clouds_i are produced from them (round 1 launched this per date and
// re-read the whole cube T times: 9.6 GB of DRAM traffic at T = 24).  forest: optional [HW] 0/1 mask (ESA WorldCover
// forest, cloud_removal.py:1254-1257), nullptr = no forest anywhere.
struct CloudWinAll { CloudWin w[CT_MAX]; };
__global__ void __launch_bounds__(128) k_cloud_refs(const float* __restrict__ img, const unsigned char* __restrict__ shadows,
                                                    StaticRefs s, const unsigned char* __restrict__ forest, int T, int HW,
                                                    const __grid_constant__ CloudWinAll wins,
                                                    float* __restrict__ rc_out /*[T][HW][3]*/, float* __restrict__ thr_out /*[T][HW]*/,
                                                    unsigned char* __restrict__ ci_out /*[T][HW]*/, int* __restrict__ ci_count /*[T]*/) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = p < HW;
  if (!live) p = HW - 1;                              // keep the warp whole for the ballots below
  float ref[CT_MAX][3];
  float own[3];
  for (int tt = 0; tt < T; ++tt) {
    const float* q = img + ((int64_t)tt * HW + p) * 10;
    bool shd = (T > 2) && shadows[(int64_t)tt * HW + p];
    for (int k = 0; k < 3; ++k) ref[tt][k] = shd ? nanf("") : q[k];
  }
  const bool in_forest = forest && forest[p] == 1;
  for (int t = 0; t < T; ++t) {
    const CloudWin& w = wins.w[t];
    float up[3], rc[3];
    if (T > 2) {
      for (int k = 0; k < 3; ++k) {
        float m = nanf("");
        for (int tt = w.others_lo; tt < w.others_hi; ++tt) { float x = ref[tt][k]; if (!isnan(x)) m = isnan(m) ? x : fminf(m, x); }
        up[k] = m;
      }
      if (isnan(up[0])) for (int k = 0; k < 3; ++k) up[k] = s.p25[p * 3 + k];
      for (int k = 0; k < 3; ++k) {
        float m = nanf("");
        for (int c = 0; c < w.nclose; ++c) { int ci_ = w.close[c]; if (ci_ < 0) ci_ += T;   /* Python negative index */
          float x = ref[ci_][k]; if (!isnan(x)) m = isnan(m) ? x : fminf(m, x); }
        rc[k] = m;
      }
      // progressive widening (:1385-1394): the image-level `any NaN` test only bounds the number of
      // rounds; per pixel the value is the first widened window that holds a valid date
      int lo_i = w.close[0], hi_i = w.close[w.nclose - 1];
      for (int it = 0; it < 10; ++it) {
        lo_i = lo_i - 1 > 0 ? lo_i - 1 : 0; hi_i = hi_i + 1 < T ? hi_i + 1 : T;
        for (int k = 0; k < 3; ++k) {
          if (!isnan(rc[k])) continue;
          float m = nanf("");
          for (int tt = lo_i; tt < hi_i; ++tt) { if (tt == t) continue; float x = ref[tt][k]; if (!isnan(x)) m = isnan(m) ? x : fminf(m, x); }
          rc[k] = m;
        }
      }
      for (int k = 0; k < 3; ++k) if (isnan(rc[k])) rc[k] = s.minrgb[p * 3 + k];
    } else {
      for (int k = 0; k < 3; ++k) { float m = INFINITY; for (int tt = 0; tt < T; ++tt) m = fminf(m, ref[tt][k]); rc[k] = m; up[k] = m; }
    }
    float thr = __fadd_rn(__fdiv_rn(__fdiv_rn(rc[0], 0.02f), 100.f), 0.005f);
    thr = fminf(thr, 0.10f); thr = fmaxf(thr, 0.05f);
    if (in_forest) thr = __fsub_rn(thr, 0.02f);       // close_thresh[forest_mask == 1] -= 0.02 (:1415)
    thr = fmaxf(thr, 0.04f);
    { const float* q = img + ((int64_t)t * HW + p) * 10; own[0] = q[0]; own[1] = q[1]; own[2] = q[2]; }     // L1 hit: read above
    const bool ci = (__fsub_rn(own[0], up[0]) > 0.08f) && (__fsub_rn(own[1], up[1]) > 0.08f) && (__fsub_rn(own[2], up[2]) > 0.07f);
    const unsigned bal = __ballot_sync(0xffffffffu, live && ci);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(ci_count + t, __popc(bal));
    if (live) {
      const int64_t i = (int64_t)t * HW + p;
      ci_out[i] = ci;
      rc_out[i * 3 + 0] = rc[0]; rc_out[i * 3 + 1] = rc[1]; rc_out[i * 3 + 2] = rc[2];
      thr_out[i] = thr;
    }
  }
}

// clouds_close of every still-active date for its current modifier (float32 arithmetic of `thr + mod + 0.01`), + counts
struct CloseMods { float mod[CT_MAX]; int active[CT_MAX]; };
__global__ void __launch_bounds__(256) k_cloud_close(const float* __restrict__ img, int HW, const float* __restrict__ rc_all,
                                                     const float* __restrict__ thr_all, const __grid_constant__ CloseMods m,
                                                     unsigned char* __restrict__ cc_all, int* __restrict__ count /*[T]*/) {
  const int t = blockIdx.y;
  if (!m.active[t]) return;
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned char v = 0;
  if (p < HW) {
    const int64_t i = (int64_t)t * HW + p;
    const float* x = img + i * 10;
    const float* rc = rc_all + i * 3;
    const float mod = m.mod[t];
    float a = __fadd_rn(__fadd_rn(thr_all[i], mod), 0.01f), b = __fadd_rn(thr_all[i], mod);
    v = (__fsub_rn(x[0], rc[0]) > a) && (__fsub_rn(x[1], rc[1]) > a) && (__fsub_rn(x[2], rc[2]) > b);
    cc_all[i] = v;
  }
  unsigned bal = __ballot_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(count + t, __popc(bal));
}

__global__ void __launch_bounds__(256) k_count(const unsigned char* __restrict__ m, int n, int* __restrict__ count) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned char v = p < n ? m[p] : 0;
  unsigned bal = __ballot_sync(0xffffffffu, v != 0);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(count, __popc(bal));
}

__global__ void __launch_bounds__(256) k_count_dates(const unsigned char* __restrict__ m, int HW, int* __restrict__ count) {
  int p = blockIdx.x * blockDim.x + threadIdx.x; const int t = blockIdx.y;
  unsigned char v = p < HW ? m[(int64_t)t * HW + p] : 0;
  unsigned bal = __ballot_sync(0xffffffffu, v != 0);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(count + t, __popc(bal));
}

// cc &= (sum rgb < 0.75), every date
__global__ void __launch_bounds__(256) k_cc_bright(const float* __restrict__ img, int64_t N, unsigned char* __restrict__ cc) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float* x = img + i * 10;
  cc[i] = cc[i] && (__fadd_rn(__fadd_rn(x[0], x[1]), x[2]) < 0.75f);
}
// clouds = max(clouds_i, clouds_close) where clouds_close is the eroded version outside forest (:1443-1447)
__global__ void __launch_bounds__(256) k_clouds_join(const unsigned char* __restrict__ ci, const unsigned char* __restrict__ cc,
                                                     const unsigned char* __restrict__ cc_eroded, const unsigned char* __restrict__ forest,
                                                     int HW, int64_t N, unsigned char* __restrict__ clouds) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const bool in_forest = forest && forest[i % HW] != 0;
  clouds[i] = ci[i] | (in_forest ? cc[i] : cc_eroded[i]);
}
__global__ void __launch_bounds__(256) k_or(const unsigned char* a, const unsigned char* b, unsigned char* o, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = a[i] | b[i];
}

// ---------------------------------------------------------------------------------------------
// stage D: brightness z-score clouds (:1458-1481) and whiteness filter (:1484-1492)
// ---------------------------------------------------------------------------------------------
// np.nanmedian of the brightness (B2 + B3 + B4) over the pixels with clouds == 0 && shadows == 0 of every date: the
// selected values (others: NaN, which sorts last) go through the GPU-wide radix select of stc_select.cu -- round 1 ran
// one block per date (24 blocks on 148 SMs, 3.2 ms).
__global__ void __launch_bounds__(256) k_bright_vals(const float* __restrict__ img, const unsigned char* __restrict__ clouds,
                                                     const unsigned char* __restrict__ shadows, int HW, float* __restrict__ vals,
                                                     int* __restrict__ nvalid) {
  const int t = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false; float v = 0.f;
  if (p < HW) {
    const int64_t i = (int64_t)t * HW + p;
    ok = !clouds[i] && !shadows[i];
    const float* x = img + i * 10;
    v = __fadd_rn(__fadd_rn(x[0], x[1]), x[2]);
    ok = ok && !isnan(v);
    vals[i] = ok ? v : __uint_as_float(0x7fc00000u);
  }
  const unsigned bal = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(nvalid + t, __popc(bal));
}
__global__ void k_median_ks(const int* __restrict__ nvalid, int T, int* __restrict__ ks) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) ks[t * SEL_MAX_COLS] = nvalid[t] > 0 ? (nvalid[t] - 1) / 2 : 0;
}
__global__ void k_median_finish(const float* __restrict__ pairs, const int* __restrict__ nvalid, int T, float* __restrict__ med) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int n = nvalid[t];
  const float a = pairs[(t * SEL_MAX_COLS) * 2], b = pairs[(t * SEL_MAX_COLS) * 2 + 1];
  med[t] = n == 0 ? nanf("") : (n & 1) ? a : __fmul_rn(__fadd_rn(a, b), 0.5f);
}

// ---- NumPy-exact float32 mean / std of a masked selection ------------------------------------
// np.mean / np.std / np.nanmean / np.nanstd on `arr[mask]` (a compacted, contiguous float32 vector)
// add with NumPy's pairwise summation: blocks of <= 128 elements summed with 8 strided accumulators,
// halves split at a multiple of 8.  To be bit-identical the selection is compacted in row-major order
// (k_compact), the recursion's leaves are enumerated, leaf sums are computed in parallel and then
// combined in recursion order.
// kind 0: brightness ratio over clouds==0 (or every pixel when all_px[t]); NaN -> 0 and not counted
// kind 1: 1/blue over clouds==0      kind 2: mean rgb over clouds==0     kind 3: ptp(rgb) over clouds==0
// The selection of one date is compacted in row-major order by all SMs: per 1024-pixel chunk the selected / valid counts
// (k_compact_count), an exclusive scan of the chunk counts per date (k_compact_scan), and the ordered scatter
// (k_compact_scatter).  Round 1 walked each date with ONE block (24 blocks on 148 SMs, 0.8 ms per call).
__device__ __forceinline__ bool compact_value(const float* __restrict__ img, const unsigned char* __restrict__ clouds,
                                              const float* __restrict__ water, const float* __restrict__ medb, bool everything, int HW,
                                              int kind, int t, int p, float& v, bool& valid) {
  valid = false; v = 0.f;
  if (p >= HW) return false;
  const bool sel = everything || !clouds[(int64_t)t * HW + p];
  if (!sel) return false;
  const float* x = img + ((int64_t)t * HW + p) * 10;
  if (kind == 0) { float r = __fdiv_rn(__fadd_rn(__fadd_rn(x[0], x[1]), x[2]), medb[t]); v = (water[p] > 0.f) ? 1.f : r; }
  else if (kind == 1) v = __fdiv_rn(1.f, x[0]);
  else if (kind == 2) v = __fdiv_rn(__fadd_rn(__fadd_rn(x[0], x[1]), x[2]), 3.f);
  else v = __fsub_rn(fmaxf(fmaxf(x[0], x[1]), x[2]), fminf(fminf(x[0], x[1]), x[2]));
  valid = true;
  if (kind == 0 && isnan(v)) { v = 0.f; valid = false; }
  return true;
}
__global__ void __launch_bounds__(1024) k_compact_count(const float* __restrict__ img, const unsigned char* __restrict__ clouds,
                                                        const float* __restrict__ water, const float* __restrict__ medb,
                                                        const int* __restrict__ all_px, int HW, int kind, int2* __restrict__ chunk_cnt) {
  const int t = blockIdx.y, p = blockIdx.x * 1024 + threadIdx.x;
  const bool everything = (kind == 0) && all_px && all_px[t];
  float v; bool valid;
  const bool sel = compact_value(img, clouds, water, medb, everything, HW, kind, t, p, v, valid);
  __shared__ int s_sel, s_val;
  if (threadIdx.x == 0) { s_sel = 0; s_val = 0; }
  __syncthreads();
  const unsigned bal = __ballot_sync(0xffffffffu, sel), balv = __ballot_sync(0xffffffffu, valid);
  if ((threadIdx.x & 31) == 0) { if (bal) atomicAdd(&s_sel, __popc(bal)); if (balv) atomicAdd(&s_val, __popc(balv)); }
  __syncthreads();
  if (threadIdx.x == 0) chunk_cnt[(int64_t)t * gridDim.x + blockIdx.x] = make_int2(s_sel, s_val);
}
__global__ void __launch_bounds__(1024) k_compact_scan(const int2* __restrict__ chunk_cnt, int chunks, int* __restrict__ chunk_base,
                                                       int* __restrict__ cnts /*[T][2]: slots, valid*/) {
  const int t = blockIdx.x;
  __shared__ int wtot[32]; __shared__ int carry, vsum;
  if (threadIdx.x == 0) { carry = 0; vsum = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int c0 = 0; c0 < chunks; c0 += 1024) {
    const int c = c0 + threadIdx.x;
    const int2 v = c < chunks ? chunk_cnt[(int64_t)t * chunks + c] : make_int2(0, 0);
    int incl = v.x;
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    int vs = v.y;
    for (int o = 16; o > 0; o >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, o);
    if (lane == 31) wtot[wid] = incl;
    if (lane == 0 && vs) atomicAdd(&vsum, vs);
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < wid; ++w) woff += wtot[w];
    if (c < chunks) chunk_base[(int64_t)t * chunks + c] = carry + woff + incl - v.x;
    __syncthreads();
    if (threadIdx.x == 0) { int sum = 0; for (int w = 0; w < 32; ++w) sum += wtot[w]; carry += sum; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { cnts[2 * t] = carry; cnts[2 * t + 1] = vsum; }
}
__global__ void __launch_bounds__(1024) k_compact_scatter(const float* __restrict__ img, const unsigned char* __restrict__ clouds,
                                                          const float* __restrict__ water, const float* __restrict__ medb,
                                                          const int* __restrict__ all_px, int HW, int kind,
                                                          const int* __restrict__ chunk_base, float* __restrict__ vals) {
  const int t = blockIdx.y, p = blockIdx.x * 1024 + threadIdx.x;
  const bool everything = (kind == 0) && all_px && all_px[t];
  float v; bool valid;
  const bool sel = compact_value(img, clouds, water, medb, everything, HW, kind, t, p, v, valid);
  __shared__ int wtot[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, sel);
  if (lane == 0) wtot[wid] = __popc(bal);
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < wid; ++w) woff += wtot[w];
  if (sel) vals[(int64_t)t * HW + chunk_base[(int64_t)t * gridDim.x + blockIdx.x] + woff + __popc(bal & ((1u << lane) - 1u))] = v;
}

__device__ __forceinline__ float np_leaf_sum(const float* a, int n, int pass, float mean) {
  auto g = [&](int i) { float v = a[i]; if (pass) { float d = __fsub_rn(v, mean); v = __fmul_rn(d, d); } return v; };
  if (n < 8) { float r = 0.f; for (int i = 0; i < n; ++i) r = __fadd_rn(r, g(i)); return r; }
  float r[8];
  for (int k = 0; k < 8; ++k) r[k] = g(k);
  int i = 8;
  for (; i < n - (n % 8); i += 8) for (int k = 0; k < 8; ++k) r[k] = __fadd_rn(r[k], g(i + k));
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])), __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, g(i));
  return res;
}
// out[t] = {mean, std} as float32 exactly as NumPy computes them (finite inputs; a date whose
// brightness median is NaN has no valid slot and yields NaN, as np.nanmean / np.nanstd do).
// np.add.reduce over a contiguous float32 vector is a recursion: n > 128 splits into (n/2 rounded down to a multiple
// of 8, rest), leaves use 8 strided accumulators.  The recursion tree depends only on n, so it is built breadth first
// (k_np_tree: one block per date, one thread per node, block-wide scan per level); the leaves are summed by the WHOLE GPU
// (k_np_leaves: eight lanes per leaf = NumPy's eight accumulators, so a leaf is read in coalesced 32-byte steps and the
// lanes are combined in NumPy's order); k_np_up combines the inner nodes level by level from the bottom, left + right.
// History: replaying the recursion on one thread took 3.4 ms per call; one block per date doing everything (each thread
// walking its own 512-byte leaf) 0.22 ms; split like this the two leaf passes use 148 SMs instead of 12-24.
struct NpTree { int lvl_start[40]; int n_lvl; int total; };
__global__ void __launch_bounds__(1024) k_np_tree(const int* __restrict__ cnts, int node_cap, int2* __restrict__ nodes_all,
                                                  int* __restrict__ child_all, NpTree* __restrict__ trees, float* __restrict__ out /*[T][2]*/) {
  const int t = blockIdx.x;
  int2* node = nodes_all + (int64_t)t * node_cap;        // (offset, length)
  int* child = child_all + (int64_t)t * node_cap;        // index of the left child (right = +1), -1 for a leaf
  const int n = cnts[2 * t], nvalid = cnts[2 * t + 1];
  __shared__ int lvl_start[40]; __shared__ int n_lvl; __shared__ int wtot[32]; __shared__ int base_s;
  if (nvalid == 0) {
    if (threadIdx.x == 0) { out[2 * t] = nanf(""); out[2 * t + 1] = nanf(""); trees[t].n_lvl = 0; trees[t].total = 0; }
    return;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) { node[0] = make_int2(0, n); lvl_start[0] = 0; lvl_start[1] = 1; n_lvl = 1; }
  __syncthreads();
  for (int d = 0; d < 38; ++d) {
    const int ls = lvl_start[d], le = lvl_start[d + 1];
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    for (int i0 = ls; i0 < le; i0 += 1024) {
      const int i = i0 + threadIdx.x;
      int2 f = make_int2(0, 0);
      bool inner = false;
      if (i < le) { f = node[i]; inner = f.y > 128; }
      const unsigned bal = __ballot_sync(0xffffffffu, inner);
      if (lane == 0) wtot[wid] = __popc(bal);
      __syncthreads();
      int woff = 0;
      for (int w = 0; w < wid; ++w) woff += wtot[w];
      if (i < le) {
        if (inner) {
          const int c = le + 2 * (base_s + woff + __popc(bal & ((1u << lane) - 1u)));
          int n2 = f.y / 2; n2 -= n2 % 8;
          child[i] = c;
          node[c] = make_int2(f.x, n2);
          node[c + 1] = make_int2(f.x + n2, f.y - n2);
        } else child[i] = -1;
      }
      __syncthreads();
      if (threadIdx.x == 0) { int sum = 0; for (int w = 0; w < 32; ++w) sum += wtot[w]; base_s += sum; }
      __syncthreads();
    }
    if (threadIdx.x == 0) { lvl_start[d + 2] = le + 2 * base_s; if (base_s > 0) n_lvl = d + 2; }
    __syncthreads();
    if (base_s == 0) break;
  }
  if (threadIdx.x < 40) trees[t].lvl_start[threadIdx.x] = lvl_start[threadIdx.x];
  if (threadIdx.x == 0) { trees[t].n_lvl = n_lvl; trees[t].total = lvl_start[n_lvl]; }
}
// pass 0: leaf sums of the values; pass 1: of (value - mean)^2 with the mean k_np_up stored in out[2 t]
__global__ void __launch_bounds__(256) k_np_leaves(const float* __restrict__ vals, int HW, int node_cap, const int2* __restrict__ nodes_all,
                                                   const int* __restrict__ child_all, const NpTree* __restrict__ trees,
                                                   const float* __restrict__ out, int pass, float* __restrict__ val_all) {
  const int t = blockIdx.y;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, k = threadIdx.x & 7;
  const bool live = i < trees[t].total && child_all[(int64_t)t * node_cap + i] < 0;      // the same for the eight lanes of a leaf
  const int2 nd = live ? nodes_all[(int64_t)t * node_cap + i] : make_int2(0, 0);
  const float* a = vals + (int64_t)t * HW + nd.x;
  const int n = nd.y;
  const float mean = pass ? out[2 * t] : 0.f;
  auto g = [&](int idx) { float v = a[idx]; if (pass) { const float d = __fsub_rn(v, mean); v = __fmul_rn(d, d); } return v; };
  // NumPy's leaf: r[k] = a[k] + a[8 + k] + ... (lane k), then ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)), then the
  // n % 8 trailing values one by one; fewer than eight values: a plain running sum.  The shuffles sit outside every branch
  // (groups of one warp differ in n).
  float r = 0.f;
  const int full = n - (n % 8);
  if (n >= 8) {
    r = g(k);
    for (int j = 8; j < full; j += 8) r = __fadd_rn(r, g(j + k));
  }
  r = __fadd_rn(r, __shfl_down_sync(0xffffffffu, r, 1, 8));
  r = __fadd_rn(r, __shfl_down_sync(0xffffffffu, r, 2, 8));
  r = __fadd_rn(r, __shfl_down_sync(0xffffffffu, r, 4, 8));
  float res = r;
  if (k == 0) for (int j = (n >= 8 ? full : 0); j < n; ++j) res = __fadd_rn(res, g(j));
  if (live && k == 0) val_all[(int64_t)t * node_cap + i] = res;
}
__global__ void __launch_bounds__(1024) k_np_up(const int* __restrict__ cnts, int node_cap, const int* __restrict__ child_all,
                                                const NpTree* __restrict__ trees, int pass, float* __restrict__ val_all, float* __restrict__ out) {
  const int t = blockIdx.x;
  const NpTree& tr = trees[t];
  if (tr.total == 0) return;                                  // no valid value: k_np_tree wrote NaN
  const int* child = child_all + (int64_t)t * node_cap;
  float* val = val_all + (int64_t)t * node_cap;
  for (int d = tr.n_lvl - 2; d >= 0; --d) {               // the deepest level holds leaves only
    for (int i = tr.lvl_start[d] + threadIdx.x; i < tr.lvl_start[d + 1]; i += 1024) {
      const int c = child[i];
      if (c >= 0) val[i] = __fadd_rn(val[c], val[c + 1]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float r = (float)((double)val[0] / (double)cnts[2 * t + 1]);     // float32 / intp -> float64 divide -> float32
    if (pass == 0) out[2 * t] = r; else out[2 * t + 1] = __fsqrt_rn(r);
  }
}

// brightness_clouds[t] = (z > 3.5) * (water < 0)
__global__ void __launch_bounds__(256) k_bright_clouds(const float* __restrict__ img, const float* __restrict__ water,
                                                       const float* __restrict__ medb, const float* __restrict__ mom, int T, int HW,
                                                       unsigned char* __restrict__ bc) {
  int p = blockIdx.x * blockDim.x + threadIdx.x; const int t = blockIdx.y;
  if (p >= HW) return;
  const float* x = img + ((int64_t)t * HW + p) * 10;
  float ratio = __fdiv_rn(__fadd_rn(__fadd_rn(x[0], x[1]), x[2]), medb[t]);
  if (water[p] > 0.f) ratio = 1.f;
  float mean = mom[t * 2], sd = mom[t * 2 + 1];
  float z = __fdiv_rn(__fsub_rn(ratio, mean), sd);
  bc[(int64_t)t * HW + p] = (z > 3.5f) && (water[p] < 0.f);
}

// multi = sum_t (bc - clouds) > 0 ; bc[t][multi > 1] = 0 ; clouds = max(clouds, bc) ; whiteness filter
__global__ void __launch_bounds__(256) k_bright_merge(const float* __restrict__ img, const unsigned char* __restrict__ bc, int T, int HW,
                                                      unsigned char* __restrict__ clouds) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  int multi = 0;
  for (int t = 0; t < T; ++t) multi += (bc[(int64_t)t * HW + p] && !clouds[(int64_t)t * HW + p]);
  for (int t = 0; t < T; ++t) {
    unsigned char c = clouds[(int64_t)t * HW + p] | ((multi > 1) ? 0 : bc[(int64_t)t * HW + p]);
    const float* x = img + ((int64_t)t * HW + p) * 10;
    float mb = __fdiv_rn(__fadd_rn(__fadd_rn(x[0], x[1]), x[2]), 3.f);
    float vr = __fsub_rn(fmaxf(fmaxf(x[0], x[1]), x[2]), fminf(fminf(x[0], x[1]), x[2]));
    bool fp = (mb < 0.4f) && (__fdiv_rn(vr, mb) > 0.5f);
    clouds[(int64_t)t * HW + p] = c && !fp;
  }
}

// ---------------------------------------------------------------------------------------------
// stage E: false-positive removal (:1514-1551)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_nsr(const float* __restrict__ img, unsigned char* __restrict__ nsr, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = img + i * 10;
  nsr[i] = __fdiv_rn(p[3], __fadd_rn(p[8], 0.01f)) < 0.75f;
}
// nsr[t][water<0] = 0 ; clouds[t][nsr & isnt_cloud] = 0 ; water/B11 candidate for the next stage
__global__ void __launch_bounds__(256) k_fp1(const float* __restrict__ img, const float* __restrict__ water, int T, int HW,
                                             unsigned char* __restrict__ nsr, unsigned char* __restrict__ clouds,
                                             unsigned char* __restrict__ wfp) {
  int p = blockIdx.x * blockDim.x + threadIdx.x; const int t = blockIdx.y;
  if (p >= HW) return;
  const int lo = t - 1 > 0 ? t - 1 : 0, hi = t + 2 < T ? t + 2 : T;
  float bmin = INFINITY;
  for (int tt = lo; tt < hi; ++tt) { const float* q = img + ((int64_t)tt * HW + p) * 10; bmin = fminf(bmin, fminf(fminf(q[0], q[1]), q[2])); }
  const float* x = img + ((int64_t)t * HW + p) * 10;
  float bi = __fdiv_rn(__fadd_rn(__fadd_rn(x[0], x[1]), x[2]), 3.f);
  bool isnt = __fsub_rn(bi, bmin) < 0.4f;
  int64_t i = (int64_t)t * HW + p;
  unsigned char ns = nsr[i]; if (water[p] < 0.f) ns = 0; nsr[i] = ns;
  if (ns && isnt) clouds[i] = 0;
  wfp[i] = (water[p] > 0.f) && (x[8] < 0.11f);
}
__global__ void __launch_bounds__(256) k_clear_where(unsigned char* __restrict__ clouds, const unsigned char* __restrict__ m, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && m[i]) clouds[i] = 0;
}
__global__ void __launch_bounds__(256) k_winsum_lt(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, int T, int H,
                                                   int W, int thresh) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)T * H * W) return;
  int x = (int)(idx % W); int64_t r = idx / W; int y = (int)(r % H); int t = (int)(r / H);
  const unsigned char* m = in + (int64_t)t * H * W;
  out[idx] = (winsum3(m, y, x, H, W) < thresh) ? 0 : m[(int64_t)y * W + x];
}
__global__ void __launch_bounds__(256) k_dark(const float* __restrict__ img, unsigned char* __restrict__ o, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = img + i * 10;
  o[i] = __fadd_rn(__fadd_rn(p[0], p[1]), p[2]) < 0.21f;
}
// `clouds[i][brightness_threshold.astype(uint8)] = 0` (:1546-1551) is integer fancy indexing: it zeroes
// ROW 0 of the date if any mask pixel is 0 and ROW 1 if any is 1.  flags[t] = {any zero, any one}.
__global__ void __launch_bounds__(256) k_any01(const unsigned char* __restrict__ m, int HW, int* __restrict__ flags) {
  int p = blockIdx.x * blockDim.x + threadIdx.x; const int t = blockIdx.y;
  if (p >= HW) return;
  if (m[(int64_t)t * HW + p]) flags[2 * t + 1] = 1; else flags[2 * t] = 1;
}
__global__ void __launch_bounds__(256) k_zero_rows(unsigned char* __restrict__ clouds, const int* __restrict__ flags, int H, int W) {
  int x = blockIdx.x * blockDim.x + threadIdx.x; const int t = blockIdx.y;
  if (x >= W) return;
  if (flags[2 * t]) clouds[((int64_t)t * H + 0) * W + x] = 0;
  if (flags[2 * t + 1] && H > 1) clouds[((int64_t)t * H + 1) * W + x] = 0;
}

// clouds / shadows of date t are cleared where fcps > 0 and the pixel is not much brighter than its darkest neighbour date
// (:1500-1511)
__global__ void __launch_bounds__(256) k_fcps_remove(const float* __restrict__ img, const unsigned char* __restrict__ fcps, int T, int HW,
                                                     unsigned char* __restrict__ clouds, unsigned char* __restrict__ shadows) {
  int p = blockIdx.x * blockDim.x + threadIdx.x; const int t = blockIdx.y;
  if (p >= HW) return;
  const int64_t i = (int64_t)t * HW + p;
  if (!fcps[i]) return;
  const int lo = t - 1 > 0 ? t - 1 : 0, hi = t + 2 < T ? t + 2 : T;
  float bmin = INFINITY;
  for (int tt = lo; tt < hi; ++tt) { const float* q = img + ((int64_t)tt * HW + p) * 10; bmin = fminf(bmin, fminf(fminf(q[0], q[1]), q[2])); }
  const float* x = img + i * 10;
  const float bi = __fdiv_rn(__fadd_rn(__fadd_rn(x[0], x[1]), x[2]), 3.f);
  if (__fsub_rn(bi, bmin) < 0.4f) { clouds[i] = 0; shadows[i] = 0; }
}
// brightness_threshold * (1 - forest_mask) (:1549)
__global__ void __launch_bounds__(256) k_clear_forest(unsigned char* __restrict__ m, const unsigned char* __restrict__ forest, int HW, int64_t N) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N && forest[i % HW] != 0) m[i] = 0;       // 1 - forest is 0 where forest == 1 (the raster is 0/1)
}
// urban / non-urban split of the eroded clouds (:1592-1597): pf5 [HW] is the dilated potential-false-positive mask
__global__ void __launch_bounds__(256) k_split_urban(const unsigned char* __restrict__ c, const unsigned char* __restrict__ pf5, int HW, int64_t N,
                                                     unsigned char* __restrict__ urban, unsigned char* __restrict__ nonurban) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const bool u = pf5[i % HW] != 0;
  urban[i] = c[i] && u; nonurban[i] = c[i] && !u;
}
// clouds = non_urban + urban (:1612) is 2 where both are set; the 2s only matter for the image means of the next two
// stages (they are clipped at :1649), so they are kept as a second mask: clouds = a | b, two = a & b
__global__ void __launch_bounds__(256) k_sum_masks(const unsigned char* __restrict__ a, const unsigned char* __restrict__ b, int64_t N,
                                                   unsigned char* __restrict__ clouds, unsigned char* __restrict__ two) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  clouds[i] = a[i] | b[i]; two[i] = a[i] & b[i];
}

// ---------------------------------------------------------------------------------------------
// stage F: shape clean-up (:1590-1612)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_split_size(const unsigned char* __restrict__ c, unsigned char* __restrict__ large,
                                                    unsigned char* __restrict__ small, int T, int H, int W) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)T * H * W) return;
  int x = (int)(idx % W); int64_t r = idx / W; int y = (int)(r % H); int t = (int)(r / H);
  const unsigned char* m = c + (int64_t)t * H * W;
  int ws = winsum3(m, y, x, H, W);
  unsigned char v = m[(int64_t)y * W + x];
  large[idx] = (ws < 6) ? 0 : v;
  small[idx] = (ws >= 6) ? 0 : v;
}
__global__ void __launch_bounds__(256) k_and_or_dem(unsigned char* __restrict__ shadows, const unsigned char* __restrict__ near,
                                                    const float* __restrict__ dem, int HW) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  if (!(near[p] || dem[p] >= 30.f)) shadows[p] = 0;
}
// dark-blue shadow candidates: 1/B2 > ref && B8A < 0.17 (:1638-1648)
// all dates in one launch (blockIdx.y = date): inactive dates (np.mean(clouds) >= 0.9, :1641) give an empty candidate mask,
// which the opening that follows leaves empty and the final OR ignores -- the same as skipping the date
__global__ void __launch_bounds__(256) k_darkblue_all(const float* __restrict__ img, int HW, const float* __restrict__ refs,
                                                      const int* __restrict__ active, unsigned char* __restrict__ o) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
  if (p >= HW) return;
  const float* x = img + ((int64_t)t * HW + p) * 10;
  o[(int64_t)t * HW + p] = active[t] && (__fdiv_rn(1.f, x[0]) > refs[t]) && (x[7] < 0.17f);
}
__global__ void __launch_bounds__(256) k_or_nowater_all(unsigned char* __restrict__ clouds, const unsigned char* __restrict__ s,
                                                        const float* __restrict__ water, int HW) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const int64_t i = (int64_t)blockIdx.y * HW + p;
  if (s[i] && !(water[p] > 0.f)) clouds[i] = 1;
}
__global__ void __launch_bounds__(256) k_fill(unsigned char* p, int64_t n, unsigned char v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void __launch_bounds__(256) k_to_float(const unsigned char* __restrict__ in, float* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] ? 1.f : 0.f;
}

float median_host(std::vector<float> v) {            // np.median of a short float32 list (NaN propagates)
  for (float x : v) if (x != x) return x;
  std::sort(v.begin(), v.end());
  size_t n = v.size();
  volatile float s = (n & 1) ? v[n / 2] : v[n / 2 - 1] + v[n / 2];
  return (n & 1) ? s : s / 2.f;
}

}  // namespace

#include <algorithm>

// windows of cloud_removal.py:1266-1273 and :1343-1361 (integer logic, host)
static void shadow_window(int t, int T, int& lo, int& hi) {
  lo = t - 4 > 0 ? t - 4 : 0; hi = t + 3 < T ? t + 3 : T;
  if (hi - lo == 3) { if (hi == T) lo = lo - 1 > 0 ? lo - 1 : 0; if (lo == 0) hi = hi + 1 < T ? hi + 1 : T; }
}
static CloudWin cloud_window(int t, int T) {
  CloudWin w; memset(&w, 0, sizeof(w));
  int lo = t - 2 > 0 ? t - 2 : 0, hi = t + 3 < T ? t + 3 : T;
  if (hi - lo == 3) { if (hi == T) lo = lo - 2 > 0 ? lo - 2 : 0; if (lo == 0) hi = hi + 2 < T ? hi + 2 : T; }
  w.others_lo = lo; w.others_hi = hi;
  int c0 = t - 1 > 0 ? t - 1 : 0, c1 = t + 1 < T - 1 ? t + 1 : T - 1;
  if (c1 - c0 < 2) { if (c0 == 0) { c0 += 1; c1 += 1; } else { c0 -= 1; c1 -= 1; } }
  if (c1 >= T - 2 && T > 3) { w.close[0] = c0 - 1; w.close[1] = c0; w.close[2] = c1; w.nclose = 3; }
  else { w.close[0] = c0; w.close[1] = c1; w.nclose = 2; }
  return w;
}

// shared with stc_cloudfill.cu: binary dilation / erosion-by-dilation on [frames][H][W] uint8 masks
void maskop_dilate(stc_ctx* ctx, const unsigned char* in, unsigned char* out, int frames, int H, int W, int k, int conn, int inv_in,
                   int inv_out, int three_d) {
  morph_dilate_dev(ctx, in, out, frames, H, W, k, conn, inv_in, inv_out, three_d);
}

#define LAUNCH1D(kern, n, ...) do { TraceScope ts_(ctx, #kern); kern<<<cdiv((n), 256), 256, 0, ctx->stream>>>(__VA_ARGS__); ctx->launches++; } while (0)

// Device-resident core: img_dev [T,H,W,10] float32, dem_dev [H,W]; clouds_dev [T,H,W] float32, fcps_dev [T,H,W] uint8.
// Returns with its kernels enqueued on ctx->stream (it synchronises internally where the reference's control flow needs
// scalars on the host: the adaptive threshold loop, the plausibility tests, the haze flags).
int cloud_masks_dev(stc_ctx* ctx, const float* img, const float* dem, int T, int H, int W, const unsigned char* forest_dev,
                    const unsigned char* urban_core_dev, const unsigned char* urban_near_dev, float* clouds_dev, unsigned char* fcps_dev,
                    uint8_t* stage_host, int stage_id) {
  if (!ctx) return STC_ERR_ARG;
  if (!img || !dem || !clouds_dev || !fcps_dev || T < 1 || T > CT_MAX || H < 3 || W < 3)
    STC_FAIL(STC_ERR_ARG, "cloud_masks: bad argument (1 <= T <= 32)");
  const int HW = H * W; const int64_t N = (int64_t)T * HW;
  Buf d_clm, d_a, d_b, d_c, d_sh, d_cl, d_bc, d_nsr, d_water, d_allref, d_minb4, d_p25, d_minrgb, d_rc, d_thr,
      d_ci, d_cc, d_cnt, d_win, d_med, d_mom, d_flags, d_all, d_vals, d_leaves, d_leafsum, d_child, d_cnts, d_ccnt, d_cbase, d_trees;
  for (Buf* b : {&d_clm, &d_a, &d_b, &d_c, &d_sh, &d_cl, &d_bc, &d_nsr}) STC_CUDA(stc_dmalloc(&b->p, N));
  STC_CUDA(stc_dmalloc(&d_water.p, HW * 4)); STC_CUDA(stc_dmalloc(&d_allref.p, HW * 16)); STC_CUDA(stc_dmalloc(&d_minb4.p, HW * 16));
  STC_CUDA(stc_dmalloc(&d_p25.p, HW * 12)); STC_CUDA(stc_dmalloc(&d_minrgb.p, HW * 12)); STC_CUDA(stc_dmalloc(&d_rc.p, HW * 12));
  STC_CUDA(stc_dmalloc(&d_thr.p, HW * 4)); STC_CUDA(stc_dmalloc(&d_ci.p, HW)); STC_CUDA(stc_dmalloc(&d_cc.p, HW));
  STC_CUDA(stc_dmalloc(&d_cnt.p, 64)); STC_CUDA(stc_dmalloc(&d_win.p, 2 * CT_MAX * 4)); STC_CUDA(stc_dmalloc(&d_med.p, CT_MAX * 4));
  STC_CUDA(stc_dmalloc(&d_mom.p, CT_MAX * 2 * 4)); STC_CUDA(stc_dmalloc(&d_flags.p, 2 * CT_MAX * 4)); STC_CUDA(stc_dmalloc(&d_all.p, CT_MAX * 4));
  const int node_cap = 2 * (HW / 56 + 8) + 2;            // a pairwise leaf holds 58..128 values; a binary tree has < 2 x leaves nodes
  STC_CUDA(stc_dmalloc(&d_ccnt.p, (size_t)T * cdiv(HW, 1024) * 8)); STC_CUDA(stc_dmalloc(&d_cbase.p, (size_t)T * cdiv(HW, 1024) * 4));
  STC_CUDA(stc_dmalloc(&d_vals.p, N * 4)); STC_CUDA(stc_dmalloc(&d_leaves.p, (size_t)T * node_cap * 8));
  STC_CUDA(stc_dmalloc(&d_trees.p, (size_t)CT_MAX * sizeof(NpTree))); STC_CUDA(stc_dmalloc(&d_leafsum.p, (size_t)T * node_cap * 4)); STC_CUDA(stc_dmalloc(&d_child.p, (size_t)T * node_cap * 4)); STC_CUDA(stc_dmalloc(&d_cnts.p, CT_MAX * 2 * 4));
  unsigned char *clm = (unsigned char*)d_clm.p, *ta = (unsigned char*)d_a.p, *tb = (unsigned char*)d_b.p, *tc = (unsigned char*)d_c.p,
                *sh = (unsigned char*)d_sh.p, *cl = (unsigned char*)d_cl.p, *bc = (unsigned char*)d_bc.p, *nsr = (unsigned char*)d_nsr.p;
  StaticRefs sr{(float*)d_water.p, (float*)d_allref.p, (float*)d_minb4.p, (float*)d_p25.p, (float*)d_minrgb.p};
  const float* water = sr.water;
  const unsigned char* forest = forest_dev;            // [HW] 0/1 or nullptr (no forestmask.tif: zeros, :1256-1257)
  auto dilate = [&](const unsigned char* in, unsigned char* out, int64_t frames, int k, int conn, int inv_in, int inv_out, int three_d) {
    morph_dilate_dev(ctx, in, out, (int)frames, H, W, k, conn, inv_in, inv_out, three_d);
  };
  auto dump = [&](int id, const unsigned char* src) -> int {     // stage taps for the tests
    if (stage_host && stage_id == id) {
      STC_CUDA(cudaMemcpyAsync(stage_host, src, N, cudaMemcpyDeviceToHost, ctx->stream));
      STC_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return STC_OK;
  };
  // NumPy-exact float32 {mean, std} of f_kind over the clear pixels of every date -> mom_h[t*2..], cnt_h[t*2..]
  std::vector<float> mom_h(2 * CT_MAX); std::vector<int> cnt_h(2 * CT_MAX);
  auto moments = [&](int kind, const float* medb, const int* all_px, bool to_host) -> int {
    const int chunks = cdiv(HW, 1024);
    { TraceScope ts_(ctx, "k_compact_count"); k_compact_count<<<dim3(chunks, T), 1024, 0, ctx->stream>>>(img, cl, water, medb, all_px, HW, kind, (int2*)d_ccnt.p); }
    { TraceScope ts_(ctx, "k_compact_scan"); k_compact_scan<<<T, 1024, 0, ctx->stream>>>((const int2*)d_ccnt.p, chunks, (int*)d_cbase.p, (int*)d_cnts.p); }
    { TraceScope ts_(ctx, "k_compact_scatter"); k_compact_scatter<<<dim3(chunks, T), 1024, 0, ctx->stream>>>(img, cl, water, medb, all_px, HW, kind, (const int*)d_cbase.p, (float*)d_vals.p); }
    ctx->launches += 2;
    {
      const dim3 gl(cdiv((int64_t)node_cap * 8, 256), T);
      const int2* nodes = (const int2*)d_leaves.p; const int* child = (const int*)d_child.p; float* vsum = (float*)d_leafsum.p;
      const NpTree* trees = (const NpTree*)d_trees.p; const int* cn = (const int*)d_cnts.p; float* mom = (float*)d_mom.p;
      { TraceScope ts_(ctx, "k_np_tree"); k_np_tree<<<T, 1024, 0, ctx->stream>>>(cn, node_cap, (int2*)d_leaves.p, (int*)d_child.p, (NpTree*)d_trees.p, mom); }
      for (int pass = 0; pass < 2; ++pass) {
        { TraceScope ts_(ctx, "k_np_leaves"); k_np_leaves<<<gl, 256, 0, ctx->stream>>>((const float*)d_vals.p, HW, node_cap, nodes, child, trees, mom, pass, vsum); }
        { TraceScope ts_(ctx, "k_np_up"); k_np_up<<<T, 1024, 0, ctx->stream>>>(cn, node_cap, child, trees, pass, vsum, mom); }
      }
    }
    ctx->launches += 6;
    if (to_host) {
      STC_CUDA(cudaMemcpyAsync(mom_h.data(), d_mom.p, T * 8, cudaMemcpyDeviceToHost, ctx->stream));
      STC_CUDA(cudaMemcpyAsync(cnt_h.data(), d_cnts.p, T * 8, cudaMemcpyDeviceToHost, ctx->stream));
      STC_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return STC_OK;
  };
  int rc_;

  // ---- A: Hollstein mask (erode 2, dilate 10), static per-pixel references ----
  LAUNCH1D(k_hollstein, N, img, ta, N);
  dilate(ta, tb, T, 2, 1, 1, 1, 0);
  dilate(tb, clm, T, 10, 1, 0, 0, 0);
  {
    TraceScope ts_(ctx, "k_static_refs");
    if (T <= 8) k_static_refs<8><<<cdiv(HW, 128), 128, 0, ctx->stream>>>(img, clm, T, HW, sr);
    else if (T <= 16) k_static_refs<16><<<cdiv(HW, 128), 128, 0, ctx->stream>>>(img, clm, T, HW, sr);
    else k_static_refs<32><<<cdiv(HW, 128), 128, 0, ctx->stream>>>(img, clm, T, HW, sr);
  }
  ctx->launches++;
  if ((rc_ = dump(1, clm))) return rc_;

  // ---- B: shadows ----
  {
    int wl[2 * CT_MAX];
    int wmax = 1;
    for (int t = 0; t < T; ++t) { shadow_window(t, T, wl[t], wl[CT_MAX + t]); wmax = std::max(wmax, wl[CT_MAX + t] - wl[t]); }
    const void* staged = ctx_stage(ctx, wl, sizeof(wl));
    if (!staged) STC_FAIL(STC_ERR_NOMEM, "cloud_masks: pinned staging");
    STC_CUDA(cudaMemcpyAsync(d_win.p, staged, sizeof(wl), cudaMemcpyHostToDevice, ctx->stream));
    {
      TraceScope ts_(ctx, "k_shadow_candidates");
      const dim3 gs(cdiv(HW, 128), T);
      const int* wlo = (const int*)d_win.p; const int* whi = wlo + CT_MAX;
      if (wmax <= 8) k_shadow_candidates<8><<<gs, 128, 0, ctx->stream>>>(img, clm, dem, sr, T, HW, wlo, whi, ta);
      else if (wmax <= 16) k_shadow_candidates<16><<<gs, 128, 0, ctx->stream>>>(img, clm, dem, sr, T, HW, wlo, whi, ta);
      else k_shadow_candidates<32><<<gs, 128, 0, ctx->stream>>>(img, clm, dem, sr, T, HW, wlo, whi, ta);
    }
    ctx->launches++;
    if ((rc_ = dump(2, ta))) return rc_;
    dilate(ta, tb, T, 2, 1, 1, 1, 0);
    dilate(tb, tc, T, 3, 1, 0, 0, 0);
    STC_CUDA(cudaMemsetAsync(d_all.p, 0, CT_MAX * 4, ctx->stream));
    { TraceScope ts_(ctx, "k_count_dates"); k_count_dates<<<dim3(cdiv(HW, 256), T), 256, 0, ctx->stream>>>(tc, HW, (int*)d_all.p); } ctx->launches++;
    if ((rc_ = morph_edt_grow_dev(ctx, tc, sh, T, H, W, 5, (const int*)d_all.p))) return rc_;
    if ((rc_ = dump(3, sh))) return rc_;
  }

  // ---- C: clouds, all dates per launch; the adaptive threshold loop (:1425-1440) advances every still-active date per round ----
  {
    Buf d_rcall, d_thrall, d_ciall, d_ccall, d_cnti, d_cntc;
    STC_CUDA(stc_dmalloc(&d_rcall.p, N * 12)); STC_CUDA(stc_dmalloc(&d_thrall.p, N * 4)); STC_CUDA(stc_dmalloc(&d_ciall.p, N));
    STC_CUDA(stc_dmalloc(&d_ccall.p, N)); STC_CUDA(stc_dmalloc(&d_cnti.p, CT_MAX * 4)); STC_CUDA(stc_dmalloc(&d_cntc.p, CT_MAX * 4));
    CloudWinAll wins; memset(&wins, 0, sizeof(wins));
    for (int t = 0; t < T; ++t) wins.w[t] = cloud_window(t, T);
    STC_CUDA(cudaMemsetAsync(d_cnti.p, 0, CT_MAX * 4, ctx->stream));
    { TraceScope ts_(ctx, "k_cloud_refs"); k_cloud_refs<<<cdiv(HW, 128), 128, 0, ctx->stream>>>(img, sh, sr, forest, T, HW, wins, (float*)d_rcall.p, (float*)d_thrall.p,
                                                                    (unsigned char*)d_ciall.p, (int*)d_cnti.p); }
    ctx->launches++;
    int cnt_i[CT_MAX], cnt_c[CT_MAX];
    STC_CUDA(cudaMemcpyAsync(cnt_i, d_cnti.p, T * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CloseMods cm; memset(&cm, 0, sizeof(cm));
    double mod[CT_MAX];
    for (int t = 0; t < T; ++t) { cm.active[t] = 1; mod[t] = 0.0; }
    bool any = true;
    while (any) {
      for (int t = 0; t < T; ++t) cm.mod[t] = (float)mod[t];
      STC_CUDA(cudaMemsetAsync(d_cntc.p, 0, CT_MAX * 4, ctx->stream));
      { TraceScope ts_(ctx, "k_cloud_close"); k_cloud_close<<<dim3(cdiv(HW, 256), T), 256, 0, ctx->stream>>>(img, HW, (const float*)d_rcall.p, (const float*)d_thrall.p, cm,
                                                                                  (unsigned char*)d_ccall.p, (int*)d_cntc.p); }
      ctx->launches++;
      STC_CUDA(cudaMemcpyAsync(cnt_c, d_cntc.p, T * 4, cudaMemcpyDeviceToHost, ctx->stream));
      STC_CUDA(cudaStreamSynchronize(ctx->stream));
      any = false;
      for (int t = 0; t < T; ++t) {
        if (!cm.active[t]) continue;
        const double mean_i = (double)cnt_i[t] / HW, mean_c = (double)cnt_c[t] / HW;
        mod[t] += 0.0025;
        if (!((mean_c - mean_i) > 0.075)) cm.active[t] = 0; else any = true;
        if (mod[t] > 10.0) STC_FAIL(STC_ERR_STATE, "cloud_masks: threshold loop did not converge");
      }
    }
    LAUNCH1D(k_cc_bright, N, img, N, (unsigned char*)d_ccall.p);
    dilate((const unsigned char*)d_ccall.p, ta, T, 2, 1, 1, 1, 0);              // erode 2 (applies outside forest)
    LAUNCH1D(k_clouds_join, N, (const unsigned char*)d_ciall.p, (const unsigned char*)d_ccall.p, ta, forest, HW, N, cl);
  }
  if ((rc_ = dump(4, cl))) return rc_;

  // ---- D: brightness z-score + whiteness ----
  {
    {
      Buf d_nv, d_ks, d_pairs;
      STC_CUDA(stc_dmalloc(&d_nv.p, CT_MAX * 4)); STC_CUDA(stc_dmalloc(&d_ks.p, (size_t)T * SEL_MAX_COLS * 4)); STC_CUDA(stc_dmalloc(&d_pairs.p, (size_t)T * SEL_MAX_COLS * 8));
      STC_CUDA(cudaMemsetAsync(d_nv.p, 0, CT_MAX * 4, ctx->stream));
      STC_CUDA(cudaMemsetAsync(d_ks.p, 0, (size_t)T * SEL_MAX_COLS * 4, ctx->stream));
      { TraceScope ts_(ctx, "k_bright_vals"); k_bright_vals<<<dim3(cdiv(HW, 256), T), 256, 0, ctx->stream>>>(img, cl, sh, HW, (float*)d_vals.p, (int*)d_nv.p); }
      k_median_ks<<<1, 32, 0, ctx->stream>>>((const int*)d_nv.p, T, (int*)d_ks.p);
      std::vector<SelJob> jobs(T);
      for (int t = 0; t < T; ++t) jobs[t] = SelJob{(const float*)d_vals.p + (int64_t)t * HW, HW, 1, 1};
      if ((rc_ = select_ranks_dev(ctx, jobs.data(), T, (const int*)d_ks.p, (float*)d_pairs.p))) return rc_;
      k_median_finish<<<1, 32, 0, ctx->stream>>>((const float*)d_pairs.p, (const int*)d_nv.p, T, (float*)d_med.p);
      ctx->launches += 3;
    }
    // `if np.sum(clouds[i] < 0.90)`: select clear pixels when any exists, else every pixel
    int allpx[CT_MAX];
    STC_CUDA(cudaMemsetAsync(d_all.p, 0, CT_MAX * 4, ctx->stream));
    { TraceScope ts_(ctx, "k_count_dates"); k_count_dates<<<dim3(cdiv(HW, 256), T), 256, 0, ctx->stream>>>(cl, HW, (int*)d_all.p); } ctx->launches++;
    STC_CUDA(cudaMemcpyAsync(allpx, d_all.p, T * 4, cudaMemcpyDeviceToHost, ctx->stream));
    STC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int t = 0; t < T; ++t) allpx[t] = (allpx[t] == HW);
    STC_CUDA(cudaMemcpyAsync(d_all.p, allpx, T * 4, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc_ = moments(0, (const float*)d_med.p, (const int*)d_all.p, false))) return rc_;
    { TraceScope ts_(ctx, "k_bright_clouds"); k_bright_clouds<<<dim3(cdiv(HW, 256), T), 256, 0, ctx->stream>>>(img, water, (const float*)d_med.p, (const float*)d_mom.p, T, HW, bc); }
    ctx->launches++;
    LAUNCH1D(k_bright_merge, HW, img, bc, T, HW, cl);
    STC_CUDA(cudaStreamSynchronize(ctx->stream));   // allpx is a stack buffer
    if ((rc_ = dump(5, cl))) return rc_;
  }

  // ---- E: false positives ----
  const bool urban = urban_core_dev && urban_near_dev;      // no urbanmask.tif: pfps == 0 -> fcps == 0 (:1133-1135)
  Buf d_fc0, d_pf, d_pf5, d_two;
  if (urban) {
    STC_CUDA(stc_dmalloc(&d_fc0.p, N)); STC_CUDA(stc_dmalloc(&d_pf.p, HW)); STC_CUDA(stc_dmalloc(&d_pf5.p, HW)); STC_CUDA(stc_dmalloc(&d_two.p, N));
    if ((rc_ = pfcp_detect_dev(ctx, img, dem, urban_core_dev, urban_near_dev, T, H, W, (unsigned char*)d_fc0.p, (unsigned char*)d_pf.p))) return rc_;
    if ((rc_ = dump(9, (const unsigned char*)d_fc0.p))) return rc_;
    { TraceScope ts_(ctx, "k_fcps_remove"); k_fcps_remove<<<dim3(cdiv(HW, 256), T), 256, 0, ctx->stream>>>(img, (const unsigned char*)d_fc0.p, T, HW, cl, sh); }
    ctx->launches++;
  } else if (stage_host && stage_id == 9) {
    memset(stage_host, 0, (size_t)N);
  }
  LAUNCH1D(k_nsr, N, img, ta, N);
  dilate(ta, nsr, T, 3, 1, 0, 0, 1);                                            // 3-D dilation (:1518)
  { TraceScope ts_(ctx, "k_fp1"); k_fp1<<<dim3(cdiv(HW, 256), T), 256, 0, ctx->stream>>>(img, water, T, HW, nsr, cl, ta); } ctx->launches++;
  dilate(ta, tb, T, 10, 1, 0, 0, 0);
  LAUNCH1D(k_clear_where, N, cl, tb, N);
  LAUNCH1D(k_winsum_lt, N, cl, ta, T, H, W, 5);
  STC_CUDA(cudaMemcpyAsync(cl, ta, N, cudaMemcpyDeviceToDevice, ctx->stream));
  LAUNCH1D(k_dark, N, img, ta, N);
  dilate(ta, tb, T, 3, 1, 0, 0, 0);
  if (forest) LAUNCH1D(k_clear_forest, N, tb, forest, HW, N);
  STC_CUDA(cudaMemsetAsync(d_flags.p, 0, 2 * CT_MAX * 4, ctx->stream));
  { TraceScope ts_(ctx, "k_any01"); k_any01<<<dim3(cdiv(HW, 256), T), 256, 0, ctx->stream>>>(tb, HW, (int*)d_flags.p); } ctx->launches++;
  { TraceScope ts_(ctx, "k_zero_rows"); k_zero_rows<<<dim3(cdiv(W, 256), T), 256, 0, ctx->stream>>>(cl, (const int*)d_flags.p, H, W); } ctx->launches++;
  if ((rc_ = dump(6, cl))) return rc_;

  // ---- F: shape clean-up: urban clouds eroded by 3, the others grown by size (:1590-1612) ----
  dilate(cl, ta, T, 1, 1, 1, 1, 0);                                             // erode 1
  Buf d_urb;
  if (urban) {
    STC_CUDA(stc_dmalloc(&d_urb.p, N));
    dilate((const unsigned char*)d_pf.p, (unsigned char*)d_pf5.p, 1, 5, 1, 0, 0, 0);      // pfcps[i] dilated 5 (the same raster for every date)
    LAUNCH1D(k_split_urban, N, ta, (const unsigned char*)d_pf5.p, HW, N, tb, tc);          // tb = urban, tc = non-urban
    dilate(tb, (unsigned char*)d_urb.p, T, 3, 1, 1, 1, 0);                                // urban clouds: erode 3
    STC_CUDA(cudaMemcpyAsync(ta, tc, N, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  LAUNCH1D(k_split_size, N, ta, tb, tc, T, H, W);                               // tb = large, tc = small
  dilate(tc, cl, T, 1, 1, 0, 0, 0);
  dilate(tb, ta, T, 5, 1, 0, 0, 0);
  LAUNCH1D(k_or, N, cl, ta, tb, N);
  STC_CUDA(cudaMemsetAsync(d_all.p, 0, CT_MAX * 4, ctx->stream));
  { TraceScope ts_(ctx, "k_count_dates"); k_count_dates<<<dim3(cdiv(HW, 256), T), 256, 0, ctx->stream>>>(tb, HW, (int*)d_all.p); } ctx->launches++;
  if ((rc_ = morph_edt_grow_dev(ctx, tb, cl, T, H, W, 3, (const int*)d_all.p))) return rc_;
  if (urban) {
    STC_CUDA(cudaMemcpyAsync(ta, cl, N, cudaMemcpyDeviceToDevice, ctx->stream));
    LAUNCH1D(k_sum_masks, N, ta, (const unsigned char*)d_urb.p, N, cl, (unsigned char*)d_two.p);
  }
  if ((rc_ = dump(7, cl))) return rc_;

  // ---- G: shadow plausibility (:1617-1626), per date on scalar means ----
  {
    // the counts of every date in three launches and ONE synchronisation (the first version counted, copied and synchronised
    // date by date: 3 n launches and n round trips); a date's shadows change only through its own restrict_far
    Buf d_g;
    STC_CUDA(stc_dmalloc(&d_g.p, 2 * CT_MAX * 4));
    int* g_sh = (int*)d_g.p; int* g_cl = g_sh + CT_MAX;
    STC_CUDA(cudaMemsetAsync(d_g.p, 0, 2 * CT_MAX * 4, ctx->stream));
    const dim3 gd(cdiv(HW, 256), T);
    { TraceScope ts_(ctx, "k_count_dates"); k_count_dates<<<gd, 256, 0, ctx->stream>>>(sh, HW, g_sh); } ctx->launches++;
    { TraceScope ts_(ctx, "k_count_dates"); k_count_dates<<<gd, 256, 0, ctx->stream>>>(cl, HW, g_cl); } ctx->launches++;
    if (urban) { TraceScope ts_(ctx, "k_count_dates"); k_count_dates<<<gd, 256, 0, ctx->stream>>>((const unsigned char*)d_two.p, HW, g_cl); ctx->launches++; }   // value-2 pixels count twice
    int cnt_g[2 * CT_MAX];
    STC_CUDA(cudaMemcpyAsync(cnt_g, d_g.p, 2 * CT_MAX * 4, cudaMemcpyDeviceToHost, ctx->stream));
    STC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int t = 0; t < T; ++t) {
      // np.mean of a float32 0/1 array: exact for these sizes
      const float ms = (float)((double)cnt_g[t] / HW), mc = (float)((double)cnt_g[CT_MAX + t] / HW);
      auto restrict_far = [&]() -> int {
        dilate(cl + (int64_t)t * HW, ta, 1, 50, 1, 0, 0, 0);                      // 50 cross iterations = L1 radius 50
        LAUNCH1D(k_and_or_dem, HW, sh + (int64_t)t * HW, ta, dem, HW);
        return STC_OK;
      };
      const bool far1 = ms > (mc + 0.3f) && mc < 0.3f;
      if (far1) restrict_far();
      if (mc < 0.05f) {
        float ms2 = ms;
        if (far1) {                                                               // the shadows of this date just changed: count again
          int c1 = 0;
          STC_CUDA(cudaMemsetAsync(d_cnt.p, 0, 4, ctx->stream));
          LAUNCH1D(k_count, HW, sh + (int64_t)t * HW, HW, (int*)d_cnt.p);
          STC_CUDA(cudaMemcpyAsync(&c1, d_cnt.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
          STC_CUDA(cudaStreamSynchronize(ctx->stream));
          ms2 = (float)((double)c1 / HW);
        }
        if ((ms2 / mc) > 3.f) restrict_far();                                     // mc == 0: inf > 3 (or nan: false), as NumPy
      }
    }
  }
  LAUNCH1D(k_or, N, cl, sh, cl, N);
  if (urban) { LAUNCH1D(k_or, N, nsr, (const unsigned char*)d_fc0.p, ta, N); dilate(ta, tc, T, 2, 1, 0, 0, 1); }   // fcps = dilate3d(max(fcps, nsr), 2)
  else dilate(nsr, tc, T, 2, 1, 0, 0, 1);                                       // fcps == 0
  STC_CUDA(cudaMemcpyAsync(fcps_dev, tc, N, cudaMemcpyDeviceToDevice, ctx->stream));
  int two_cnt[CT_MAX] = {0};                                                    // pixels where clouds == 2 (urban + grown non-urban cloud)
  if (urban) {
    STC_CUDA(cudaMemsetAsync(d_all.p, 0, CT_MAX * 4, ctx->stream));
    { TraceScope ts_(ctx, "k_count_dates"); k_count_dates<<<dim3(cdiv(HW, 256), T), 256, 0, ctx->stream>>>((const unsigned char*)d_two.p, HW, (int*)d_all.p); } ctx->launches++;
    STC_CUDA(cudaMemcpyAsync(two_cnt, d_all.p, T * 4, cudaMemcpyDeviceToHost, ctx->stream));
    STC_CUDA(cudaStreamSynchronize(ctx->stream));
  }

  // ---- H: dark-blue shadow recovery (:1638-1648) ----
  {
    if ((rc_ = moments(1, nullptr, nullptr, true))) return rc_;
    // every date in one launch per step (6 n small launches before)
    struct { float ref[CT_MAX]; int active[CT_MAX]; } db;
    memset(&db, 0, sizeof(db));
    bool any_db = false;
    for (int t = 0; t < T; ++t) {
      float frac = (float)((double)(HW - cnt_h[2 * t] + two_cnt[t]) / HW);      // np.mean(clouds[t]) as float32 (a 2 counts twice)
      if (!(frac < 0.9f)) continue;
      volatile float two_sd = 2.f * mom_h[2 * t + 1];
      db.ref[t] = mom_h[2 * t] + two_sd; db.active[t] = 1; any_db = true;
    }
    if (any_db) {
      Buf d_db;
      STC_CUDA(stc_dmalloc(&d_db.p, sizeof(db)));
      const void* staged = ctx_stage(ctx, &db, sizeof(db));
      if (!staged) STC_FAIL(STC_ERR_NOMEM, "cloud_masks: pinned staging");
      STC_CUDA(cudaMemcpyAsync(d_db.p, staged, sizeof(db), cudaMemcpyHostToDevice, ctx->stream));
      const dim3 gd(cdiv(HW, 256), T);
      { TraceScope ts_(ctx, "k_darkblue_all"); k_darkblue_all<<<gd, 256, 0, ctx->stream>>>(img, HW, (const float*)d_db.p, (const int*)((const char*)d_db.p + sizeof(db.ref)), ta); } ctx->launches++;
      dilate(ta, tb, T, 2, 1, 1, 1, 0);
      dilate(tb, ta, T, 2, 1, 0, 0, 0);
      { TraceScope ts_(ctx, "k_or_nowater_all"); k_or_nowater_all<<<gd, 256, 0, ctx->stream>>>(cl, ta, water, HW); } ctx->launches++;
    }
    if ((rc_ = dump(8, cl))) return rc_;
  }

  // ---- I: haze (:1652-1676), float32 list arithmetic as NumPy does it ----
  {
    std::vector<float> meanb, stdb, stdw;
    if ((rc_ = moments(2, nullptr, nullptr, true))) return rc_;
    std::vector<float> mb(mom_h.begin(), mom_h.begin() + 2 * T); std::vector<int> nclear(T);
    for (int t = 0; t < T; ++t) nclear[t] = cnt_h[2 * t];
    if ((rc_ = moments(3, nullptr, nullptr, true))) return rc_;
    for (int t = 0; t < T; ++t)
      if (nclear[t] > 0) { meanb.push_back(mb[2 * t]); stdb.push_back(mb[2 * t + 1]); stdw.push_back(mom_h[2 * t + 1]); }
    if (!meanb.empty()) {
      const float m1 = median_host(meanb), m2 = median_host(stdb), m3 = median_host(stdw);
      for (size_t k = 0; k < meanb.size(); ++k) {        // the reference indexes `haze` by list position, not by date (:1674-1676)
        volatile float hb = meanb[k] / m1, hs = stdb[k] / m2, hw = stdw[k] / m3;
        bool haze = ((hb >= 1.5f) && (hs <= 0.67f) && (hw < 1.f)) || ((hb >= 1.3f) && (hs <= 0.5f));
        if (haze) LAUNCH1D(k_fill, HW, cl + (int64_t)k * HW, (int64_t)HW, (unsigned char)1);
      }
    }
  }
  LAUNCH1D(k_to_float, N, cl, clouds_dev, N);
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

extern "C" int stc_cloud_masks_host(stc_ctx* ctx, const float* img_host, const float* dem_host, int T, int H, int W,
                                    float* clouds_host, uint8_t* fcps_host, uint8_t* stage_host, int stage_id) {
  if (!ctx) return STC_ERR_ARG;
  if (!img_host || !dem_host || !clouds_host || !fcps_host || T < 1 || T > CT_MAX || H < 3 || W < 3)
    STC_FAIL(STC_ERR_ARG, "cloud_masks: bad argument (1 <= T <= 32)");
  const int64_t N = (int64_t)T * H * W;
  Buf d_img, d_dem, d_out, d_fcps;
  STC_CUDA(stc_dmalloc(&d_img.p, N * 40)); STC_CUDA(stc_dmalloc(&d_dem.p, (size_t)H * W * 4));
  STC_CUDA(stc_dmalloc(&d_out.p, N * 4)); STC_CUDA(stc_dmalloc(&d_fcps.p, N));
  STC_CUDA(cudaMemcpyAsync(d_img.p, img_host, N * 40, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(d_dem.p, dem_host, (size_t)H * W * 4, cudaMemcpyHostToDevice, ctx->stream));
  const bool anc = ctx->anc_H == H && ctx->anc_W == W;
  int rc = cloud_masks_dev(ctx, (const float*)d_img.p, (const float*)d_dem.p, T, H, W, anc ? ctx->anc_forest : nullptr,
                           anc ? ctx->anc_urban_core : nullptr, anc ? ctx->anc_urban_near : nullptr, (float*)d_out.p, (unsigned char*)d_fcps.p,
                           stage_host, stage_id);
  if (rc) return rc;
  STC_CUDA(cudaMemcpyAsync(fcps_host, d_fcps.p, N, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(clouds_host, d_out.p, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}


// Ancillary rasters for the masks of the tile that follows (cloud_removal.py:1131-1135 `urbanmask.tif`, :1254-1257
// `forestmask.tif`): the caller reads the raster windows (I/O) and passes them at tile resolution -- forest [H,W] 0/1 =
// adjust_cloudmask_in_forests(...), urban_core / urban_near [H,W] 0/1 = the two resized rasters of mask_nonurban_areas
// (:745-753: dilated once, dilated five more times).  NULL pointers clear a mask; masks stay set until replaced and are
// used by stc_cloud_masks_host / stc_tile_run_host calls whose H, W match.
extern "C" int stc_set_ancillary_masks_host(stc_ctx* ctx, const uint8_t* forest_host, const uint8_t* urban_core_host,
                                            const uint8_t* urban_near_host, int H, int W) {
  if (!ctx) return STC_ERR_ARG;
  if ((!urban_core_host) != (!urban_near_host)) STC_FAIL(STC_ERR_ARG, "ancillary masks: urban_core and urban_near come together");
  if ((forest_host || urban_core_host) && (H < 3 || W < 3)) STC_FAIL(STC_ERR_ARG, "ancillary masks: bad shape");
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  for (unsigned char** p : {&ctx->anc_forest, &ctx->anc_urban_core, &ctx->anc_urban_near}) { if (*p) { cudaFree(*p); *p = nullptr; } }
  ctx->anc_H = ctx->anc_W = 0;
  if (!forest_host && !urban_core_host) return STC_OK;
  const size_t HW = (size_t)H * W;
  auto up = [&](const uint8_t* src, unsigned char** dst) -> int {
    if (!src) return STC_OK;
    STC_CUDA(cudaMalloc((void**)dst, HW));
    STC_CUDA(cudaMemcpyAsync(*dst, src, HW, cudaMemcpyHostToDevice, ctx->stream));
    return STC_OK;
  };
  int rc;
  if ((rc = up(forest_host, &ctx->anc_forest)) || (rc = up(urban_core_host, &ctx->anc_urban_core)) || (rc = up(urban_near_host, &ctx->anc_urban_near))) return rc;
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->anc_H = H; ctx->anc_W = W;
  return STC_OK;
}
