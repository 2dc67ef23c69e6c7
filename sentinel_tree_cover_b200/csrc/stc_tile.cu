// Whole-tile, device-resident driver: the body of the reference's main loop
// (/root/reference/src/download_and_predict_job.py:1995-2020)
//     process_tile (:640-997) -> superresolve_large_tile (:95-147) -> process_subtiles (:1125-1486)
//     -> load_mosaic_predictions (:1515-1641)
// behind ONE C call: the raw uint16 cubes go up once, the uint8 tile comes back once, every intermediate array lives
// in pooled device memory (stc_pool.cu) and only integers / per-date scalars cross to the host, where the reference's
// own control flow on scalars is replayed (date screening and np.delete bookkeeping, the retry loops on per-date
// fractions, the 12 x n regrid / Whittaker operator built from the surviving dates, the window table, the
// median(r)/r multipliers of the mosaic).  The array arithmetic is the SAME device code the per-stage entry points
// run (cloud_masks_dev, remove_clouds_dev, pre_feather_dev, tf_*_dev, sr_forward_dev, model_predict_dev, ...), so the
// per-stage parity tests carry over; tests/test_tile_chain.py checks the chain against the mirrored Python drivers.
#include "stc_common.cuh"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>

// ---- device cores implemented next to their kernels ----
int codec_to_float32_dev(stc_ctx* ctx, const uint16_t* in_dev, int64_t n, float* out_dev);
int codec_convert_to_db_dev(stc_ctx* ctx, const float* in_dev, int64_t n, float min_db, float* out_dev);
int tp_s1_fill_dev(stc_ctx* ctx, float* s1_dev, int m, int len);
int tp_median5_dev(stc_ctx* ctx, const float* in_dev, int H, int W, float* out_dev);
int tp_clm_pairs_dev(stc_ctx* ctx, float* clm_dev, int n, int HW);
int tp_snow_dev(stc_ctx* ctx, const float* s2_dev, int n, int H, int W, int* per_date_dev, unsigned char* low_tmp, unsigned char* snow_dev);
int tp_count_gt_dev(stc_ctx* ctx, const float* data_dev, int nseg, int len, float thresh, int* counts_dev);
int tp_elementwise_dev(stc_ctx* ctx, float* x_dev, int64_t n, int mode, float a, float b);
int tp_max_masked_dev(stc_ctx* ctx, float* a_dev, const float* b_dev, const unsigned char* zero_dev, int64_t n);
int tp_count_lt_axis0_dev(stc_ctx* ctx, const float* data_dev, int n, int64_t len, float thresh, int* out_dev);
int interp_build_sentinel2_dev(stc_ctx* ctx, const float* s2_10_dev, const float* s2_20_dev, int n, int h, int w, float* out_dev);
int interp_missing_counts_dev(stc_ctx* ctx, const float* arr_dev, int n, int HW, int C, int* bad_px_dev, int* nan_vals_dev);
int cloud_masks_dev(stc_ctx* ctx, const float* img, const float* dem, int T, int H, int W, const unsigned char* forest_dev,
                    const unsigned char* urban_core_dev, const unsigned char* urban_near_dev, float* clouds_dev, unsigned char* fcps_dev,
                    uint8_t* stage_host, int stage_id);
int remove_clouds_dev(stc_ctx* ctx, float* tiles, const float* probs_dev, const unsigned char* pfcps_dev, int n, int H, int W,
                      uint32_t* mt_state, float* areas, int32_t* to_remove_host, float* mosaic_out_dev, int clip_when_all_kept,
                      int32_t* clipped_out);
int tf_s2_medians_dev(stc_ctx* ctx, float* s2, int n, int H, int W, float* median14_dev, int32_t* bad_px_host, int64_t* nan_total_host);
int tf_smooth_quarterly_dev(stc_ctx* ctx, float* s2, int n, int H, int W, const float* M_host, const float* s1_dev, float* s2_monthly_dev,
                            float* s2_quarterly_dev, float* s1_quarterly_dev, float* s1_median_dev, int32_t* nan_after_host, int skip_fill);
int tf_process_subtiles_dev(stc_ctx* ctx, const float* s2q, const float* s1q, const float* s2m, const float* s1m, const float* dem,
                            const int* clr, int H, int W, int nt, const int32_t* windows_host, int S, int T, int length,
                            int force_no_data, const double* min17, const double* max17, float* out_dev, int32_t* no_data_host,
                            float* early_dev, float* late_dev);
int post_np_sum_dev(stc_ctx* ctx, const float* data_dev, int nseg, int len, int mode, float* sum_dev, int* valid_dev);

// =====================================================================================================================
// Host logic (integers and a handful of doubles per tile).  Exposed through stc_*_plan test hooks so that the CPU suite
// can hold it against the NumPy mirrors (regrid.py, windows.py), which are themselves pinned to the reference.
// =====================================================================================================================
namespace tilehost {

// ---- adjust_shape (:260-310) for one axis: out[i] = in[clamp(i + shift, 0, L-1)], out_len as the reference leaves it ----
struct AxisPlan { int shift, out_len; };
AxisPlan adjust_axis(int L, int target) {
  AxisPlan p{0, L};
  if (L < target) {
    const int pad = (target - L) / 2;
    if (pad == 0) { p.shift = -1; p.out_len = L + 1; }              // np.pad(..., (1, 0), 'edge')
    else { p.shift = -pad; p.out_len = L + 2 * pad; }                // np.pad(..., (pad, pad), 'edge')
  } else if (L > target) {
    const int pad = (L - target) / 2; const bool even = (L - target) % 2 == 0;
    if (pad == 0) { p.shift = 1; p.out_len = L - 1; }                // arr[:, 1:]
    else if (even) { p.shift = pad; p.out_len = L - 2 * pad; }       // arr[:, pad:-pad]
    else {                                                           // arr[:, floor(pad/2):-ceil(pad/2)]
      const int a = pad / 2, b = (pad + 1) / 2;
      p.shift = a; p.out_len = L - a - b;
    }
  }
  return p;
}

// ---- calculate_and_save_best_images (src/downloading/utils.py:176-347) as the 24 x n matrix G of regrid.py ----
// Throws std::runtime_error where the reference raises (empty neighbour sets, ambiguous duplicate dates).
static void neighbour_weights(const std::vector<long>& off, long day_min, long day_max, std::vector<long>& before,
                              std::vector<long>& after, std::vector<double>& wb, std::vector<double>& wa) {
  before.clear(); after.clear();
  {
    std::vector<long> lt; for (long o : off) if (o < 5) lt.push_back(o);
    if (lt.size() > 2) lt.erase(lt.begin(), lt.end() - 2);
    if (!lt.empty()) { long mx = *std::max_element(lt.begin(), lt.end()); for (long o : lt) if (o > mx - 100) before.push_back(o); }
    std::vector<long> ge; for (long o : off) if (o >= -5) ge.push_back(o);
    if (ge.size() > 2) ge.resize(2);
    if (!ge.empty()) { long mn = *std::min_element(ge.begin(), ge.end()); for (long o : ge) if (o < mn + 100) after.push_back(o); }
  }
  long wrap_b = 0, wrap_a = 0;
  if (before.empty()) {
    if (day_min >= 90) { before.assign(1, off.back()); wrap_b = 365; }
    else before = after;
  }
  if (after.empty()) {
    if (day_max <= 270) { after.assign(1, off.front()); wrap_a = 365; }
    else after = before;
  }
  if (before.empty() || after.empty()) throw std::runtime_error("regrid: no neighbouring image (the reference raises IndexError)");
  std::vector<double> db, da;
  for (long b : before) db.push_back(std::max(std::fabs((double)(b - wrap_b)), 1.0));
  for (long a : after) da.push_back(std::max(std::fabs((double)(a + wrap_a)), 1.0));
  const double span = std::max(db.back() + da.front(), 2.0);
  wb.clear(); wa.clear();
  for (double d : db) wb.push_back(std::fabs(1.0 - d / span));
  for (double d : da) wa.push_back(std::fabs(1.0 - d / span));
  if (wb.size() == 2) wb[0] = std::fabs((db[1] / db[0]) * wb[1]);
  if (wa.size() == 2) wa[1] = std::fabs((da[0] / da[1]) * wa[0]);
  // np.sum of <= 2 float64 values: left to right
  double sb = 0.0; for (double v : wb) sb += v;
  double sa = 0.0; for (double v : wa) sa += v;
  const double tot = sb + sa;
  for (double& v : wb) v /= tot;
  for (double& v : wa) v /= tot;
}

// G [24][n] float32
void regrid_matrix(const std::vector<long>& dates_in, std::vector<float>& G) {
  const int n = (int)dates_in.size();
  if (n < 1) throw std::runtime_error("regrid: no dates");
  std::vector<long> dates(dates_in);
  for (long& d : dates) if (d < -100) d = ((d % 365) + 365) % 365;       // Python modulo
  const long dmin = *std::min_element(dates.begin(), dates.end()), dmax = *std::max_element(dates.begin(), dates.end());
  G.assign((size_t)24 * n, 0.f);
  std::vector<long> off(n), before, after;
  std::vector<double> wb, wa;
  for (int r = 0; r < 24; ++r) {
    const long day = 15L * r;
    for (int i = 0; i < n; ++i) off[i] = dates[i] - day;
    neighbour_weights(off, dmin, dmax, before, after, wb, wa);
    auto match = [&](const std::vector<long>& sel) {
      std::vector<int> idx;
      for (int i = 0; i < n; ++i) for (long s : sel) if (dates[i] == day + s) { idx.push_back(i); break; }
      return idx;                                                       // ascending, unique
    };
    std::vector<int> ib = match(before), ia = match(after);
    if (ib.size() > 2) ib.resize(2);
    if (ia.size() > 2) ia.erase(ia.begin(), ia.end() - 2);
    // (eye[idx] * w.astype(float32)[:, None]).sum(0) with NumPy broadcasting between len(idx) and len(w)
    auto accumulate = [&](const std::vector<int>& idx, const std::vector<double>& w, std::vector<float>& row) {
      const size_t a = idx.size(), b = w.size();
      if (a != b && a != 1 && b != 1) throw std::runtime_error("regrid: duplicate image dates (the reference raises a broadcasting error)");
      const size_t m = std::max(a, b);
      std::fill(row.begin(), row.end(), 0.f);
      for (size_t k = 0; k < m; ++k) {
        const int col = idx[a == 1 ? 0 : k]; const float wk = (float)w[b == 1 ? 0 : k];
        row[col] = row[col] + wk;                                        // float32 sum over the rows, in order
      }
    };
    std::vector<float> rb(n), ra(n);
    if (ib.empty() || ia.empty()) {                                      // eye[[]] sums to zeros of length n; an empty weight set cannot happen
      if (ib.empty()) std::fill(rb.begin(), rb.end(), 0.f); else accumulate(ib, wb, rb);
      if (ia.empty()) std::fill(ra.begin(), ra.end(), 0.f); else accumulate(ia, wa, ra);
    } else { accumulate(ib, wb, rb); accumulate(ia, wa, ra); }
    for (int i = 0; i < n; ++i) G[(size_t)r * n + i] = rb[i] + ra[i];
  }
}

// A S (12 x 24, double): S = (I + 100 D2'D2)^-1 (whittaker_smoother.py:10-42), A = mean of consecutive pairs (:64-67)
const double* pair_mean_whittaker() {
  static double AS[12 * 24];
  static bool ready = false;
  if (ready) return AS;
  const int n = 24;
  long double C[24][48];
  for (int i = 0; i < n; ++i) for (int j = 0; j < 2 * n; ++j) C[i][j] = (j == i || j == n + i) ? 1.0L : 0.0L;
  for (int r = 0; r < n - 2; ++r) {            // D2 rows (1, -2, 1): coef += 100 * d d'
    const long double d[3] = {1.0L, -2.0L, 1.0L};
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) C[r + a][r + b] += 100.0L * d[a] * d[b];
  }
  for (int c = 0; c < n; ++c) {                 // Gauss-Jordan with partial pivoting
    int piv = c;
    for (int r = c + 1; r < n; ++r) if (fabsl(C[r][c]) > fabsl(C[piv][c])) piv = r;
    if (piv != c) for (int j = 0; j < 2 * n; ++j) std::swap(C[c][j], C[piv][j]);
    const long double inv = 1.0L / C[c][c];
    for (int j = 0; j < 2 * n; ++j) C[c][j] *= inv;
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const long double f = C[r][c];
      if (f != 0.0L) for (int j = 0; j < 2 * n; ++j) C[r][j] -= f * C[c][j];
    }
  }
  for (int o = 0; o < 12; ++o)
    for (int j = 0; j < n; ++j) AS[o * 24 + j] = (double)(0.5L * C[2 * o][n + j] + 0.5L * C[2 * o + 1][n + j]);
  ready = true;
  return AS;
}

// M = A S G (12 x n) float32: regrid.monthly_operator
void monthly_operator(const std::vector<long>& dates, std::vector<float>& M) {
  std::vector<float> G;
  regrid_matrix(dates, G);
  const int n = (int)dates.size();
  const double* AS = pair_mean_whittaker();
  M.assign((size_t)12 * n, 0.f);
  for (int o = 0; o < 12; ++o)
    for (int i = 0; i < n; ++i) {
      double acc = 0.0;
      for (int j = 0; j < 24; ++j) acc += AS[o * 24 + j] * (double)G[(size_t)j * n + i];
      M[(size_t)o * n + i] = (float)acc;
    }
}

// ---- process_subtiles window bookkeeping (:1295-1317; windows.subtile_windows) ----
static std::vector<long> window_starts(long L, long size, int n_rows) {
  const long gap = (long)std::ceil((double)(L - size) / (double)(n_rows - 1));
  std::vector<long> s;
  if (gap > 0) for (long v = 0; v < L - size; v += gap) s.push_back(v);
  s.push_back(L - size);
  return s;
}
// folder [nt][4], arr [nt][4]
void subtile_windows(long Lx, long Ly, long size, int n_rows, std::vector<long>& folder, std::vector<long>& arr) {
  const long diff = 7;
  std::vector<long> sx = window_starts(Lx, size, n_rows), sy = window_starts(Ly, size, n_rows);
  const size_t nt = sx.size() * sy.size();
  // column-wise np.sort of the cartesian table, then the second column re-tiled from its unique values
  std::vector<long> c0, c1;
  for (size_t i = 0; i < sy.size(); ++i) for (size_t j = 0; j < sx.size(); ++j) { c0.push_back(sx[j]); c1.push_back(sy[i]); }
  std::sort(c0.begin(), c0.end()); std::sort(c1.begin(), c1.end());
  std::vector<long> uy(c1); uy.erase(std::unique(uy.begin(), uy.end()), uy.end());
  const size_t reps = nt / uy.size();
  if (reps * uy.size() != nt) throw std::runtime_error("windows: duplicate window starts (the reference fails in its assignment)");
  folder.assign(nt * 4, size);
  for (size_t k = 0; k < nt; ++k) { folder[k * 4] = c0[k]; folder[k * 4 + 1] = uy[k % uy.size()]; }
  arr = folder;
  long n_x = 0, n_y = 0;
  for (size_t k = 0; k < nt; ++k) { n_x += arr[k * 4] == 0; n_y += arr[k * 4 + 1] == 0; }
  std::vector<long> grow_x(nt, 2 * diff);
  if (2 * n_x > (long)nt) {
    std::fill(grow_x.begin(), grow_x.end(), 0);
    for (long k = 0; k < n_x; ++k) grow_x[k] += diff;
    for (long k = (long)nt - n_x; k < (long)nt; ++k) grow_x[k] += diff;
    for (long k = n_x; k < (long)nt - n_x; ++k) grow_x[k] += 2 * diff;
  } else {
    for (long k = 0; k < n_x; ++k) grow_x[k] = diff;
    for (long k = (long)nt - n_x; k < (long)nt; ++k) grow_x[k] = diff;
  }
  for (size_t k = 0; k < nt; ++k) {
    const bool edge_y = (k % n_y == 0) || ((k + 1) % n_y == 0);
    arr[k * 4 + 2] += grow_x[k];
    arr[k * 4 + 3] += edge_y ? diff : 2 * diff;
    if ((long)k >= n_x) arr[k * 4] -= diff;
    arr[k * 4 + 1] -= diff;
    for (int c = 0; c < 4; ++c) if (arr[k * 4 + c] < 0) arr[k * 4 + c] = 0;
  }
}

// superresolve_large_tile window starts (:122-123)
std::vector<int> superres_windows(int L, int wsize) {
  std::vector<int> s;
  for (int x = 0; x < L - wsize; x += wsize) s.push_back(x);
  s.push_back(L - wsize);
  return s;
}

// np.median of a short float32 list: NaN propagates, even length = float32 mean of the middle pair
float median_f32(std::vector<float> v) {
  for (float x : v) if (x != x) return x;
  std::sort(v.begin(), v.end());
  const size_t n = v.size();
  if (n & 1) return v[n / 2];
  volatile float s = v[n / 2 - 1] + v[n / 2];
  return s / 2.f;
}

}  // namespace tilehost

// =====================================================================================================================
// Kernels that only the chain needs (data movement)
// =====================================================================================================================
namespace {

// adjust_shape on [frames][Hin][Win][C] -> [frames][Hout][Wout][C]: edge padding / cropping as index clamps
template <typename T>
__global__ void __launch_bounds__(256) k_adjust(const T* __restrict__ in, int Hin, int Win, int C, T* __restrict__ out, int Hout, int Wout,
                                                int shy, int shx, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C); int64_t p = i / C;
  const int x = (int)(p % Wout); p /= Wout;
  const int y = (int)(p % Hout); const int64_t f = p / Hout;
  int sy = y + shy; sy = sy < 0 ? 0 : (sy >= Hin ? Hin - 1 : sy);
  int sx = x + shx; sx = sx < 0 ? 0 : (sx >= Win ? Win - 1 : sx);
  out[i] = in[((f * Hin + sy) * Win + sx) * C + c];
}
// clm.repeat(2, axis=1).repeat(2, axis=2) as float32 (:687)
__global__ void __launch_bounds__(256) k_clm_upsample(const unsigned char* __restrict__ in, int h, int w, float* __restrict__ out, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int W = 2 * w, H = 2 * h;
  const int x = (int)(i % W); const int y = (int)((i / W) % H); const int64_t f = i / ((int64_t)W * H);
  out[i] = (float)in[(f * h + (y >> 1)) * w + (x >> 1)];
}
// clm[fcps] = 0 (:843)
__global__ void __launch_bounds__(256) k_zero_where(float* __restrict__ a, const unsigned char* __restrict__ m, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && m[i]) a[i] = 0.f;
}
// np.clip(probs, 0, 1) into a second buffer (id_areas_to_interp works on a clipped copy, cloud_removal.py:774-798)
__global__ void __launch_bounds__(256) k_clip_copy(const float* __restrict__ in, float* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = in[i];
  out[i] = isnan(v) ? v : fminf(fmaxf(v, 0.f), 1.f);
}

// ---- superresolve_large_tile windows ----
struct SrWin { int from_band, x, y; };      // source: the tile (0) or the pre-resolution bottom band (1); window origin in that source
__device__ __forceinline__ int reflect4(int i, int len) {     // np.pad(..., 4, 'reflect') index of padded coordinate i - 4
  int j = i - 4;
  if (j < 0) j = -j;
  if (j >= len) j = 2 * (len - 1) - j;
  return j;
}
// out [nw][n][ws+8][ws+8][10]
__global__ void __launch_bounds__(256) k_sr_gather(const float* __restrict__ arr, const float* __restrict__ band, const SrWin* __restrict__ wins,
                                                   int n, int H, int W, int band_rows, int ws, float* __restrict__ out) {
  const int P = ws + 8;
  const int w = blockIdx.z, t = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * P) return;
  const SrWin sw = wins[w];
  const int r = sw.x + reflect4(i / P, ws), c = sw.y + reflect4(i % P, ws);
  const float* src = sw.from_band ? band + (((int64_t)t * band_rows + r) * W + c) * 10 : arr + (((int64_t)t * H + r) * W + c) * 10;
  float* o = out + ((((int64_t)w * n + t) * P * P) + i) * 10;
#pragma unroll
  for (int k = 0; k < 10; ++k) o[k] = src[k];
}
// src[..., 4:] = resolved[:, 4:-4, 4:-4, :]
__global__ void __launch_bounds__(256) k_sr_scatter(const float* __restrict__ res, const SrWin* __restrict__ wins, int n, int H, int W,
                                                    int band_rows, int ws, float* __restrict__ arr, float* __restrict__ band) {
  const int P = ws + 8;
  const int w = blockIdx.z, t = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ws * ws) return;
  const SrWin sw = wins[w];
  const int rr = i / ws, cc = i % ws;
  const float* s = res + ((((int64_t)w * n + t) * P + rr + 4) * P + cc + 4) * 6;
  float* d = sw.from_band ? band + (((int64_t)t * band_rows + sw.x + rr) * W + sw.y + cc) * 10 : arr + (((int64_t)t * H + sw.x + rr) * W + sw.y + cc) * 10;
#pragma unroll
  for (int k = 0; k < 6; ++k) d[4 + k] = s[k];
}
// rows [x0, x0 + rows) of arr <-> band [n][rows][W][10]
__global__ void __launch_bounds__(256) k_band_copy(float* __restrict__ arr, float* __restrict__ band, int n, int H, int W, int x0, int rows,
                                                   int to_band) {
  const int64_t per = (int64_t)rows * W * 10;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per * n) return;
  const int64_t t = i / per, r = i % per;
  float* a = arr + ((int64_t)t * H + x0) * W * 10 + r;
  if (to_band) band[i] = *a; else *a = band[i];
}

struct TimeMarks {
  bool on; stc_ctx* ctx; double t0;
  static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
  explicit TimeMarks(stc_ctx* c) : on(getenv("STC_TILE_TIMING") != nullptr), ctx(c), t0(now()) {}
  void mark(const char* what) {
    if (!on) return;
    cudaStreamSynchronize(ctx->stream);
    const double t = now();
    fprintf(stderr, "[tile_run] %-40s %8.2f ms\n", what, t - t0);
    t0 = t;
  }
};

#define TL_CHECK(call) do { int rc__ = (call); if (rc__) return rc__; } while (0)
#define TL_LAUNCH(kern, n, ...) do { TraceScope ts_(ctx, #kern); kern<<<cdiv((n), 256), 256, 0, ctx->stream>>>(__VA_ARGS__); ctx->launches++; } while (0)

// np.delete(x, idx, axis=0) on a device array of `n` slabs: the kept slabs move down in order
int compact_slabs(stc_ctx* ctx, void* base, size_t slab_bytes, int n, const std::vector<int>& keep) {
  for (size_t k = 0; k < keep.size(); ++k)
    if ((int)k != keep[k])
      STC_CUDA(cudaMemcpyAsync((char*)base + k * slab_bytes, (char*)base + (size_t)keep[k] * slab_bytes, slab_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  (void)n;
  return STC_OK;
}

}  // namespace

// =====================================================================================================================
// The chain
// =====================================================================================================================
namespace {

struct TileState {
  int n = 0, H = 0, W = 0;                       // dates alive, 10 m grid
  std::vector<long> dates;
  PoolBuf s2, s1, dem, clm, cloudshad, fcps, interp, clipbuf, fa, fb, fsums, cnt;
  bool have_clm = false;
};

// cloud_removal.identify_clouds_shadows + the Sen2Cor merge (:840-846, 872-876, ...)
int tile_masks(stc_ctx* ctx, TileState& S, bool first) {
  const int64_t N = (int64_t)S.n * S.H * S.W;
  const bool anc = ctx->anc_H == S.H && ctx->anc_W == S.W;
  TL_CHECK(cloud_masks_dev(ctx, S.s2.as<float>(), S.dem.as<float>(), S.n, S.H, S.W, anc ? ctx->anc_forest : nullptr, anc ? ctx->anc_urban_core : nullptr,
                           anc ? ctx->anc_urban_near : nullptr, S.cloudshad.as<float>(), S.fcps.as<unsigned char>(), nullptr, 0));
  if (S.have_clm) {
    if (first) TL_LAUNCH(k_zero_where, N, S.clm.as<float>(), S.fcps.as<unsigned char>(), N);      // clm[fcps] = 0.
    TL_CHECK(tp_max_masked_dev(ctx, S.cloudshad.as<float>(), S.clm.as<float>(), nullptr, N));      // np.maximum(cloudshad, clm)
  }
  return STC_OK;
}
// cloud_removal.id_areas_to_interp (closing 15) on np.clip(probs, 0, 1)
int tile_feather(stc_ctx* ctx, TileState& S) {
  const int64_t N = (int64_t)S.n * S.H * S.W;
  TL_LAUNCH(k_clip_copy, N, S.cloudshad.as<float>(), S.clipbuf.as<float>(), N);
  return pre_feather_dev(ctx, S.clipbuf.as<float>(), S.n, S.H, S.W, 15, S.fa.as<float>(), S.fb.as<float>(), S.fsums.as<float>(), S.interp.as<float>());
}
// np.delete of dates from every per-date array
int tile_delete(stc_ctx* ctx, TileState& S, const std::vector<int>& remove, bool with_interp) {
  if (remove.empty()) return STC_OK;
  std::vector<char> gone(S.n, 0);
  for (int r : remove) gone[r] = 1;
  std::vector<int> keep; std::vector<long> nd;
  for (int t = 0; t < S.n; ++t) if (!gone[t]) { keep.push_back(t); nd.push_back(S.dates[t]); }
  const size_t HW = (size_t)S.H * S.W;
  TL_CHECK(compact_slabs(ctx, S.s2.p, HW * 40, S.n, keep));
  if (with_interp) TL_CHECK(compact_slabs(ctx, S.interp.p, HW * 4, S.n, keep));
  if (S.have_clm) TL_CHECK(compact_slabs(ctx, S.clm.p, HW * 4, S.n, keep));
  S.dates = nd; S.n = (int)keep.size();
  return STC_OK;
}

}  // namespace

extern "C" int stc_tile_run_host(stc_ctx* ctx, const uint16_t* s2_10_host, int n, int h10, int w10, const uint16_t* s2_20_host, int h20,
                                 int w20, const uint16_t* s1_host, int m1, int hs, int ws, const float* dem_host, int hd, int wd,
                                 const uint8_t* clm_host, const int32_t* dates_host, uint32_t* mt_state, int make_shadow, int superresolve,
                                 int size, int length, const double* min17, const double* max17, const float* gauss_host,
                                 uint8_t* out_host, int out_h, int out_w, int32_t* dates_kept_host, int32_t* n_kept_host,
                                 float* subtile_preds_host) {
  if (!ctx) return STC_ERR_ARG;
  if (!s2_10_host || !s2_20_host || !s1_host || !dem_host || !dates_host || !mt_state || !min17 || !max17 || !out_host || n < 1 || n > 32 ||
      h10 < 1 || w10 < 1 || h20 < 2 || w20 < 2 || m1 != 12 || hs < 1 || ws < 1 || hd < 1 || wd < 1 || length != 4 || size < 14)
    STC_FAIL(STC_ERR_ARG, "tile_run: bad argument (1 <= n <= 32 dates, 12 Sentinel-1 composites, length 4)");
  const int H = 2 * h20, W = 2 * w20;                 // width / height of :725-726
  const int64_t HW = (int64_t)H * W;
  if (superresolve && (H < 110 || W < 110)) STC_FAIL(STC_ERR_ARG, "tile_run: superresolve_large_tile needs at least 110 x 110 px");
  if (H <= size || W <= size) STC_FAIL(STC_ERR_ARG, "tile_run: the tile must be larger than one subtile");
  TimeMarks tm(ctx);
  TileState S; S.n = n; S.H = H; S.W = W; S.dates.assign(dates_host, dates_host + n);
  using tilehost::AxisPlan;

  // ------------------------------------------------------------------ process_tile (:640-997) ----
  // uploads: raw uint16 cubes + DEM (+ Sen2Cor mask), ~ (8 + 3 + 2) bytes per 10 m pixel and date
  PoolBuf u10, u20, us1, fdem0, u10adj, f10, f20, uclm;
  const int64_t n10 = (int64_t)n * h10 * w10 * 4, n20 = (int64_t)n * h20 * w20 * 6, ns1 = (int64_t)m1 * hs * ws * 2;
  STC_CUDA(u10.alloc(n10 * 2)); STC_CUDA(u20.alloc(n20 * 2)); STC_CUDA(us1.alloc(ns1 * 2)); STC_CUDA(fdem0.alloc((size_t)hd * wd * 4));
  STC_CUDA(cudaMemcpyAsync(u10.p, s2_10_host, n10 * 2, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(u20.p, s2_20_host, n20 * 2, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(us1.p, s1_host, ns1 * 2, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(fdem0.p, dem_host, (size_t)hd * wd * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (clm_host) {
    STC_CUDA(uclm.alloc((size_t)n * h20 * w20));
    STC_CUDA(cudaMemcpyAsync(uclm.p, clm_host, (size_t)n * h20 * w20, cudaMemcpyHostToDevice, ctx->stream));
  }
  tm.mark("upload raw cubes");

  // Sentinel-1: / 65535, saturated-value fill, dB on both polarisations (:699-709), then adjust_shape
  {
    PoolBuf s1raw;
    STC_CUDA(s1raw.alloc(ns1 * 4));
    TL_CHECK(codec_to_float32_dev(ctx, us1.as<uint16_t>(), ns1, s1raw.as<float>()));
    TL_CHECK(tp_s1_fill_dev(ctx, s1raw.as<float>(), m1, hs * ws * 2));
    TL_CHECK(codec_convert_to_db_dev(ctx, s1raw.as<float>(), ns1, 22.f, s1raw.as<float>()));
    const AxisPlan ay = tilehost::adjust_axis(hs, H), ax = tilehost::adjust_axis(ws, W);
    if (ay.out_len != H || ax.out_len != W) STC_FAIL(STC_ERR_ARG, "tile_run: Sentinel-1 shape cannot be aligned to the 10 m grid (adjust_shape leaves a mismatch; the reference fails in its slicing)");
    STC_CUDA(S.s1.alloc((size_t)m1 * HW * 8));
    const int64_t tot = (int64_t)m1 * HW * 2;
    { TraceScope ts_(ctx, "k_adjust"); k_adjust<float><<<cdiv(tot, 256), 256, 0, ctx->stream>>>(s1raw.as<float>(), hs, ws, 2, S.s1.as<float>(), H, W, ay.shift, ax.shift, tot); } ctx->launches++;
  }
  // DEM: 5 x 5 median filter (:713), adjust_shape
  {
    PoolBuf dmed;
    STC_CUDA(dmed.alloc((size_t)hd * wd * 4));
    TL_CHECK(tp_median5_dev(ctx, fdem0.as<float>(), hd, wd, dmed.as<float>()));
    const AxisPlan ay = tilehost::adjust_axis(hd, H), ax = tilehost::adjust_axis(wd, W);
    if (ay.out_len != H || ax.out_len != W) STC_FAIL(STC_ERR_ARG, "tile_run: DEM shape cannot be aligned to the 10 m grid (adjust_shape)");
    STC_CUDA(S.dem.alloc((size_t)HW * 4));
    { TraceScope ts_(ctx, "k_adjust"); k_adjust<float><<<cdiv(HW, 256), 256, 0, ctx->stream>>>(dmed.as<float>(), hd, wd, 1, S.dem.as<float>(), H, W, ay.shift, ax.shift, HW); } ctx->launches++;
  }
  // Sentinel-2: adjust_shape of the 10 m bands (pure data movement, done on the uint16 samples), decode, 20 m -> 10 m stack
  {
    const AxisPlan ay = tilehost::adjust_axis(h10, H), ax = tilehost::adjust_axis(w10, W);
    if (ay.out_len != H || ax.out_len != W) STC_FAIL(STC_ERR_ARG, "tile_run: 10 m bands cannot be aligned to twice the 20 m grid (adjust_shape)");
    const uint16_t* src10 = u10.as<uint16_t>();
    if (h10 != H || w10 != W) {
      STC_CUDA(u10adj.alloc((size_t)n * HW * 8));
      const int64_t tot = (int64_t)n * HW * 4;
      { TraceScope ts_(ctx, "k_adjust"); k_adjust<uint16_t><<<cdiv(tot, 256), 256, 0, ctx->stream>>>(src10, h10, w10, 4, u10adj.as<uint16_t>(), H, W, ay.shift, ax.shift, tot); } ctx->launches++;
      src10 = u10adj.as<uint16_t>();
    }
    STC_CUDA(f10.alloc((size_t)n * HW * 16)); STC_CUDA(f20.alloc((size_t)n20 * 4)); STC_CUDA(S.s2.alloc((size_t)n * HW * 40));
    TL_CHECK(codec_to_float32_dev(ctx, src10, (int64_t)n * HW * 4, f10.as<float>()));
    TL_CHECK(codec_to_float32_dev(ctx, u20.as<uint16_t>(), n20, f20.as<float>()));
    TL_CHECK(interp_build_sentinel2_dev(ctx, f10.as<float>(), f20.as<float>(), n, h20, w20, S.s2.as<float>()));
  }
  if (clm_host) {                     // Sen2Cor mask: repeat x2, consecutive-date rule (:686-695)
    STC_CUDA(S.clm.alloc((size_t)n * HW * 4));
    TL_LAUNCH(k_clm_upsample, (int64_t)n * HW, uclm.as<unsigned char>(), h20, w20, S.clm.as<float>(), (int64_t)n * HW);
    TL_CHECK(tp_clm_pairs_dev(ctx, S.clm.as<float>(), n, (int)HW));
    S.have_clm = true;
  }
  STC_CUDA(S.cnt.alloc(4 * 64 * 4));
  tm.mark("decode + S1 + DEM + 20m->10m stack");

  // missing-pixel screening id_missing_px(sentinel2, 2) (:786-794)
  std::vector<int> h_cnt(128);
  {
    STC_CUDA(cudaMemsetAsync(S.cnt.p, 0, 2 * 64 * 4, ctx->stream));
    TL_CHECK(interp_missing_counts_dev(ctx, S.s2.as<float>(), S.n, (int)HW, 10, S.cnt.as<int>(), S.cnt.as<int>() + 64));
    STC_CUDA(cudaMemcpyAsync(h_cnt.data(), S.cnt.p, 64 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    STC_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<int> rm;
    for (int t = 0; t < S.n; ++t) if ((double)h_cnt[t] >= ((double)H * (double)H) / 2.0) rm.push_back(t);
    TL_CHECK(tile_delete(ctx, S, rm, false));
    if (S.n < 1) STC_FAIL(STC_ERR_STATE, "tile_run: every date is missing more than half of its pixels");
  }
  // snow screening (:808-838; the removal rule only fires for more than 10 snowy dates)
  {
    PoolBuf low, snow;
    STC_CUDA(low.alloc(HW)); STC_CUDA(snow.alloc(HW));
    TL_CHECK(tp_snow_dev(ctx, S.s2.as<float>(), S.n, H, W, S.cnt.as<int>(), low.as<unsigned char>(), snow.as<unsigned char>()));
    STC_CUDA(cudaMemcpyAsync(h_cnt.data(), S.cnt.p, S.n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    STC_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<int> rm;
    for (int t = 0; t < S.n; ++t) if ((double)h_cnt[t] / (double)HW > 0.25) rm.push_back(t);
    if (rm.size() > 10) TL_CHECK(tile_delete(ctx, S, rm, false));
  }
  tm.mark("missing-px + snow screening");

  const int n_alloc = S.n;
  STC_CUDA(S.interp.alloc((size_t)n_alloc * HW * 4));
  bool clipped = false;
  if (make_shadow) {
    STC_CUDA(S.cloudshad.alloc((size_t)n_alloc * HW * 4)); STC_CUDA(S.fcps.alloc((size_t)n_alloc * HW));
    STC_CUDA(S.clipbuf.alloc((size_t)n_alloc * HW * 4)); STC_CUDA(S.fa.alloc((size_t)n_alloc * HW * 4)); STC_CUDA(S.fb.alloc((size_t)n_alloc * HW * 4));
    STC_CUDA(S.fsums.alloc(64 * 4));
    auto frac_gt0 = [&](std::vector<int>& rm, double thresh) -> int {          // np.mean(interp > 0, axis=(1, 2)) > thresh
      TL_CHECK(tp_count_gt_dev(ctx, S.interp.as<float>(), S.n, (int)HW, 0.f, S.cnt.as<int>()));
      STC_CUDA(cudaMemcpyAsync(h_cnt.data(), S.cnt.p, S.n * 4, cudaMemcpyDeviceToHost, ctx->stream));
      STC_CUDA(cudaStreamSynchronize(ctx->stream));
      rm.clear();
      for (int t = 0; t < S.n; ++t) if ((double)h_cnt[t] / (double)HW > thresh) rm.push_back(t);
      return STC_OK;
    };
    TL_CHECK(tile_masks(ctx, S, true));
    TL_CHECK(tile_feather(ctx, S));
    tm.mark("cloud masks + feather");
    for (int attempt = 0; attempt < 3; ++attempt) {                             // :863-929, three identical screening rounds
      std::vector<int> rm;
      TL_CHECK(frac_gt0(rm, 0.9));
      if (!rm.empty()) {
        TL_CHECK(tile_delete(ctx, S, rm, true));
        if (S.n < 1) STC_FAIL(STC_ERR_STATE, "tile_run: every date is more than 90 % cloud");
        TL_CHECK(tile_masks(ctx, S, false));
        if (attempt < 2) TL_CHECK(tile_feather(ctx, S));
      }
    }
    TL_CHECK(tile_feather(ctx, S));                                             // :917
    tm.mark("cloud screening rounds + feather");
    std::vector<int32_t> to_remove(S.n, 0);
    int32_t clip_flag = 0;
    TL_CHECK(remove_clouds_dev(ctx, S.s2.as<float>(), S.cloudshad.as<float>(), S.fcps.as<unsigned char>(), S.n, H, W, mt_state,
                               S.interp.as<float>(), to_remove.data(), nullptr, 1, &clip_flag));
    clipped = clip_flag != 0;
    tm.mark("remove_cloud_and_shadows");
    std::vector<int> rm;
    for (int t = 0; t < S.n; ++t) if (to_remove[t]) rm.push_back(t);
    if (!rm.empty()) {                                                          // :972-990
      TL_CHECK(tile_delete(ctx, S, rm, true));
      if (S.n < 1) STC_FAIL(STC_ERR_STATE, "tile_run: every date was fully interpolated");
      TL_CHECK(tile_masks(ctx, S, false));
      TL_CHECK(tile_feather(ctx, S));
    }
  } else {
    STC_CUDA(cudaMemsetAsync(S.interp.p, 0, (size_t)S.n * HW * 4, ctx->stream));
  }
  TL_CHECK(tp_elementwise_dev(ctx, S.dem.as<float>(), HW, 1, 90.f, 0.f));                       // dem / 90 (:995)
  if (!clipped) TL_CHECK(tp_elementwise_dev(ctx, S.s2.as<float>(), (int64_t)S.n * HW * 10, 0, 0.f, 1.f));   // np.clip(sentinel2, 0, 1)
  tm.mark("process_tile tail");

  // ------------------------------------------------------------------ superresolve_large_tile (:95-147) ----
  if (superresolve) {
    const int wsz = 110, P = wsz + 8;
    std::vector<int> xs = tilehost::superres_windows(H, wsz), ys = tilehost::superres_windows(W, wsz);
    const int xb = xs.back(), band_rows = H - xb;            // bottom band = arr[:, xs[-1]:]
    std::vector<SrWin> first, second;
    const bool last_overlaps = ys.size() > 1 && ys.back() < ys[ys.size() - 2] + wsz;
    for (int x : xs)
      for (int y : ys) {
        if (y == ys.back() && x != xb) continue;             // :133-143 never reach the other right-edge windows
        SrWin w{x == xb ? 1 : 0, x == xb ? 0 : x, y};
        if (last_overlaps && x == xb && y == ys.back()) second.push_back(w); else first.push_back(w);
      }
    PoolBuf band, wins, gat, res;
    STC_CUDA(band.alloc((size_t)S.n * band_rows * W * 40));
    const int64_t bt = (int64_t)S.n * band_rows * W * 10;
    TL_LAUNCH(k_band_copy, bt, S.s2.as<float>(), band.as<float>(), S.n, H, W, xb, band_rows, 1);
    auto run = [&](const std::vector<SrWin>& wl) -> int {
      if (wl.empty()) return STC_OK;
      const int nw = (int)wl.size();
      STC_CUDA(wins.alloc(wl.size() * sizeof(SrWin)));
      STC_CUDA(cudaMemcpyAsync(wins.p, wl.data(), wl.size() * sizeof(SrWin), cudaMemcpyHostToDevice, ctx->stream));
      STC_CUDA(gat.alloc((size_t)nw * S.n * P * P * 40)); STC_CUDA(res.alloc((size_t)nw * S.n * P * P * 24));
      { TraceScope ts_(ctx, "k_sr_gather"); k_sr_gather<<<dim3(cdiv(P * P, 256), S.n, nw), 256, 0, ctx->stream>>>(S.s2.as<float>(), band.as<float>(), wins.as<SrWin>(), S.n, H, W, band_rows,
                                                                            wsz, gat.as<float>()); }
      ctx->launches++;
      TL_CHECK(sr_forward_dev(ctx, gat.as<float>(), nullptr, nw * S.n, P, P, res.as<float>()));
      { TraceScope ts_(ctx, "k_sr_scatter"); k_sr_scatter<<<dim3(cdiv(wsz * wsz, 256), S.n, nw), 256, 0, ctx->stream>>>(res.as<float>(), wins.as<SrWin>(), S.n, H, W, band_rows, wsz,
                                                                                 S.s2.as<float>(), band.as<float>()); }
      ctx->launches++;
      STC_CUDA(cudaStreamSynchronize(ctx->stream));          // wl is a host vector
      return STC_OK;
    };
    TL_CHECK(run(first));
    TL_CHECK(run(second));
    // the reference's final write loop copies every bottom-row window out of the (now resolved) band copy: together they cover it
    TL_LAUNCH(k_band_copy, bt, S.s2.as<float>(), band.as<float>(), S.n, H, W, xb, band_rows, 0);
    tm.mark("superresolve_large_tile");
  }

  // ------------------------------------------------------------------ process_subtiles (:1125-1486) ----
  PoolBuf med14, s2q, s1q, s1m, clear, preds;
  STC_CUDA(med14.alloc((size_t)HW * 56));
  {
    std::vector<int32_t> bad(S.n); int64_t nan_total = 0;
    TL_CHECK(tf_s2_medians_dev(ctx, S.s2.as<float>(), S.n, H, W, med14.as<float>(), bad.data(), &nan_total));
    std::vector<int> rm;
    for (int t = 0; t < S.n; ++t) if ((double)bad[t] >= ((double)H * (double)H) / 10.0) rm.push_back(t);       // id_missing_px(arr, 10)
    const bool keep_clm = S.have_clm; S.have_clm = false;                                                      // only s2 / interp / dates from here on
    TL_CHECK(tile_delete(ctx, S, rm, true));
    S.have_clm = keep_clm;
    if (S.n < 1) STC_FAIL(STC_ERR_STATE, "tile_run: no usable date left for smoothing");
  }
  tm.mark("medians of the raw dates");
  STC_CUDA(s2q.alloc((size_t)4 * HW * 56)); STC_CUDA(s1q.alloc((size_t)4 * HW * 8)); STC_CUDA(s1m.alloc((size_t)HW * 8));
  {
    std::vector<float> M;
    auto build_M = [&]() -> int {
      try { tilehost::monthly_operator(S.dates, M); }
      catch (const std::exception& e) { STC_FAIL(STC_ERR_STATE, std::string("tile_run: ") + e.what()); }
      return STC_OK;
    };
    TL_CHECK(build_M());
    std::vector<int32_t> nan_after(S.n, 0);
    TL_CHECK(tf_smooth_quarterly_dev(ctx, S.s2.as<float>(), S.n, H, W, M.data(), S.s1.as<float>(), nullptr, s2q.as<float>(), s1q.as<float>(),
                                     s1m.as<float>(), nan_after.data(), 0));
    std::vector<int> rm;
    for (int t = 0; t < S.n; ++t) if (nan_after[t] > 0) rm.push_back(t);
    if (!rm.empty()) {                                       // deal_w_missing_px :1048-1053: drop the NaN dates, rebuild the operator
      const bool keep_clm = S.have_clm; S.have_clm = false;
      TL_CHECK(tile_delete(ctx, S, rm, true));
      S.have_clm = keep_clm;
      if (S.n < 1) STC_FAIL(STC_ERR_STATE, "tile_run: no NaN-free date left for smoothing");
      TL_CHECK(build_M());
      nan_after.assign(S.n, 0);
      TL_CHECK(tf_smooth_quarterly_dev(ctx, S.s2.as<float>(), S.n, H, W, M.data(), S.s1.as<float>(), nullptr, s2q.as<float>(), s1q.as<float>(),
                                       s1m.as<float>(), nan_after.data(), 1));
    }
  }
  tm.mark("smooth + quarterly composites");
  // window table (:1295-1317, :1369-1388)
  std::vector<long> folder, arrw;
  try { tilehost::subtile_windows(H, W, size, size != 222 ? 6 : 7, folder, arrw); }
  catch (const std::exception& e) { STC_FAIL(STC_ERR_STATE, std::string("tile_run: ") + e.what()); }
  const int nt = (int)(folder.size() / 4);
  if (nt > 64) STC_FAIL(STC_ERR_ARG, "tile_run: more than 64 subtiles");
  std::vector<int32_t> table((size_t)nt * 12, 0);
  {
    int pad_u = -1, pad_d = -1;                              // deliberately persistent across subtiles, like the reference's loop variables
    for (int t = 0; t < nt; ++t) {
      const int start_x = (int)arrw[t * 4], start_y = (int)arrw[t * 4 + 1];
      const int nr = std::min(start_x + (int)arrw[t * 4 + 2], H) - start_x, nc = std::min(start_y + (int)arrw[t * 4 + 3], W) - start_y;
      int32_t* row = table.data() + (size_t)t * 12;
      row[0] = start_x; row[1] = start_y; row[2] = nr; row[3] = nc;
      if (nc == size + 7) {
        pad_u = start_y == 0 ? 7 : 0; pad_d = start_y != 0 ? 7 : 0;
        row[6] = pad_u; row[7] = pad_d; row[10] = pad_u; row[11] = pad_d;
      }
      if (nr == size + 7) {
        if (pad_u < 0) STC_FAIL(STC_ERR_STATE, "tile_run: the reference raises NameError here (pad_u read before assignment, :1388)");
        row[4] = start_x == 0 ? 7 : 0; row[5] = start_x != 0 ? 7 : 0;
        row[8] = pad_u; row[9] = pad_d;
      }
    }
  }
  STC_CUDA(clear.alloc((size_t)HW * 4)); STC_CUDA(preds.alloc((size_t)nt * size * size * 4));
  TL_CHECK(tp_count_lt_axis0_dev(ctx, S.interp.as<float>(), S.n, HW, 0.33f, clear.as<int>()));           // np.sum(interp < 0.33, axis=0)
  std::vector<int32_t> no_data(nt, 0);
  TL_CHECK(tf_process_subtiles_dev(ctx, s2q.as<float>(), s1q.as<float>(), med14.as<float>(), s1m.as<float>(), S.dem.as<float>(), clear.as<int>(),
                                   H, W, nt, table.data(), size, 4, length, S.dates.size() < 2 ? 1 : 0, min17, max17, preds.as<float>(), no_data.data(),
                                   nullptr, nullptr));
  tm.mark("subtile gather + forward + post-filters");
  if (subtile_preds_host) STC_CUDA(cudaMemcpyAsync(subtile_preds_host, preds.p, (size_t)nt * size * size * 4, cudaMemcpyDeviceToHost, ctx->stream));

  // ------------------------------------------------------------------ load_mosaic_predictions (:1515-1641) ----
  // files are processed/<folder_y>/<folder_x>.npy; the mosaic walks x = folder_y directories, y = folder_x files.  The
  // reference's layer order is its os.listdir order (file-system dependent); here: ascending (x, y).
  {
    const int SS = size * size;
    std::vector<int> order(nt);
    for (int t = 0; t < nt; ++t) order[t] = t;
    std::sort(order.begin(), order.end(), [&](int a, int b) {
      if (folder[a * 4 + 1] != folder[b * 4 + 1]) return folder[a * 4 + 1] < folder[b * 4 + 1];
      return folder[a * 4] < folder[b * 4];
    });
    std::vector<int32_t> xs(nt), ys(nt);
    long max_x = 0, max_y = 0;
    for (int k = 0; k < nt; ++k) { xs[k] = (int32_t)folder[order[k] * 4 + 1]; ys[k] = (int32_t)folder[order[k] * 4]; max_x = std::max<long>(max_x, xs[k]); max_y = std::max<long>(max_y, ys[k]); }
    const int Hc = (int)max_x + size, Wc = (int)max_y + size;
    if (Hc != out_h || Wc != out_w) STC_FAIL(STC_ERR_ARG, "tile_run: the mosaic is " + std::to_string(Hc) + " x " + std::to_string(Wc) + " px, the output buffer is not");
    PoolBuf P, dxs, dys, dpl, sums, valid, diffs, dg, dm, tmp, outd;
    STC_CUDA(P.alloc((size_t)nt * SS * 4)); STC_CUDA(dxs.alloc(nt * 4)); STC_CUDA(dys.alloc(nt * 4)); STC_CUDA(dpl.alloc(nt * 4));
    STC_CUDA(sums.alloc(nt * 4)); STC_CUDA(valid.alloc(nt * 4)); STC_CUDA(diffs.alloc((size_t)nt * SS * 4)); STC_CUDA(dg.alloc((size_t)SS * 4));
    STC_CUDA(dm.alloc(nt * 4)); STC_CUDA(tmp.alloc((size_t)Hc * Wc)); STC_CUDA(outd.alloc((size_t)Hc * Wc));
    bool identity = true;
    for (int k = 0; k < nt; ++k) identity = identity && order[k] == k;
    if (identity) STC_CUDA(cudaMemcpyAsync(P.p, preds.p, (size_t)nt * SS * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    else for (int k = 0; k < nt; ++k)
      STC_CUDA(cudaMemcpyAsync(P.as<float>() + (size_t)k * SS, preds.as<float>() + (size_t)order[k] * SS, (size_t)SS * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    STC_CUDA(cudaMemcpyAsync(dxs.p, xs.data(), nt * 4, cudaMemcpyHostToDevice, ctx->stream));
    STC_CUDA(cudaMemcpyAsync(dys.p, ys.data(), nt * 4, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<float> gauss((size_t)SS);
    if (gauss_host) memcpy(gauss.data(), gauss_host, (size_t)SS * 4);
    else {                                                   // fspecial_gauss(size, 36) (:1489-1501), float64 then cast
      const int lo_py = (int)std::floor(-(double)size / 2.0) + 1;      // np.mgrid[-size//2 + 1 : size//2 + 1]
      for (int i = 0; i < size; ++i)
        for (int j = 0; j < size; ++j) {
          const double x = lo_py + i, y = lo_py + j;
          gauss[(size_t)i * size + j] = (float)std::exp(-((x * x + y * y) / (2.0 * 36.0 * 36.0)));
        }
    }
    STC_CUDA(cudaMemcpyAsync(dg.p, gauss.data(), (size_t)SS * 4, cudaMemcpyHostToDevice, ctx->stream));
    // :1570-1573: values < 255 are scaled by 100; a subtile whose sum equals S*S*255 is skipped
    std::vector<float> h_sums(nt); std::vector<int32_t> h_valid(nt), placed(nt);
    TL_CHECK(post_np_sum_dev(ctx, P.as<float>(), nt, SS, 1, sums.as<float>(), valid.as<int>()));
    STC_CUDA(cudaMemcpyAsync(h_sums.data(), sums.p, nt * 4, cudaMemcpyDeviceToHost, ctx->stream));
    STC_CUDA(cudaStreamSynchronize(ctx->stream));
    bool all_placed = true;
    for (int k = 0; k < nt; ++k) { placed[k] = h_sums[k] < (float)((double)SS * 255.0) ? 1 : 0; all_placed = all_placed && placed[k]; }
    STC_CUDA(cudaMemcpyAsync(dpl.p, placed.data(), nt * 4, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<float> mult(nt, 1.f);
    if (all_placed) {                                        // an unplaced subtile makes calc_overlap raise -> no reweighting (:1597-1608)
      TL_CHECK(pre_gauss_mosaic_dev(ctx, P.as<float>(), dxs.as<int>(), dys.as<int>(), dpl.as<int>(), nullptr, nullptr, diffs.as<float>(), 0, nt, size,
                                    0, 0, nullptr, nullptr));
      TL_CHECK(post_np_sum_dev(ctx, diffs.as<float>(), nt, SS, 2, sums.as<float>(), valid.as<int>()));
      STC_CUDA(cudaMemcpyAsync(h_sums.data(), sums.p, nt * 4, cudaMemcpyDeviceToHost, ctx->stream));
      STC_CUDA(cudaMemcpyAsync(h_valid.data(), valid.p, nt * 4, cudaMemcpyDeviceToHost, ctx->stream));
      STC_CUDA(cudaStreamSynchronize(ctx->stream));
      std::vector<float> ratios(nt);
      for (int k = 0; k < nt; ++k) ratios[k] = (float)((double)h_sums[k] / (double)h_valid[k]);          // np.nanmean's scalar path
      const float med = tilehost::median_f32(ratios);
      for (int k = 0; k < nt; ++k) { volatile float q = med / ratios[k]; mult[k] = q > 1.5f ? 1.5f : q; }
    }
    STC_CUDA(cudaMemcpyAsync(dm.p, mult.data(), nt * 4, cudaMemcpyHostToDevice, ctx->stream));
    TL_CHECK(pre_gauss_mosaic_dev(ctx, P.as<float>(), dxs.as<int>(), dys.as<int>(), dpl.as<int>(), dg.as<float>(), dm.as<float>(), nullptr, 1, nt,
                                  size, Hc, Wc, tmp.as<unsigned char>(), outd.as<unsigned char>()));
    STC_CUDA(cudaMemcpyAsync(out_host, outd.p, (size_t)Hc * Wc, cudaMemcpyDeviceToHost, ctx->stream));
    STC_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (n_kept_host) *n_kept_host = S.n;
  if (dates_kept_host) for (int t = 0; t < S.n; ++t) dates_kept_host[t] = (int32_t)S.dates[t];
  tm.mark("mosaic + download");
  STC_CUDA(cudaGetLastError());
  return STC_OK;
}

// ---- host-logic test hooks (no device, no context) ----
extern "C" int stc_monthly_operator_plan(const int32_t* dates, int n, float* G_out /*[24,n] or NULL*/, float* M_out /*[12,n] or NULL*/) {
  if (!dates || n < 1) return STC_ERR_ARG;
  try {
    std::vector<long> d(dates, dates + n);
    std::vector<float> G, M;
    if (G_out) { tilehost::regrid_matrix(d, G); memcpy(G_out, G.data(), G.size() * 4); }
    if (M_out) { tilehost::monthly_operator(d, M); memcpy(M_out, M.data(), M.size() * 4); }
  } catch (const std::exception&) { return STC_ERR_STATE; }
  return STC_OK;
}
extern "C" int stc_subtile_windows_plan(int Lx, int Ly, int size, int n_rows, int32_t* folder_out /*[nt,4]*/, int32_t* array_out /*[nt,4]*/, int cap) {
  try {
    std::vector<long> f, a;
    tilehost::subtile_windows(Lx, Ly, size, n_rows, f, a);
    const int nt = (int)(f.size() / 4);
    if (nt > cap) return STC_ERR_ARG;
    for (size_t i = 0; i < f.size(); ++i) { if (folder_out) folder_out[i] = (int32_t)f[i]; if (array_out) array_out[i] = (int32_t)a[i]; }
    return nt;
  } catch (const std::exception&) { return STC_ERR_STATE; }
}
extern "C" int stc_adjust_shape_plan(int len, int target, int32_t* shift_out, int32_t* out_len) {
  if (len < 1 || target < 1 || !shift_out || !out_len) return STC_ERR_ARG;
  const tilehost::AxisPlan p = tilehost::adjust_axis(len, target);
  *shift_out = p.shift; *out_len = p.out_len;
  return STC_OK;
}
