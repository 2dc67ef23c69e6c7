// Batched, multi-block exact order statistics (np.median / np.percentile building block).
//
// The reference takes medians and percentiles of whole-image vectors in many places (20 per date in make_aligned_mosaic,
// cloud_removal.py:598-677; EVI percentiles of the fit sample :455-467; clear-sky brightness :1458-1481; the Sentinel-1
// fill, download_and_predict_job.py:702-705).  Round 1 ran one thread block per vector (MSB-first radix select, 4-5
// passes of one block over 4e5 values): 12-24 blocks on 148 SMs, 3 % of the HBM roofline, the largest share of the
// preprocessing chain.  Here every pass is spread over the whole GPU, and there are three of them (key digits of 11, 11
// and 10 bits; the first version of this file read the data five times: four 8-bit digits plus a successor pass):
//   hist  : grid (chunks, jobs); each block histograms the next digit of its slice of the job (all <= 16 interleaved
//           columns of a row-major matrix at once, so [K][10] band matrices are read coalesced, once per pass) in shared
//           memory and flushes the non-zero bins to the job's global histogram.  A thread keeps ONE column (the block
//           strides by a multiple of the column count), so the loop has no division and the last pass can keep, per thread,
//           the smallest key ABOVE the selected 22-bit bucket;
//   pick  : one warp per (job, column) scans the bins, fixes the digit and the remaining rank.  After the last digit a bin
//           is a single key value, so the same scan yields the number of keys <= x_(k) and the next larger key inside the
//           bucket; with the smallest key above the bucket that is everything NumPy's even-length median and
//           linear-interpolation percentile need (x_(k+1)) -- no successor pass over the data.
// Results are exact (bit patterns), independent of the block schedule (integer atomics only).
// Keys: order-preserving map of the float32 bits; +NaN (0x7fc00000, what producers write for "not selected") sorts last,
// so ranks below the number of valid values never see it.
#include "stc_common.cuh"
#include "stc_select.cuh"

namespace {

constexpr int SEL_BINS = 2048;            // bins of the widest digit
constexpr int SEL_ITERS = 128;            // elements per thread, block and pass

struct SelState { unsigned prefix; int kth; int cnt_le; unsigned min_gt; };

__device__ __forceinline__ unsigned sel_key(float v) {
  unsigned u = __float_as_uint(v);
  return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float sel_unkey(unsigned u) {
  u ^= (u >> 31) ? 0x80000000u : 0xffffffffu;
  return __uint_as_float(u);
}
__host__ __device__ __forceinline__ int sel_stride(int cols) { return (512 / cols) * cols; }    // active threads = elements per iteration

__global__ void __launch_bounds__(256) k_sel_init(const int* __restrict__ ks, SelState* __restrict__ st, int* __restrict__ hist, int total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) { st[i].prefix = 0; st[i].kth = ks[i]; st[i].cnt_le = 0; st[i].min_gt = 0xffffffffu; }
  if (i < (int64_t)total * SEL_BINS) hist[i] = 0;
  for (int64_t q = i + (int64_t)gridDim.x * blockDim.x; q < (int64_t)total * SEL_BINS; q += (int64_t)gridDim.x * blockDim.x) hist[q] = 0;
}

// digit = (key >> shift) & (nbins - 1); keys must match the job column's prefix in the bits above the digit.
// 512 threads, eight independent loads in flight per thread: with one load per thread and iteration the pass ran at
// 1-2 TB/s (latency-bound: 80 KB of histogram per block leave room for two blocks per SM).
constexpr int SEL_THREADS = 512;
constexpr int SEL_UNROLL = 8;
template <bool LAST>
__global__ void __launch_bounds__(SEL_THREADS) k_sel_hist(const SelJob* __restrict__ jobs, SelState* __restrict__ st, int* __restrict__ hist, int shift, int nbins) {
  const SelJob j = jobs[blockIdx.y];
  const int S = sel_stride(j.cols);
  const int64_t total = (int64_t)j.rows * j.cols;
  const int64_t e0 = (int64_t)blockIdx.x * S * SEL_ITERS;
  if (e0 >= total) return;
  const int64_t e1 = e0 + (int64_t)S * SEL_ITERS < total ? e0 + (int64_t)S * SEL_ITERS : total;
  extern __shared__ int h[];                         // [cols][nbins]
  __shared__ unsigned pre[SEL_MAX_COLS], mn[SEL_MAX_COLS];
  for (int i = threadIdx.x; i < j.cols * nbins; i += blockDim.x) h[i] = 0;
  if (threadIdx.x < j.cols) { pre[threadIdx.x] = st[blockIdx.y * SEL_MAX_COLS + threadIdx.x].prefix; mn[threadIdx.x] = 0xffffffffu; }
  __syncthreads();
  const unsigned lanes = __ballot_sync(0xffffffffu, (int)threadIdx.x < S);       // the warp's lanes that work (all but the block's last few)
  if ((int)threadIdx.x < S) {
    const int c = (int)threadIdx.x % j.cols;         // e0 and S are multiples of cols: the column never changes
    const int hi_shift = shift + (LAST ? 10 : 11);   // bits above the digit (the digit of the last pass is 10 bits wide)
    const unsigned want = (hi_shift >= 32) ? 0u : (pre[c] >> hi_shift);
    int* hc = h + c * nbins;
    unsigned above = 0xffffffffu;
    const unsigned dmask = (unsigned)nbins - 1u;
    const float* src = j.data + (j.ld == j.cols ? e0 + threadIdx.x : ((e0 + threadIdx.x) / j.cols) * (int64_t)j.ld + c);
    const int64_t step = (j.ld == j.cols) ? S : (int64_t)(S / j.cols) * j.ld;
    // a single column puts all lanes of a warp on the same few bins in the first pass: lanes with equal digits are merged
    const bool merge = j.cols <= 2;
    const int rounds = (int)((e1 - e0 + (int64_t)S * SEL_UNROLL - 1) / ((int64_t)S * SEL_UNROLL));      // the same for every thread
    int64_t e = e0 + threadIdx.x;
    for (int k = 0; k < rounds; ++k, e += (int64_t)S * SEL_UNROLL, src += step * SEL_UNROLL) {
      unsigned u[SEL_UNROLL];
#pragma unroll
      for (int q = 0; q < SEL_UNROLL; ++q) u[q] = (e + (int64_t)q * S < e1) ? sel_key(src[q * step]) : 0u;
#pragma unroll
      for (int q = 0; q < SEL_UNROLL; ++q) {
        const bool in = e + (int64_t)q * S < e1;
        const unsigned top = (hi_shift >= 32) ? 0u : (u[q] >> hi_shift);
        const bool hit = in && top == want;
        const int bin = (int)((u[q] >> shift) & dmask);
        if (merge) {
          const unsigned act = __ballot_sync(lanes, hit);
          if (hit) {
            const unsigned peers = __match_any_sync(act, c * nbins + bin);
            if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hc[bin], __popc(peers));
          }
        } else if (hit) atomicAdd(&hc[bin], 1);
        if (LAST && in && top > want && u[q] < above) above = u[q];
      }
    }
    if (LAST && above != 0xffffffffu) atomicMin(&mn[c], above);
  }
  __syncthreads();
  int* g = hist + (int64_t)blockIdx.y * SEL_MAX_COLS * SEL_BINS;
  for (int i = threadIdx.x; i < j.cols * nbins; i += blockDim.x) {
    const int v = h[i];
    if (v) atomicAdd(g + (i / nbins) * SEL_BINS + (i % nbins), v);
  }
  if (LAST && threadIdx.x < j.cols && mn[threadIdx.x] != 0xffffffffu)
    atomicMin(&st[blockIdx.y * SEL_MAX_COLS + threadIdx.x].min_gt, mn[threadIdx.x]);
}

// one warp per (job, column)
template <bool LAST>
__global__ void __launch_bounds__(256) k_sel_pick(const SelJob* __restrict__ jobs, SelState* __restrict__ st, int* __restrict__ hist,
                                                  const int* __restrict__ ks, int shift, int nbins, int njobs) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= njobs * SEL_MAX_COLS) return;
  if ((i % SEL_MAX_COLS) >= jobs[i / SEL_MAX_COLS].cols) return;
  int* g = hist + (int64_t)i * SEL_BINS;
  const int per = nbins / 32, b0 = lane * per;
  int mine = 0;
  for (int q = 0; q < per; ++q) mine += g[b0 + q];
  int incl = mine;
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  const int excl = incl - mine;
  const int kk0 = st[i].kth;
  // the lane whose bins hold rank kk0 (the last lane when the rank is out of range: same clamp as a serial walk)
  const unsigned has = __ballot_sync(0xffffffffu, incl > kk0);
  const int L = has ? (__ffs(has) - 1) : 31;
  int b = 0, kk = 0, gb = 0, nxt = -1;
  if (lane == L) {
    kk = kk0 - excl; b = b0;
    while (b < b0 + per - 1 && kk >= g[b]) { kk -= g[b]; ++b; }
    if (!has) { while (b < nbins - 1 && kk >= g[b]) { kk -= g[b]; ++b; } }
    gb = g[b];
    if (LAST) for (int q = b + 1; q < b0 + per; ++q) if (g[q]) { nxt = q; break; }
  }
  b = __shfl_sync(0xffffffffu, b, L); kk = __shfl_sync(0xffffffffu, kk, L); gb = __shfl_sync(0xffffffffu, gb, L);
  if (LAST) {
    // first non-empty bin above b: in lane L's remaining bins, else the first non-empty bin of the lowest later lane
    if (lane > L) { nxt = -1; for (int q = 0; q < per; ++q) if (g[b0 + q]) { nxt = b0 + q; break; } }
    unsigned cand = (lane >= L && nxt >= 0) ? (unsigned)nxt : 0xffffffffu;
    for (int o = 16; o > 0; o >>= 1) { const unsigned t = __shfl_xor_sync(0xffffffffu, cand, o); cand = t < cand ? t : cand; }
    if (lane == 0) {
      const unsigned a = st[i].prefix | ((unsigned)b << shift);
      st[i].prefix = a; st[i].kth = kk;
      st[i].cnt_le = ks[i] - kk + gb;                              // keys < a: ks - kk; keys == a: gb
      if (cand != 0xffffffffu) st[i].min_gt = (st[i].prefix & ~((unsigned)nbins - 1u)) | cand;     // inside the bucket: smaller than anything above it
    }
  } else if (lane == 0) {
    st[i].kth = kk; st[i].prefix |= ((unsigned)b << shift);
  }
  __syncwarp();
  for (int q = lane; q < nbins; q += 32) g[q] = 0;
}

// out[i] = {x_(k), x_(k+1)}: the next statistic repeats x_(k) when more than k+1 values are <= it, otherwise it is the
// smallest larger value (x_(k) again when there is none)
__global__ void __launch_bounds__(256) k_sel_finish(const SelJob* __restrict__ jobs, const SelState* __restrict__ st, const int* __restrict__ ks,
                                                    float* __restrict__ out, int njobs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= njobs * SEL_MAX_COLS) return;
  if ((i % SEL_MAX_COLS) >= jobs[i / SEL_MAX_COLS].cols) return;
  const float a = sel_unkey(st[i].prefix);
  float b = a;
  if (st[i].cnt_le < ks[i] + 2 && st[i].min_gt != 0xffffffffu) b = sel_unkey(st[i].min_gt);
  out[2 * i] = a; out[2 * i + 1] = b;
}

}  // namespace

int select_ranks_dev(stc_ctx* ctx, const SelJob* jobs_host, int njobs, const int* ks_dev, float* out_dev) {
  if (njobs < 1) return STC_OK;
  int64_t max_total = 0;
  for (int j = 0; j < njobs; ++j) {
    if (jobs_host[j].cols < 1 || jobs_host[j].cols > SEL_MAX_COLS || jobs_host[j].rows < 1) STC_FAIL(STC_ERR_ARG, "select: bad job");
    max_total = std::max<int64_t>(max_total, (int64_t)jobs_host[j].rows * jobs_host[j].cols);
  }
  const int slots = njobs * SEL_MAX_COLS;
  PoolBuf jobs, st, hist;
  STC_CUDA(jobs.alloc((size_t)njobs * sizeof(SelJob))); STC_CUDA(st.alloc((size_t)slots * sizeof(SelState))); STC_CUDA(hist.alloc((size_t)slots * SEL_BINS * 4));
  const void* staged = ctx_stage(ctx, jobs_host, (size_t)njobs * sizeof(SelJob));
  if (!staged) STC_FAIL(STC_ERR_NOMEM, "select: pinned staging");
  STC_CUDA(cudaMemcpyAsync(jobs.p, staged, (size_t)njobs * sizeof(SelJob), cudaMemcpyHostToDevice, ctx->stream));
  int max_cols = 1; int64_t max_blocks = 1;
  for (int j = 0; j < njobs; ++j) {
    max_cols = std::max(max_cols, jobs_host[j].cols);
    max_blocks = std::max<int64_t>(max_blocks, cdiv((int64_t)jobs_host[j].rows * jobs_host[j].cols, (int64_t)sel_stride(jobs_host[j].cols) * SEL_ITERS));
  }
  const dim3 grid((unsigned)max_blocks, njobs);
  const int pick_blocks = cdiv((int64_t)slots * 32, 256);
  static bool configured = false;
  if (!configured) {
    STC_CUDA(cudaFuncSetAttribute(k_sel_hist<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SEL_MAX_COLS * SEL_BINS * 4));
    STC_CUDA(cudaFuncSetAttribute(k_sel_hist<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SEL_MAX_COLS * SEL_BINS * 4));
    configured = true;
  }
  { TraceScope ts_(ctx, "k_sel_init"); k_sel_init<<<cdiv((int64_t)slots * 64, 256), 256, 0, ctx->stream>>>(ks_dev, st.as<SelState>(), hist.as<int>(), slots); }
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = pass == 0 ? 21 : pass == 1 ? 10 : 0, nbins = pass == 2 ? 1024 : 2048;
    const size_t smem = (size_t)max_cols * nbins * 4;
    if (pass < 2) {
      { TraceScope ts_(ctx, "k_sel_hist"); k_sel_hist<false><<<grid, SEL_THREADS, smem, ctx->stream>>>(jobs.as<SelJob>(), st.as<SelState>(), hist.as<int>(), shift, nbins); }
      { TraceScope ts_(ctx, "k_sel_pick"); k_sel_pick<false><<<pick_blocks, 256, 0, ctx->stream>>>(jobs.as<SelJob>(), st.as<SelState>(), hist.as<int>(), ks_dev, shift, nbins, njobs); }
    } else {
      { TraceScope ts_(ctx, "k_sel_hist"); k_sel_hist<true><<<grid, SEL_THREADS, smem, ctx->stream>>>(jobs.as<SelJob>(), st.as<SelState>(), hist.as<int>(), shift, nbins); }
      { TraceScope ts_(ctx, "k_sel_pick"); k_sel_pick<true><<<pick_blocks, 256, 0, ctx->stream>>>(jobs.as<SelJob>(), st.as<SelState>(), hist.as<int>(), ks_dev, shift, nbins, njobs); }
    }
  }
  { TraceScope ts_(ctx, "k_sel_finish"); k_sel_finish<<<cdiv(slots, 256), 256, 0, ctx->stream>>>(jobs.as<SelJob>(), st.as<SelState>(), ks_dev, out_dev, njobs); }
  STC_CUDA(cudaGetLastError());
  ctx->launches += 8;
  return STC_OK;
}

// Host-buffer entry (include/stc.h): the order statistics of rank ks[c] and ks[c] + 1 of every column of a row-major
// matrix -- the exact building block under np.median / np.percentile.  One job; mainly a test surface for the select.
extern "C" int stc_order_stats_host(stc_ctx* ctx, const float* data, int64_t rows, int cols, int64_t ld, const int32_t* ks, float* out) {
  if (!ctx) return STC_ERR_ARG;
  if (!data || !ks || !out || rows < 1 || cols < 1 || cols > SEL_MAX_COLS || ld < cols || rows * ld >= (1ll << 31))
    STC_FAIL(STC_ERR_ARG, "order_stats: bad arguments (1 <= cols <= 16, ld >= cols, rows * ld < 2^31)");
  for (int c = 0; c < cols; ++c)
    if (ks[c] < 0 || ks[c] >= rows) STC_FAIL(STC_ERR_ARG, "order_stats: rank out of range");
  PoolBuf d_data, d_ks, d_out;
  const size_t nbytes = (size_t)((rows - 1) * ld + cols) * 4;
  STC_CUDA(d_data.alloc(nbytes)); STC_CUDA(d_ks.alloc(SEL_MAX_COLS * 4)); STC_CUDA(d_out.alloc(SEL_MAX_COLS * 8));
  int kslots[SEL_MAX_COLS] = {0};
  for (int c = 0; c < cols; ++c) kslots[c] = ks[c];
  STC_CUDA(cudaMemcpyAsync(d_data.p, data, nbytes, cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaMemcpyAsync(d_ks.p, kslots, sizeof(kslots), cudaMemcpyHostToDevice, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));                      // kslots is a stack array
  SelJob job{(const float*)d_data.p, (int)rows, cols, (int)ld};
  int rc = select_ranks_dev(ctx, &job, 1, d_ks.as<int>(), d_out.as<float>());
  if (rc) return rc;
  STC_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)cols * 8, cudaMemcpyDeviceToHost, ctx->stream));
  STC_CUDA(cudaStreamSynchronize(ctx->stream));
  return STC_OK;
}
