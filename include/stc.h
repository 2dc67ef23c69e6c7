/* libstc.so -- C ABI of the B200-native sentinel-tree-cover hot path.
 *
 * The reference has no FFI: its boundary is Python module-level functions in
 * src/download_and_predict_job.py that close over two tf.Session globals
 * (:1785-1826).  Each entry point below cites the reference function it
 * replaces; sentinel_tree_cover_b200/api.py is the ctypes binding and mirrors
 * the reference signatures (INTEGRATION.md shows the maintainer-side patch).
 *
 * Conventions: every call returns 0 on success or a negative error code;
 * stc_last_error(ctx) gives the message.  All arrays are C-contiguous,
 * channel-innermost (NHWC) exactly as the reference's NumPy arrays.  Plain
 * `*_host` pointers are host memory (pageable or pinned); `*_dev` pointers are
 * device memory obtained from stc_malloc.  A context is bound to one CUDA
 * device + one stream and is not thread-safe (the reference is single-threaded,
 * one session per process); multi-GPU = one context per device/process.
 * There is no CPU fallback: stc_create fails if no sm_100 device is present.
 */
#ifndef STC_H_
#define STC_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct stc_ctx stc_ctx;

#define STC_OK 0
#define STC_ERR_CUDA (-1)
#define STC_ERR_ARG (-2)
#define STC_ERR_STATE (-3)
#define STC_ERR_NOMEM (-4)

/* ---- context (replaces the two tf.compat.v1.Session globals,
 *      src/download_and_predict_job.py:1785-1826) ------------------------- */
int stc_create(int device, stc_ctx** out);
void stc_destroy(stc_ctx* ctx);
const char* stc_last_error(stc_ctx* ctx);
const char* stc_version(void);
/* Number of kernels this context has launched so far (bench.py gpu_launches). */
int64_t stc_launch_count(stc_ctx* ctx);
/* 0 = tcgen05 implicit-GEMM convolutions (default), 1 = SIMT verification
 * kernel (same data layout and fp16 operands; used by tests to bisect). */
int stc_set_conv_impl(stc_ctx* ctx, int impl);

/* ---- weights: canonical tensors of predict_graph-*.pb / superresolve_graph.pb
 *      (names in sentinel_tree_cover_b200/weights.py).  Set every tensor, then
 *      finalize (packs fp16 operand layouts on the device).  `which`: 0 =
 *      predict graph, 1 = super-resolve graph. --------------------------- */
int stc_set_weight(stc_ctx* ctx, const char* name, const float* data, int64_t n);
int stc_finalize_weights(stc_ctx* ctx, int which);

/* ---- device memory helpers (so the host side needs no CUDA binding) ---- */
int stc_malloc(stc_ctx* ctx, size_t bytes, void** dptr);
int stc_free(stc_ctx* ctx, void* dptr);
int stc_malloc_host(stc_ctx* ctx, size_t bytes, void** hptr);
/* same, with cudaHostAllocWriteCombined when write_combined != 0 (upload-only staging buffers) */
int stc_malloc_host_flags(stc_ctx* ctx, size_t bytes, int write_combined, void** hptr); /* pinned */
int stc_free_host(stc_ctx* ctx, void* hptr);
int stc_h2d(stc_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int stc_d2h(stc_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
int stc_sync(stc_ctx* ctx);
/* CUDA-event timing on the context's stream (bench.py): begin/end return ms. */
int stc_timer_begin(stc_ctx* ctx);
int stc_timer_end(stc_ctx* ctx, float* ms);
/* Accumulated device time (ms) and launch count of the convolution kernels
 * since the last reset -- measured with CUDA events around each conv launch
 * when enabled (adds no sync; read after stc_sync). */
int stc_conv_timing(stc_ctx* ctx, int enable_reset, float* total_ms, int64_t* launches);
/* Same, restricted to one kernel instantiation: N output channels, GroupNorm groups
 * accumulated (0 = none) and epilogue mode (0 plain, 1 partial-conv+Swish, 2 Swish, 3 GRU
 * candidate, 4 bias, 5 bias+ReLU).  Call before the reset of stc_conv_timing. */
int stc_conv_timing_kind(stc_ctx* ctx, int N, int groups, int mode, float* total_ms, int64_t* launches);
/* Kernel timeline of the model path (profiling aid): enable != 0 starts recording CUDA events around every kernel
 * of the forward; enable == 0 stops, synchronises and writes `label,slot,start_ms,end_ms` rows to csv_path.
 * A start stamp is the time the kernel became the head of its stream, not the time its first block ran. */
int stc_trace(stc_ctx* ctx, int enable, const char* csv_path);

/* ---- model forward: predict_subtile (src/download_and_predict_job.py:328-369)
 *      = sess.run(predict_logits, {predict_inp: x[B,T+1,H,W,17], predict_length})
 *      x: frames 0..T-1 sequence, frame T the median frame (pb:strided_slice,
 *      pb:strided_slice_1).  `length` is the uniform sequence length
 *      (np.full(B, args.length), :354).  If normalize != 0, normalize_subtile
 *      (:316-325) is applied first with min17/max17.  out: [B,H-14,W-14].
 *      Limits (STC_ERR_ARG otherwise): H and W multiples of 4 and >= 28 (independent: the released graphs are square,
 *      76 / 124 / 172 / 220; the border re-segmentation pass, src/resegment_tiles_wide.py:182-222,478, runs 220 x 684
 *      windows through the same call), 1 <= length <= T.  The patch entry points (stc_predict_patches_*) take 12 months
 *      x 13 bands, H == W.
 *      Date-axis limit of the preprocessing entry points: n <= 32 dates per tile (registers hold a pixel's time series;
 *      the reference has no limit, its date selection leaves <= 24). ---- */
int stc_predict_host(stc_ctx* ctx, const float* x_host, int B, int T, int H, int W, int length,
                     int normalize, const double* min17, const double* max17, float* out_host);
/* Same forward, additionally returning the two feature taps of the --gen_feats path
 * (src/download_and_predict_job.py:1429-1431,1807-1809): early = pb:gru_drop/drop_block2d/cond/Merge (the
 * bidirectional ConvGRU output, 64 ch) centre-cropped to the output size like predict_subtile does (:360-362),
 * late = pb:csse_out_mul/mul (the last block's sSE output, 64 ch).  early/late: [B,H-14,W-14,64] float32;
 * probs_host [B,H-14,W-14] may be NULL. */
int stc_predict_feats_host(stc_ctx* ctx, const float* x_host, int B, int T, int H, int W, int length,
                           int normalize, const double* min17, const double* max17,
                           float* probs_host, float* early_host, float* late_host);
int stc_predict_dev(stc_ctx* ctx, const float* x_dev, int B, int T, int H, int W, int length,
                    int normalize, const double* min17, const double* max17, float* out_dev);

/* ---- assemble (process_subtiles :1274-1283 quarterly medians, :1152-1160 /
 *      :1174 median frame, :1398-1407 17-channel layout; indices
 *      src/preprocessing/indices.py:4-54 as in the 13-band contract of
 *      src/download_and_predict_job_multiyear.py:794-838).
 *      in : monthly [B,12,H,W,13]  (10 S2 bands, DEM, S1 VV, S1 VH)
 *      out: [B,5,H,W,17] frames 0-3 = median of months (3q..3q+2), frame 4 =
 *      median over the 12 months; channels [0:10] S2,[10] DEM,[11:13] S1,
 *      [13:17] EVI,BI,MSAVI2,GRNDVI (computed per month, then the same medians). */
int stc_assemble_dev(stc_ctx* ctx, const float* monthly_dev, int B, int H, int W, float* out_dev);
int stc_assemble_host(stc_ctx* ctx, const float* monthly_host, int B, int H, int W, float* out_host);

/* ---- fused tile path used by the throughput benchmark:
 *      assemble -> normalize_subtile -> predict, monthly [B,12,H,W,13] -> [B,H-14,W-14] */
int stc_predict_patches_host(stc_ctx* ctx, const float* monthly_host, int B, int H, int W,
                             const double* min17, const double* max17, float* out_host);
int stc_predict_patches_dev(stc_ctx* ctx, const float* monthly_dev, int B, int H, int W,
                            const double* min17, const double* max17, float* out_dev);

/* Same, for uint16 patches: integer input follows predict_subtile's `subtile / 65535.`
 * convention (:345-347); halves the host->device bytes of the tile path. */
int stc_predict_patches_u16_host(stc_ctx* ctx, const uint16_t* monthly_host, int B, int H, int W,
                                 const double* min17, const double* max17, float* out_host);
int stc_predict_patches_u16_dev(stc_ctx* ctx, const uint16_t* monthly_dev, int B, int H, int W,
                                const double* min17, const double* max17, float* out_dev);

/* ---- temporal regrid + Whittaker + monthly mean as one linear operator
 *      out[12,P,C] = M[12,n] . in[n,P,C]   (P = H*W pixels)
 *      replaces calculate_and_save_best_images (src/downloading/utils.py:176-347)
 *      + Smoother.interpolate_array (src/preprocessing/whittaker_smoother.py:44-69);
 *      M is built on the host from the image dates (regrid.py). ------------- */
int stc_temporal_matmul_host(stc_ctx* ctx, const float* in_host, const float* M_host,
                             int n_in, int n_out, int64_t inner, float* out_host);
int stc_temporal_matmul_dev(stc_ctx* ctx, const float* in_dev, const float* M_host,
                            int n_in, int n_out, int64_t inner, float* out_dev);

/* ---- band indices, make_indices (src/download_and_predict_job.py:998-1006):
 *      in [npix,C>=10] -> out [npix,4] = EVI,BI,MSAVI2,GRNDVI ------------- */
int stc_indices_host(stc_ctx* ctx, const float* in_host, int64_t npix, int C, float* out_host);

/* ---- temporal median over the leading axis (np.median(axis=0),
 *      process_subtiles :1152-1160; n <= 32) ------------------------------ */
int stc_temporal_median_host(stc_ctx* ctx, const float* in_host, int n, int64_t inner, float* out_host);

/* ---- 20 m -> 10 m super-resolution: one sess.run of superresolve_graph.pb
 *      (src/download_and_predict_job.py:110-120 `_worker_fn` body without the
 *      pad/crop): x [N,H,W,10] (already reflect-padded by the caller),
 *      bilinear [N,H,W,6] -> out [N,H,W,6] = pb:Add_2 -------------------- */
int stc_superresolve_host(stc_ctx* ctx, const float* x_host, const float* bilinear_host,
                          int N, int H, int W, float* out_host);
/* Device-resident variant (x_dev [N,H,W,10], out_dev [N,H,W,6]; bilinear_dev may be NULL = x[..., 4:]); asynchronous. */
int stc_superresolve_dev(stc_ctx* ctx, const float* x_dev, const float* bilinear_dev, int N, int H, int W, float* out_dev);

/* ---- Gaussian overlap-blend mosaic, load_mosaic_predictions depth == 1
 *      (src/download_and_predict_job.py:1515-1641).  preds [n,S,S]: subtile predictions as
 *      saved (probabilities 0..1, 255 = no data), in the reference's layer order; xs/ys [n]
 *      canvas offsets; placed [n] = 0 for all-255 subtiles (:1573).  Step 1 returns, per
 *      subtile, the calc_overlap map |nanmean(others) - self| [n,S,S] (:1503-1512, NaN where
 *      nothing overlaps); the host takes its nanmean and forms the capped multipliers
 *      median(r)/r (:1603-1606) with NumPy; step 2 blends with fspecial_gauss weights (gauss [S,S] float32),
 *      applies the uint8 / no-data rules and the 10x 3x3 dilation of 255 (:1609-1640). ---- */
int stc_mosaic_diffs_host(stc_ctx* ctx, const float* preds_host, const int32_t* xs, const int32_t* ys,
                          const int32_t* placed, int n, int S, float* diffs_host);
int stc_gauss_mosaic_host(stc_ctx* ctx, const float* preds_host, const int32_t* xs, const int32_t* ys,
                          const int32_t* placed, const float* gauss_host, const float* mult_host,
                          int n, int S, int out_h, int out_w, uint8_t* out_host);
/* Feature mosaic: load_mosaic_predictions with depth > 1 (src/download_and_predict_job.py:1540-1592,1628-1635).
 * feats [n,S,S,D] int16 as saved by process_subtiles --gen_feats (feats/<y>/<x>.npy), layers in the reference's
 * os.listdir order; plain Gaussian weights normalised over the layers, nansum, int16 truncation.
 * out: [D,out_h,out_w] int16.  (The reference walks the depth axis 8 channels at a time for memory reasons only.) */
int stc_feature_mosaic_host(stc_ctx* ctx, const int16_t* feats_host, const int32_t* xs, const int32_t* ys,
                            const float* gauss_host, int n, int S, int D, int out_h, int out_w, int16_t* out_host);

/* ---- np.sum of `nseg` contiguous float32 segments of length `len` in NumPy's pairwise order (bit-identical
 *      to np.sum on a contiguous array): mode 0 plain; mode 1 values < 255 are multiplied by 100 first (the
 *      in-place scaling + "is this subtile all no-data" sum of load_mosaic_predictions,
 *      src/download_and_predict_job.py:1570-1573); mode 2 NaN -> 0 and valid[s] = number of non-NaN values
 *      (np.nanmean of calc_overlap's difference maps, :1503-1512).  valid_host may be NULL. ---- */
int stc_np_sum_host(stc_ctx* ctx, const float* data_host, int nseg, int len, int mode, float* sum_host, int32_t* valid_host);

/* ---- normalize_subtile (src/download_and_predict_job.py:316-325), in place on x [npx,C] float32:
 *      clip to [min,max] then (x - (max+min)/2) / ((max-min)/2), constants are Python floats (double). ---- */
int stc_normalize_host(stc_ctx* ctx, float* x_host, int64_t npx, int C, const double* mins, const double* maxs);

/* ---- identify_bright_bare_surfaces (src/download_and_predict_job.py:1099-1122): img [F,H,W,C>=9] float32 ->
 *      ramp [(H-14),(W-14)] float64 = min(EDT,3)/3 of the opened bright-surface mask, cropped by 7. ---- */
int stc_bright_bare_host(stc_ctx* ctx, const float* img_host, int F, int H, int W, int C, double* ramp_host);

/* ---- post-filters of the subtile loop (src/download_and_predict_job.py:1408-1409,1451-1483): bright-bare
 *      attenuation, no-image block vote -> 255 (S == 158: 4x4 blocks of 40 px, > 25 %; S == 142: 9x9 blocks of
 *      16 px, > 75 %), np.around(.., 3).  preds [S,S] float32, img [F,S+14,S+14,C] (stack BEFORE
 *      normalisation), min_clear [S+14,S+14] float32 (min_clear_images_per_date before its [6:-6] crop)
 *      -> out [S,S] float32. ---- */
int stc_postprocess_subtile_host(stc_ctx* ctx, const float* preds_host, const float* img_host, const float* min_clear_host,
                                 int S, int F, int C, float* out_host);

/* ---- array work of process_tile's front half (src/download_and_predict_job.py:684-832, 995-997) ----
 * stc_s1_fill_host:        s1 [m,H,W,C] float32 in place: values == 1 of every date are replaced by np.median over
 *                          ALL values of that date (`s1_i[s1_i == 1] = np.median(s1_i[s1_i < 65535], axis=0)`, :702-705).
 * stc_median_filter5_host: scipy.ndimage.median_filter(dem, size=5) (mode 'reflect'), :713.
 * stc_clm_pairs_host:      Sen2Cor mask [n,H,W] float32 in place: in date order, a pixel set in two consecutive
 *                          dates is cleared in both (:688-695).
 * stc_snow_host:           snow_filter(sentinel2) > 0 (:808-825): per_date[t] = flagged pixels of date t;
 *                          snow [H,W] uint8 = 1 - binary_dilation(mean_t < 0.7, iterations=2) (:827-829).
 * stc_count_gt_host:       counts[s] = #(data[s,:] > thresh) (the np.mean(interp > 0, axis=(1,2)) tests, :863,880,...).
 * stc_count_lt_axis0_host: out[i] = #(data[t,i] < thresh over the n dates) (`np.sum(interp_tile < 0.33, axis=0)`, :1355).
 * stc_elementwise_host:    mode 0 np.clip(x, a, b) in place (:996), mode 1 x / a in place (`dem / 90`, :995),
 *                          mode 2 NaN -> a in place (interpolate_na_vals, src/preprocessing/interpolation.py:42-56: the
 *                          NaN-propagating median of a column that holds a NaN is NaN, reset to 0, so every NaN becomes 0).
 * stc_max_masked_host:     a = np.maximum(a, b) after `b[zero] = 0` (clm[fcps] = 0; cloudshad = max(cloudshad, clm), :842-845);
 *                          zero may be NULL. ---- */
int stc_s1_fill_host(stc_ctx* ctx, float* s1_host, int m, int H, int W, int C);
int stc_median_filter5_host(stc_ctx* ctx, const float* in_host, int H, int W, float* out_host);
int stc_clm_pairs_host(stc_ctx* ctx, float* clm_host, int n, int H, int W);
int stc_snow_host(stc_ctx* ctx, const float* s2_host, int n, int H, int W, int32_t* per_date_host, uint8_t* snow_host);
int stc_count_gt_host(stc_ctx* ctx, const float* data_host, int nseg, int len, float thresh, int32_t* counts_host);
int stc_count_lt_axis0_host(stc_ctx* ctx, const float* data_host, int n, int64_t len, float thresh, int32_t* out_host);
int stc_elementwise_host(stc_ctx* ctx, float* x_host, int64_t n, int mode, float a, float b);
int stc_max_masked_host(stc_ctx* ctx, float* a_host, const float* b_host, const uint8_t* zero_host, int64_t n);

/* ---- fused, device-resident steps of process_subtiles (src/download_and_predict_job.py:1125-1486): one upload per entry,
 *      the kernels of the unfused entry points chained on the device, only the kept results copied back.
 * stc_s2_medians_host: s2 [n,H,W,10] -> median14 [H,W,14] = np.median over the raw dates of the 10 bands and of
 *      EVI/BI/MSAVI2/GRNDVI (:1151-1159); bad_px[n] as stc_missing_px_host; *nan_total = NaN values found, in which case
 *      they are set to 0 (interpolate_na_vals :1148) and s2_host is rewritten.
 * stc_smooth_quarterly_host: s2 [n,H,W,10] (dates already screened) -> running-median fill of the 0/1 sentinels
 *      (deal_w_missing_px :1039-1047), indices, the 12 x n regrid/Whittaker/monthly operator M (smooth_large_tile :1057-1096)
 *      -> s2_monthly [12,H,W,14] (optional), s2_quarterly [4,H,W,14] = medians of months 0-2, 3-5, 6-8, 9-11 (:1274-1276,
 *      optional); s1 [12,H,W,2] (optional) -> s1_quarterly [4,H,W,2], s1_median [H,W,2] (:1174, :1277-1278).
 *      nan_after[n]: NaN values per date after the fill; if any is non-zero nothing else is computed and the caller
 *      drops those dates first (:1048-1053).
 * stc_predict_postprocess_host: x [B,T+1,H,H,17] un-normalised subtile stacks, min_clear [B,H,H] float32, no_data[B]
 *      -> out [B,H-14,H-14]: normalize_subtile + forward (:1420-1421) for the batch, 255 fill for no_data subtiles
 *      (:1417), then the post-filters of stc_postprocess_subtile_host (:1451-1483), all on the device. ---- */
int stc_s2_medians_host(stc_ctx* ctx, float* s2_host, int n, int H, int W, float* median14_host, int32_t* bad_px_host,
                        int64_t* nan_total_host);
int stc_smooth_quarterly_host(stc_ctx* ctx, const float* s2_host, int n, int H, int W, const float* M_host, const float* s1_host,
                              float* s2_monthly_host, float* s2_quarterly_host, float* s1_quarterly_host, float* s1_median_host,
                              int32_t* nan_after_host);
int stc_predict_postprocess_host(stc_ctx* ctx, const float* x_host, const float* min_clear_host, const int32_t* no_data_host, int B,
                                 int T, int H, int length, const double* min17, const double* max17, float* out_host);

/* ---- the subtile loop of process_subtiles (:1345-1486) for one tile in one call: the windows are gathered on the device
 *      (reference reflect padding of edge subtiles, :1369-1388) from the quarterly composites s2q [T,H,W,14], s1q [T,H,W,2],
 *      the medians s2med [H,W,14], s1med [H,W,2], dem [H,W] and clear [H,W] int32 (= np.sum(interp < 0.33, axis=0)); 17-channel
 *      stacks, the no-image test np.percentile(min_clear, 50) < 1 on the unpadded window (or force_no_data, :1409), the
 *      normalised batched forward and the post-filters follow on the device.  windows [nt][12] int32 per subtile:
 *      row0, col0, rows, cols of the array window; data pads (rows before/after, cols before/after); min_clear pads (same
 *      order; the reference pads that map with its own, partly stale, variables).  out [nt,S,S] float32, no_data [nt]. ---- */
int stc_process_subtiles_host(stc_ctx* ctx, const float* s2q_host, const float* s1q_host, const float* s2med_host,
                              const float* s1med_host, const float* dem_host, const int32_t* clear_host, int H, int W, int nt,
                              const int32_t* windows_host, int S, int T, int length, int force_no_data,
                              const double* min17, const double* max17, float* out_host, int32_t* no_data_host);

/* Same with the two --gen_feats taps of the forward (stc_predict_feats_host) for every subtile: early / late
 * [nt,S,S,64] float32 (src/download_and_predict_job.py:1429-1431). */
int stc_process_subtiles_feats_host(stc_ctx* ctx, const float* s2q_host, const float* s1q_host, const float* s2med_host,
                                    const float* s1med_host, const float* dem_host, const int32_t* clear_host, int H, int W, int nt,
                                    const int32_t* windows_host, int S, int T, int length, int force_no_data,
                                    const double* min17, const double* max17, float* out_host, int32_t* no_data_host,
                                    float* early_host, float* late_host);

/* ---- storage codecs and the Sentinel-1 dB transform.
 *      to_float32 (src/tof/tof_downloading.py:64-72): uint16 -> x/65535 float32;
 *      to_int16 (:51-61): trunc(clip(x,0,1)*65535) -> uint16;
 *      convert_to_db (src/download_and_predict_job.py:74-89): 10*log10(x+1/65535), floored at
 *      -min_db, rescaled to [0,1]. ------------------------------------------------------ */
int stc_to_float32_host(stc_ctx* ctx, const uint16_t* in_host, int64_t n, float* out_host);
int stc_to_uint16_host(stc_ctx* ctx, const float* in_host, int64_t n, uint16_t* out_host);
/* float_to_int16 (src/download_and_predict_job.py:174-180): NaN -> -32768, clip to +-32.768 (for precision 1000),
 * x * precision, int16 truncation -- the storage codec of the --gen_feats feature stacks. */
int stc_float_to_int16_host(stc_ctx* ctx, const float* in_host, int64_t n, int precision, int16_t* out_host);
int stc_convert_to_db_host(stc_ctx* ctx, const float* in_host, int64_t n, float min_db, float* out_host);

/* ---- cloud-mask feathering, id_areas_to_interp (src/preprocessing/cloud_removal.py:774-798,
 *      closing_size 15) and the same stage of remove_cloud_and_shadows (:913-921, size 20):
 *      per date with sum(mask) > 0:  a = 1 - min(EDT(1-mask),12)/12; a<0.2 -> 0;
 *      grey_closing(a, size) (SciPy 'reflect' semantics).  masks/out [n,H,W] float32. ---- */
int stc_feather_host(stc_ctx* ctx, const float* masks_host, int n, int H, int W, int closing_size, float* out_host);
/* ---- scipy.ndimage.binary_dilation(x, iterations=k) with the 4-connected cross
 *      (connectivity 1) or generate_binary_structure(2,2) (connectivity 2), border_value 0.
 *      in/out [n,H,W] uint8 (0/1). ----------------------------------------------------- */
int stc_binary_dilate_host(stc_ctx* ctx, const uint8_t* in_host, int n, int H, int W, int iterations,
                           int connectivity, uint8_t* out_host);

/* ---- cloud / shadow removal: remove_cloud_and_shadows(tiles, probs, shadows, image_dates, pfcps, sentinel1)
 *      (src/preprocessing/cloud_removal.py:888-973; make_aligned_mosaic :578-699, align_interp_array_randomforest
 *      :316-575, calculate_clouds_in_mosaic :703-732).  tiles [n,H,W,10] float32 are rewritten IN PLACE
 *      (cloudy pixels blended with the NNLS-aligned cloud-free mosaic), probs [n,H,W] float32 is the cloud|shadow
 *      mask, pfcps [n,H,W] uint8 the false-positive mask of stc_cloud_masks_host (only date 0 is read, as in the
 *      reference).  mt_state: the 624 MT19937 words + position of Python's `random.getstate()[1]`; the
 *      reference draws its training sample with random.shuffle, and the same generator is advanced here, so a
 *      pinned random.seed reproduces the reference's sample; the advanced state is written back.
 *      areas_out [n,H,W] float32 = the feathered interpolation weights (+ residual mosaic clouds, clipped to 1);
 *      to_remove_out[n] = 1 for dates that are 100 % interpolated; mosaic_out [H,W,10] optional (may be NULL).
 *      1 <= n <= 32.  Situations in which the reference itself raises (a date with <= 1 % usable pixels, an empty
 *      fit set, a one-element EVI stratum, NNLS not converging) return STC_ERR_STATE with the cause in
 *      stc_last_error. ---- */
int stc_remove_clouds_host(stc_ctx* ctx, float* tiles_host, const float* probs_host, const uint8_t* pfcps_host, int n, int H, int W,
                           uint32_t* mt_state, float* areas_out_host, int32_t* to_remove_out_host, float* mosaic_out_host);
/* Same, for process_tile: when no date has to be removed the cube is clipped to [0, 1] before it is copied back
 * (np.clip(sentinel2, 0, 1), src/download_and_predict_job.py:996, is the next statement on that path) and
 * *clipped_out = 1; otherwise the cube comes back unclipped (the masks are recomputed on it first, :972-990). */
/* Test hook (host only, no device): Python's random.shuffle replayed on data[0..n) from the generator state mt_state
 * (624 MT19937 words + position, `random.getstate()[1]`); the advanced state is written back.  data == NULL walks the
 * generator through the same draws without shuffling anything (the state a following shuffle would start from). */
int stc_py_shuffle(uint32_t* mt_state, int32_t* data, int64_t n);
int stc_remove_clouds_clip_host(stc_ctx* ctx, float* tiles_host, const float* probs_host, const uint8_t* pfcps_host, int n, int H, int W,
                                uint32_t* mt_state, float* areas_out_host, int32_t* to_remove_out_host, int32_t* clipped_out);

/* ---- exact squared Euclidean distance to the nearest non-zero pixel of `target`, searched
 *      within `radius` (radius^2+1 where none): the capped distance_transform_edt call sites
 *      (cap 3: identify_bright_bare_surfaces src/download_and_predict_job.py:1117-1119;
 *       cap 3/5: cloud_removal.py:1333-1336,1608-1611).  target [n,H,W] uint8 -> int32. ---- */
int stc_edt_sq_host(stc_ctx* ctx, const uint8_t* target_host, int n, int H, int W, int radius, int32_t* out_host);

/* ---- missing-pixel bookkeeping: id_missing_px (src/preprocessing/interpolation.py:5-23) and the NaN-date
 *      test of deal_w_missing_px (src/download_and_predict_job.py:1048).  arr [n,H,W,C] float32.
 *      bad_px[t] = number of pixels of date t with more than one of the first 10 bands == 0 or >= 1;
 *      nan_vals[t] = number of NaN values of date t.  The caller applies the thresholds. ---- */
int stc_missing_px_host(stc_ctx* ctx, const float* arr_host, int n, int H, int W, int C, int32_t* bad_px_host,
                        int32_t* nan_vals_host);

/* ---- temporal median fill of the 0 / 1 sentinels, deal_w_missing_px (src/download_and_predict_job.py:1039-1047):
 *      in date order, every value == 0 is replaced by the CURRENT np.median over dates of its pixel-band
 *      column (earlier dates already filled), then the same for values == 1.  arr [n,H,W,C] float32 in
 *      place, n <= 96; nan_vals[t] = NaN values of date t afterwards. ---- */
int stc_median_fill_host(stc_ctx* ctx, float* arr_host, int n, int H, int W, int C, int32_t* nan_vals_host);

/* ---- 20 m / 40 m -> 10 m band stack of process_tile (src/download_and_predict_job.py:743-782):
 *      out[...,0:4] = s2_10; out[...,4:8] = bilinear x2 of the four 20 m bands; out[...,8:10] = 2x2 mean
 *      pool then bilinear of the two 40 m bands, with the odd-shape cases of :760-782.  Bilinear =
 *      skimage.transform.resize(order=1) = scipy.ndimage.zoom(order=1, mode='mirror', grid_mode=True).
 *      s2_10 [n,2h,2w,4], s2_20 [n,h,w,6] -> out [n,2h,2w,10] float32. ---- */
int stc_build_sentinel2_host(stc_ctx* ctx, const float* s2_10_host, const float* s2_20_host, int n, int h, int w,
                             float* out_host);

/* ---- multi-temporal cloud / shadow mask: identify_clouds_shadows(img, dem, bbx)
 *      (src/preprocessing/cloud_removal.py:1215-1677) in the configuration the reference tree runs in
 *      (forestmask.tif / urbanmask.tif absent: forest and potential-false-positive masks are zero).
 *      img [T,H,W,10] float32 reflectance (1 <= T <= 32), dem [H,W] float32.
 *      clouds_host [T,H,W] float32 0/1 (clouds | shadows | haze dates), fcps_host [T,H,W] uint8
 *      (the dilated NIR/SWIR false-positive mask the reference returns second).
 *      stage_host / stage_id: test tap -- when stage_host != NULL the uint8 [T,H,W] mask after stage
 *      `stage_id` is copied out (1 Hollstein, 2 raw shadows, 3 cleaned shadows, 4 raw clouds,
 *      5 +brightness/whiteness, 6 after false-positive removal, 7 after shape clean-up, 8 before haze). ---- */
int stc_cloud_masks_host(stc_ctx* ctx, const float* img_host, const float* dem_host, int T, int H, int W,
                         float* clouds_host, uint8_t* fcps_host, uint8_t* stage_host, int stage_id);
/* Ancillary rasters of the tile whose masks are computed next (replaces the file reads of
 * src/preprocessing/cloud_removal.py:735-771: mask_nonurban_areas("urbanmask.tif", ...) inside detect_pfcp :1131-1135 and
 * adjust_cloudmask_in_forests("forestmask.tif", ...) :1254-1257).  Reading the .tif window is the caller's I/O; the arrays
 * come in at tile resolution, uint8 0/1: forest [H,W] = the function's return value; urban_core / urban_near [H,W] = the
 * two resized rasters of mask_nonurban_areas (dilated once :745-747; dilated five more times :751-752).  NULL = that
 * raster is absent (the reference's except-branch: zeros).  The masks stay set until replaced and apply to
 * stc_cloud_masks_host / stc_tile_run_host calls with the same H, W. */
int stc_set_ancillary_masks_host(stc_ctx* ctx, const uint8_t* forest_host, const uint8_t* urban_core_host, const uint8_t* urban_near_host,
                                 int H, int W);

/* ---- one whole tile, device-resident: the body of the reference's main loop
 *      (src/download_and_predict_job.py:1995-2020)  process_tile (:640-997, make_shadow) -> superresolve_large_tile
 *      (:95-147) -> process_subtiles (:1125-1486, prediction path, length 4) -> load_mosaic_predictions (:1515-1641)
 *      in ONE call: the raw cubes are uploaded once, the uint8 tile is downloaded once, every intermediate stays in
 *      pooled device memory and only per-date scalars / integers reach the host, where the reference's control flow on
 *      scalars is replayed (date screening, retry loops, the regrid / Whittaker operator of the surviving dates, window
 *      table, mosaic multipliers).  Same device code as the per-stage entry points above.
 *      s2_10 [n,h10,w10,4], s2_20 [n,h20,w20,6], s1 [12,hs,ws,2] uint16 as stored in raw/s2_10, raw/s2_20, raw/s1
 *      (x / 65535); dem [hd,wd] float32 (raw/misc/dem); clm [n,h20,w20] uint8 Sen2Cor mask (raw/clouds/cloudmask) or
 *      NULL; dates [n] day-of-year of the Sentinel-2 images (raw/misc/s2_dates).  Shapes are aligned to (2*h20, 2*w20)
 *      like adjust_shape (:260-310).  mt_state: Python's MT19937 state (see stc_remove_clouds_host), advanced in place.
 *      gauss: float32 [size,size] fspecial_gauss(size, 36) (:1489) or NULL (computed here).
 *      out [out_h,out_w] uint8 with out_h = 2*w20, out_w = 2*h20 (the mosaic's axes are the file axes y, x; :1515-1641);
 *      layers are blended in ascending (x, y) order (the reference uses its os.listdir order).
 *      dates_kept [n] / n_kept: the image dates that survived; subtile_preds [nt,size,size] float32 (optional) = the
 *      per-subtile arrays process_subtiles would have saved, in window-table order (nt = 36).
 *      Situations in which the reference raises return STC_ERR_STATE with the cause in stc_last_error. ---- */
int stc_tile_run_host(stc_ctx* ctx, const uint16_t* s2_10_host, int n, int h10, int w10, const uint16_t* s2_20_host, int h20,
                      int w20, const uint16_t* s1_host, int m1, int hs, int ws, const float* dem_host, int hd, int wd,
                      const uint8_t* clm_host, const int32_t* dates_host, uint32_t* mt_state, int make_shadow, int superresolve,
                      int size, int length, const double* min17, const double* max17, const float* gauss_host,
                      uint8_t* out_host, int out_h, int out_w, int32_t* dates_kept_host, int32_t* n_kept_host,
                      float* subtile_preds_host);
/* Host-logic hooks of the tile chain (no device, no context; used by the CPU tests to hold the C++ date / window logic
 * against the NumPy mirrors regrid.py / windows.py, which are pinned to the reference):
 *   stc_monthly_operator_plan: G [24,n] = calculate_and_save_best_images weights (src/downloading/utils.py:176-347),
 *                              M [12,n] = pair-mean . Whittaker . G (whittaker_smoother.py:10-69); either may be NULL.
 *   stc_subtile_windows_plan : the window tables of process_subtiles (:1295-1317); returns the number of subtiles.
 *   stc_adjust_shape_plan    : adjust_shape (:260-310) for one axis as out[i] = in[clamp(i + shift, 0, len - 1)]. */
int stc_monthly_operator_plan(const int32_t* dates, int n, float* G_out, float* M_out);
int stc_subtile_windows_plan(int Lx, int Ly, int size, int n_rows, int32_t* folder_out, int32_t* array_out, int cap);
int stc_adjust_shape_plan(int len, int target, int32_t* shift_out, int32_t* out_len);
/* Device-memory pool statistics (scratch buffers of all entry points are recycled, not cudaMalloc'd per call). */
int stc_pool_info(stc_ctx* ctx, int64_t* hits, int64_t* misses, int64_t* cached_bytes, int64_t* total_bytes);

/* ---- debug: copy an internal activation buffer of the last stc_predict_*
 *      call to the host as float32 NHWC (interior only).  Names: "ccin",
 *      "cat2", "p1", "cat1", "p2", "u2in", "u3in".  Returns the number of
 *      floats written (or needed if out_host == NULL). --------------------- */
int64_t stc_debug_read(stc_ctx* ctx, const char* name, float* out_host);

/* ---- region-scale pieces: a region is a grid of R x C overlapping patches
 *      (P-px windows every `stride` px; the reference mosaics one 6 x 6 km tile at a time, :1515-1641, this is its
 *      generalisation).  stc_region_gather_dev cuts B patch windows [B,T,P,P,Cc] out of a device-resident canvas
 *      band [T,Hc,Wc,Cc] (wrap != 0: the canvas is periodic, used to synthesise a region from a small base cube); the
 *      window origins ys/xs [B] are DEVICE arrays so that the call is asynchronous (the caller validates them);
 *      stc_region_blend_dev blends patch outputs preds [rows_have,C,S,S] (first grid row r_first of R) into the
 *      uint8 canvas rows [y0,y1): Gaussian weights gauss [S,S], sum(w * p*100) / sum(w) in row-major patch order,
 *      uint8 truncation, <= 15 -> 0, uncovered -> 255.  Patch (r,c) covers canvas rows r*stride+margin .. +S. ---- */
int stc_region_gather_dev(stc_ctx* ctx, const float* canvas_dev, int T, int Hc, int Wc, int Cc, int wrap,
                          const int32_t* ys_dev, const int32_t* xs_dev, int B, int P, float* out_dev);
/* Forward over n patch windows of the canvas (gather + fused front end + model), `batch` windows at a time with the
 * gather of the next batch overlapped; preds_dev [n, P-14, P-14] float32.  Synchronous (returns when preds_dev is complete). */
int stc_region_predict_dev(stc_ctx* ctx, const float* canvas_dev, int T, int Hc, int Wc, int Cc, int wrap,
                           const int32_t* ys_dev, const int32_t* xs_dev, int n, int batch, int P,
                           const double* min17, const double* max17, float* preds_dev);
int stc_region_blend_dev(stc_ctx* ctx, const float* preds_dev, int r_first, int rows_have, int R, int C, int S, int stride,
                         int margin, const float* gauss_host, int y0, int y1, int Wc, uint8_t* out_host);

/* ---- seam re-segmentation pass (src/resegment_tiles_wide.py) ----
 * align_subtile_histograms(array) :284-345: arr [T,H,W,C] float32 (a border subtile: right edge of a tile | left edge of its
 * neighbour), transformed in place: per time step and band the two column halves (split at `half` = (SIZE + 14) // 2) are
 * rescaled to their common mean / standard deviation (non-water pixels), kept only where that shrinks the jump across
 * column `seam_col` = SIZE // 2 + 7.  applied_out [T] (optional): 1 where the transform was kept. */
int stc_align_histograms_host(stc_ctx* ctx, float* arr_host, int T, int H, int W, int C, int half, int seam_col, int32_t* applied_out);

/* ---- exact order statistics (the building block under every np.median / np.percentile of the path:
 * cloud_removal.py:455-467, 598-677, 1458-1481; download_and_predict_job.py:702-705) ----
 * data: row-major [rows][ld] float32, the first `cols` (<= 16) entries of a row are the columns.  For every column c:
 * out[2c] = the value of 0-based rank ks[c] in sorted order, out[2c + 1] = the value of rank ks[c] + 1 (repeats out[2c] at
 * the last rank).  NaN sorts last.  Bit-exact (the values are elements of the input); three passes over the data. */
int stc_order_stats_host(stc_ctx* ctx, const float* data, int64_t rows, int cols, int64_t ld, const int32_t* ks, float* out);

/* ---- on-disk product: write_tif (src/downloading/io.py:229-263; called per tile at src/download_and_predict_job.py:2033-2036
 * and by the border pass, src/resegment_tiles_wide.py:1240 ff) ----
 * Single-band uint8 GeoTIFF, LZW-compressed strips, EPSG:4326, north-up: what rasterio writes for `driver='GTiff', count=1,
 * dtype='uint8', compress='lzw', crs='+proj=longlat +datum=WGS84 +no_defs', transform=from_bounds(west, south, east, north,
 * width=cols, height=rows)`.  img: row-major [rows][cols] (the reference passes arr.T).  Host-only (no ctx, no device work).
 * STC_ERR_ARG: null / empty / west >= east / south >= north / more than 2^31 pixels; STC_ERR_STATE: the file could not be
 * written (it is written under `<path>.part` and renamed); STC_ERR_NOMEM.
 * stc_geotiff_encode_u8 returns the same bytes in a malloc'ed buffer (*out_buf, *out_len; release with stc_geotiff_free),
 * for callers that upload the product instead of keeping a local file (the reference's next step is an S3 upload). */
int stc_write_geotiff_u8(const char* path, const uint8_t* img, int rows, int cols, double west, double south, double east,
                         double north);
int stc_geotiff_encode_u8(const uint8_t* img, int rows, int cols, double west, double south, double east, double north,
                          uint8_t** out_buf, int64_t* out_len);
void stc_geotiff_free(uint8_t* buf);

/* ---- reading a tile product back: load_tif (src/resegment_tiles_wide.py:713-751, `rasterio.open(tif).read(1)` of a _FINAL /
 * _POST / _SMOOTH* product before the border pass) ----
 * Classic TIFF (either byte order), 8 bits per sample, strips or tiles, chunky or planar samples, compression none / LZW /
 * PackBits, predictor 1 or 2: what this library, GDAL (compress='lzw'), libtiff and OpenCV write for a uint8 raster.
 * band: 1-based like rasterio.  *out_img: malloc'ed row-major [rows][cols] (release with stc_geotiff_free).  bounds4 (may be
 * NULL): west, south, east, north from ModelPixelScale + ModelTiepoint, NaN when the file carries none.  Host-only.
 * STC_ERR_ARG: null / band out of range; STC_ERR_STATE: unreadable, unsupported (BigTIFF, Deflate, ZSTD, other sample widths)
 * or corrupt file; STC_ERR_NOMEM. */
int stc_read_geotiff_u8(const char* path, int band, uint8_t** out_img, int* rows, int* cols, double* bounds4);
int stc_geotiff_decode_u8(const uint8_t* file, int64_t len, int band, uint8_t** out_img, int* rows, int* cols, double* bounds4);

#ifdef __cplusplus
}
#endif
#endif /* STC_H_ */
