"""Border re-segmentation pass: the second production caller of the hot-path kernels
(/root/reference/src/resegment_tiles_wide.py).  Two neighbouring tiles are re-processed TOGETHER
along their shared edge so that the tree-cover map has no seam: joint cloud removal, regridding + smoothing, an optional
histogram alignment of the two halves, re-prediction of the border subtiles and a check of the seam statistics.

Mirrored here, same names / arguments / return values, arrays in and out, every array operation on the GPU through libstc:
    align_dates                 :242-263   host integers (which dates the two tiles share within a day)
    make_tiles_right_neighb     :267-281   host integers (the border window table)
    check_if_artifact           :675-712   host scalars on two 1-D edge profiles (is there a visible seam?)
    preprocess_tile             :619-672   missing-px screening -> cloud / shadow masks -> feather -> cloud removal
    regularize_and_smooth       :772-791   dates -> 24 steps -> Whittaker -> 12 months, all bands
    align_subtile_histograms    :284-345   stc_align_histograms_host
    adjust_predictions          :348-357
    balance_seam_predictions    :541-553   the left/right mean correction applied to a border prediction
NOT mirrored (said plainly in DESIGN.md): the S3 / GeoTIFF / .hkl plumbing of resegment_border (:846-1166), and the wide
border model itself -- the reference predicts the seam with an UNRELEASED 220 x 684 graph (`retrain-combined-ca-220-684`,
:1604) while the released graphs and this library's forward are square; recreate_resegmented_tifs (:1240) builds on it."""
import numpy as np

from . import api as _api
from . import regrid as _regrid


def align_dates(tile_date, neighb_date):
    """:242-263.  A date survives when the other tile has an image within one day; a date equal to its predecessor
    (np.diff == 0, with a zero prepended) is dropped as a duplicate.  Returns (to_rm_tile, to_rm_neighb, min_images_left)."""
    t = np.asarray(tile_date)
    nb = np.asarray(neighb_date)

    def far(a, b):
        return [i for i, d in enumerate(a) if np.min(np.abs(d - b)) > 1]

    def dup(a):
        return list(np.flatnonzero(np.diff(a, prepend=0) == 0))
    rm_t = far(t, nb) + dup(t)
    rm_n = far(nb, t) + dup(nb)
    return rm_t, rm_n, int(min(len(t) - len(rm_t), len(nb) - len(rm_n)))


def make_tiles_right_neighb(tiles_folder_x, tiles_folder_y, size, size_y):
    """:267-281: window table of the border strip (one column of `size`-wide windows stepping down the seam).  Returns
    (tiles_array, tiles_folder) int arrays [n, 4] = (x, y, width, height) exactly as the reference builds them, including its
    column-wise sort and the re-tiling of the y column."""
    fx, fy = np.asarray(tiles_folder_x), np.asarray(tiles_folder_y)
    pairs = np.stack([np.repeat(fx, len(fy)), np.tile(fy, len(fx))], 1)          # cartesian(tiles_folder_x, tiles_folder_y)
    folder = np.sort(np.hstack([pairs, np.full_like(pairs, size + 7)]), axis=0)  # np.sort(axis=0): every column on its own
    uy = np.unique(folder[:, 1])
    folder[:, 1] = np.tile(uy, len(folder) // len(uy))
    arr = folder.copy()
    arr[1:, 1] -= 7
    arr[:, 0] = 0
    arr[:, 2] = size + 14
    arr[:, 3] = size_y + 7
    arr[1:-1, 3] += 7
    return arr, folder


def check_if_artifact(tile, neighb):
    """:675-712 on the two uint8 tree-cover rasters (NaN where > 100): compares the last column of `tile` with the first column
    of `neighb` in 10-row bins.  Returns 1 when the seam is visible.  (The reference prints a module-global x, y here.)"""
    def bins(col):
        col = np.pad(np.asarray(col, np.float64), (10 - (col.shape[0] % 10)) // 2, constant_values=np.nan)
        with np.errstate(all="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                return np.nanmean(np.reshape(col, (col.shape[0] // 10, 10)), axis=1)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        right_mean, left_mean = np.nanmean(neighb[:, :3]), np.nanmean(tile[:, -3:])
        right, left = bins(neighb[:, 0]), bins(tile[:, -1])
        d = np.abs(right - left)
        frac20, frac125 = np.nanmean(d > 20), np.nanmean(d > 12.5)
        frac_l, frac_r = np.nanmean(d[:15] > 17.5), np.nanmean(d[-15:] > 17.5)
    jump = abs(right_mean - left_mean)
    a = jump > 6
    b = (frac125 > 0.5) and (jump > 1)
    c = ((frac20 > 0.3) or (frac_l > 0.5) or (frac_r > 0.5)) and (jump > 1)
    return 1 if (a or b or c) else 0


def preprocess_tile(arr, dates, interp, clm, fname, dem, bbx, sess, forest_mask=None, urban_mask=None):
    """:619-672, same arguments (+ sess): returns (arr, interp, dates).  `interp` in and `fname` are unused by the reference
    too.  Python's global `random` state is consumed by the cloud removal exactly as in the reference."""
    del interp, fname
    arr = np.ascontiguousarray(arr, np.float32)
    dates = np.asarray(dates)
    missing = _api.id_missing_px(arr, 20, sess)
    if len(missing) > 0:
        dates = np.delete(dates, missing)
        arr = np.delete(arr, missing, 0)
    cld, fcps = _api.identify_clouds_shadows(arr, dem, bbx, sess, forest_mask=forest_mask, urban_mask=urban_mask)
    if clm is not None:
        if len(missing) > 0:
            clm = np.delete(clm, missing, 0)
        if np.asarray(clm).shape == np.asarray(fcps).shape == np.asarray(cld).shape:       # the reference's try/except guards this
            clm[fcps] = 0.
            cld = np.maximum(clm, cld)
    interp = _api.id_areas_to_interp(arr, cld, cld, dates, fcps, sess)
    to_remove = np.argwhere(np.mean(interp == 1, axis=(1, 2)) > 0.95)
    if len(to_remove) > 0:
        cld = np.delete(cld, to_remove, axis=0)
        dates = np.delete(dates, to_remove)
        arr = np.ascontiguousarray(np.delete(arr, to_remove, axis=0))
        cld, fcps = _api.identify_clouds_shadows(arr, dem, bbx, sess, forest_mask=forest_mask, urban_mask=urban_mask)
    arr, interp2, _ = _api.remove_cloud_and_shadows(arr, cld, cld, dates, fcps, None, sess=sess)
    return arr, interp2, dates


def regularize_and_smooth(arr, dates, sess):
    """:772-791: every 2-band window goes through calculate_and_save_best_images + Smoother; the linear operator is the
    same for all of them, so the whole cube is one 12 x n temporal product (K1).  Returns [12, H, W, C] float32."""
    M, _ = _regrid.monthly_operator(np.asarray(dates))
    return sess.temporal_matmul(np.ascontiguousarray(arr, np.float32), M)


def align_subtile_histograms(array, sess, size):
    """:284-345 (the reference reads SIZE from a module global).  `array` [T, H, size + 14, C] float32 is transformed in
    place and returned, like the reference."""
    a = np.ascontiguousarray(array, np.float32)
    T, H, W, C = a.shape
    sess._check(sess.lib.stc_align_histograms_host(sess.h, _api._dptr(a), T, H, W, C, (size + 14) // 2, size // 2 + 7, None))
    if a is not array:
        array[...] = a
    return array


def adjust_predictions(preds, ref):
    """:348-357: match the mean / standard deviation of `preds` to `ref` (nan-aware), clip to [0, 1]."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        mult = np.nanstd(ref) / np.nanstd(preds)
        add = np.nanmean(ref) - np.nanmean(preds) * mult
    return np.clip(preds * mult + add, 0, 1)


def balance_seam_predictions(preds, size):
    """:538-553: when the two 4-column strips on either side of the seam differ by more than 0.15 in mean, the confident
    pixels (> 0.05) of each half are shifted by half the difference of the half means; returns the (clipped) array."""
    preds = np.array(preds, copy=True)
    lm = np.mean(preds[:, (size - 8) // 2: size // 2])
    rm = np.mean(preds[:, size // 2: (size + 8) // 2])
    if abs(lm - rm) > 0.15:
        left, right = preds[:, : size // 2], preds[:, size // 2:]
        shift = (np.mean(right[right > 0.05]) - np.mean(left[left > 0.05])) / 2
        left[left > 0.05] += shift
        right[right > 0.05] -= shift
        preds = np.clip(preds, 0, 1)
    return preds
