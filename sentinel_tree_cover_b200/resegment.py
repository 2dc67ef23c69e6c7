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
    assemble_border_subtile     :455-489   reflect padding of the edge windows + the 17-channel [LEN + 1, SIZE_Y + 14, SIZE + 14] stack
    predict_subtile             :182-222   clip / normalise / forward of one RECTANGULAR border window (220 x 684 -> 206 x 670)
    predict_border_subtile      :411-553   window cut-out -> NaN-date removal -> histogram alignment -> assemble -> predict -> balance
    process_subtiles            :360-616   the loop over the border window table (arrays in, {window: prediction} out)
    seam_prediction_accepted    :536-611   is the border prediction inside the range the two tiles' own maps allow?
    mosaic_subtiles             :1169-1237  weighted mean of one kind of subtile layers + the kind's blending weight (host)
    recreate_resegmented_tifs   :1240-1547  re-mosaic of a tile from its normal subtiles and border strips (host, .npy files)
NOT mirrored (said plainly in DESIGN.md): the S3 / GeoTIFF / .hkl plumbing of resegment_border (:846-1166).  The reference predicts the seam with an UNRELEASED weight set
(`retrain-combined-ca-220-684`, :1605); the network is fully convolutional, so the forward here takes the 220 x 684 window
with whatever weights the session holds (the released 172-px set in the tests, checked against the float32 restatement of
the same graph evaluated at the same rectangular size)."""
import os

import numpy as np

from . import api as _api
from . import regrid as _regrid

# :1664-1676 -- the 17 band ranges of the border model; identical to download_and_predict_job.py:1828-1844 except for the DEM
# maximum (0.509... here, 0.4 there).
MIN_ALL = list(_api.MIN_ALL)
MAX_ALL = list(_api.MAX_ALL)
MAX_ALL[10] = 0.509269855802243


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


def make_tiles_right_neighb(tiles_folder_x, tiles_folder_y, size, size_y, edge="right"):
    """:267-281: window table of the border strip (one column of `size`-wide windows stepping down the seam).  Returns
    (tiles_array, tiles_folder) int arrays [n, 4] = (x, y, width, height) exactly as the reference builds them, including its
    column-wise sort and the re-tiling of the y column.  edge="up": the table of the NORTH seam
    (src/resegment_tiles_north_wide.py:253-266; `size_y` is that file's SIZE_X): windows stepping along x, SIZE + 14 tall."""
    fx, fy = np.asarray(tiles_folder_x), np.asarray(tiles_folder_y)
    pairs = np.stack([np.repeat(fx, len(fy)), np.tile(fy, len(fx))], 1)          # cartesian(tiles_folder_x, tiles_folder_y)
    folder = np.sort(np.hstack([pairs, np.full_like(pairs, size + 7)]), axis=0)  # np.sort(axis=0): every column on its own
    uy = np.unique(folder[:, 1])
    folder[:, 1] = np.tile(uy, len(folder) // len(uy))
    arr = folder.copy()
    step, fixed = (1, 0) if edge == "right" else (0, 1)        # which column steps along the seam, which is pinned to 0
    arr[1:, step] -= 7
    arr[:, fixed] = 0
    arr[:, 2 + fixed] = size + 14
    arr[:, 2 + step] = size_y + 7
    arr[1:-1, 2 + step] += 7
    return arr, folder


def check_if_artifact(tile, neighb, edge="right"):
    """:675-712 on the two uint8 tree-cover rasters (NaN where > 100): compares the last column of `tile` with the first column
    of `neighb` in 10-row bins.  Returns 1 when the seam is visible.  (The reference prints a module-global x, y here.)
    edge="up": the north seam (src/resegment_tiles_north_wide.py:661-700): first row of `tile` against the last row of
    `neighb`, both cut to the shorter one, and that file's thresholds."""
    import warnings

    def bins(col):
        col = np.pad(np.asarray(col, np.float64), (10 - (col.shape[0] % 10)) // 2, constant_values=np.nan)
        return np.nanmean(np.reshape(col, (col.shape[0] // 10, 10)), axis=1)
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore", RuntimeWarning)
        if edge == "right":
            right_mean, left_mean = np.nanmean(neighb[:, :3]), np.nanmean(tile[:, -3:])
            right, left = bins(neighb[:, 0]), bins(tile[:, -1])
            t_all, t_half, t_end, jump_a, jump_bc, f_half, f_all = 20, 12.5, 17.5, 6, 1, 0.5, 0.3
        else:
            right_mean, left_mean = np.nanmean(neighb[-3:]), np.nanmean(tile[:3])
            right, left = neighb[-1], tile[0]
            right = right[:left.shape[0]]
            left = left[:right.shape[0]]
            right, left = bins(right), bins(left)
            t_all, t_half, t_end, jump_a, jump_bc, f_half, f_all = 25, 15, 15, 7, 0, 0.6, 0.25
        d = np.abs(right - left)
        frac_all, frac_half = np.nanmean(d > t_all), np.nanmean(d > t_half)
        frac_l, frac_r = np.nanmean(d[:15] > t_end), np.nanmean(d[-15:] > t_end)
    jump = abs(right_mean - left_mean)
    a = jump > jump_a
    b = (frac_half > f_half) and (jump > jump_bc)
    c = (frac_all > f_all) or (frac_l > 0.5) or (frac_r > 0.5)
    if edge == "right":
        c = c and (jump > jump_bc)
    return 1 if (a or b or c) else 0


def preprocess_tile(arr, dates, interp, clm, fname, dem, bbx, sess, forest_mask=None, urban_mask=None, edge="right"):
    """:619-672, same arguments (+ sess): returns (arr, interp, dates).  `interp` in and `fname` are unused by the reference
    too.  Python's global `random` state is consumed by the cloud removal exactly as in the reference.
    edge="up": the north-seam file's version (src/resegment_tiles_north_wide.py:573-625): a date is dropped when a FIFTH of its
    pixels is missing (a twentieth here), and the Sen2Cor mask loses its false-positive pixels only when a date was dropped."""
    del interp, fname
    arr = np.ascontiguousarray(arr, np.float32)
    dates = np.asarray(dates)
    missing = _api.id_missing_px(arr, 20 if edge == "right" else 5, sess)
    if len(missing) > 0:
        dates = np.delete(dates, missing)
        arr = np.delete(arr, missing, 0)
    cld, fcps = _api.identify_clouds_shadows(arr, dem, bbx, sess, forest_mask=forest_mask, urban_mask=urban_mask)
    if clm is not None:
        if len(missing) > 0:
            clm = np.delete(clm, missing, 0)
        if np.asarray(clm).shape == np.asarray(fcps).shape == np.asarray(cld).shape:       # the reference's try/except guards this
            if edge == "right" or len(missing) > 0:
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


def assemble_border_subtile(subtile_s2, s1_subtile, dem_subtile, median_s2, median_s1, start_y, size, size_y, length=4):
    """:455-489.  A window cut at the tile edge is 7 pixels short on one side (make_tiles_right_neighb gives every window
    SIZE + 14 columns and SIZE_Y + 7 rows at the two ends of the seam); the reference mirrors the missing rows in
    (np.pad 'reflect': below the first window, above the others -- and the same for a short column axis) and stacks
    [10 S2 bands | DEM | 2 S1 bands | 4 indices] for the `length` quarterly frames plus the median frame.
    subtile_s2 [length, h, w, 14], s1_subtile [length, h, w, 2], dem_subtile [1, h, w], median_s2 [1, h, w, 14],
    median_s1 [1, h, w, 2]  ->  float32 [length + 1, size_y + 14, size + 14, 17]."""
    first = (start_y == 0)

    def grow(a, axis):
        pads = [(0, 0)] * a.ndim
        pads[axis] = ((7, 0) if not first else (0, 7)) if axis == 2 else ((7, 0) if first else (0, 7))
        return np.pad(a, pads, 'reflect')
    arrs = [subtile_s2, s1_subtile, dem_subtile, median_s2, median_s1]
    if subtile_s2.shape[2] == size + 7:                 # :455-463 (pad_u / pad_d are applied to axis 2 in the reference)
        arrs = [grow(a, 2) for a in arrs]
    if arrs[0].shape[1] == size_y + 7:                  # :466-474
        arrs = [grow(a, 1) for a in arrs]
    s2, s1, dem, ms2, ms1 = arrs
    out = np.empty((length + 1, size_y + 14, size + 14, 17), np.float32)
    out[:-1, ..., :10] = s2[..., :10]
    out[:-1, ..., 11:13] = s1
    out[:-1, ..., 13:] = s2[..., 10:]
    out[:, ..., 10] = dem.repeat(length + 1, axis=0)
    out[-1, ..., :10] = ms2[0, ..., :10]
    out[-1, ..., 11:13] = ms1[0]
    out[-1, ..., 13:] = ms2[0, ..., 10:]
    return out


def predict_subtile(subtile, sess, size=None):
    """:182-222.  subtile [LEN + 1, SIZE_Y + 14, SIZE + 14, 17] (quarterly frames + median frame; uint16 storage is
    rescaled by 1 / 65535 first) is clipped to the 17 band ranges, normalised to [-1, 1] and run through the ConvGRU / U-Net
    forward at ITS OWN height and width: one rectangular launch sequence, no tiling into squares, so the seam the pass
    exists to remove is not re-introduced inside the window.  Returns float32 [SIZE_Y, SIZE].  An all-zero window gives the
    reference's integer 255 fill, which is SIZE x SIZE there (:220) -- `size` defaults to the window's own width - 14."""
    subtile = np.asarray(subtile)
    if np.sum(subtile) != 0:
        if not isinstance(subtile.flat[0], np.floating):
            assert np.max(subtile) > 1
            subtile = sess.to_float32(np.ascontiguousarray(subtile).astype(np.uint16, copy=False))
        batch_x = np.ascontiguousarray(subtile[np.newaxis], np.float32)
        return np.float32(sess.predict(batch_x, length=subtile.shape[0] - 1, normalize=True, mins=MIN_ALL, maxs=MAX_ALL).squeeze())
    size = subtile.shape[2] - 14 if size is None else size
    return np.full((size, size), 255)


def predict_border_subtile(s2, s1, s2_median, s1_median, dem, interp, dates, window, sess, size, size_y, hist_align=True,
                           forward=None):
    """:411-553 for one row of the border window table (make_tiles_right_neighb): cut the window out of the joint
    two-tile arrays (s2 [LEN, H, W, 14] quarterly composites with indices, s1 [LEN, H, W, 2], medians [1, H, W, *],
    dem [H, W], interp [n, H, W], dates [n]), drop NaN dates from the bookkeeping, align the two halves' histograms,
    assemble, predict, balance the two sides of the seam.  `forward` (tests) replaces predict_subtile(x, sess, size).
    Returns (preds [size_y, size] float32 -- or the integer 255 fill when fewer than two dates are left, :508-510 --,
    dates_left)."""
    start_x, start_y, w, h = (int(v) for v in window[:4])
    ys, xs = slice(start_y, start_y + h), slice(start_x, start_x + w)
    subset = np.copy(s2[:, ys, xs, :])
    med_s2 = np.copy(s2_median[:, ys, xs, :])
    dates_tile = np.copy(dates)
    nan_dates = np.argwhere(np.sum(np.isnan(subset), axis=(1, 2, 3)) > 0).flatten()
    if len(nan_dates) > 0:                              # :439-444
        dates_tile = np.delete(dates_tile, nan_dates)
        subset = np.delete(subset, nan_dates, 0)
    if hist_align:                                      # in place, so the stack below is the aligned one (:446-451)
        subset = align_subtile_histograms(subset, sess, size)
        med_s2 = align_subtile_histograms(med_s2, sess, size)
    x = assemble_border_subtile(subset, s1[:, ys, xs, :], dem[np.newaxis, ys, xs], med_s2, s1_median[:, ys, xs, :],
                                start_y, size, size_y, length=s2.shape[0])
    if len(dates_tile) < 2:
        return np.full((size_y, size), 255), dates_tile
    preds = forward(x) if forward is not None else predict_subtile(x, sess, size)
    return balance_seam_predictions(preds, size), dates_tile


def process_subtiles(s2, dates, interp, s1, dem, sess, tiles_folder, tiles_array, right_all, left_all, size, size_y,
                     hist_align=True, length=4, forward=None):
    """:360-616 without the file system: s2 [12, H, size + 14, 14] monthly composites with indices of the JOINT strip (the
    tile's last and the neighbour's first (size + 14) // 2 columns), s1 [12, H, size + 14, 2], dem [H, size + 14],
    interp [n, H, size + 14], dates [n]; right_all / left_all = the two tiles' own tree-cover maps next to the seam
    (percent, [H, size // 2]).  NaN -> 0 (interpolate_na_vals), annual and quarterly medians on the GPU, then every window
    of the table goes through predict_border_subtile; a prediction the acceptance rule rejects is left out, exactly like
    the reference does not write its file.  Returns {(folder_y, folder_x): preds} -- the arrays the reference np.save's to
    `processed/right<folder_y>/<folder_x>.npy` and to the neighbour's `processed/0/left<folder_x>.npy`."""
    from . import tile as _tile
    if sess is None:
        raise RuntimeError("process_subtiles needs an StcSession (sess=...); there is no CPU path")
    s2 = _tile.nan_to_zero(s2, sess)
    s1 = np.ascontiguousarray(s1, np.float32)
    s2_median = sess.temporal_median(s2)[np.newaxis]
    s1_median = sess.temporal_median(s1)[np.newaxis]
    if length == 4:                                     # :402-407: medians of the four 3-month groups
        s2 = np.stack([sess.temporal_median(s2[3 * q: 3 * q + 3]) for q in range(4)])
        s1 = np.stack([sess.temporal_median(s1[3 * q: 3 * q + 3]) for q in range(4)])
    out = {}
    for folder, window in zip(np.asarray(tiles_folder), np.asarray(tiles_array)):
        preds, _ = predict_border_subtile(s2, s1, s2_median, s1_median, dem, interp, dates, window, sess, size, size_y,
                                          hist_align=hist_align, forward=forward)
        start_y = int(window[1])
        if seam_prediction_accepted(preds, left_all[start_y:start_y + size_y], right_all[start_y:start_y + size_y], size):
            out[(int(folder[0]), int(folder[1]))] = preds
    return out


def seam_prediction_accepted(preds, left_rows, right_rows, size):
    """:536-611: the border prediction is written out only if its mean tree cover (percent) lies within 15 points of the
    range spanned by the two tiles' own maps next to the seam (left_rows = the tile's last SIZE // 2 columns over the
    window's rows, right_rows = the neighbour's first SIZE // 2; percent, NaN = no data), or if neither map has data.
    (The three in-range / out-of-range branches of the reference all save the same array; only the last `else` skips.)"""
    import warnings
    if np.max(preds) >= 255:
        return True
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        lo_ref = np.nanmean(left_rows[:, :100])
        hi_ref = np.nanmean(right_rows[:, -100:])
        src = 100 * np.nanmean(preds)
    lo, hi = np.minimum(lo_ref, hi_ref), np.maximum(lo_ref, hi_ref)
    if src <= lo - 15 or src >= hi + 15 or (lo - 15 <= src <= hi + 15):
        return True
    return bool(np.isnan(lo) and np.isnan(hi))



# ---- re-mosaic of a tile's subtile predictions after the border pass (:1164-1547) -----------------------------------------
# Host logic over small arrays (one 618 x 618 tile, a few dozen layers): file listing, float weights, one weighted mean.
# It runs once per tile pair after the GPU work, like the window tables above.

def resize_linear(img, shape, anti_aliasing=None):
    """`skimage.transform.resize(img, shape, order=1)` as scikit-image >= 0.19 computes it (the package is not a dependency
    here; this restates its published algorithm over SciPy, which is what scikit-image itself calls): when an axis shrinks, a
    Gaussian of sigma (in / out - 1) / 2 along that axis first (`anti_aliasing`, on by default for non-boolean input), then
    `scipy.ndimage.zoom(order=1, mode='mirror', grid_mode=True)` ('mirror' is ndimage's name for skimage's 'reflect')."""
    import scipy.ndimage as ndi
    img = np.asarray(img)
    if img.dtype not in (np.float32, np.float64):
        img = img.astype(np.float64)
    factors = np.divide(img.shape, shape)
    if anti_aliasing is None:
        anti_aliasing = bool(np.any(factors > 1))
    if anti_aliasing:
        img = ndi.gaussian_filter(img, np.maximum(0, (factors - 1) / 2), cval=0, mode="mirror")
    return ndi.zoom(img, [o / float(i) for o, i in zip(shape, img.shape)], order=1, mode="mirror", cval=0, grid_mode=True)


def adjust_resegment(res, mults, n):
    """:1164-1166."""
    return res * np.maximum(np.sum(mults[..., :n], axis=-1), 1.)


# which edge flags feather which end of the ramp, and where the zero half goes, per border kind (:1192-1235)
#            flip the ramp, flag for columns [:300], flag for columns [-300:], zeros first, transpose
_RAMP_PLAN = {"r": (False, "up", "down", True, None),
              "l": (True, "up", "down", False, None),
              "u": (True, "left", "right", False, "T"),
              "d": (False, "right", "left", True, "T-flip")}


def mosaic_subtiles(preds, mults, na, kind, left, right, up, down, size=670, resize=None, feather_power=1.33):
    """:1169-1237, same arguments (+ `size` = the reference's global SIZE, `resize` = the bilinear resize to use, default
    `resize_linear`).  preds / mults [X, Y, n] float32 layers (NaN / 0 where a layer has no data; both are modified in place
    like the reference does), na [X, Y, 1].  Returns (weighted mean of the layers [X, Y], blending weight of this kind [X, Y]):
    a Gaussian for the normal subtiles, for a border kind a ramp (x / half) ** 1.2 rising towards the shared edge over
    `size // 2` pixels, feathered by (t / 300) ** 1.33 over 300 pixels at the ends where another border strip exists
    (`feather_power` = 1.5 in the north-seam file, src/resegment_tiles_north_wide.py:1155, otherwise identical)."""
    resize = resize or resize_linear
    preds[np.tile(na, (1, 1, preds.shape[-1])) > 0] = np.nan
    mults[np.isnan(preds)] = 0.
    with np.errstate(invalid="ignore", divide="ignore"):
        mults = mults / np.sum(mults, axis=-1)[..., np.newaxis]
    preds = np.nansum(preds * mults, axis=-1)
    X, Y = preds.shape
    half = size // 2
    if kind == "n":
        m = resize(_api.fspecial_gauss(X, X / 5.25), (X, Y))
    else:
        flip, flag_lo, flag_hi, zeros_first, post = _RAMP_PLAN[kind]
        present = {"left": left is not None, "right": right is not None, "up": up is not None, "down": down is not None}
        feather = np.tile((np.arange(0, 300, 1) / 300) ** feather_power, (half, 1))
        m = (np.ones((half, Y)) * (np.arange(0, half, 1) / half)[:, np.newaxis]) ** 1.2
        if flip:
            m = np.flipud(m)
        m = np.copy(m)
        if present[flag_lo]:
            m[:, :300] *= feather
        if present[flag_hi]:
            m[:, -300:] *= np.fliplr(feather)
        m = resize(m, (half, Y))
        zeros = np.zeros((X - half, Y))
        m = np.concatenate([zeros, m] if zeros_first else [m, zeros], axis=0)
        if post == "T":
            m = m.T
        elif post == "T-flip":
            m = np.flipud(m.T)
        m = resize(m, (X, Y))
    m[np.isnan(preds)] = 0.
    return preds, m


def _gauss_sigma(subtile_size, border):
    """:1301-1311 (normal subtiles), :1343-1352 (border strips): the Gaussian width for a subtile size."""
    table = {208: 44, 216: 44, 348: 85, 412: 95}
    if subtile_size in table:
        return table[subtile_size]
    if border:
        return 150 if (subtile_size == 588 or subtile_size >= 620) else 28
    return 38 if subtile_size == 168 else 28


# per border kind: file-name prefix, which axis of the TRANSPOSED prediction is halved, which half is kept
_STRIP_PLAN = {"l": ("left", 0, 1), "r": ("", 0, 0), "u": ("up", 1, 1), "d": ("down", 1, 0)}


def _list_subtile_files(out_folder):
    """The reference's os.listdir walk (:1254-1271): ({kind: [(x, y, path)]} in listing order, number of .npy files)."""
    entries = [e for e in os.listdir(out_folder) if ".DS" not in e]
    right_dirs = [e for e in entries if "right" in e]
    x_dirs = [e for e in entries if "right" not in e and len(os.listdir(os.path.join(out_folder, e))) > 0]
    found = {k: [] for k in "nlrud"}
    for xd in x_dirs:
        for f in os.listdir(os.path.join(out_folder, xd)):
            if ".DS" in f:
                continue
            kind = "l" if "left" in f else "d" if "down" in f else "u" if "up" in f else "n"
            y = int(f[len(_STRIP_PLAN[kind][0]) if kind != "n" else 0:-4])
            found[kind].append((int(xd), y, os.path.join(out_folder, xd, f)))
    for xd in right_dirs:
        for f in os.listdir(os.path.join(out_folder, xd)):
            if ".DS" not in f:
                found["r"].append((int(xd[5:]), int(f[:-4]), os.path.join(out_folder, xd, f)))
    n_files = sum(len(v) for v in found.values())
    return found, n_files


def recreate_resegmented_tifs(out_folder, shape, size=670, resize=None, edge="right"):
    """:1240-1547, same arguments (+ `size` = the reference's global SIZE, + `resize`) and return value `(preds, sums)`.
    edge="up": the north-seam file's version (src/resegment_tiles_north_wide.py: the same function, its mosaic_subtiles
    feathers with exponent 1.5 instead of 1.33).
    Reads `<out_folder>/<x>/<y>.npy` (normal subtiles), `<x>/left<y>.npy`, `<x>/up<y>.npy`, `<x>/down<y>.npy` and
    `right<x>/<y>.npy` (border strips written by the border pass of this tile and of its neighbours; a strip file holds the
    window across the seam, of which this tile owns one half), builds one layer stack per kind, takes the weighted mean inside
    each kind (`mosaic_subtiles`) and blends the five with the kind weights.  preds: float64 [shape[1], shape[0]], tree cover
    in percent, 255 = no data (no normal prediction, or a border strip without data where strips exist)."""
    resize = resize or resize_linear
    X, Y = int(shape[1]), int(shape[0])
    found, n_files = _list_subtile_files(out_folder)
    n_border = sum(len(found[k]) for k in "lrud")
    stacks = {}
    covered = {"n": np.zeros((X, Y)), "b": np.zeros((X, Y))}                # how many subtiles / strips cover a pixel
    nodata = {"n": np.zeros((X, Y)), "b": np.zeros((X, Y))}                 # ... of which without data there
    for kind in "nlrud":
        files = found[kind]
        if kind != "n" and not files:
            continue
        n_layers = (n_files - n_border) if kind == "n" else len(files)
        P = np.full((X, Y, n_layers), np.nan, dtype=np.float32)
        M = np.full((X, Y, n_layers), 0, dtype=np.float32)
        i = 0
        for (x0, y0, path) in files:
            raw = np.load(path)
            if kind == "n":
                sx, sy = raw.shape[1], raw.shape[0]
                sub = max(sx, sy)
                has_data = np.sum(raw) < sx * sy * 255
                if not has_data:
                    continue
                p = (raw * 100).T.astype(np.float32)
                if (x0 + sx - 1) < X and (y0 + sy - 1) < Y:
                    w = _api.fspecial_gauss(sub, _gauss_sigma(sub, False))
                    w[p > 100] = 0.
                    P[x0:x0 + sx, y0:y0 + sy, i] = p
                    M[x0:x0 + sx, y0:y0 + sy, i] = w
                    group = "n"
                else:
                    i += 1
                    continue
            else:
                _, axis, keep = _STRIP_PLAN[kind]
                sx = raw.shape[1] // 2 if axis == 0 else raw.shape[1]
                sy = raw.shape[0] // 2 if axis == 1 else raw.shape[0]
                sub = max(2 * sx if axis == 0 else sx, 2 * sy if axis == 1 else sy)
                cut = sx if axis == 0 else sy
                half_of = (lambda a: (a[cut:] if keep else a[:cut]) if axis == 0 else (a[:, cut:] if keep else a[:, :cut]))
                w = resize(half_of(_api.fspecial_gauss(sub, _gauss_sigma(sub, True))), (sx, sy))
                has_data = np.sum(raw) < sx * sy * 255
                if not has_data:
                    i += 1
                    continue
                p = half_of((raw * 100).T.astype(np.float32))
                P[x0:x0 + sx, y0:y0 + sy, i] = p
                P[x0:x0 + sx, y0:y0 + sy, :][p > 100] = 255.          # :1361: every layer of the kind, at the strip's no-data px
                w[p > 100] = 0.
                M[x0:x0 + sx, y0:y0 + sy, i] = w
                group = "b"
            covered[group][x0:x0 + sx, y0:y0 + sy] += 1.
            nodata[group][x0:x0 + sx, y0:y0 + sy] += (p > 100)
            i += 1
        stacks[kind] = (P, M)

    # :1494-1501: no data where no normal subtile has data; under border strips also where any strip lacks data
    normal_valid = covered["n"] - nodata["n"]
    na = np.zeros((X, Y))
    na[(covered["b"] == 0) & (normal_valid == 0)] = 1.
    na[(covered["b"] > 0) & ((normal_valid == 0) | (nodata["b"] > 0))] = 1.
    na = na[..., np.newaxis]
    layers = {k: (stacks[k][0] if k in stacks else None) for k in "lrud"}
    out = {}
    for kind in "nrlud":
        if kind in stacks:
            P, M = stacks[kind]
            pk, mk = mosaic_subtiles(P, M, na, kind, layers["l"], layers["r"], layers["u"], layers["d"], size=size, resize=resize,
                                     feather_power=1.33 if edge == "right" else 1.5)
            if kind != "n":
                mk[np.sum(~np.isnan(P), axis=-1) == 0] = 0.
            out[kind] = (pk, mk)
        else:
            out[kind] = (np.zeros_like(out["n"][0]), np.zeros_like(out["n"][0]))
    (pn, mn), (pl, ml), (pr, mr), (pu, mu), (pd_, md) = (out[k] for k in "nlrud")
    with np.errstate(invalid="ignore", divide="ignore"):
        sums = (ml + mr + mu + md + mn)
        preds = (pl * (ml / sums)) + (pd_ * (md / sums))
        preds = preds + (pr * (mr / sums)) + (pu * (mu / sums))
        preds = preds + (pn * (mn / sums))
    preds[np.isnan(preds)] = 255.
    preds[na.squeeze() == 1.] = 255.
    return preds, sums


def seam_smooth_diff(predictions_left, predictions_right):
    """:1763-1770: mean absolute difference (percent) between the tile's last 8 rows and the neighbour's first 8 rows of the
    re-mosaicked maps, each averaged over the 8 rows first; 255 = no data is ignored.  NaN if no column has data on both sides."""
    import warnings
    right = np.array(predictions_right[:8, :], np.float32)
    left = np.array(predictions_left[-8:, :], np.float32)
    right[right == 255] = np.nan
    left[left == 255] = np.nan
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        return np.nanmean(abs(np.nanmean(right, axis=0) - np.nanmean(left, axis=0)))


def write_smoothed_pair(predictions_left, predictions_right, bbx, neighb_bbx, x, y, path_to_tile, path_to_right, smooth_diff, diff):
    """:1796-1816: if the seam got no worse than 20 points over the difference before the border pass (`diff`, NaN -> 100,
    :1771), write both re-mosaicked tiles as `<x>X<y>Y_SMOOTH_X.tif` (`_SMOOTH_XY` when a Y pass has already written one)
    through the libstc GeoTIFF writer.  Returns the two file names, or None when the pair is rejected.  Uploading and the
    cleanup of the working folders (:1806,1818) are the caller's."""
    diff = 100 if np.isnan(diff) else diff
    if not (smooth_diff < (diff + 20) or np.isnan(smooth_diff)):
        return None
    files = []
    for preds, box, tx, folder in ((predictions_left, bbx, x, path_to_tile), (predictions_right, neighb_bbx, str(int(x) + 1), path_to_right)):
        stem = "%s/%sX%sY" % (folder, str(tx), str(y))
        redo = os.path.exists(stem + "_SMOOTH_XY.tif") or os.path.exists(stem + "_SMOOTH_Y.tif")
        files.append(_api.write_tif(preds, box, tx, y, folder, "_SMOOTH_XY" if redo else "_SMOOTH_X"))
    return files


def load_tif(tile_id, local_path, edge="right"):
    """:713-751, same arguments and return value `(raster, is_smooth_y)`: the tile's current tree-cover product, picked in the
    reference's order of preference -- a `_SMOOTH_XY` product, else `_SMOOTH_X`, else `_SMOOTH_Y` (any `_SMOOTH*`: is_smooth_y
    = 1 for XY / Y), else `_FINAL`, else `_POST` -- and read as band 1, uint8 [rows, cols], by libstc's TIFF reader
    (api.read_tif; the reference uses rasterio).  Raises IndexError like the reference when the folder holds no product.
    edge="up" (src/resegment_tiles_north_wide.py:703-742): same choice, but the flag says whether ANY file with "SMOOTH" in
    its name exists in the folder."""
    dir_i = f"{local_path}/{tile_id[0]}/{tile_id[1]}/"
    chosen, is_smooth_y = [], 0
    if os.path.exists(dir_i):
        files = [f for f in os.listdir(dir_i) if os.path.splitext(f)[-1] == '.tif']
        by_kind = {k: [f for f in files if k in f] for k in ("_SMOOTH_XY", "_SMOOTH_X", "_SMOOTH_Y", "_SMOOTH", "_FINAL", "_POST")}
        if by_kind["_SMOOTH"]:
            # ("_SMOOTH_X" is a substring of "_SMOOTH_XY": the XY test comes first, as in the reference)
            for kind, flag in (("_SMOOTH_XY", 1), ("_SMOOTH_X", 0), ("_SMOOTH_Y", 1)):
                if by_kind[kind]:
                    chosen, is_smooth_y = by_kind[kind], flag
                    break
            else:
                chosen = files                              # a `_SMOOTH` name of no known kind: the reference keeps every .tif
        elif by_kind["_FINAL"]:
            chosen = by_kind["_FINAL"]
        else:
            chosen = by_kind["_POST"]
    if edge != "right":
        is_smooth_y = 1 if (os.path.exists(dir_i) and any("SMOOTH" in f for f in os.listdir(dir_i))) else 0
    return _api.read_tif(os.path.join(dir_i, chosen[0])), is_smooth_y


def concatenate_s2_files(s2, s2_neighb):
    """:754-763: centre-crop the wider of the two tiles' cubes along axis 2 so that both have the neighbour's / tile's width."""
    s2_diff = s2.shape[2] - s2_neighb.shape[2]
    if s2_diff > 0:
        s2 = s2[:, :, s2_diff // 2: -(s2_diff // 2), :]
    if s2_diff < 0:
        s2_neighb = s2_neighb[:, :, - (s2_diff // 2): (s2_diff // 2)]
    return s2, s2_neighb
