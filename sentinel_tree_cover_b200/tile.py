"""Host-side mirror of `process_tile` (raw tile arrays -> analysis-ready cube),
/root/reference/src/download_and_predict_job.py:640-997, and of `adjust_shape` (:260-310).

Everything that is arithmetic on the tile arrays is a GPU call through StcSession (codecs, dB, DEM
median filter, 20 m -> 10 m upsampling, missing-pixel / snow counts, cloud masks, feathering, cloud
removal, clip).  What stays here is what the reference also does with scalars and indices: file naming,
np.delete of dropped dates, np.pad / slicing in adjust_shape (data movement), the retry loops on
per-date fractions.  File I/O is out of scope: `loader(path)` stands in for `hkl.load` and `exists(path)`
for `os.path.exists` (defaults: hickle if importable / os.path.exists).
There is no CPU path: every function needs an StcSession."""
import os
import numpy as np

from . import api as _api


def _check(sess, rc):
    sess._check(rc)


def _ptr(a):
    return a.ctypes.data_as(_api.C.c_void_p)


# ---- thin session helpers for the tile-prep entry points (include/stc.h) ----
def s1_fill(s1, sess):
    a = np.array(s1, dtype=np.float32, order="C")      # always a copy: the caller's array is left alone
    m, H, W, Cc = a.shape
    _check(sess, sess.lib.stc_s1_fill_host(sess.h, _ptr(a), m, H, W, Cc))
    return a


def median_filter5(dem, sess):
    a = np.ascontiguousarray(dem, np.float32)
    out = np.empty_like(a)
    _check(sess, sess.lib.stc_median_filter5_host(sess.h, _ptr(a), a.shape[0], a.shape[1], _ptr(out)))
    return out


def clm_pairs(clm, sess):
    a = np.array(clm, dtype=np.float32, order="C")      # always a copy: the caller's array is left alone
    _check(sess, sess.lib.stc_clm_pairs_host(sess.h, _ptr(a), a.shape[0], a.shape[1], a.shape[2]))
    return a


def snow_mask(sentinel2, sess):
    """-> (flagged pixels per date int32[n], snow int64 [H,W] = 1 - binary_dilation(mean_t < 0.7, 2))."""
    a = np.ascontiguousarray(sentinel2[..., :10], np.float32)
    n, H, W, _ = a.shape
    per_date = np.zeros(n, np.int32)
    snow = np.empty((H, W), np.uint8)
    _check(sess, sess.lib.stc_snow_host(sess.h, _ptr(a), n, H, W, _ptr(per_date), _ptr(snow)))
    return per_date, snow.astype(np.int64)


def count_gt(data, thresh, sess):
    a = np.ascontiguousarray(data, np.float32)
    n = a.shape[0]
    a2 = a.reshape(n, -1)
    out = np.zeros(n, np.int32)
    _check(sess, sess.lib.stc_count_gt_host(sess.h, _ptr(a2), n, a2.shape[1], float(thresh), _ptr(out)))
    return out


def clip01(x, sess):
    a = np.array(x, dtype=np.float32, order="C")      # always a copy: the caller's array is left alone
    _check(sess, sess.lib.stc_elementwise_host(sess.h, _ptr(a), a.size, 0, 0.0, 1.0))
    return a


def divide(x, d, sess):
    a = np.array(x, dtype=np.float32, order="C")      # always a copy: the caller's array is left alone
    _check(sess, sess.lib.stc_elementwise_host(sess.h, _ptr(a), a.size, 1, float(d), 0.0))
    return a


def max_masked(a, b, zero, sess):
    """np.maximum(a, b) after b[zero] = 0 (float32)."""
    a = np.ascontiguousarray(a, np.float32).copy()
    b = np.ascontiguousarray(b, np.float32)
    z = None if zero is None else np.ascontiguousarray(np.asarray(zero) != 0, np.uint8)
    _check(sess, sess.lib.stc_max_masked_host(sess.h, _ptr(a), _ptr(b), None if z is None else _ptr(z), a.size))
    return a


def _axis_plan(length, target):
    """One axis of adjust_shape (:260-310) as (shift, out_len): out[i] = in[clip(i + shift, 0, length - 1)].
    Padding is 'edge' padding, i.e. an index clamp; the reference's rules for how much it pads or crops (and the cases
    in which it leaves the axis one pixel off) are reproduced, see csrc/stc_tile.cu: tilehost::adjust_axis."""
    shift, out_len = _api.C.c_int32(0), _api.C.c_int32(0)
    rc = _api.load_library().stc_adjust_shape_plan(int(length), int(target), _api.C.byref(shift), _api.C.byref(out_len))
    if rc != 0:
        raise ValueError("adjust_shape: bad axis length %r / target %r" % (length, target))
    return shift.value, out_len.value


def adjust_shape(arr, width, height):
    """:260-310 -- pad ('edge') or centre-crop axes 1 / 2 to (width, height): a gather along each axis with clamped
    indices (the device chain runs the same plan as a kernel, stc_tile.cu: k_adjust)."""
    if arr.ndim == 3:
        arr = arr[:, :, :, np.newaxis]
    elif arr.ndim == 2:
        arr = arr[np.newaxis, :, :, np.newaxis]
    for axis, target in ((1, width), (2, height)):
        shift, out_len = _axis_plan(arr.shape[axis], target)
        if shift != 0 or out_len != arr.shape[axis]:
            idx = np.clip(np.arange(out_len) + shift, 0, arr.shape[axis] - 1)
            arr = np.take(arr, idx, axis=axis)
    return arr.squeeze()


def _default_loader(path):
    import hickle as hkl            # not shipped with this package; pass `loader=` otherwise
    return hkl.load(path)


def process_tile(x, y, data, local_path, bbx, make_shadow=False, sess=None, loader=None, exists=os.path.exists,
                 forest_mask=None, urban_mask=None):
    """:640-997, same arguments and return tuple
    `(sentinel2, image_dates, interp, s1, dem, cloudshad, snow)`."""
    if sess is None:
        raise RuntimeError("process_tile needs an StcSession (sess=...); there is no CPU path")
    load = loader or _default_loader
    import time as _time
    _t = [_time.perf_counter()]
    _trace = os.environ.get("STC_TILE_TIMING")

    def _mark(name):
        if _trace:
            _t.append(_time.perf_counter())
            print("[process_tile] %-32s %7.1f ms" % (name, (_t[-1] - _t[-2]) * 1e3), file=__import__("sys").stderr)
    x = str(int(x)); y = str(int(y))
    x = x[:-2] if ".0" in x else x
    y = y[:-2] if ".0" in y else y
    folder = f"{local_path}{str(x)}/{str(y)}/"
    tile_idx = f'{str(x)}X{str(y)}Y'
    clouds_file = f'{folder}raw/clouds/clouds_{tile_idx}.hkl'
    cloud_mask_file = f'{folder}raw/clouds/cloudmask_{tile_idx}.hkl'
    s1_file = f'{folder}raw/s1/{tile_idx}.hkl'
    s2_10_file = f'{folder}raw/s2_10/{tile_idx}.hkl'
    s2_20_file = f'{folder}raw/s2_20/{tile_idx}.hkl'
    s2_dates_file = f'{folder}raw/misc/s2_dates_{tile_idx}.hkl'
    dem_file = f'{folder}raw/misc/dem_{tile_idx}.hkl'

    clouds = load(clouds_file)
    if exists(cloud_mask_file):
        clm = np.asarray(load(cloud_mask_file)).repeat(2, axis=1).repeat(2, axis=2)      # :687
        clm = clm_pairs(clm, sess)                                                        # :688-695
    else:
        clm = None

    s1 = np.asarray(load(s1_file))                                                        # :699-700 np.float32(s1) / 65535
    if np.issubdtype(s1.dtype, np.integer) and s1.min() >= 0 and s1.max() <= 65535:
        s1 = sess.to_float32(s1.astype(np.uint16, copy=False))
    else:
        s1 = divide(np.float32(s1), 65535, sess)
    s1 = s1_fill(s1, sess)                                                                # :702-705
    s1[..., -1] = sess.convert_to_db(np.ascontiguousarray(s1[..., -1]), 22)               # :707-708
    s1[..., -2] = sess.convert_to_db(np.ascontiguousarray(s1[..., -2]), 22)
    s1 = s1.astype(np.float32)
    _mark("load + S1 decode/fill/dB")

    s2_10 = _api.to_float32(load(s2_10_file), sess)
    s2_20 = _api.to_float32(load(s2_20_file), sess)
    dem = median_filter5(load(dem_file), sess)                                            # :713
    image_dates = load(s2_dates_file)

    width = s2_20.shape[1] * 2
    height = s2_20.shape[2] * 2
    s1 = adjust_shape(s1, width, height)
    s2_10 = adjust_shape(s2_10, width, height)
    dem = adjust_shape(dem, width, height)
    if len(s2_10.shape) == 3:
        s2_10 = s2_10[np.newaxis]
    if len(s2_20.shape) == 3:
        s2_20 = s2_20[np.newaxis]

    _mark("S2 decode + DEM filter + adjust")
    sentinel2 = sess.build_sentinel2(s2_10, s2_20)                                        # :743-782
    _mark("build_sentinel2")

    missing_px = _api.id_missing_px(sentinel2, 2, sess)                                   # :786
    if len(missing_px) > 0:
        if clouds.shape[0] == len(image_dates):
            clouds = np.delete(clouds, missing_px, axis=0)
        image_dates = np.delete(image_dates, missing_px)
        sentinel2 = np.delete(sentinel2, missing_px, axis=0)
        if clm is not None:
            clm = np.delete(clm, missing_px, axis=0)

    HW = sentinel2.shape[1] * sentinel2.shape[2]
    snow_per_date, snow = snow_mask(sentinel2, sess)                                      # :808-829
    mean_snow_per_img = snow_per_date / HW
    to_remove = np.argwhere(mean_snow_per_img > 0.25).flatten()
    if len(to_remove) > 10:                                                               # :831 ("currently defunct")
        if clouds.shape[0] == len(image_dates):
            clouds = np.delete(clouds, to_remove, axis=0)
        image_dates = np.delete(image_dates, to_remove)
        sentinel2 = np.delete(sentinel2, to_remove, axis=0)
        if clm is not None:
            clm = np.delete(clm, to_remove, axis=0)
    # interpolation.interpolate_missing_vals (:833) is a no-op: its guard `s2 >= 1 and s2 == 0` is never true

    _mark("missing px + snow")

    def masks(first):
        nonlocal clm
        cloudshad, fcps = _api.identify_clouds_shadows(sentinel2, dem, bbx, sess, forest_mask=forest_mask, urban_mask=urban_mask)
        if clm is not None:
            if clm.shape == np.asarray(fcps).shape == np.asarray(cloudshad).shape:       # the reference's try/except guards a shape mismatch
                if first:
                    clm[fcps] = 0.                                                        # :843 (in place: later rounds see it)
                cloudshad = max_masked(cloudshad, clm, None, sess)                        # :844 / :874 / ...; libstc errors propagate
        return cloudshad, fcps

    def frac_gt0(a):
        return count_gt(a, 0.0, sess) / HW

    if make_shadow:
        cloudshad, fcps = masks(True)
        interp = _api.id_areas_to_interp(sentinel2, cloudshad, cloudshad, image_dates, fcps, sess)
        for attempt in range(3):                                                          # :863-929: three identical retry rounds
            to_remove = np.argwhere(frac_gt0(interp) > 0.9).flatten()
            if len(to_remove) > 0:
                if attempt == 2 or clouds.shape[0] == len(image_dates):
                    clouds = np.delete(clouds, to_remove, axis=0)
                image_dates = np.delete(image_dates, to_remove)
                sentinel2 = np.delete(sentinel2, to_remove, axis=0)
                interp = np.delete(interp, to_remove, axis=0)
                if clm is not None:
                    clm = np.delete(clm, to_remove, axis=0)
                cloudshad, fcps = masks(False)
                if attempt < 2:
                    interp = _api.id_areas_to_interp(sentinel2, cloudshad, cloudshad, image_dates, fcps, sess)
        _mark("cloud masks + feather + screening")
        interp = _api.id_areas_to_interp(sentinel2, cloudshad, cloudshad, image_dates, fcps, sess)   # :917
        if not (isinstance(sentinel2, np.ndarray) and sentinel2.dtype == np.float32 and sentinel2.flags.c_contiguous):
            sentinel2 = np.ascontiguousarray(sentinel2, np.float32)
        _, interp, to_remove, clipped = _api.remove_cloud_and_shadows(sentinel2, cloudshad, cloudshad, image_dates, fcps, None, sess=sess,
                                                                      clip_when_all_kept=True)
        _mark("remove_cloud_and_shadows")
        if len(to_remove) > 0:                                                            # :972-990
            clouds = np.delete(clouds, to_remove, axis=0)
            image_dates = np.delete(image_dates, to_remove)
            sentinel2 = np.delete(sentinel2, to_remove, axis=0)
            interp = np.delete(interp, to_remove, axis=0)
            if clm is not None:
                clm = np.delete(clm, to_remove, axis=0)
            cloudshad, fcps = masks(False)
            interp = _api.id_areas_to_interp(sentinel2, cloudshad, cloudshad, image_dates, fcps, sess)
    else:
        clipped = False
        interp = np.zeros((sentinel2.shape[0], sentinel2.shape[1], sentinel2.shape[2]), dtype=np.float32)
        cloudshad = np.zeros((sentinel2.shape[0], sentinel2.shape[1], sentinel2.shape[2]), dtype=np.float32)

    dem = divide(dem, 90, sess)                                                           # :995
    if not clipped:                                                                       # :996 (already done on the device
        sentinel2 = clip01(sentinel2, sess)                                               #  when remove_clouds kept all dates)
    _mark("clip + dem scale")
    return sentinel2, image_dates, interp, s1, dem, cloudshad, snow


def nan_to_zero(x, sess):
    """interpolation.interpolate_na_vals (src/preprocessing/interpolation.py:42-56): NaN -> the temporal median,
    which is NaN (np/bn.median propagate NaN) and therefore reset to 0 -- i.e. every NaN becomes 0."""
    a = np.array(x, dtype=np.float32, order="C")
    _check(sess, sess.lib.stc_elementwise_host(sess.h, _ptr(a), a.size, 2, 0.0, 0.0))
    return a


def count_lt_axis0(data, thresh, sess):
    a = np.ascontiguousarray(data, np.float32)
    n = a.shape[0]
    out = np.empty(a.shape[1:], np.int32)
    _check(sess, sess.lib.stc_count_lt_axis0_host(sess.h, _ptr(a), n, a.size // n, float(thresh), _ptr(out)))
    return out


def process_subtiles(x, y, s2=None, dates=None, interp=None, s1=None, dem=None, sess=None, bbx=None, size=158, train_bbx=None,
                     local_path="", length=4, gen_feats=False):
    """:1125-1486 for the prediction path (args.process, optionally --gen_feats; no --gen_composite /
    --make_training_data): medians, smoothing, quarterly composites, 6x6 overlapping windows with the
    reference's edge padding, 17-channel assembly, prediction, post-filters, one
    `<local_path><x>/<y>/processed/<folder_y>/<folder_x>.npy` per subtile (float32, 255 = no data).
    B200 shape of the loop: all subtiles of the tile go through ONE batched forward.
    gen_feats (:1429-1448): for every subtile with data, `feats/<folder_y>/<folder_x>.npy` = int16 x1000 of
    [early features 0..31 | late features 0..31] (one more batched forward that returns both taps).
    Not reproduced (I/O products, out of scope): ard_ndmi.hkl, ard_dates.npy, the composite GeoTIFF / ARD
    uploads (:1161-1205).  `bbx` / `train_bbx` are accepted and unused, like in the prediction path."""
    if sess is None:
        raise RuntimeError("process_subtiles needs an StcSession (sess=...); there is no CPU path")
    SIZE = size
    x = str(int(x)); y = str(int(y))
    x = x[:-2] if ".0" in x else x
    y = y[:-2] if ".0" in y else y
    import time as _time
    _t = [_time.perf_counter()]
    _trace = os.environ.get("STC_TILE_TIMING")

    def _mark(name):
        if _trace:
            _t.append(_time.perf_counter())
            print("[process_subtiles] %-28s %7.1f ms" % (name, (_t[-1] - _t[-2]) * 1e3), file=__import__("sys").stderr)

    # :1148-1159 NaN -> 0, medians of the raw dates (bands + indices), missing-pixel counts: one upload
    s2 = np.array(s2, dtype=np.float32, order="C")
    n0, Hh, Ww, _ = s2.shape
    s2_median = np.empty((Hh, Ww, 14), np.float32)
    bad_px = np.zeros(n0, np.int32)
    nan_total = _api.C.c_int64(0)
    _check(sess, sess.lib.stc_s2_medians_host(sess.h, _ptr(s2), n0, Hh, Ww, _ptr(s2_median), _ptr(bad_px), _api.C.byref(nan_total)))
    _mark("nan fill + medians + indices")
    # :1171 smooth_large_tile (deal_w_missing_px -> indices -> regrid / Whittaker / monthly mean) and :1174, :1274-1278
    # the quarterly / annual medians, fused on the device; the date screening stays here (scalars)
    missing = np.argwhere(bad_px >= (s2.shape[1] ** 2) / 10).flatten()                   # id_missing_px(arr, 10)
    if len(missing) > 0:
        dates = np.delete(dates, missing)
        s2 = np.delete(s2, missing, 0)
        interp = np.delete(interp, missing, 0)
    s1 = np.ascontiguousarray(s1, np.float32)
    fused = length == 4 and s1.shape[0] == 12
    if fused:
        try:
            M, _ = _api._regrid.monthly_operator(dates)
        except Exception:
            M = None
        fused = M is not None
    if fused:
        M = np.ascontiguousarray(M, np.float32)
        s2q = np.empty((4, Hh, Ww, 14), np.float32)
        s1q = np.empty((4, Hh, Ww, 2), np.float32)
        s1_median = np.empty((1, Hh, Ww, 2), np.float32)
        nan_after = np.zeros(s2.shape[0], np.int32)
        s2c = np.ascontiguousarray(s2)
        _check(sess, sess.lib.stc_smooth_quarterly_host(sess.h, _ptr(s2c), s2c.shape[0], Hh, Ww, _ptr(M), _ptr(s1), None, _ptr(s2q), _ptr(s1q),
                                                        _ptr(s1_median), _ptr(nan_after)))
        fused = not nan_after.any()
    if fused:
        s2, s1 = s2q, s1q
        s2_median = s2_median[np.newaxis]
        _mark("smooth + quarterly (fused)")
    else:                      # NaN dates after the fill, no usable dates, or length != 4: statement-by-statement path
        s2, dates, interp = _api.smooth_large_tile(s2, dates, interp, sess)
        s2_median = s2_median[np.newaxis]
        s1_median = sess.temporal_median(s1)[np.newaxis].astype(np.float32)
        if length == 4:
            s2 = np.stack([sess.temporal_median(s2[3 * q:3 * q + 3]) for q in range(4)])
            s1 = np.stack([sess.temporal_median(s1[3 * q:3 * q + 3]) for q in range(4)])
        elif length == 1:
            s2 = np.repeat(sess.temporal_median(s2)[np.newaxis], 4, axis=0)
            s1 = np.repeat(sess.temporal_median(s1)[np.newaxis], 4, axis=0)
        _mark("smooth + quarterly (unfused)")

    from .windows import subtile_windows
    tiles_folder, tiles_array = subtile_windows(s1.shape[1], s1.shape[2], size, 6 if SIZE != 222 else 7)
    path = f'{local_path}{str(x)}/{str(y)}/processed/'
    clear_all = count_lt_axis0(interp, 0.33, sess)                                        # np.sum(interp < 0.33, axis=0), whole tile
    _mark("quarterly medians + counts")

    # window table only (integers); gather, stacks, no-image test, forward and post-filters run on the device
    # (oracle/subtiles_ref.host_gather_stacks is the statement-by-statement NumPy restatement the tests compare against)
    if s2.shape[0] != length or s1.shape[0] != length:
        raise ValueError("process_subtiles: %d / %d composite frames for length %d" % (s2.shape[0], s1.shape[0], length))
    table = subtile_table(tiles_array, s2.shape[1], s2.shape[2], SIZE)
    outputs = [f"{path}{str(tf[1])}/{str(tf[0])}.npy" for tf in tiles_folder]
    out = np.empty((len(table), SIZE, SIZE), np.float32)
    flags = np.zeros(len(table), np.int32)
    mn, mnp = _api._f64(sess.min_all)
    mx, mxp = _api._f64(sess.max_all)
    s2 = np.ascontiguousarray(s2, np.float32); s1 = np.ascontiguousarray(s1, np.float32)
    dem32 = np.ascontiguousarray(dem, np.float32)
    s2m = np.ascontiguousarray(s2_median[0], np.float32); s1m = np.ascontiguousarray(s1_median[0], np.float32)
    args = (sess.h, _ptr(s2), _ptr(s1), _ptr(s2m), _ptr(s1m), _ptr(dem32), _ptr(clear_all), s2.shape[1], s2.shape[2], len(table), _ptr(table),
            SIZE, length, length, int(len(dates) < 2), mnp, mxp, _ptr(out), _ptr(flags))
    if gen_feats:
        early = np.empty((len(table), SIZE, SIZE, 64), np.float32)
        late = np.empty((len(table), SIZE, SIZE, 64), np.float32)
        _check(sess, sess.lib.stc_process_subtiles_feats_host(*args, _ptr(early), _ptr(late)))
    else:
        _check(sess, sess.lib.stc_process_subtiles_host(*args))
    _mark("gather + forward + post-filters (device)")
    for i in range(len(table)):
        os.makedirs(os.path.realpath(os.path.dirname(outputs[i])), exist_ok=True)
        np.save(outputs[i], out[i])
    _mark("save")
    if gen_feats:                                                                         # :1429-1446
        live = [i for i in range(len(table)) if not flags[i]]
        if live:
            both = sess.float_to_int16(np.concatenate([early[live][..., :32], late[live][..., :32]], axis=-1))
            root = f'{local_path}{str(x)}/{str(y)}/'
            for k, i in enumerate(live):
                out_f = outputs[i].replace(path, root + "feats/")
                os.makedirs(os.path.realpath(os.path.dirname(out_f)), exist_ok=True)
                np.save(out_f, both[k])
            os.makedirs(os.path.realpath(root + "raw/feats/"), exist_ok=True)
            os.makedirs(os.path.realpath(root + "ard/"), exist_ok=True)
        _mark("features")


def subtile_table(tiles_array, H, W, SIZE):
    """The integer description of the subtile loop's slicing and edge padding (:1345-1388) that the device gather
    consumes: per subtile (row0, col0, rows, cols, data pads rows-before/after, cols-before/after, min_clear pads in the
    same order).  The reference pads min_clear in its first-axis block with the (pad_u, pad_d) of the second-axis block --
    whatever those loop variables currently hold -- which is kept."""
    table = np.zeros((len(tiles_array), 12), np.int32)
    pad_u = pad_d = None
    for t, (start_x, start_y, nr, nc) in enumerate(np.asarray(tiles_array).astype(int).tolist()):
        nr = min(start_x + nr, H) - start_x
        nc = min(start_y + nc, W) - start_y
        row = table[t]
        row[:4] = (start_x, start_y, nr, nc)
        if nc == SIZE + 7:                                       # second array axis (:1369-1377)
            pad_u, pad_d = (7, 0) if start_y == 0 else (0, 7)
            row[6:8] = (pad_u, pad_d)
            row[10:12] = (pad_u, pad_d)
        if nr == SIZE + 7:                                       # first array axis (:1378-1388)
            if pad_u is None:
                raise NameError("pad_u")                         # the reference reads the variable before any assignment here
            row[4:6] = (7, 0) if start_x == 0 else (0, 7)
            row[8:10] = (pad_u, pad_d)
    return table


def run_tile(x, y, local_path, sess, bbx=None, loader=None, exists=os.path.exists, make_shadow=True, size=158, length=4,
             return_subtiles=False):
    """The body of the reference's main loop for one tile (:1995-2020) in ONE device-resident call
    (StcSession.run_tile -> stc_tile_run_host): raw/*.hkl arrays -> uint8 tree-cover tile.  Same file naming and
    `loader` / `exists` stand-ins as process_tile.  `bbx` is accepted and unused (see api.identify_clouds_shadows)."""
    del bbx
    load = loader or _default_loader
    x = str(int(x)); y = str(int(y))
    folder = f"{local_path}{x}/{y}/"
    tile_idx = f'{x}X{y}Y'
    cloud_mask_file = f'{folder}raw/clouds/cloudmask_{tile_idx}.hkl'
    clm = np.asarray(load(cloud_mask_file)) if exists(cloud_mask_file) else None
    s1 = np.asarray(load(f'{folder}raw/s1/{tile_idx}.hkl'))
    s2_10 = np.asarray(load(f'{folder}raw/s2_10/{tile_idx}.hkl'))
    s2_20 = np.asarray(load(f'{folder}raw/s2_20/{tile_idx}.hkl'))
    for name, a in (("s1", s1), ("s2_10", s2_10), ("s2_20", s2_20)):
        if not (np.issubdtype(a.dtype, np.integer) and a.min() >= 0 and a.max() <= 65535):
            raise ValueError("run_tile: raw/%s is not a uint16-range integer array; use process_tile for float storage" % name)
    dem = np.asarray(load(f'{folder}raw/misc/dem_{tile_idx}.hkl'), np.float32)
    dates = np.asarray(load(f'{folder}raw/misc/s2_dates_{tile_idx}.hkl'))
    if s2_10.ndim == 3:
        s2_10 = s2_10[np.newaxis]
    if s2_20.ndim == 3:
        s2_20 = s2_20[np.newaxis]
    return sess.run_tile(s2_10, s2_20, s1, dem, dates, clm=None if clm is None else (clm != 0), make_shadow=make_shadow, size=size,
                         length=length, return_subtiles=return_subtiles)
