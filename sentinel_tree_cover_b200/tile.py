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
    a = np.ascontiguousarray(s1, np.float32)
    m, H, W, Cc = a.shape
    _check(sess, sess.lib.stc_s1_fill_host(sess.h, _ptr(a), m, H, W, Cc))
    return a


def median_filter5(dem, sess):
    a = np.ascontiguousarray(dem, np.float32)
    out = np.empty_like(a)
    _check(sess, sess.lib.stc_median_filter5_host(sess.h, _ptr(a), a.shape[0], a.shape[1], _ptr(out)))
    return out


def clm_pairs(clm, sess):
    a = np.ascontiguousarray(clm, np.float32)
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
    a = np.ascontiguousarray(x, np.float32)
    _check(sess, sess.lib.stc_elementwise_host(sess.h, _ptr(a), a.size, 0, 0.0, 1.0))
    return a


def divide(x, d, sess):
    a = np.ascontiguousarray(x, np.float32)
    _check(sess, sess.lib.stc_elementwise_host(sess.h, _ptr(a), a.size, 1, float(d), 0.0))
    return a


def max_masked(a, b, zero, sess):
    """np.maximum(a, b) after b[zero] = 0 (float32)."""
    a = np.ascontiguousarray(a, np.float32).copy()
    b = np.ascontiguousarray(b, np.float32)
    z = None if zero is None else np.ascontiguousarray(np.asarray(zero) != 0, np.uint8)
    _check(sess, sess.lib.stc_max_masked_host(sess.h, _ptr(a), _ptr(b), None if z is None else _ptr(z), a.size))
    return a


def adjust_shape(arr, width, height):
    """:260-310 -- pad ('edge') or centre-crop axes 1/2 to (width, height); pure data movement."""
    arr = arr[:, :, :, np.newaxis] if len(arr.shape) == 3 else arr
    arr = arr[np.newaxis, :, :, np.newaxis] if len(arr.shape) == 2 else arr
    if arr.shape[1] < width:
        pad_amt = (width - arr.shape[1]) // 2
        if pad_amt == 0:
            arr = np.pad(arr, ((0, 0), (1, pad_amt), (0, 0), (0, 0)), 'edge')
        else:
            arr = np.pad(arr, ((0, 0), (pad_amt, pad_amt), (0, 0), (0, 0)), 'edge')
    if arr.shape[2] < height:
        pad_amt = (height - arr.shape[2]) // 2
        if pad_amt == 0:
            arr = np.pad(arr, ((0, 0), (0, 0), (1, 0), (0, 0)), 'edge')
        else:
            arr = np.pad(arr, ((0, 0), (0, 0), (pad_amt, pad_amt), (0, 0)), 'edge')
    if arr.shape[1] > width:
        pad_amt = (arr.shape[1] - width) // 2
        even = (arr.shape[1] - width) % 2 == 0
        if pad_amt == 0:
            arr = arr[:, 1:, ...]
        elif even:
            arr = arr[:, int(pad_amt):-int(pad_amt), ...]
        else:
            arr = arr[:, int(np.floor(pad_amt / 2)):-int(np.ceil(pad_amt / 2)), ...]
    if arr.shape[2] > height:
        pad_amt = (arr.shape[2] - height) // 2
        even = (arr.shape[2] - height) % 2 == 0
        if pad_amt == 0:
            arr = arr[:, :, 1:, :]
        elif even:
            arr = arr[:, :, int(pad_amt):-int(pad_amt), ...]
        else:
            arr = arr[:, :, int(np.floor(pad_amt / 2)):-int(np.ceil(pad_amt / 2)), ...]
    return arr.squeeze()


def _default_loader(path):
    import hickle as hkl            # not shipped with this package; pass `loader=` otherwise
    return hkl.load(path)


def process_tile(x, y, data, local_path, bbx, make_shadow=False, sess=None, loader=None, exists=os.path.exists):
    """:640-997, same arguments and return tuple
    `(sentinel2, image_dates, interp, s1, dem, cloudshad, snow)`."""
    if sess is None:
        raise RuntimeError("process_tile needs an StcSession (sess=...); there is no CPU path")
    load = loader or _default_loader
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

    sentinel2 = sess.build_sentinel2(s2_10, s2_20)                                        # :743-782

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

    def masks(first):
        nonlocal clm
        cloudshad, fcps = _api.identify_clouds_shadows(sentinel2, dem, bbx, sess)
        if clm is not None:
            try:
                if first:
                    clm[fcps] = 0.                                                        # :843 (in place: later rounds see it)
                cloudshad = max_masked(cloudshad, clm, None, sess)                        # :844 / :874 / ...
            except Exception:
                pass
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
        interp = _api.id_areas_to_interp(sentinel2, cloudshad, cloudshad, image_dates, fcps, sess)   # :917
        if not (isinstance(sentinel2, np.ndarray) and sentinel2.dtype == np.float32 and sentinel2.flags.c_contiguous):
            sentinel2 = np.ascontiguousarray(sentinel2, np.float32)
        _, interp, to_remove = _api.remove_cloud_and_shadows(sentinel2, cloudshad, cloudshad, image_dates, fcps, None, sess=sess)
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
        interp = np.zeros((sentinel2.shape[0], sentinel2.shape[1], sentinel2.shape[2]), dtype=np.float32)
        cloudshad = np.zeros((sentinel2.shape[0], sentinel2.shape[1], sentinel2.shape[2]), dtype=np.float32)

    dem = divide(dem, 90, sess)                                                           # :995
    sentinel2 = clip01(sentinel2, sess)                                                   # :996
    return sentinel2, image_dates, interp, s1, dem, cloudshad, snow
