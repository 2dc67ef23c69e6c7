"""TEST INFRASTRUCTURE (oracle) -- NumPy/SciPy restatement of cloud / shadow removal,
/root/reference/src/preprocessing/cloud_removal.py:888-973 (`remove_cloud_and_shadows`) with
`make_aligned_mosaic` (:578-699, randomforest=False), `align_interp_array_randomforest` (:316-575,
linregress=True, equibatch) and `calculate_clouds_in_mosaic` (:703-732).
sklearn.LinearRegression(positive=True, fit_intercept=False) is scipy.optimize.nnls on (X, y)
followed by `X @ coef_`; the oracle calls nnls directly.  Sampling uses Python's global `random`
exactly as the reference does, so a pinned `random.seed` reproduces it.
Pinned against the reference function executed through oracle/refshim.py
(tests/test_cloud_fill.py).  Tests only; never imported by the product."""
import random
import numpy as np
from scipy.ndimage import binary_dilation as dil, distance_transform_edt as edt, grey_closing
from scipy.optimize import nnls


def feather(probs, size=20):
    """:908-921 (size 20) / id_areas_to_interp :783-795 (size 15)."""
    a = np.copy(probs).astype(np.float32)
    for t in range(a.shape[0]):
        if np.sum(a[t]) > 0:
            b = edt(1 - a[t])
            b[b > 12] = 12
            b = 1 - (b / 12)
            b[b < 0.2] = 0.
            a[t] = grey_closing(b, size=size)
    return a.astype(np.float32)


def _ndwi(x):
    return (x[..., 1] - x[..., 3]) / (x[..., 1] + x[..., 3])


def make_aligned_mosaic(arr, interp):
    """:578-699.  Mutates `interp` (dates that fail the 1000-pixel test are set to 1)."""
    n, H, W, C = arr.shape
    with np.errstate(all="ignore"):
        water = np.median(_ndwi(arr), axis=0) > 0
        water = dil(1 - water, iterations=2)
        water = dil(1 - water, iterations=5)
        mosaic = np.zeros((H, W, C), np.float32)
        divisor = np.sum(1 - interp, axis=0)[..., None]
        for i in range(n):
            mask_i = np.logical_and(interp[i] < 0.25, water == 0)
            ref = np.zeros((H, W, C), np.float32)
            cnt = np.zeros((H, W, C), np.float32)
            for b in range(n):
                if b != i:
                    m = np.logical_and(np.logical_and(interp[i] < 0.25, interp[b] < 1), water == 0)
                    ref[m * mask_i] += arr[b][m * mask_i]
                    cnt[m * mask_i] += 1
            ref = ref / cnt
            mask_i[cnt[..., 0] == 0] = 0
            src = arr[i][mask_i]
            ref = ref.reshape(H * W, C)
            ref = ref[~np.isnan(ref).any(axis=1)]
            if src.shape[0] > 1000 and ref.shape[0] > 1000:
                src = src[:ref.shape[0]]
                ref = ref[:src.shape[0]]
                mean_ref, std_ref = np.nanmedian(ref, axis=0), np.nanstd(ref, axis=0)
                mean_src, std_src = np.nanmedian(src, axis=0), np.nanstd(src, axis=0)
                mult = std_ref / std_src
                add = mean_ref - mean_src * mult
                x = np.copy(arr[i])
                x[water == 0] = x[water == 0] * mult + add
                mosaic = mosaic + (1 - interp[i][..., None]) * x
            elif np.mean(water < 0.9):
                interp[i] = 1.
        divisor[divisor < 0] = 0.
        mosaic = mosaic / divisor
        mosaic[np.isnan(mosaic)] = np.percentile(arr, 10, axis=0)[np.isnan(mosaic)]
        mosaic = np.maximum(mosaic, np.min(arr, axis=0))
        mosaic = np.minimum(mosaic, np.max(arr, axis=0))
    return mosaic


def snow_filter(arr):
    """:348-370 (probability form)."""
    with np.errstate(all="ignore"):
        ndsi = (arr[..., 1] - arr[..., 8]) / (arr[..., 1] + arr[..., 8])
        ndsi[ndsi < 0.10] = 0.
        ndsi[ndsi > 0.42] = 0.42
        p = (ndsi - 0.1) / 0.32
        p[arr[..., 3] < 0.10] = 0.
        p[np.logical_and(arr[..., 3] > 0.35, p > 0)] = 1.
        p[arr[..., 0] < 0.10] = 0.
        p[np.logical_and(arr[..., 0] > 0.22, p > 0)] = 1.
        p[(arr[..., 0] / arr[..., 2]) < 0.75] = 0.
    return p


def _evi(x):
    e = 2.5 * ((x[..., 3] - x[..., 2]) / (x[..., 3] + (6 * x[..., 2]) - (7.5 * x[..., 0]) + 1))
    return np.clip(e, -1.5, 1.5)


def sample_indices(evi, n_rows):
    """:447-491: EVI-percentile strata + oversampled tails, shuffled with Python's `random`."""
    n_samples = np.minimum(90000, n_rows)
    n_i = n_samples // 5
    b2, b20, b40, b60, b80, b98 = [np.percentile(evi, q) for q in (2, 20, 40, 60, 80, 98)]
    p2 = np.argwhere(evi < b2).squeeze()
    p20 = np.argwhere(evi < b20).squeeze()
    p40 = np.argwhere(np.logical_and(evi >= b20, evi < b40)).squeeze()
    p60 = np.argwhere(np.logical_and(evi >= b40, evi < b60)).squeeze()
    p80 = np.argwhere(np.logical_and(evi >= b60, evi < b80)).squeeze()
    p100 = np.argwhere(evi >= b80).squeeze()
    p98 = np.argwhere(evi >= b98).squeeze()
    p98 = np.repeat(p98, 10)
    p2 = np.repeat(p2, 10)
    for p in (p2, p98, p20, p40, p60, p80, p100):
        random.shuffle(p)
    s = np.concatenate([p2, p20[:n_i], p40[:n_i], p60[:n_i], p80[:n_i], p100[:n_i], p98])
    random.shuffle(s)
    return s[:n_rows]


def align_date(interp_array, array, date, interp, mosaic, water_mask, taps=None):
    """:316-575 for one date; returns the (H,W,10) array that replaces the cloudy pixels."""
    n = array.shape[0]
    snow = np.mean(snow_filter(array), axis=0)[..., None]
    a = interp[date]
    if not (np.sum(a > 0) > 0 and np.sum(a == 0) > 0):
        return interp_array
    if not (np.mean(np.logical_and(a < 1, water_mask <= 1)) > 0.01):
        raise UnboundLocalError("to_remove")          # the reference falls through to an unbound name (:575)
    n_cur = np.sum(np.logical_and(a == 0, water_mask <= 1))
    if n_cur > 40000:
        lo, hi = max(date, 0), date + 1
    else:
        lo = max(date - 2, 0) if date == n - 1 else max(date - 1, 0)
        hi = min(date + 2, n)
    areas, mos = [], []
    for t in range(lo, hi):
        req = np.logical_and(interp[t] == 0, water_mask < 1)
        areas.append(np.concatenate([array[t], snow], axis=-1)[req])
        mos.append(np.concatenate([mosaic, snow], axis=-1)[req])
    if n_cur > 40000:
        areas, mos = areas[0], mos[0]
    else:
        areas, mos = np.concatenate(areas, axis=0), np.concatenate(mos, axis=0)
    s = sample_indices(_evi(areas), mos.shape[0])
    mos, areas = mos[s], areas[s]
    out = np.copy(interp_array)
    feats = np.concatenate([interp_array, snow], axis=-1).reshape(-1, 11)
    sel = np.logical_and(a > 0, water_mask <= 1)
    coefs = []
    for band in range(10):
        train_x = np.copy(mos)
        mos[..., band] = np.clip(mos[..., band], 0.005, 1)
        coef = nnls(train_x, areas[..., band])[0]
        coefs.append(coef)
        pred = (feats @ coef + 0.0).reshape(a.shape)
        out[sel, band] = pred[sel]
    if taps is not None:
        taps.setdefault("coef", {})[date] = np.array(coefs)
        taps.setdefault("sample", {})[date] = s
    return out


def clouds_in_mosaic(mosaic, interp, pfcps):
    """:703-732."""
    only1 = np.sum(1 - (interp > 0), axis=0).squeeze() < 2
    if len(pfcps.shape) == 3 and pfcps.shape[0] > 1:
        pfcps = pfcps[0]
    pfcps = dil(pfcps, iterations=10)
    only1 = np.maximum(only1, pfcps.squeeze())
    if np.sum(only1) == np.prod(only1.shape):
        return np.zeros_like(only1)
    rb = np.percentile(mosaic[..., 0][~only1], 99)
    rr = np.percentile(mosaic[..., 2][~only1], 99)
    c = (mosaic[..., 0] > rb) * (mosaic[..., 2] > rr) * only1 * (np.sum(mosaic[..., :3], axis=-1) < 1)
    c[pfcps.squeeze() > 0] = 0.
    c = dil(1 - c, iterations=3)
    return dil(1 - c, iterations=8)


def remove_cloud_and_shadows(tiles, probs, pfcps, taps=None):
    """:888-973 (shadows / image_dates / sentinel1 arguments are unused by the reference body)."""
    areas = feather(probs, 20)
    mosaic = make_aligned_mosaic(tiles, areas)
    with np.errstate(all="ignore"):
        water = _ndwi(np.median(tiles, axis=0)) > 0.0
    if taps is not None:
        taps["areas"], taps["mosaic"], taps["water"] = areas.copy(), mosaic.copy(), water.copy()
    to_remove = []
    for date in range(tiles.shape[0]):
        ia = np.zeros_like(tiles[date])
        ia[areas[date] > 0] = mosaic[areas[date] > 0]
        ia = align_date(ia, tiles, date, areas, mosaic, water, taps)
        tiles[date] = tiles[date] * (1 - areas[date][..., None]) + ia * areas[date][..., None]
        if np.mean(areas[date] == 1) == 1:
            to_remove.append(date)
    areas = areas + clouds_in_mosaic(mosaic, areas, pfcps)[None]
    areas[areas > 1] = 1.
    return tiles, areas, to_remove
