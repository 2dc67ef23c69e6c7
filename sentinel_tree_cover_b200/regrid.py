"""Host-side date logic of the temporal stack: everything that depends only on the
image dates (a handful of integers per tile) stays in NumPy and is folded into ONE
12 x n matrix that the GPU applies to every pixel-band column (stc_temporal_matmul).

  regrid_matrix(dates)   G (24 x n): irregular dates -> 24 steps (day 0,15,...,345),
                         <=2 images before / <=2 after with inverse-distance weights and
                         year wrap  -- calculate_and_save_best_images,
                         /root/reference/src/downloading/utils.py:176-347
  whittaker_matrix()     S (24 x 24) = (I + 100 D2'D2)^-1 -- Smoother.__init__/smooth,
                         /root/reference/src/preprocessing/whittaker_smoother.py:10-42
  pair_mean_matrix()     A (12 x 24): mean of consecutive pairs -- :64-67
  monthly_operator(d)    M = A S G (12 x n), float32
(The array work of deal_w_missing_px / id_missing_px runs on the GPU: api.deal_w_missing_px.)
"""
import numpy as np

GRID_DAYS = np.arange(0, 360, 15)


def _neighbour_weights(offsets, day_min, day_max):
    """For one grid day: `offsets` = image_day - grid_day in image order.  Returns
    (before_offsets, after_offsets, before_w, after_w) following utils.py:219-278."""
    before = offsets[offsets < 5][-2:]
    if before.size:
        before = before[before > before.max() - 100]
    after = offsets[offsets >= -5][:2]
    if after.size:
        after = after[after < after.min() + 100]
    wrap_b = wrap_a = 0
    if before.size == 0:
        if day_min >= 90:                 # nothing early in the year: borrow the last image, one year back
            before, wrap_b = offsets[-1:], 365
        else:
            before = after
    if after.size == 0:
        if day_max <= 270:                # nothing late in the year: borrow the first image, one year ahead
            after, wrap_a = offsets[:1], 365
        else:
            after = before
    db = np.maximum(np.abs(before - wrap_b), 1.0)
    da = np.maximum(np.abs(after + wrap_a), 1.0)
    span = max(db[-1] + da[0], 2)
    wb = np.abs(1 - db / span)
    wa = np.abs(1 - da / span)
    if wb.size == 2:
        wb[0] = abs((db[1] / db[0]) * wb[1])
    if wa.size == 2:
        wa[1] = abs((da[0] / da[1]) * wa[0])
    tot = wb.sum() + wa.sum()
    return before, after, wb / tot, wa / tot


def regrid_matrix(dates):
    """G [24, n] float32 and max_distance (utils.py:304-311)."""
    dates = np.array(dates)
    dates[dates < -100] = dates[dates < -100] % 365
    n = dates.shape[0]
    eye = np.eye(n, dtype=np.float32)
    G = np.zeros((GRID_DAYS.size, n), np.float32)
    max_distance = 0
    for r, day in enumerate(GRID_DAYS):
        off = dates - day
        before, after, wb, wa = _neighbour_weights(off, dates.min(), dates.max())
        # images whose date equals one of the selected dates (value match, as the reference does)
        ib = sorted(set(np.flatnonzero(np.isin(dates, day + before)).tolist()))[:2]
        ia = sorted(set(np.flatnonzero(np.isin(dates, day + after)).tolist()))[-2:]
        # same broadcasting rule as `img_bands[idx] * ratio[:, None, None, None]` then sum(axis=0)
        row = (eye[ib] * wb.astype(np.float32)[:, None]).sum(0) + (eye[ia] * wa.astype(np.float32)[:, None]).sum(0)
        G[r] = row
        sel = np.concatenate([day + before, day + after]).ravel()
        if sel.size == 2:
            max_distance = max(max_distance, int(sel[1] - sel[0]))
    return G, max_distance


def whittaker_matrix(size=24, lmbd=100.0):
    D = np.zeros((size - 2, size))
    for i in range(size - 2):
        D[i, i:i + 3] = (1.0, -2.0, 1.0)
    coef = np.eye(size) + lmbd * (D.T @ D)
    return np.linalg.inv(coef)


def pair_mean_matrix(size=24, outsize=12):
    k = size // outsize
    A = np.zeros((outsize, size))
    for o in range(outsize):
        A[o, o * k:(o + 1) * k] = 1.0 / k
    return A


def monthly_operator(dates):
    G, max_distance = regrid_matrix(dates)
    M = pair_mean_matrix() @ whittaker_matrix() @ G.astype(np.float64)
    return M.astype(np.float32), max_distance


def s1_monthly_operator(dates):
    """process_sentinel_1_tile (/root/reference/src/tof/tof_downloading.py:75-95): regrid to 24 steps,
    then np.median of consecutive pairs -- the median of two values is their mean, so the whole
    stage is the 12 x n operator A G (no Whittaker smoothing for Sentinel-1)."""
    G, max_distance = regrid_matrix(dates)
    return (pair_mean_matrix() @ G.astype(np.float64)).astype(np.float32), max_distance


