"""TEST INFRASTRUCTURE (oracle) -- NumPy/SciPy restatement of the reference's temporal
preprocessing arithmetic.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this; never the product path.

Pinned against the reference's own functions executed through oracle/refshim.py
(tools/make_golden.py -> tests/golden/preproc_*.npz; tests/test_oracle_preproc.py).
All file:line citations are relative to /root/reference.
"""
import numpy as np
import scipy.sparse as sparse
from scipy.sparse.linalg import splu


# ---- src/preprocessing/indices.py:4-54 -------------------------------------------------
def grndvi(x):
    nir, green, red = np.clip(x[..., 3], 0., 1), np.clip(x[..., 1], 0., 1), np.clip(x[..., 2], 0., 1)
    return (nir - (green + red)) / ((nir + (green + red)) + 1e-5)


def evi(x):
    BLUE, RED, NIR = np.clip(x[..., 0], 0, 1), np.clip(x[..., 2], 0, 1), np.clip(x[..., 3], 0, 1)
    return np.clip(2.5 * ((NIR - RED) / (NIR + (6 * RED) - (7.5 * BLUE) + 1)), -1.5, 1.5)


def msavi2(x):
    RED, NIR = np.clip(x[..., 2], 0, 1), np.clip(x[..., 3], 0, 1)
    s = (2 * NIR + 1) ** 2 - 8 * (NIR - RED)
    s[s < 0] = 0.
    return np.clip((2 * NIR + 1 - np.sqrt(s)) / 2, -1, 1)


def bi(x):
    B11, B4, B8, B2 = (np.clip(x[..., 8], 0, 1), np.clip(x[..., 2], 0, 1), np.clip(x[..., 3], 0, 1), np.clip(x[..., 0], 0, 1))
    return np.clip(((B11 + B4) - (B8 + B2)) / (((B11 + B4) + (B8 + B2)) + 1e-5), -1, 1)


def make_indices(arr):
    """src/download_and_predict_job.py:998-1006 -> [...,4] = EVI, BI, MSAVI2, GRNDVI (float32)."""
    out = np.zeros(arr.shape[:-1] + (4,), np.float32)
    out[..., 0], out[..., 1], out[..., 2], out[..., 3] = evi(arr), bi(arr), msavi2(arr), grndvi(arr)
    return out


# ---- src/download_and_predict_job.py:316-325 -------------------------------------------
def normalize_subtile(subtile, min_all, max_all):
    subtile = subtile.copy()
    for band in range(subtile.shape[-1]):
        mins, maxs = min_all[band], max_all[band]
        subtile[..., band] = np.clip(subtile[..., band], mins, maxs)
        subtile[..., band] = (subtile[..., band] - (maxs + mins) / 2) / ((maxs - mins) / 2)
    return subtile


# ---- medians + 17-channel frame layout --------------------------------------------------
def assemble(monthly):
    """monthly [B,12,H,W,13] -> [B,5,H,W,17].  Quarterly = np.median over month triples
    (process_subtiles :1274-1278), frame 4 = np.median over the 12 steps (:1152-1160,1174);
    channel layout [0:10] S2, [10] DEM, [11:13] S1, [13:17] indices (:1398-1407); indices
    are computed per month from the 13-band cube as in
    src/download_and_predict_job_multiyear.py:808-813, then reduced by the same medians."""
    m = np.asarray(monthly, np.float32)
    B, n, H, W, _ = m.shape
    full = np.concatenate([m, make_indices(m)], axis=-1)            # [B,12,H,W,17]
    q = np.median(full.reshape(B, 4, 3, H, W, 17), axis=2)
    med = np.median(full, axis=1, keepdims=True)
    return np.concatenate([q, med], axis=1).astype(np.float32)


# ---- src/preprocessing/whittaker_smoother.py:10-69 --------------------------------------
class WhittakerRef:
    def __init__(self, lmbd=100, size=24):
        d = np.zeros(5, np.float32)
        d[2] = 1.
        for _ in range(2):
            d = d[:-1] - d[1:]
        E = sparse.eye(size, format='csc', dtype=np.float32)
        D = sparse.diags(d, np.arange(3), (size - 2, size), dtype=np.float32)
        self.lu = splu(E + D.conj().T.dot(D) * lmbd)
        self.size = size

    def interpolate_array(self, x):
        """x [24,H,W,C] -> [12,H,W,C]: splu solve per column, mean of consecutive pairs."""
        shp = x.shape
        z = self.lu.solve(np.array(x.reshape(self.size, -1)))
        z = z.reshape((12, self.size // 12) + shp[1:])
        return np.mean(z, axis=1)


def regrid_apply(G, arr):
    """keep_steps = sum_i w_i * img_i with float32 weights (src/downloading/utils.py:313-345)."""
    G = np.asarray(G, np.float32)
    out = np.zeros((G.shape[0],) + arr.shape[1:], np.float32)
    for r in range(G.shape[0]):
        nz = np.flatnonzero(G[r])
        out[r] = np.sum(arr[nz] * G[r, nz][:, None, None, None], axis=0)
    return out


def smooth_stack(arr, G):
    """smooth_large_tile numeric core (src/download_and_predict_job.py:1057-1096) given the
    24 x n regrid matrix: bands and indices regridded + Whittaker-smoothed separately."""
    sm = WhittakerRef()
    bands = sm.interpolate_array(regrid_apply(G, arr))
    if arr.shape[-1] != 10:
        return bands.astype(np.float32)
    idx = sm.interpolate_array(regrid_apply(G, make_indices(arr)))
    return np.concatenate([bands, idx], -1).astype(np.float32)


def temporal_matmul(M, arr):
    return np.einsum('on,n...->o...', np.asarray(M, np.float64), np.asarray(arr, np.float64)).astype(np.float32)


# ---- synthetic inputs --------------------------------------------
from sentinel_tree_cover_b200.synth import synth_monthly  # noqa: E402,F401  (the generator lives with the package: bench.py uses it too)


def synth_model_input(B, H, seed, T1=5):
    """Normalised model input [B,T1,H,H,17] in [-1,1] with spatial structure."""
    r = np.random.default_rng(seed)
    k = H // 4 + 1
    base = r.uniform(-1, 1, (B, 1, k, k, 17)).astype(np.float32)
    base = np.repeat(np.repeat(base, 4, 2), 4, 3)[:, :, :H, :H]
    x = base + 0.15 * r.standard_normal((B, T1, H, H, 17)).astype(np.float32)
    return np.ascontiguousarray(np.clip(x, -1, 1), np.float32)


# ---- src/download_and_predict_job.py:1489-1641 (depth == 1) ----------------------------
def fspecial_gauss(size, sigma):
    x, y = np.mgrid[-size // 2 + 1:size // 2 + 1, -size // 2 + 1:size // 2 + 1]
    return np.exp(-((x ** 2 + y ** 2) / (2.0 * sigma ** 2)))


def mosaic(preds, xs, ys, out_shape, sigma=36, dilate_iters=10):
    """load_mosaic_predictions for depth == 1, given the subtile arrays in the reference's layer
    order.  (bn.nanmean -> np.nanmean as in the import shim.)"""
    from scipy.ndimage import binary_dilation, generate_binary_structure
    n = len(preds)
    S = np.asarray(preds[0]).shape[0]
    predictions = np.full((out_shape[0], out_shape[1], n), np.nan, dtype=np.float32)
    mults = np.full((out_shape[0], out_shape[1], n), 0, dtype=np.float32)
    for i in range(n):
        prediction = np.array(preds[i])
        prediction[prediction < 255] = prediction[prediction < 255] * 100
        if np.sum(prediction) < S * S * 255:
            prediction = prediction.T.astype(np.float32)
            predictions[xs[i]:xs[i] + S, ys[i]:ys[i] + S, i] = prediction
            f = fspecial_gauss(S, sigma)
            f[prediction > 100] = 0.
            mults[xs[i]:xs[i] + S, ys[i]:ys[i] + S, i] = f
    ratios = np.zeros(n, np.float32)
    mults[np.isnan(predictions)] = 0.
    try:
        with np.errstate(all="ignore"):
            for i in range(n):
                sub = predictions[..., i]
                others = np.delete(predictions, i, -1)
                others = others[~np.isnan(sub)].reshape((S, S, n - 1))
                remove = np.argwhere(np.sum(np.isnan(others), axis=(0, 1)) == (S * S)).flatten()
                others = np.nanmean(np.delete(others, remove, -1), axis=-1)
                sub = sub[~np.isnan(sub)].reshape((S, S))
                ratios[i] = np.nanmean(abs(others - sub))
            multipliers = np.median(ratios) / ratios
            multipliers[multipliers > 1.5] = 1.5
        for i in range(n):
            mults[..., i] *= multipliers[i]
    except Exception:
        pass
    with np.errstate(all="ignore"):
        predictions[predictions > 100] = np.nan
        mults = mults / np.sum(mults, axis=-1)[..., np.newaxis]
        nan_count = np.sum(np.isnan(predictions), axis=2)
        out = np.nansum(predictions * mults, axis=-1)
        out[nan_count == n] = np.nan
        out[np.isnan(out)] = 255.
        out = out.astype(np.uint8)
    out[out <= .15 * 100] = 0.
    out[out > 100] = 255.
    no_images = binary_dilation(out == 255, structure=generate_binary_structure(2, 2), iterations=dilate_iters)
    out[no_images] = 255
    return out


def synth_subtile_preds(L=618, S=158, seed=0, nodata_blocks=True, all_nodata=()):
    """Synthetic saved subtile predictions for an L x L tile: smooth field + per-subtile bias,
    optional 40x40 no-data blocks (255) and whole no-data subtiles (int 255 fill, :366)."""
    from sentinel_tree_cover_b200.windows import subtile_windows
    r = np.random.default_rng(seed)
    folder, _ = subtile_windows(L, L, S)
    yy, xx = np.mgrid[0:L, 0:L]
    field = 0.5 + 0.45 * np.sin(xx / 37.0) * np.cos(yy / 53.0)
    preds, xs, ys = [], [], []
    for k, (x0, y0, _, _) in enumerate(folder):
        if k in all_nodata:
            p = np.full((S, S), 255)
        else:
            p = field[x0:x0 + S, y0:y0 + S].T + r.normal(0, 0.03, (S, S)) + r.normal(0, 0.02)
            p = np.around(np.clip(p, 0.001, 0.999), 3).astype(np.float32)
            if nodata_blocks and k % 7 == 3:
                p[40:80, 80:120] = 255.
        preds.append(p); xs.append(int(x0)); ys.append(int(y0))
    return preds, xs, ys


def float_to_int16(arr, precision=1000):
    """src/download_and_predict_job.py:174-180."""
    arr = np.array(arr, copy=True)
    arr[np.isnan(arr)] = -32768
    arr = np.clip(arr, (-32768 / precision), (32767 / precision))
    arr = arr * precision
    return np.int16(arr)


def mosaic_feats(feats, xs, ys, out_shape, depth, sigma=36):
    """load_mosaic_predictions for depth > 1 (src/download_and_predict_job.py:1540-1592,1628-1635), given the
    saved int16 feature stacks [S,S,>=depth] in the reference's layer order: 8 channels at a time, plain Gaussian
    weights normalised over the layer axis, nansum, int16 truncation."""
    S = feats[0].shape[0]
    N = len(feats)
    output = np.full((depth,) + tuple(out_shape), 0., dtype=np.int16)
    for start in np.arange(0, depth, 8):
        end = start + 8
        predictions = np.full((8, out_shape[0], out_shape[1], N), np.nan, dtype=np.float32)
        mults = np.full((1, out_shape[0], out_shape[1], N), 0, dtype=np.float32)
        for i, (f, x, y) in enumerate(zip(feats, xs, ys)):
            prediction = f[..., start:end]
            prediction = (prediction).T.astype(np.float32)
            predictions[:, x:x + S, y:y + S, i] = prediction
            mults[:, x:x + S, y:y + S, i] = fspecial_gauss(S, sigma)
        with np.errstate(all="ignore"):
            mults = mults / np.sum(mults, axis=-1)[..., np.newaxis]
            p = np.nansum(predictions * mults, axis=-1)
            output[start:end] = np.int16(p)
    return output


def synth_subtile_feats(L=250, S=158, D=16, seed=0):
    """Synthetic saved int16 feature stacks for an L x L tile (smooth fields + noise, x1000 codes)."""
    from sentinel_tree_cover_b200.windows import subtile_windows
    r = np.random.default_rng(seed)
    folder, _ = subtile_windows(L, L, S)
    yy, xx = np.mgrid[0:L, 0:L]
    feats, xs, ys = [], [], []
    for k, (x0, y0, _, _) in enumerate(folder):
        f = np.empty((S, S, D), np.float32)
        for c in range(D):
            field = np.sin(xx / (17.0 + c)) * np.cos(yy / (29.0 - c)) * (1.0 + 0.3 * c)
            f[..., c] = field[x0:x0 + S, y0:y0 + S].T + r.normal(0, 0.05, (S, S))
        feats.append(float_to_int16(f)); xs.append(int(x0)); ys.append(int(y0))
    return feats, xs, ys
