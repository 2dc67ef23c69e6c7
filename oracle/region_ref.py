"""TEST INFRASTRUCTURE (oracle): region-scale grid bookkeeping and Gaussian overlap blend, the generalisation of
load_mosaic_predictions (/root/reference/src/download_and_predict_job.py:1515-1641, fspecial_gauss :1489-1501) from the
6 x 6 subtile mosaic of one tile to an R x C patch grid.  NumPy, float32, patches
accumulated in row-major grid order -- the order the CUDA kernel reproduces.  There is no reference function at this
scale (the reference mosaics tile by tile); the per-pixel rules (x100, Gaussian weights sigma 36, uint8 truncation,
<= 15 -> 0, 255 = no data) are the reference's."""
import numpy as np

from oracle.preproc_ref import fspecial_gauss


def canvas_size(n, patch, stride):
    return (n - 1) * stride + patch


def synth_canvas_patch(base, y0, x0, patch):
    """Window of the periodic synthetic canvas: canvas[t, y, x, c] = base[t, y % Pb, x % Pb, c]."""
    ys = np.arange(y0, y0 + patch) % base.shape[1]
    xs = np.arange(x0, x0 + patch) % base.shape[2]
    return base[:, ys][:, :, xs]


def blend_region(preds, stride, margin=7, sigma=36, rows=None):
    """preds [R, C, S, S] float32 probabilities -> uint8 canvas [(R-1)*stride + S + 2*margin] x [...C...] (or the
    canvas rows `rows` = (y0, y1) only)."""
    R, C, S, _ = preds.shape
    Hc, Wc = canvas_size(R, S + 2 * margin, stride), canvas_size(C, S + 2 * margin, stride)
    num = np.zeros((Hc, Wc), np.float32)
    den = np.zeros((Hc, Wc), np.float32)
    w = fspecial_gauss(S, sigma).astype(np.float32)
    for r in range(R):
        for c in range(C):
            ys, xs = r * stride + margin, c * stride + margin
            v = preds[r, c].astype(np.float32) * np.float32(100.0)
            num[ys:ys + S, xs:xs + S] += w * v
            den[ys:ys + S, xs:xs + S] += w
    with np.errstate(all="ignore"):
        q = num / den
    out = np.where(den > 0, q, 0).astype(np.uint8)
    out[out <= 15] = 0
    out[out > 100] = 255
    out[~(den > 0)] = 255
    if rows is not None:
        out = out[rows[0]:rows[1]]
    return out
