"""TEST INFRASTRUCTURE (oracle) -- 20 m / 40 m band upsampling of process_tile
(/root/reference/src/download_and_predict_job.py:734-782) and missing-pixel handling
(src/preprocessing/interpolation.py:5-23, src/download_and_predict_job.py:1031-1054).

`skimage.transform.resize(img, shape, 1)` is a third-party call (scikit-image, unpinned in
requirements.txt and absent from this image).  Its published algorithm for float input when
upsampling (scikit-image >= 0.19, skimage/transform/_warps.py `resize`): no anti-aliasing filter,
`scipy.ndimage.zoom(img, [out/in per axis], order=1, mode='mirror', grid_mode=True)` (skimage's
default mode 'reflect' is ndimage's 'mirror'), then a clip to the input range (a no-op for
bilinear weights).  SciPy IS installed here, so the oracle calls that ndimage routine directly;
parity with scikit-image itself is UNPINNED (no skimage to run).  Tests only.
"""
import numpy as np
from scipy import ndimage as ndi


def resize_bilinear(img, shape):
    img = np.asarray(img)
    factors = np.divide(img.shape, shape)
    zoom = [1 / f for f in factors]
    out = ndi.zoom(img, zoom, order=1, mode="mirror", grid_mode=True)
    assert out.shape == tuple(shape), (out.shape, shape)
    return out


def build_sentinel2(s2_10, s2_20):
    """:743-782: (n,W,H,4) 10 m + (n,W/2,H/2,6) 20 m -> (n,W,H,10) float32."""
    width, height = s2_20.shape[1] * 2, s2_20.shape[2] * 2
    out = np.zeros((s2_10.shape[0], width, height, 10), np.float32)
    out[..., :4] = s2_10
    for band in range(4):
        for step in range(out.shape[0]):
            out[step, ..., band + 4] = resize_bilinear(s2_20[step, ..., band], (width, height))
    for band in range(4, 6):
        for step in range(out.shape[0]):
            mid = s2_20[step, ..., band]
            ey, ex = mid.shape[0] % 2 == 0, mid.shape[1] % 2 == 0
            if ey and ex:
                m = mid.reshape(mid.shape[0] // 2, 2, mid.shape[1] // 2, 2)
                out[step, ..., band + 4] = resize_bilinear(np.mean(m, axis=(1, 3)), (width, height))
            elif not ey and not ex:
                mx, my = mid[0, :], mid[:, 0]
                m = mid[1:, 1:].reshape(mid.shape[0] // 2, 2, mid.shape[1] // 2, 2)
                out[step, 1:, 1:, band + 4] = resize_bilinear(np.mean(m, axis=(1, 3)), (width - 1, height - 1))
                out[step, 0, :, band + 4] = mx.repeat(2)
                out[step, :, 0, band + 4] = my.repeat(2)
            elif not ey:
                mx = mid[0, :]
                m = mid[1:].reshape(mid.shape[0] // 2, 2, mid.shape[1] // 2, 2)
                out[step, 1:, :, band + 4] = resize_bilinear(np.mean(m, axis=(1, 3)), (width - 1, height))
                out[step, 0, :, band + 4] = mx.repeat(2)
            else:
                my = mid[:, 0]
                m = mid[:, 1:].reshape(mid.shape[0] // 2, 2, mid.shape[1] // 2, 2)
                out[step, :, 1:, band + 4] = resize_bilinear(np.mean(m, axis=(1, 3)), (width, height - 1))
                out[step, :, 0, band + 4] = my.repeat(2)
    return out


def id_missing_px(s2, thresh=11):
    """interpolation.py:5-23."""
    bad = np.sum(s2[..., :10] == 0.0, axis=-1) + np.sum(s2[..., :10] >= 1., axis=-1)
    per_date = np.sum(bad > 1., axis=(1, 2))
    return np.argwhere(per_date >= (s2.shape[1] ** 2) / thresh).flatten()


def deal_w_missing_px(arr, dates, interp):
    """download_and_predict_job.py:1031-1054 (prints dropped)."""
    missing = id_missing_px(arr, 10)
    if len(missing) > 0:
        dates = np.delete(dates, missing)
        arr = np.delete(arr, missing, 0)
        interp = np.delete(interp, missing, 0)
    for sentinel in (0, 1):
        if np.sum(arr == sentinel) > 0:
            for i in range(arr.shape[0]):
                a = arr[i]
                a[a == sentinel] = np.median(arr, axis=0)[a == sentinel]
    bad = np.argwhere(np.sum(np.isnan(arr), axis=(1, 2, 3)) > 0).flatten()
    if len(bad) > 0:
        dates = np.delete(dates, bad)
        arr = np.delete(arr, bad, 0)
        interp = np.delete(interp, bad, 0)
    return arr, dates, interp
