"""TEST INFRASTRUCTURE (oracle) -- SciPy restatement of the cloud-mask feathering and binary
morphology call sites of /root/reference/src/preprocessing/cloud_removal.py.  Tests only."""
import numpy as np
from scipy.ndimage import distance_transform_edt as distance, grey_closing, binary_dilation, generate_binary_structure


def feather(masks, closing_size):
    """cloud_removal.py:786-796 (size 15) / :913-921 (size 20)."""
    a = np.array(masks, np.float32)
    for d in range(a.shape[0]):
        if np.sum(a[d]) > 0:
            b = distance(1 - a[d])
            b[b > 12] = 12
            b = 1 - (b / 12)
            b[b < 0.2] = 0.
            a[d] = grey_closing(b, size=closing_size)
    return a.astype(np.float32)


def dilate(x, iterations, connectivity):
    st = generate_binary_structure(2, connectivity)
    x = np.asarray(x) != 0
    if x.ndim == 2:
        return binary_dilation(x, structure=st, iterations=iterations)
    return np.stack([binary_dilation(m, structure=st, iterations=iterations) for m in x])


def synth_cloud_masks(n, H, W, seed):
    r = np.random.default_rng(seed)
    m = np.zeros((n, H, W), np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    for d in range(n):
        if d % 4 == 3:
            continue                     # a cloud-free date exercises the sum()==0 guard
        for _ in range(r.integers(1, 5)):
            cy, cx, rad = r.integers(0, H), r.integers(0, W), r.integers(2, 14)
            m[d][(yy - cy) ** 2 + (xx - cx) ** 2 <= rad * rad] = 1
    return m
