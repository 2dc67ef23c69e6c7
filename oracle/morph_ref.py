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


def identify_bright_bare_surfaces(img):
    """src/download_and_predict_job.py:1099-1122 (SciPy restatement)."""
    BLUE, RED, NIR = np.clip(img[..., 0], 0, 1), np.clip(img[..., 2], 0, 1), np.clip(img[..., 3], 0, 1)
    evis = np.clip(2.5 * ((NIR - RED) / (NIR + (6 * RED) - (7.5 * BLUE) + 1)), -1.5, 1.5)
    c = (img[..., 3] / (img[..., 8] + 0.01)) < 0.9
    c = c * (np.mean(img[..., :3], axis=-1) > 0.2)
    c = c * (evis < 0.3)
    bs = np.sum(c, axis=0) > 1
    bs = binary_dilation(1 - bs, iterations=2)
    bs = binary_dilation(1 - bs, iterations=1)
    blurred = distance(1 - bs)
    blurred[blurred > 3] = 3
    return (blurred / 3)[7:-7, 7:-7]


def postprocess_subtile(preds, subtile_all, min_clear, size=158):
    """src/download_and_predict_job.py:1408-1409,1451-1483 (size 158 branch)."""
    bright = identify_bright_bare_surfaces(subtile_all)
    preds = np.array(preds, copy=True)
    st2 = generate_binary_structure(2, 2)
    no_images = min_clear[6:-6, 6:-6] < 1
    no_images = 1 - binary_dilation(1 - no_images, structure=st2, iterations=6)
    no_images = binary_dilation(no_images, structure=st2, iterations=6)
    no_images = np.reshape(no_images, (4, 40, 4, 40))
    no_images = np.sum(no_images, axis=(1, 3)) > (40 * 40) * 0.25
    no_images = no_images.repeat(40, axis=0).repeat(40, axis=1)[1:-1, 1:-1]
    preds[no_images] = 255.
    preds = np.around(preds * bright, 3)
    return preds.astype(np.float32)


def synth_subtile_stack(seed, size=158):
    """(5, size+14, size+14, 17) un-normalised frame stack with a few bright bare patches."""
    r = np.random.default_rng(seed)
    S = size + 14
    x = r.uniform(0.02, 0.3, (5, S, S, 17)).astype(np.float32)
    x[..., 3] = r.uniform(0.25, 0.5, (5, S, S))          # NIR
    x[..., 8] = r.uniform(0.1, 0.25, (5, S, S))          # SWIR
    yy, xx = np.mgrid[0:S, 0:S]
    for _ in range(4):
        cy, cx, rad = r.integers(10, S - 10), r.integers(10, S - 10), r.integers(3, 12)
        blob = (yy - cy) ** 2 + (xx - cx) ** 2 <= rad * rad
        x[:, blob, 0:3] = r.uniform(0.25, 0.4)           # bright RGB
        x[:, blob, 3] = 0.2; x[:, blob, 8] = 0.4         # NIR/SWIR < 0.9
    clear = r.integers(0, 4, (S, S))
    clear[40:110, 20:90] = 0                             # a region without cloud-free images
    return x, clear
