"""Generate tests/golden/resegment.npz: outputs of the REFERENCE's border re-segmentation functions
(/root/reference/src/resegment_tiles_wide.py, through oracle/refshim.py, this container only) on seeded inputs.
  align_dates :242, make_tiles_right_neighb :267, check_if_artifact :675, align_subtile_histograms :284,
  adjust_predictions :348, regularize_and_smooth :772, preprocess_tile :619 (random.seed pinned).
Usage: python tools/make_golden_resegment.py"""
import os, sys, random, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim, cloud_ref

DATE_SETS = [([10, 40, 70, 100, 130, 160], [10, 41, 75, 100, 131, 200]), ([5, 5, 30, 60, 90], [6, 30, 30, 61, 120, 150]),
             ([15, 45, 75], [200, 230]), ([0, 20, 40, 60, 80, 100, 120], [0, 20, 40, 60, 80, 100, 120])]
HIST_CASES = [(4, 60, 80, 14, 81), (1, 50, 64, 14, 82), (4, 72, 100, 14, 83)]     # (T, H, SIZE, C, seed): W = SIZE + 14
PRE_CASES = [(8, 96, 104, 91, 7, 0), (6, 80, 90, 92, 8, 1)]                        # (T, H, W, seed, random seed, with clm)


def hist_input(T, H, size, C, seed):
    """Two halves with different gain / offset and a lake, like the two sides of a real seam."""
    r = np.random.default_rng(seed)
    W = size + 14
    base = r.uniform(0.05, 0.4, (1, H, W, C)) + 0.03 * r.standard_normal((T, H, W, C))
    base[:, :, W // 2:, :] = base[:, :, W // 2:, :] * r.uniform(0.8, 1.25, (T, 1, 1, C)) + r.uniform(-0.03, 0.03, (T, 1, 1, C))
    base[:, H // 3: H // 2, 5:25, 1] = 0.2; base[:, H // 3: H // 2, 5:25, 3] = 0.05        # water: NDWI > 0.1
    if T > 1:
        base[1, 3, 4, 2] = np.nan
    return base.astype(np.float32)


def main():
    m = refshim.ref("resegment_tiles_wide")
    os.chdir(tempfile.mkdtemp())
    out = {}
    for i, (a, b) in enumerate(DATE_SETS):
        rt, rn, mn = m.align_dates(list(a), list(b))
        out["dates_%d" % i] = np.array([len(rt), len(rn), mn] + [int(v) for v in rt] + [int(v) for v in rn], np.int64)
    m.SIZE, m.SIZE_Y = 670, 206
    ta, tf = m.make_tiles_right_neighb(np.array([0]), np.array([0, 138, 276, 412]))
    out["tiles_array"], out["tiles_folder"] = np.asarray(ta, np.int64), np.asarray(tf, np.int64)
    m.x, m.y = 0, 0
    r = np.random.default_rng(3)
    arts = []
    for k in range(6):
        tile = r.uniform(0, 100, (618, 40)).astype(np.float32); nb = tile[:, ::-1] + r.normal(0, [1, 4, 9, 15, 25, 2][k], tile.shape).astype(np.float32)
        if k == 5:
            nb += 8
        tile[r.random(tile.shape) < 0.02] = np.nan
        arts.append(m.check_if_artifact(tile, nb))
    out["artifact"] = np.array(arts, np.int64)
    for i, (T, H, size, C, seed) in enumerate(HIST_CASES):
        m.SIZE = size
        x = hist_input(T, H, size, C, seed)
        y = m.align_subtile_histograms(np.copy(x))
        out["hist_%d" % i] = y.astype(np.float32)
        print("hist", i, "changed time steps:", [bool((y[t] != x[t]).any()) for t in range(T)], flush=True)
    p = r.uniform(0, 1, (40, 50)).astype(np.float32); ref = r.uniform(0.2, 0.9, (40, 50)).astype(np.float32); ref[3, 4] = np.nan
    out["adjust"] = np.asarray(m.adjust_predictions(np.copy(p), ref), np.float32)
    img, _ = cloud_ref.synth_cloudy_cube(9, 40, 44, 95)
    dates = np.array([12, 40, 75, 101, 140, 170, 220, 260, 320])
    out["regsmooth"] = np.asarray(m.regularize_and_smooth(np.copy(img), dates), np.float32)
    for i, (T, H, W, seed, rseed, with_clm) in enumerate(PRE_CASES):
        img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed)
        dts = (np.arange(T) * (330 // T) + 10).astype(np.int64)
        clm = None
        if with_clm:
            clm = np.zeros((T, H, W), np.float32); clm[2, 10:30, 20:50] = 1.
        random.seed(rseed)
        arr, interp2, d2 = m.preprocess_tile(np.copy(img), np.copy(dts), None, clm, "tile", np.copy(dem), None)
        out["pre_dates_%d" % i] = np.asarray(d2, np.int64)
        out["pre_interp_%d" % i] = np.asarray(interp2, np.float32).astype(np.float16)
        out["pre_arr_sum_%d" % i] = np.array([float(np.asarray(arr, np.float64).sum())])
        idx = np.argwhere((np.asarray(arr) != img[: arr.shape[0]]).any(-1)) if arr.shape == img.shape else np.zeros((0, 3), np.int64)
        step = max(1, len(idx) // 5000)
        out["pre_idx_%d" % i] = idx[::step].astype(np.int32)
        out["pre_val_%d" % i] = np.asarray(arr)[tuple(idx[::step].T)].astype(np.float32)
        out["pre_next_random_%d" % i] = np.array([random.random()])
        print("pre", i, arr.shape, "changed px", len(idx), flush=True)
    path = os.path.join(ROOT, "tests", "golden", "resegment.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
