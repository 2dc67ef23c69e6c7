"""Generate tests/golden/north.npz: outputs of the host-only functions of the NORTH-seam variant of the border pass
(/root/reference/src/resegment_tiles_north_wide.py, through oracle/refshim.py, this container only):
make_tiles_right_neighb :253, check_if_artifact :661, mosaic_subtiles / recreate_resegmented_tifs :1144-1522 (feather exponent
1.5), load_tif :703 (stub rasterio = Pillow).  The array functions of that file share their kernels with the east-seam file.
Usage: python tools/make_golden_north.py"""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util
spec = importlib.util.spec_from_file_location("mk_recreate", os.path.join(ROOT, "tools", "make_golden_recreate.py"))
MKR = importlib.util.module_from_spec(spec); spec.loader.exec_module(MKR)

ART_SIGMAS = [1, 4, 9, 15, 25, 2, 6, 12]
TIF_SETS = (["_POST"], ["_POST", "_FINAL"], ["_FINAL", "_SMOOTH_Y"], ["_FINAL", "_SMOOTH_X", "_SMOOTH_Y"], ["_FINAL", "_SMOOTH_X", "_SMOOTH_XY"])


def artifact_inputs():
    r = np.random.default_rng(17)
    for k, sg in enumerate(ART_SIGMAS):
        tile = r.uniform(0, 100, (40, 618 - 4 * (k % 3))).astype(np.float32)
        nb = tile[::-1, :610] + r.normal(0, sg, (40, 610)).astype(np.float32)
        if k == 5:
            nb += 8
        tile[r.random(tile.shape) < 0.02] = np.nan
        yield tile, nb


def main():
    from oracle import refshim
    from PIL import Image
    from sentinel_tree_cover_b200 import api
    m = refshim.ref("resegment_tiles_north_wide")
    out = {}
    m.SIZE, m.SIZE_X = 670, 206
    ta, tf = m.make_tiles_right_neighb(np.array([0, 138, 276, 412]), np.array([0]))
    out["tiles_array"], out["tiles_folder"] = np.asarray(ta, np.int64), np.asarray(tf, np.int64)
    m.x, m.y = 0, 0
    out["artifact"] = np.array([m.check_if_artifact(t, n) for t, n in artifact_inputs()], np.int64)
    print("artifact flags", out["artifact"].tolist())
    case = MKR.CASES[0]
    folder = tempfile.mkdtemp() + "/"
    MKR.write_case(folder, case)
    m.SIZE = case[1]
    preds, sums = m.recreate_resegmented_tifs(folder, case[2])
    st = MKR.SAMPLE[1]
    out["recreate_preds_sample"] = preds[::st[0], ::st[1]].copy(); out["recreate_sums_sample"] = sums[::st[0], ::st[1]].copy()

    class _DS:
        def __init__(self, f): self.f = f
        def read(self, band): return np.array(Image.open(self.f))
    m.rasterio.open = lambda f: _DS(f)
    r = np.random.default_rng(3)
    root = tempfile.mkdtemp()
    flags = []
    for case_i, names in enumerate(TIF_SETS):
        d = os.path.join(root, str(case_i), "7") + "/"
        os.makedirs(d)
        for s in names:
            api.write_tif(r.integers(0, 101, (40, 44)).astype(np.uint8), [10.0, 5.0, 10.1, 5.1], case_i, 7, d, s)
        _, flag = m.load_tif((str(case_i), "7"), root)
        flags.append(flag)
    out["load_tif_flags"] = np.array(flags, np.int64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "north.npz"), **out)


if __name__ == "__main__":
    main()
