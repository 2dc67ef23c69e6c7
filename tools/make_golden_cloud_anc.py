"""Generate tests/golden/cloud_masks_anc.npz: identify_clouds_shadows of the REFERENCE (through oracle/refshim.py, this
container only) WITH the two ancillary rasters it normally reads from urbanmask.tif / forestmask.tif
(cloud_removal.py:735-771).  rasterio is absent here, so the two loader functions are monkey-patched to do what they do
after the window read -- dilate, resize, combine -- on seeded synthetic rasters (oracle.cloud_ref.rasters_to_masks); every
other line of detect_pfcp (:1109-1212) and of the forest-threshold rules (:1412-1416, :1443-1447, :1546-1551) is the
reference's own code.  Usage: python tools/make_golden_cloud_anc.py"""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim, cloud_ref

# (T, H, W, seed, with_forest, with_urban)
CASES = [(9, 96, 104, 71, 1, 1), (12, 128, 120, 72, 1, 0), (6, 90, 92, 73, 0, 1), (24, 200, 208, 74, 1, 1), (5, 75, 81, 75, 1, 1)]


def synth_rasters(H, W, seed):
    """16x coarser 0/1 rasters (ESA WorldCover at 160 m against 10 m pixels)."""
    r = np.random.default_rng(seed + 500)
    h, w = max(2, H // 16 + 1), max(2, W // 16 + 1)
    forest = np.zeros((h, w), bool)
    forest[: max(1, h // 3), :] = True                              # a forested band + scattered stands
    forest |= r.random((h, w)) < 0.04
    urban = np.zeros((h, w), bool)
    urban[int(h * 0.2):int(h * 0.45) + 1, int(w * 0.5):int(w * 0.8) + 1] = True
    urban &= r.random((h, w)) < 0.7
    return forest, urban


def main():
    cr = refshim.ref("preprocessing.cloud_removal")
    os.chdir(tempfile.mkdtemp())
    out = {"cases": np.array(CASES, np.int32)}
    for i, (T, H, W, seed, wf, wu) in enumerate(CASES):
        img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed, urban=True)
        frst, urb = synth_rasters(H, W, seed)
        forest, core, near = cloud_ref.rasters_to_masks(frst if wf else None, urb if wu else None, (H, W))

        def mask_nonurban_areas(file, bbx, pfcps):
            if not wu:
                raise FileNotFoundError(file)
            pfcps[core == 1] = 1.
            pfcps[near == 0] = 0.
            return pfcps

        def adjust_cloudmask_in_forests(file, bbx, pf):
            if not wf:
                raise FileNotFoundError(file)
            return forest.astype(bool)
        cr.mask_nonurban_areas, cr.adjust_cloudmask_in_forests = mask_nonurban_areas, adjust_cloudmask_in_forests
        clouds, fcps = cr.identify_clouds_shadows(np.copy(img), np.copy(dem), None)
        out["clouds_%d" % i] = np.packbits(np.asarray(clouds) > 0)
        out["fcps_%d" % i] = np.packbits(np.asarray(fcps) > 0)
        print(i, (T, H, W), "forest %.2f urban core %.2f  clouds %.4f fcps %.4f" % (
            forest.mean() if wf else 0, core.mean() if wu else 0, np.mean(clouds), np.mean(fcps)), flush=True)
        o_c, o_f = cloud_ref.identify_clouds_shadows(img, dem, forest=forest, urban=(core, near) if wu else None)
        print("ORACLE == reference:", np.array_equal(o_c > 0, np.asarray(clouds) > 0), np.array_equal(np.asarray(o_f) > 0, np.asarray(fcps) > 0),
              "values:", np.unique(clouds), flush=True)
    path = os.path.join(ROOT, "tests", "golden", "cloud_masks_anc.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
