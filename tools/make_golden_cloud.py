"""Generate tests/golden/cloud_masks.npz by running the REFERENCE function
preprocessing.cloud_removal.identify_clouds_shadows (through oracle/refshim.py, in this container only)
on seeded synthetic cubes (oracle.cloud_ref.synth_cloudy_cube).  The fixture stores the seeds / shapes
and the reference outputs as packed bits, so the inputs are regenerated, not stored.
Usage: python tools/make_golden_cloud.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim, cloud_ref

CASES = [(9, 64, 72, 11), (12, 80, 80, 12), (5, 48, 48, 13), (3, 40, 56, 14), (2, 32, 32, 15)]


def main():
    cr = refshim.ref("preprocessing.cloud_removal")
    out = {"cases": np.array(CASES, np.int32)}
    for i, (T, H, W, seed) in enumerate(CASES):
        img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed)
        clouds, fcps = cr.identify_clouds_shadows(np.copy(img), np.copy(dem), None)
        out["clouds_%d" % i] = np.packbits(np.asarray(clouds) > 0)
        out["fcps_%d" % i] = np.packbits(np.asarray(fcps) > 0)
        vals = np.unique(np.asarray(clouds))
        assert set(vals.tolist()) <= {0.0, 1.0}, vals
        print(i, (T, H, W), "cloud frac %.3f" % np.mean(clouds), "fcps frac %.3f" % np.mean(fcps))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cloud_masks.npz"), **out)


if __name__ == "__main__":
    main()
