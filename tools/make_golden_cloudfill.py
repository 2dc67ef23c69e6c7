"""Generate tests/golden/cloud_fill.npz by running the REFERENCE
preprocessing.cloud_removal.remove_cloud_and_shadows (through oracle/refshim.py, this container only)
on seeded synthetic cubes with a pinned random.seed.  Inputs are regenerated from the seeds by the
tests (oracle.cloud_ref.synth_cloudy_cube + the oracle cloud mask).
Usage: python tools/make_golden_cloudfill.py   (run from any directory; uses a scratch CWD)"""
import os, sys, random, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim, cloud_ref

CASES = [(8, 96, 96, 41, 123), (5, 64, 64, 43, 7)]


def main():
    cr = refshim.ref("preprocessing.cloud_removal")
    os.chdir(tempfile.mkdtemp())          # the reference dumps debug .npy files into the CWD
    out = {"cases": np.array(CASES, np.int32)}
    for i, (T, H, W, seed, rseed) in enumerate(CASES):
        img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed)
        clouds, fcps = cloud_ref.identify_clouds_shadows(img, dem)
        random.seed(rseed)
        tiles, areas, to_remove = cr.remove_cloud_and_shadows(np.copy(img), np.copy(clouds), np.copy(clouds), list(range(T)),
                                                              np.copy(fcps), np.zeros((H, W, 2), np.float32))
        out["areas_%d" % i] = areas.astype(np.float32)
        changed = np.argwhere((tiles != img).any(-1))
        out["changed_count_%d" % i] = np.array([len(changed)])
        out["sample_idx_%d" % i] = changed[::3].astype(np.int32)
        out["sample_val_%d" % i] = tiles[tuple(changed[::3].T)].astype(np.float32)
        out["to_remove_%d" % i] = np.array(to_remove, np.int32)
        out["next_random_%d" % i] = np.array([random.random()])
        print(i, (T, H, W), "changed px", len(changed), "areas mean %.4f" % areas.mean(), to_remove)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cloud_fill.npz"), **out)


if __name__ == "__main__":
    main()
