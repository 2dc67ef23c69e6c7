"""Generate tests/golden/large_t24.npz: outputs of the REFERENCE functions (through oracle/refshim.py, this container only)
on the series lengths and tile sizes the production job runs at -- BASELINE configs[2] is a 24-step series and the
reference's tiles are ~618 x 618 px -- because the date-window logic of identify_clouds_shadows
(cloud_removal.py:1266-1273, 1352-1363) depends on T and the small fixtures stop at T = 12 / 150 px.
  masks_*  identify_clouds_shadows(img, dem, bbx)            packed bits of clouds and fcps
  fill_*   remove_cloud_and_shadows(...) with random.seed    feather weights (subsampled + checksum), changed pixels
                                                             (subsampled), removal list, next random number
Inputs are regenerated from seeds (oracle.cloud_ref.synth_cloudy_cube).  Usage: python tools/make_golden_large.py [--quick]"""
import os, sys, random, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim, cloud_ref

MASK_CASES = [(24, 256, 256, 61), (24, 618, 618, 62), (17, 300, 280, 63)]
FILL_CASES = [(24, 256, 256, 61, 11), (12, 618, 618, 64, 12)]


def main():
    quick = "--quick" in sys.argv
    cr = refshim.ref("preprocessing.cloud_removal")
    os.chdir(tempfile.mkdtemp())          # the reference dumps debug .npy files into the CWD
    out = {"mask_cases": np.array(MASK_CASES, np.int32), "fill_cases": np.array(FILL_CASES, np.int32)}
    for i, (T, H, W, seed) in enumerate(MASK_CASES):
        if quick and H > 300:
            continue
        t0 = time.time()
        img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed)
        clouds, fcps = cr.identify_clouds_shadows(np.copy(img), np.copy(dem), None)
        out["masks_clouds_%d" % i] = np.packbits(np.asarray(clouds) > 0)
        out["masks_fcps_%d" % i] = np.packbits(np.asarray(fcps) > 0)
        print("masks", i, (T, H, W), "cloud frac %.3f fcps frac %.3f  %.1f s" % (np.mean(clouds), np.mean(fcps), time.time() - t0), flush=True)
    for i, (T, H, W, seed, rseed) in enumerate(FILL_CASES):
        if quick and H > 300:
            continue
        t0 = time.time()
        img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed)
        clouds, fcps = cr.identify_clouds_shadows(np.copy(img), np.copy(dem), None)
        random.seed(rseed)
        tiles, areas, to_remove = cr.remove_cloud_and_shadows(np.copy(img), np.copy(clouds), np.copy(clouds), list(range(T)),
                                                              np.copy(fcps), np.zeros((H, W, 2), np.float32))
        areas = np.asarray(areas, np.float32)
        out["fill_areas_sub_%d" % i] = areas[:, ::5, ::5].astype(np.float16)          # values are multiples of 1/12 in [0, 1]
        out["fill_areas_sum_%d" % i] = np.array([float(areas.sum(dtype=np.float64)), float((areas > 0).sum()), float((areas == 1).sum())])
        changed = np.argwhere((tiles != img).any(-1))
        out["fill_changed_count_%d" % i] = np.array([len(changed)])
        step = max(1, len(changed) // 20000)
        out["fill_sample_idx_%d" % i] = changed[::step].astype(np.int32)
        out["fill_sample_val_%d" % i] = tiles[tuple(changed[::step].T)].astype(np.float32)
        out["fill_tiles_sum_%d" % i] = np.array([float(tiles.sum(dtype=np.float64))])
        out["fill_to_remove_%d" % i] = np.array(to_remove, np.int32)
        out["fill_next_random_%d" % i] = np.array([random.random()])
        print("fill", i, (T, H, W), "changed px", len(changed), "areas mean %.4f" % areas.mean(), to_remove, "%.1f s" % (time.time() - t0), flush=True)
    path = os.path.join(ROOT, "tests", "golden", "large_t24_quick.npz" if quick else "large_t24.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
