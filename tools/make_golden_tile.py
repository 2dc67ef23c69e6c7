"""Generate tests/golden/process_tile.npz by running the REFERENCE process_tile
(src/download_and_predict_job.py:640-997) through oracle/refshim.py with hkl.load / os.path.exists
replaced by oracle.tile_ref.FakeStore over seeded synthetic raw tiles.  Usage: python tools/make_golden_tile.py"""
import os, sys, random, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim, tile_ref

CASES = [dict(seed=61, n=8, h=60, w=64, with_clm=False, ragged=False, rseed=9),
         dict(seed=62, n=7, h=57, w=61, with_clm=True, ragged=True, rseed=10)]


def main():
    job = refshim.ref("download_and_predict_job")
    os.chdir(tempfile.mkdtemp())
    out = {}
    real_exists = os.path.exists
    for i, c in enumerate(CASES):
        store = tile_ref.FakeStore(tile_ref.synth_raw_tile(c["seed"], c["n"], c["h"], c["w"], c["with_clm"], c["ragged"]))
        job.hkl.load = store.load
        job.os.path.exists = lambda p, s=store: s.exists(p) or real_exists(p)
        random.seed(c["rseed"])
        try:
            s2, dates, interp, s1, dem, cloudshad, snow = job.process_tile(1, 2, None, "/nonexistent/", [0, 0, 1, 1], make_shadow=True)
        finally:
            job.os.path.exists = real_exists
        out["case_%d" % i] = np.array([c["seed"], c["n"], c["h"], c["w"], int(c["with_clm"]), int(c["ragged"]), c["rseed"]], np.int32)
        out["s2_sub_%d" % i] = s2[:, ::3, ::3].astype(np.float32)
        out["s2_sum_%d" % i] = np.array([np.sum(s2, dtype=np.float64)])
        out["dates_%d" % i] = np.asarray(dates)
        out["interp_%d" % i] = interp.astype(np.float32)
        out["s1_sub_%d" % i] = s1[:, ::2, ::2].astype(np.float32)
        out["dem_%d" % i] = np.asarray(dem, np.float32)
        out["cloudshad_%d" % i] = np.packbits(np.asarray(cloudshad) > 0)
        out["snow_%d" % i] = np.asarray(snow).astype(np.int8)
        out["next_random_%d" % i] = np.array([random.random()])
        print(i, s2.shape, s2.dtype, dates, interp.shape, s1.shape, dem.shape, dem.dtype, cloudshad.shape, snow.dtype)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "process_tile.npz"), **out)


if __name__ == "__main__":
    main()
