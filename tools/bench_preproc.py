"""Timing of the preprocessing chain (n = 24 dates with cloud blobs, one tile)
through the host-buffer C-ABI: cloud masks (P5) -> feather (P6) -> cloud removal (P7) -> missing-pixel
fill (P4) -> indices + regrid/Whittaker/monthly (P8-P11), plus the 20 m -> 10 m band stack (P3).
Each line: wall ms per call (H2D + kernels + D2H, the call a user makes), algorithmic bytes
and the implied GB/s; `--cpu` adds the oracle (NumPy/SciPy restatement of the reference) on a bounded
crop for context.  One JSON object per line on stdout.
Usage (GPU box): python tools/bench_preproc.py [--n 24] [--size 620] [--reps 3] [--cpu]"""
import argparse, json, os, random, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=24)
    ap.add_argument("--size", type=int, default=620)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    from oracle import cloud_ref          # synthetic cube generator (and the CPU context arm)
    from sentinel_tree_cover_b200 import api
    sess = api.StcSession(0)
    n, S = args.n, args.size
    img, dem = cloud_ref.synth_cloudy_cube(n, S, S, 77)
    px = n * S * S
    dates = np.arange(n) * 15 + 7

    def timed(name, fn, alg_bytes, reps=args.reps):
        fn()                                          # warm-up (allocations, first launch)
        t = []
        for _ in range(reps):
            sess.sync(); t0 = time.perf_counter(); out = fn(); sess.sync(); t.append((time.perf_counter() - t0) * 1e3)
        ms = float(np.median(t))
        print(json.dumps({"stage": name, "ms": round(ms, 3), "algorithmic_MB": round(alg_bytes / 1e6, 1),
                          "GBps_e2e": round(alg_bytes / ms / 1e6, 1), "n": n, "size": S}), flush=True)
        return out

    clouds, fcps = timed("P5 identify_clouds_shadows", lambda: sess.cloud_masks(img, dem), px * 40 + px * 5)
    timed("P6 id_areas_to_interp (feather 15)", lambda: sess.feather(clouds, 15), px * 8)
    state = np.array(random.Random(1).getstate()[1], dtype=np.uint32)

    def p7():
        t = img.copy()
        return sess.remove_clouds(t, clouds, fcps, state.copy())
    timed("P7 remove_cloud_and_shadows", p7, px * 44 + px * 40)
    s20 = np.ascontiguousarray(img[:, ::2, ::2, 4:10])
    timed("P3 build_sentinel2 (20m->10m)", lambda: sess.build_sentinel2(img[..., :4], s20), px * 16 + px * 6 + px * 40)
    interp = np.zeros((n, S, S), np.float32)
    timed("P4 deal_w_missing_px", lambda: api.deal_w_missing_px(img.copy(), dates.copy(), interp, sess), px * 80)
    timed("P8-P11 smooth_large_tile", lambda: api.smooth_large_tile(img.copy(), dates.copy(), interp, sess), px * 40 + 12 * S * S * 56)
    if args.cpu:
        from oracle import cloudfill_ref
        c = 256
        sub, dsub = np.ascontiguousarray(img[:8, :c, :c]), np.ascontiguousarray(dem[:c, :c])
        t0 = time.perf_counter(); cm, fc = cloud_ref.identify_clouds_shadows(sub, dsub); t1 = time.perf_counter()
        random.seed(1)
        cloudfill_ref.remove_cloud_and_shadows(sub.copy(), cm, fc); t2 = time.perf_counter()
        scale = px / (8 * c * c)
        print(json.dumps({"stage": "cpu oracle (8 x 256^2 crop, scaled to the full cube)", "P5_ms": round((t1 - t0) * 1e3 * scale, 1),
                          "P7_ms": round((t2 - t1) * 1e3 * scale, 1), "measured_crop_ms": [round((t1 - t0) * 1e3, 1), round((t2 - t1) * 1e3, 1)]}), flush=True)


if __name__ == "__main__":
    main()
