"""Batch sweep of the batchable preprocessing kernels with device-resident inputs (BASELINE configs[4], the part of it
that has a batch axis): K1 regrid + Whittaker + monthly mean as one 12 x n operator (n = 24), the quarterly / annual
median assembly, and the DSen2 super-resolution, for B = 1 .. 4096 patches of 168 x 168 (super-resolution: B x 12 dates
capped at 3072 images).  Per line: ms per call (CUDA events), SURVEY 8d algorithmic bytes, GB/s and fraction of the
measured HBM peak; the super-resolution line also gives TFLOP/s (it is the tensor-bound one: 82,944 FLOP per pixel).
The cloud pipeline (masks / removal) works on one tile of n dates at a time and has no batch axis: tools/bench_preproc.py.
Usage (GPU box): python tools/bench_sweep.py [--max 4096]"""
import argparse, ctypes, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max", type=int, default=4096)
    a = ap.parse_args()
    from sentinel_tree_cover_b200.api import StcSession
    from sentinel_tree_cover_b200.weights import random_superresolve_weights
    from sentinel_tree_cover_b200 import regrid
    peak = 6541.1
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    sess = StcSession(0, superresolve_weights=random_superresolve_weights(0))
    H = 168
    px = H * H
    dates = np.arange(24) * 15 + 7
    M = np.ascontiguousarray(regrid.monthly_operator(dates)[0], np.float32)
    r = np.random.default_rng(0)

    def timed(fn, reps):
        fn(); sess.sync()
        sess.timer_begin()
        for _ in range(reps):
            fn()
        return sess.timer_end() / reps

    for B in (1, 4, 16, 64, 256, 1024, 4096):
        if B > a.max:
            break
        reps = 20 if B <= 64 else 5
        LIMIT = 120e9         # bytes of buffers per test (180 GB of HBM): a larger batch streams through the same buffer in parts
        # ---- K1: [24, B*px, 14] -> [12, B*px, 14]
        parts = 1
        while 36 * (B // parts) * px * 14 * 4 > LIMIT:
            parts *= 2
        inner = (B // parts) * px * 14
        d_in = sess.malloc(24 * inner * 4); d_out = sess.malloc(12 * inner * 4)

        def k1():
            for _ in range(parts):
                sess._check(sess.lib.stc_temporal_matmul_dev(sess.h, d_in, M.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 24, 12, inner, d_out))
        ms = timed(k1, reps)
        by = (24 + 12) * inner * 4 * parts
        row = {"kernel": "K1 regrid+Whittaker+monthly (12x24 operator)", "B": B, "ms": round(ms, 4), "algorithmic_MB": round(by / 1e6, 1),
               "GBps": round(by / ms / 1e6, 1), "frac_hbm": round(by / ms / 1e6 / peak, 3)}
        if parts > 1:
            row["note"] = "%d launches over a %d-patch buffer (the whole cube would be %.0f GB)" % (parts, B // parts, by / 1e9)
        print(json.dumps(row), flush=True)
        sess.free(d_in); sess.free(d_out)
        # ---- assemble: [B,12,H,W,13] -> [B,5,H,W,17]
        if B * px * (12 * 13 + 5 * 17) * 4 > LIMIT:
            print(json.dumps({"kernel": "assemble (quarterly/annual medians + indices)", "B": B, "skipped": "buffers exceed %d GB" % (LIMIT / 1e9)}), flush=True)
            continue
        reps = min(reps, 3) if B >= 4096 else reps
        d_in = sess.malloc(B * 12 * px * 13 * 4); d_out = sess.malloc(B * 5 * px * 17 * 4)
        ms = timed(lambda: sess._check(sess.lib.stc_assemble_dev(sess.h, d_in, B, H, H, d_out)), reps)
        by = B * px * (12 * 13 + 5 * 17) * 4
        print(json.dumps({"kernel": "assemble (quarterly/annual medians + indices)", "B": B, "ms": round(ms, 4), "algorithmic_MB": round(by / 1e6, 1),
                          "GBps": round(by / ms / 1e6, 1), "frac_hbm": round(by / ms / 1e6 / peak, 3)}), flush=True)
        sess.free(d_in); sess.free(d_out)
        # ---- super-resolution: N = min(12*B, 3072) images of 168 x 168 x 10
        N = min(12 * B, 3072)
        x = r.uniform(0, 0.6, (min(N, 64), H, H, 10)).astype(np.float32)
        d_in = sess.malloc(N * px * 40); d_out = sess.malloc(N * px * 24)
        for k in range(0, N, x.shape[0]):
            n = min(x.shape[0], N - k)
            sess.h2d(ctypes.c_void_p(d_in.value + k * px * 40), x[:n])
        sess.sync()
        ms = timed(lambda: sess._check(sess.lib.stc_superresolve_dev(sess.h, d_in, None, N, H, H, d_out)), max(2, reps // 2))
        by = N * px * (40 + 24)
        print(json.dumps({"kernel": "DSen2 super-resolution", "B": B, "images": N, "ms": round(ms, 4), "algorithmic_MB": round(by / 1e6, 1),
                          "GBps": round(by / ms / 1e6, 1), "frac_hbm": round(by / ms / 1e6 / peak, 3),
                          "TFLOPs": round(82944.0 * N * px / ms / 1e9, 1)}), flush=True)
        sess.free(d_in); sess.free(d_out)
    sess.close()


if __name__ == "__main__":
    main()
