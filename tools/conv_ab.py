"""A/B timing of the conv kernel variants at the bench shapes (168x168 tiles), one process per variant
because the switches (STC_CONV_V / STC_CONV_ISS / STC_CONV_WRES / STC_SINGLE_STREAM) are read once.

  python tools/conv_ab.py [--batch 128] [--steps 3] [--tag name]

Prints one JSON line: step ms, per-kind conv launch averages (CUDA events on the launching stream) and a
checksum of the probabilities so that variants can be compared for equality.  Not a bench line."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KINDS = {"gates64": (64, 16, 0), "cand32": (32, 8, 3), "n64_pscale": (64, 8, 1), "n64_swish": (64, 8, 2),
         "n128_pscale": (128, 8, 1), "n128_swish": (128, 8, 2), "n256_swish": (256, 8, 2)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--tag", default="")
    ap.add_argument("--trace", default="", help="write a kernel timeline CSV of one extra step")
    a = ap.parse_args()
    from sentinel_tree_cover_b200.api import StcSession
    from sentinel_tree_cover_b200.weights import random_predict_weights
    from sentinel_tree_cover_b200 import synth as P
    H = 168
    sess = StcSession(0, predict_weights=random_predict_weights(0))
    base = P.synth_monthly(8, H, 2000)
    host = np.empty((a.batch, 12, H, H, 13), np.float32)
    for i in range(a.batch):
        host[i] = base[i % 8]
    out = np.empty((a.batch, H - 14, H - 14), np.float32)
    d_in = sess.malloc(host.nbytes); d_out = sess.malloc(out.nbytes)
    sess.h2d(d_in, host); sess.sync()
    for _ in range(2):
        sess.predict_patches_dev(d_in, a.batch, H, H, d_out)
    sess.sync()
    sess.conv_timing(1)
    sess.timer_begin()
    for _ in range(a.steps):
        sess.predict_patches_dev(d_in, a.batch, H, H, d_out)
    ms = sess.timer_end()
    res = {"tag": a.tag, "env": {k: v for k, v in os.environ.items() if k.startswith("STC_")}, "batch": a.batch,
           "ms_per_step": ms / a.steps, "tiles_per_s": a.batch * a.steps / (ms / 1000.0)}
    for name, (n, g, m) in KINDS.items():
        t, c = sess.conv_timing_kind(n, g, m)
        if c:
            res[name] = {"avg_us": 1000.0 * t / c, "n": c, "ms_per_step": t / a.steps}
    t, c = sess.conv_timing(0)
    res["conv_ms_per_step"] = t / a.steps
    if a.trace:
        sess.sync()
        sess.trace(1)
        sess.predict_patches_dev(d_in, a.batch, H, H, d_out)
        sess.trace(0, a.trace)
    sess.d2h(out, d_out); sess.sync()
    res["checksum"] = float(out.astype(np.float64).sum())
    res["max"] = float(out.max())
    print(json.dumps(res), flush=True)
    sess.free(d_in); sess.free(d_out); sess.close()


if __name__ == "__main__":
    main()
