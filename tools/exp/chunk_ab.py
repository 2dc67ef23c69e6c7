"""A/B of the forward's sub-batch size (STC_CHUNK) on the 36 subtile patches of a chain tile; warm calls, one process."""
import os, random, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import tile_ref
from sentinel_tree_cover_b200 import api

gold = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden")
sess = api.StcSession(0, predict_weights=os.path.join(gold, "weights_predict_172.npz"), superresolve_weights=os.path.join(gold, "weights_superresolve.npz"))
raw = tile_ref.synth_raw_tile(91, n=12, h=309, w=309)
pin = {k: sess.pinned_empty(raw[k].shape, raw[k].dtype) for k in ("s2_10", "s2_20", "s1", "dem")}
for k in pin:
    pin[k][...] = raw[k]
ref = None
for ch in (32, 36, 18, 12, 9, 32):
    os.environ["STC_CHUNK"] = str(ch)
    ts = []
    for rep in range(5):
        random.seed(4)
        t0 = time.perf_counter()
        out, kept = sess.run_tile(pin["s2_10"], pin["s2_20"], pin["s1"], pin["dem"], raw["s2_dates"])
        ts.append((time.perf_counter() - t0) * 1e3)
    if ref is None:
        ref = out
    print("STC_CHUNK=%d  chain ms %s  same tile: %s" % (ch, ["%.1f" % t for t in ts], bool(np.array_equal(out, ref))), flush=True)
