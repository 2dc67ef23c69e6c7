"""Per-stage mismatch report of the CUDA cloud-mask pipeline against oracle/cloud_ref.py (GPU box)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cloud_ref
from sentinel_tree_cover_b200.api import StcSession

sess = StcSession(0)
for (T, H, W, seed) in [(9, 64, 72, 11), (12, 80, 80, 12), (3, 40, 56, 14), (2, 32, 32, 15), (7, 150, 130, 24)]:
    img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed)
    c0, f0, st = cloud_ref.identify_clouds_shadows(img, dem, stages=True)
    print("case", (T, H, W, seed))
    for name in sess.CLOUD_STAGES:
        c, f, tap = sess.cloud_masks(img, dem, stage=name)
        want = np.asarray(st[name]) > 0
        bad = (tap > 0) != want
        per_date = bad.reshape(T, -1).sum(1)
        msg = ""
        if bad.any():
            t, y, x = np.argwhere(bad)[0]
            msg = " first (t,y,x)=%s got %d want %d" % ((t, y, x), tap[t, y, x], want[t, y, x])
        print("  %-16s bad %6d  per-date %s%s" % (name, bad.sum(), per_date.tolist(), msg))
    print("  final clouds bad", int(((c > 0) != (np.asarray(c0) > 0)).sum()), "fcps bad", int((f != (np.asarray(f0) > 0)).sum()))
