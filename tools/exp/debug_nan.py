import os, sys, importlib.util, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import subtiles_ref
from sentinel_tree_cover_b200 import api
from sentinel_tree_cover_b200.tile import process_subtiles
spec = importlib.util.spec_from_file_location("mk_sub", os.path.join(ROOT, "tools", "make_golden_subtiles.py")); mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
g = np.load(os.path.join(ROOT, "tests/golden/process_subtiles_nan.npz"))
gd = os.path.join(ROOT, "tests/golden/")
sess = api.StcSession(0, predict_weights=gd + "weights_predict_172.npz", superresolve_weights=gd + "weights_superresolve.npz")
seed, n, H, W = [int(v) for v in g["case"]]
for variant in ("nan", "clean"):
    s2, dates, interp, s1, dem = subtiles_ref.synth_ard(seed, n, H, W)
    if variant == "nan":
        s2 = mk.add_nans(s2)
    root = tempfile.mkdtemp() + "/"
    process_subtiles(3, 4, s2, dates, interp, s1, dem, sess, [0, 0, 1, 1], 158, None, local_path=root, length=4)
    if variant != "nan":
        break
    path = root + "3/4/processed/"
    tot = 0
    for fy, fx in sorted(map(tuple, g["names"].tolist())):
        got = np.load(f"{path}{fy}/{fx}.npy"); want = g["pred_%d_%d" % (fy, fx)]
        m = want < 2
        d = np.abs(got - want) * m
        bad = np.argwhere(d >= 0.0015)
        tot += len(bad)
        if len(bad):
            print("subtile", (fy, fx), "n bad", len(bad), "max", float(d.max()), "rows", bad[:, 0].min(), bad[:, 0].max(), "cols", bad[:, 1].min(), bad[:, 1].max(),
                  "sample got/want", got[tuple(bad[0])], want[tuple(bad[0])])
    print("total bad px", tot)
