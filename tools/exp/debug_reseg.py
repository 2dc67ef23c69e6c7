import os, sys, random, importlib.util
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cloud_ref, cloudfill_ref as F
from sentinel_tree_cover_b200 import api, resegment as R
G = np.load(os.path.join(ROOT, "tests/golden/resegment.npz"))
spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tools/make_golden_resegment.py")); MK = importlib.util.module_from_spec(spec); spec.loader.exec_module(MK)
g = os.path.join(ROOT, "tests/golden/")
sess = api.StcSession(0, predict_weights=g + "weights_predict_172.npz", superresolve_weights=g + "weights_superresolve.npz")
for i, (T, H, W, seed, rseed, with_clm) in enumerate(MK.PRE_CASES):
    img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed)
    dts = (np.arange(T) * (330 // T) + 10).astype(np.int64)
    clm = None
    if with_clm:
        clm = np.zeros((T, H, W), np.float32); clm[2, 10:30, 20:50] = 1.
    random.seed(rseed)
    arr, interp2, d2 = R.preprocess_tile(np.copy(img), np.copy(dts), None, clm.copy() if clm is not None else None, "tile", np.copy(dem), None, sess)
    nxt = random.random()
    idx = G["pre_idx_%d" % i]; val = G["pre_val_%d" % i]
    got = np.asarray(arr)[tuple(idx.T)]
    print("case", i, "next_random equal:", nxt == float(G["pre_next_random_%d" % i][0]))
    for t in range(arr.shape[0]):
        sel = idx[:, 0] == t
        if sel.any():
            d = np.abs(got[sel] - val[sel])
            print("  date", t, "n", int(sel.sum()), "max abs", float(d.max()), "frac>1e-5", float((d > 1e-5).mean()), "per band max", np.round(d.max(0), 6))
    # the oracle restatement on the same inputs
    cld, fcps = cloud_ref.identify_clouds_shadows(img, dem)
    if clm is not None:
        c2 = clm.copy(); c2[np.asarray(fcps) > 0] = 0; cld = np.maximum(c2, cld)
    random.seed(rseed)
    o_tiles, o_areas, o_rm = F.remove_cloud_and_shadows(np.copy(img), np.copy(cld), np.copy(fcps))
    o_got = o_tiles[tuple(idx.T)]
    print("  oracle vs golden max", float(np.abs(o_got - val).max()), " oracle next_random equal", random.random() == float(G["pre_next_random_%d" % i][0]))
    print("  gpu vs oracle max", float(np.abs(np.asarray(arr) - o_tiles).max()))
    # stage-level comparison with the oracle on the same masks
    rec = []
    orig_si = F.sample_indices
    def spy(evi, n_rows):
        rec.append([float(np.percentile(evi, q)) for q in (2, 20, 40, 60, 80, 98)] + [n_rows, float(evi[:3].sum())])
        return orig_si(evi, n_rows)
    F.sample_indices = spy
    taps = {}
    random.seed(rseed)
    o_tiles, o_areas, o_rm = F.remove_cloud_and_shadows(np.copy(img), np.copy(cld), np.copy(fcps), taps)
    random.seed(rseed)
    state = np.array(random.getstate()[1], dtype=np.uint32)
    tiles = np.copy(img)
    os.environ["STC_CF_DEBUG"] = "1"
    sys.stderr.flush()
    areas, rm, mosaic = sess.remove_clouds(tiles, cld, fcps, state, want_mosaic=True)
    del os.environ["STC_CF_DEBUG"]
    F.sample_indices = orig_si
    for k, rr in enumerate(rec):
        print("  [oracle] fit", k, "evi percentiles", " ".join("%.9g" % v for v in rr[:6]), "n_rows", rr[6])
    for dd in sorted(taps["coef"].keys()):
        print("  [oracle] date", dd, "sample", taps["sample"][dd][:6].tolist(), "len", len(taps["sample"][dd]))
        for b in (0, 8):
            print("  [oracle]   band", b, "coef", " ".join("%.6g" % v for v in taps["coef"][dd][b]))
    dm = np.abs(mosaic - taps["mosaic"])
    print("  mosaic: n diff px", int((dm > 0).any(-1).sum()), "max", float(np.nanmax(dm)), "areas equal", np.array_equal(areas, o_areas))
    bad = np.argwhere((dm > 0).any(-1))
    print("  first bad mosaic px", bad[:5].tolist(), "nan in oracle mosaic", int(np.isnan(taps["mosaic"]).sum()), "nan gpu", int(np.isnan(mosaic).sum()))
    for t in range(tiles.shape[0]):
        d = np.abs(tiles[t] - o_tiles[t])
        print("   date", t, "max", float(d.max()), "n px", int((d > 1e-6).any(-1).sum()), "interp>0 px", int((o_areas[t] > 0).sum()))
