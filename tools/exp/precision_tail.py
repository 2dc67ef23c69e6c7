"""Tail of |GPU - f32 oracle| on the probability map over many patches (released weights): max, quantiles, count > 1e-3."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.model_ref import PredictRef
from oracle import preproc_ref as P, subtiles_ref
from sentinel_tree_cover_b200 import api
from sentinel_tree_cover_b200.weights import load_npz
gd = os.path.join(ROOT, "tests/golden/")
w = load_npz(gd + "weights_predict_172.npz")
sess = api.StcSession(0, predict_weights=w)
ref = PredictRef(w)
errs = []
for seed in range(6):
    # realistic inputs: cloudy-cube spectra (vegetation / soil / water), like the tile goldens
    s2, dates, interp, s1, dem = subtiles_ref.synth_ard(200 + seed, 12, 172, 172)
    m = np.concatenate([s2[:12], np.repeat(dem[None, ..., None], 12, 0), s1], -1)[None].astype(np.float32)
    x = P.normalize_subtile(P.assemble(m), api.MIN_ALL, api.MAX_ALL)
    y = sess.predict(x, length=4)[0]
    r = ref.forward(x)[0]
    errs.append(np.abs(y - r).ravel())
    m2 = P.synth_monthly(1, 172, 300 + seed)
    x2 = P.normalize_subtile(P.assemble(m2), api.MIN_ALL, api.MAX_ALL)
    errs.append(np.abs(sess.predict(x2, length=4)[0] - ref.forward(x2)[0]).ravel())
e = np.concatenate(errs)
print("pixels", e.size, "max %.3e" % e.max(), "p99.99 %.3e" % np.quantile(e, 0.9999), "p99.9 %.3e" % np.quantile(e, 0.999), "mean %.3e" % e.mean(),
      "count > 1e-3:", int((e > 1e-3).sum()), "count > 7.5e-4:", int((e > 7.5e-4).sum()))
for k, a in enumerate(errs):
    print(k, "max %.3e" % a.max())
