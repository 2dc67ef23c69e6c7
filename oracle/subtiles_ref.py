"""TEST INFRASTRUCTURE -- synthetic analysis-ready cube for process_subtiles
(/root/reference/src/download_and_predict_job.py:1125-1486) and the patches that let the reference
function run here: module globals it reads (args, SIZE, year, min_all/max_all, uploader), file outputs,
and `predict_subtile`, whose TensorFlow session is replaced by the oracle restatement of the frozen
graph.  Tests only."""
import os
import types
import numpy as np
from oracle import cloud_ref


def synth_ard(seed, n, H, W):
    """(s2 [n,H,W,10] float32 in [0,1], dates [n], interp [n,H,W] float32, s1 [12,H,W,2], dem [H,W])."""
    img, dem = cloud_ref.synth_cloudy_cube(n, H, W, seed)
    r = np.random.default_rng(seed + 5)
    s2 = np.clip(img, 0, 1).astype(np.float32)
    dates = (np.arange(n) * (340 // n) + 12).astype(np.int64)
    interp = np.zeros((n, H, W), np.float32)
    for t in range(n):                                   # a few interpolated patches, one region never clear
        y0, x0 = r.integers(0, H - 60), r.integers(0, W - 60)
        interp[t, y0:y0 + 50, x0:x0 + 60] = r.uniform(0.4, 1.0)
    interp[:, :70, W - 90:] = 1.0
    s1 = r.uniform(0.05, 0.95, (12, H, W, 2)).astype(np.float32)
    return s2, dates, interp, s1, (dem / 90).astype(np.float32)


def patch_reference(job, tmp, min_all, max_all):
    import torch
    from oracle.model_ref import PredictRef
    from sentinel_tree_cover_b200.weights import load_npz
    gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "weights_predict_172.npz")
    model = PredictRef(load_npz(gold))
    job.args = types.SimpleNamespace(local_path=tmp, length=4, process=True, gen_feats=False, gen_composite=False,
                                     make_training_data=False, s3_bucket="none")
    job.SIZE = 158
    job.year = 2020
    job.min_all, job.max_all = list(min_all), list(max_all)
    job.WRITE_MONTHLY_TIFS = False
    job.uploader = types.SimpleNamespace(upload=lambda **k: None)
    job.write_ard_to_tif = lambda *a, **k: None
    job.hkl.dump = lambda *a, **k: None
    real_remove = os.remove
    job.os.remove = lambda p: real_remove(p) if os.path.exists(p) else None
    job.predict_logits = None

    def predict_subtile(subtile, sess, op, size):
        # :328-369 with sess.run replaced by the oracle graph restatement
        if np.sum(subtile) != 0:
            batch_x = subtile[np.newaxis].astype(np.float32)
            preds = model.forward(batch_x).squeeze()
            clip = (preds.shape[0] - size) // 2
            if clip > 0:
                preds = preds[clip:-clip, clip:-clip]
            return np.float32(preds)
        return np.full((size, size), 255)
    job.predict_subtile = predict_subtile
