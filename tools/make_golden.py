"""Generate tests/golden/* by RUNNING THE REFERENCE (only works in the build container,
where /root/reference exists).  Committed together with its outputs so the fixtures
can be regenerated and audited.

  * weights_predict_172.npz / weights_superresolve.npz : tensors of the released frozen
    graphs under canonical names (sentinel_tree_cover_b200/weights.py).
  * model_172.npz : outputs of predict_graph-172.pb for seeded inputs, computed by the
    mechanical GraphDef interpreter (oracle/tfgraph_interp.py; TensorFlow is absent).
  * superresolve.npz : same for superresolve_graph.pb.
  * preproc.npz : outputs of the reference's own NumPy functions executed unmodified
    through oracle/refshim.py (indices, calculate_and_save_best_images, Smoother,
    smooth_large_tile, normalize_subtile, make_overlapping_windows, fspecial_gauss ...).
"""
import os
import sys
import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import refshim                      # noqa: E402
from oracle import preproc_ref as P             # noqa: E402
from oracle.tfgraph_interp import GraphInterpreter  # noqa: E402
from sentinel_tree_cover_b200 import weights as W   # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF = refshim.REF_ROOT
PB172 = os.path.join(REF, "models-release/master-ckpt-frozen/predict_graph-172.pb")
PBSR = os.path.join(REF, "models-release/supres-40k-swir/superresolve_graph.pb")


def main():
    os.makedirs(OUT, exist_ok=True)
    os.chdir("/tmp")  # the reference dumps debug .npy files into CWD
    # ---- weights ----
    np.savez_compressed(os.path.join(OUT, "weights_predict_172.npz"), **W.load_predict_pb(PB172))
    np.savez_compressed(os.path.join(OUT, "weights_superresolve.npz"), **W.load_superresolve_pb(PBSR))

    # ---- model (graph interpreter) ----
    g = GraphInterpreter(PB172)
    gold = {}
    for name, seed, length in (("a", 7, 4), ("b", 8, 2)):
        x = P.synth_model_input(1, 172, seed)
        y = g.run("conv2d/Sigmoid", {"Placeholder": x, "PlaceholderWithDefault": np.full(1, length, np.int64)})[0, ..., 0]
        gold["y_" + name] = y.astype(np.float32)
        gold["seed_" + name] = np.int64(seed)
        gold["length_" + name] = np.int64(length)
        print("model", name, y.shape, float(y.mean()))
    np.savez_compressed(os.path.join(OUT, "model_172.npz"), **gold)

    gs = GraphInterpreter(PBSR)
    r = np.random.default_rng(11)
    x = r.uniform(0.0, 0.6, (2, 48, 40, 10)).astype(np.float32)
    b = x[..., 4:].copy()
    y = gs.run("Add_2", {"Placeholder": x, "Placeholder_1": b})
    np.savez_compressed(os.path.join(OUT, "superresolve.npz"), x=x, y=y.astype(np.float32))

    # ---- preprocessing through the reference's own functions ----
    job = refshim.ref("download_and_predict_job")
    utils = refshim.ref("downloading.utils")
    ind = refshim.ref("preprocessing.indices")
    tofd = refshim.ref("tof.tof_downloading")
    out = {}
    r = np.random.default_rng(3)
    cube = r.uniform(-0.1, 1.1, (5, 16, 16, 10)).astype(np.float32)
    out["idx_in"] = cube
    out["idx_out"] = job.make_indices(cube)
    # date sets -> G matrices (identity trick) and max_distance
    date_sets = [np.array([10, 40, 75, 100, 140, 190, 230, 290, 340]),
                 np.array([0, 22, 105, 232, 295, 310, 330]),
                 np.array([95, 120, 150, 200, 260]),
                 np.array([5, 15, 33, 48, 61, 77, 92, 110, 125, 141, 155, 170, 188, 201, 216, 230, 246, 262, 275, 290, 307, 321, 336, 350]),
                 np.array([15, 45, 75, 105, 135, 165, 195, 225, 255, 285, 315, 345])]
    for i, d in enumerate(date_sets):
        G, md = utils.calculate_and_save_best_images(np.eye(len(d), dtype=np.float32).reshape(len(d), len(d), 1, 1), d.copy())
        out["dates_%d" % i] = d
        out["G_%d" % i] = G[..., 0, 0]
        out["maxdist_%d" % i] = np.int64(md)
    out["n_date_sets"] = np.int64(len(date_sets))
    # smooth_large_tile end to end
    arr = r.uniform(0.02, 0.45, (9, 20, 24, 10)).astype(np.float32)
    arr = np.ascontiguousarray(arr[:, :, :20])  # the reference Smoother assumes square-ish dims via dimx/dimy
    sm_out, d_out, _ = job.smooth_large_tile(arr.copy(), date_sets[0].copy(), np.zeros((9, 20, 20), np.float32))
    out["smooth_in"] = arr
    out["smooth_out"] = sm_out.astype(np.float32)
    # whittaker matrix from the reference's splu (Appendix B)
    wh = refshim.ref("preprocessing.whittaker_smoother")
    S = wh.Smoother(lmbd=100, size=24, nbands=1, dimx=1, dimy=24, average=False).smooth(np.eye(24, dtype=np.float32))
    out["whittaker_S"] = np.asarray(S, np.float32)
    # normalize_subtile (module globals min_all / max_all are bound in __main__: set them)
    from sentinel_tree_cover_b200.api import MIN_ALL, MAX_ALL
    job.min_all, job.max_all = MIN_ALL, MAX_ALL
    sub = r.uniform(-0.2, 1.0, (5, 12, 12, 17)).astype(np.float32)
    out["norm_in"] = sub
    out["norm_out"] = job.normalize_subtile(sub.copy())
    # windows
    for L in (618, 600, 316):
        gap = int(np.ceil((L - 158) / 5))
        fx = np.hstack([np.arange(0, L - 158, gap), np.array(L - 158)])
        mesh = np.meshgrid(fx, fx)
        w = np.reshape(np.concatenate(mesh).ravel(), (2, mesh[0].size)).T
        tf = np.sort(np.hstack([w, np.full_like(w, 158)]), axis=0)
        tf[:, 1] = np.tile(np.unique(tf[:, 1]), int(len(tf[:, 1]) / len(np.unique(tf[:, 1]))))
        out["win_folder_%d" % L] = tf
        out["win_array_%d" % L] = tofd.make_overlapping_windows(tf, diff=7)
    out["gauss_158_36"] = job.fspecial_gauss(158, 36)
    # 13-band legacy index path (multiyear predict_subtile :808-813): indices on a 13-band cube
    m = P.synth_monthly(1, 12, 5)
    out["monthly_in"] = m
    out["monthly_idx"] = np.stack([ind.evi(m), ind.bi(m), ind.msavi2(m), ind.grndvi(m)], -1).astype(np.float32)
    # superresolve_large_tile window quirks with a stand-in network (x -> bilinear + 0.01*mean(x))
    class FakeSess:
        def run(self, ops, feed_dict):
            xin = feed_dict["inp"]; bil = feed_dict["bil"]
            return [bil + 0.01 * xin.mean(axis=-1, keepdims=True)]
    job.superresolve_logits, job.superresolve_inp, job.superresolve_inp_bilinear = "logits", "inp", "bil"
    tile = r.uniform(0.0, 0.5, (1, 225, 231, 10)).astype(np.float32)
    res = job.superresolve_large_tile(tile.copy(), FakeSess())
    out["srtile_in"] = tile
    out["srtile_out_sub"] = res[:, ::3, ::3, 4:].astype(np.float32)
    out["srtile_out_sum"] = np.float64(res.astype(np.float64).sum())
    np.savez_compressed(os.path.join(OUT, "preproc.npz"), **out)
    mosaic_golden()
    mosaic_feats_golden()
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__" and "--mosaic-only" not in sys.argv and "--mosaic-feats-only" not in sys.argv:
    main()


def mosaic_golden():
    """load_mosaic_predictions run by the reference on synthetic subtile files."""
    import shutil, tempfile
    job = refshim.ref("download_and_predict_job")
    out = {}
    for case, kw in (("a", dict(seed=1)), ("b", dict(seed=2, all_nodata=(5,)))):
        preds, xs, ys = P.synth_subtile_preds(618, 158, **kw)
        d = tempfile.mkdtemp() + "/"
        for p, x, y in zip(preds, xs, ys):
            os.makedirs(d + str(x), exist_ok=True)
            np.save(d + str(x) + "/" + str(y) + ".npy", p)
        job.SIZE = 158
        res = job.load_mosaic_predictions(d, 1)
        # record the layer order the reference walked (os.listdir order)
        order = []
        for xt in [int(x) for x in os.listdir(d)]:
            for yt in [int(y[:-4]) for y in os.listdir(d + str(xt) + "/")]:
                order.append((xt, yt))
        out["order_" + case] = np.array(order, np.int32)
        out["out_" + case] = res
        shutil.rmtree(d)
    np.savez_compressed(os.path.join(OUT, "mosaic.npz"), **out)
    print("mosaic golden", {k: v.shape for k, v in out.items()})


def mosaic_feats_golden():
    """load_mosaic_predictions(depth=16) and float_to_int16 run by the reference on synthetic feature files."""
    import shutil, tempfile
    job = refshim.ref("download_and_predict_job")
    out = {}
    feats, xs, ys = P.synth_subtile_feats(250, 158, 16, seed=3)
    d = tempfile.mkdtemp() + "/"
    for f, x, y in zip(feats, xs, ys):
        os.makedirs(d + str(x), exist_ok=True)
        np.save(d + str(x) + "/" + str(y) + ".npy", f)
    job.SIZE = 158
    with np.errstate(all="ignore"):
        res = job.load_mosaic_predictions(d, 16)
    order = []
    for xt in [int(x) for x in os.listdir(d)]:
        for yt in [int(y[:-4]) for y in os.listdir(d + str(xt) + "/")]:
            order.append((xt, yt))
    out["order"] = np.array(order, np.int32)
    out["out"] = res
    shutil.rmtree(d)
    r = np.random.default_rng(11)
    x = (r.normal(0, 12, (64, 33)) * r.choice([0.01, 1, 5], (64, 33))).astype(np.float32)
    x[3, 4] = np.nan; x[0, 0] = 40.0; x[1, 1] = -40.0; x[2, 2] = 32.767; x[5, 5] = -32.768
    out["f2i_in"] = x
    out["f2i_out"] = job.float_to_int16(x.copy())
    np.savez_compressed(os.path.join(OUT, "mosaic_feats.npz"), **out)
    print("mosaic feats golden", {k: v.shape for k, v in out.items()})


if __name__ == "__main__" and "--mosaic-only" in sys.argv:
    mosaic_golden()
if __name__ == "__main__" and "--mosaic-feats-only" in sys.argv:
    mosaic_feats_golden()
