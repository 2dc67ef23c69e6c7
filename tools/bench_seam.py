"""Border re-segmentation forward (src/resegment_tiles_wide.py:182-222, :478): one [5, 220, 684, 17] seam window through
resegment.predict_subtile (host buffers in and out, H2D / D2H inside) vs the torch-CPU restatement of the same graph at
the same size.  Prints one JSON line.  Usage (GPU box): python tools/bench_seam.py [reps]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    from sentinel_tree_cover_b200.api import StcSession
    from sentinel_tree_cover_b200 import resegment as R, synth
    gold = os.path.join(ROOT, "tests", "golden")
    sess = StcSession(0, predict_weights=os.path.join(gold, "weights_predict_172.npz"))
    r = np.random.default_rng(0)
    mn, mx = np.float32(R.MIN_ALL), np.float32(R.MAX_ALL)
    x = (r.random((5, 220, 684, 17), dtype=np.float32) * (mx - mn) + mn).astype(np.float32)
    for _ in range(3):
        y = R.predict_subtile(x, sess)
    sess.sync()
    l0 = sess.launch_count()
    t0 = time.perf_counter()
    for _ in range(reps):
        y = R.predict_subtile(x, sess)
    sess.sync()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    launches = (sess.launch_count() - l0) // reps
    line = {"workload": "one border window [5,220,684,17] -> [206,670] (resegment.predict_subtile, host buffers)", "ms_per_window": ms,
            "windows_per_s": 1e3 / ms, "launches_per_window": int(launches), "out_mean": float(y.mean())}
    try:                                    # CPU leg: the oracle restatement at the same size (checker + baseline)
        import torch
        from oracle.model_ref import PredictRef
        from sentinel_tree_cover_b200.weights import load_npz
        m = PredictRef(load_npz(os.path.join(gold, "weights_predict_172.npz")))
        xn = ((np.clip(x, mn, mx) - ((mx + mn) / 2).astype(np.float32)) / ((mx - mn).astype(np.float32) / 2))[np.newaxis]
        m.forward(xn)
        t0 = time.perf_counter()
        ref = m.forward(xn)[0]
        cpu_s = time.perf_counter() - t0
        line.update({"cpu_port_s_per_window": cpu_s, "cpu_threads": torch.get_num_threads(),
                     "max_abs_err_vs_f32_port": float(np.abs(np.asarray(ref) - y).max())})
    except Exception as e:
        line["cpu_port_error"] = str(e)[:200]
    print(json.dumps(line), flush=True)
    sess.close()


if __name__ == "__main__":
    main()
