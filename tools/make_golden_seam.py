"""Generate tests/golden/seam.npz: the REFERENCE's border `process_subtiles`
(/root/reference/src/resegment_tiles_wide.py:360-616, through oracle/refshim.py, this container only) run on a seeded
joint strip with a stub TensorFlow session.  The stub records every `batch_x` the reference feeds to `sess.run` (the
clipped + normalised [1, LEN + 1, SIZE_Y + 14, SIZE + 14, 17] window, stored as every third pixel plus per-band sums: pins the window cut-outs, the reflect padding of
the end windows, the channel order, the medians and the resegment-specific band ranges :1664-1685) and answers with a
deterministic function of it (`stub_forward`, shared with tests/test_resegment.py), so the files the reference writes
(`processed/right<y>/<x>.npy`) also pin the left / right balance step :517-533 and the acceptance rule :536-611.
hist_align = False in case 0 (pure host logic), True in case 1 (adds align_subtile_histograms :284).
Usage: python tools/make_golden_seam.py"""
import os, sys, tempfile, textwrap, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim

SIZE, SIZE_Y, H = 32, 132, 340           # the reference asserts SIZE_Y + 14 >= 145 (:500)
CASES = [(0, False), (1, True)]            # (seed, hist_align)


def seam_inputs(seed):
    """Monthly joint strip [12, H, SIZE + 14, 14] with a gain / offset step at the seam, a few NaNs, S1, DEM, interp."""
    r = np.random.default_rng(100 + seed)
    W = SIZE + 14
    s2 = (r.uniform(0.03, 0.45, (1, H, W, 14)) + 0.02 * r.standard_normal((12, H, W, 14))).astype(np.float32)
    s2[:, :, W // 2:, :] = s2[:, :, W // 2:, :] * np.float32(1.15) + np.float32(0.01)
    s2[..., 10:] = r.uniform(-0.3, 0.6, (12, H, W, 4))
    s2[3, 5, 7, 2] = np.nan
    s2[7, 140:143, 20:24, :] = np.nan
    s1 = r.uniform(0.05, 0.9, (12, H, W, 2)).astype(np.float32)
    dem = r.uniform(0, 0.5, (H, W)).astype(np.float32)
    interp = (r.random((6, H, W)) < 0.1).astype(np.float32)
    dates = np.array([10, 60, 120, 190, 250, 330])
    left_all = r.uniform(20, 70, (H, SIZE // 2)).astype(np.float32)
    right_all = r.uniform(30, 80, (H, SIZE // 2)).astype(np.float32)
    left_all[2, 3] = np.nan
    return s2, dates, interp, s1, dem, left_all, right_all


def stub_forward(batch_x, call_index):
    """[1, 5, h, w, 17] normalised -> [1, h - 14, w - 14, 1]; every second call gets a left / right step of 0.3 so that
    the balance branch (:520-533) runs."""
    x = np.asarray(batch_x, np.float32)[0]
    p = 0.5 + 0.25 * x[1, 7:-7, 7:-7, 3] + 0.1 * x[4, 7:-7, 7:-7, 10]
    if call_index % 2 == 1:
        p[:, p.shape[1] // 2:] += np.float32(0.3)
    return np.clip(p, 0, 1).astype(np.float32)[np.newaxis, ..., np.newaxis]


class StubSession:
    def __init__(self):
        self.fed = []

    def run(self, fetch, feed_dict):
        bx = next(v for v in feed_dict.values() if np.ndim(v) == 5)
        self.fed.append(np.array(bx, np.float32))
        return stub_forward(bx, len(self.fed) - 1)


def main():
    m = refshim.ref("resegment_tiles_wide")
    src = open(m.__file__).read()
    i, j = src.index("    min_all = [0.0065"), src.index("    if os.path.exists(args.db_path)")
    exec(textwrap.dedent(src[i:j]), m.__dict__)                     # the module-level band ranges live under __main__ there
    m.SIZE, m.SIZE_Y, m.LEN = SIZE, SIZE_Y, 4
    m.predict_logits, m.predict_inp, m.predict_length = "logits", "inp", "length"
    out = {}
    for seed, hist_align in CASES:
        tmp = tempfile.mkdtemp() + "/"
        m.args = types.SimpleNamespace(local_path=tmp)
        os.makedirs(tmp + "0/0/processed/", exist_ok=True)
        s2, dates, interp, s1, dem, left_all, right_all = seam_inputs(seed)
        gap_y = int(np.ceil((H - SIZE_Y) / 3))                      # :1144-1146
        tiles_folder_y = np.hstack([np.arange(0, H - SIZE_Y, gap_y), np.array(H - SIZE_Y)])
        tiles_array, tiles_folder = m.make_tiles_right_neighb(np.array([0]), tiles_folder_y)
        sess = StubSession()
        cwd = os.getcwd(); os.chdir(tmp)
        m.process_subtiles(0, 0, np.copy(s2), np.copy(dates), np.copy(interp), np.copy(s1), np.copy(dem), sess, None,
                           tiles_folder, tiles_array, right_all, left_all, hist_align, np.zeros((H, SIZE + 14)))
        os.chdir(cwd)
        out["tiles_array_%d" % seed] = np.asarray(tiles_array, np.int64)
        out["tiles_folder_%d" % seed] = np.asarray(tiles_folder, np.int64)
        fed = np.stack([f[0] for f in sess.fed]).astype(np.float32)
        out["fed_%d" % seed] = fed[:, :, ::3, ::3, :]                # every third pixel (rows 0, 3, 6 lie in the reflect padding) ...
        out["fed_sums_%d" % seed] = fed.astype(np.float64).sum(axis=(2, 3))   # ... and a sum over ALL pixels per window / frame / band
        for t, tf in enumerate(np.asarray(tiles_folder)):
            f = "%s0/0/processed/right%d/%d.npy" % (tmp, int(tf[0]), int(tf[1]))
            if os.path.exists(f):
                out["preds_%d_%d" % (seed, t)] = np.load(f).astype(np.float32)
        print("case", seed, "windows", len(tiles_array), "fed", out["fed_%d" % seed].shape,
              "saved", sorted(k for k in out if k.startswith("preds_%d_" % seed)), flush=True)
    path = os.path.join(ROOT, "tests", "golden", "seam.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
