"""Border re-segmentation pass (sentinel_tree_cover_b200/resegment.py) against outputs of the reference's own functions
(/root/reference/src/resegment_tiles_wide.py run through oracle/refshim.py by tools/make_golden_resegment.py).
CPU: the integer / scalar host logic (date alignment, border window table, seam-artifact test, prediction adjustment).
GPU: align_subtile_histograms (rtol 1e-5: float64 accumulation of the statistics), regularize_and_smooth (1e-4, the K1
tolerance), preprocess_tile (cloud pipeline: dates, feather weights and generator state exact, filled values rtol 1e-4)."""
import importlib.util
import os
import random
import numpy as np
import pytest
from oracle import cloud_ref

HERE = os.path.dirname(__file__)
G = np.load(os.path.join(HERE, "golden", "resegment.npz"))
spec = importlib.util.spec_from_file_location("mk_reseg", os.path.join(HERE, "..", "tools", "make_golden_resegment.py"))
MK = importlib.util.module_from_spec(spec); spec.loader.exec_module(MK)


def test_align_dates_matches_reference():
    from sentinel_tree_cover_b200 import resegment as R
    for i, (a, b) in enumerate(MK.DATE_SETS):
        rt, rn, mn = R.align_dates(a, b)
        g = G["dates_%d" % i].tolist()
        assert [len(rt), len(rn), mn] == g[:3] and [int(v) for v in rt] + [int(v) for v in rn] == g[3:], i


def test_border_window_table_matches_reference():
    from sentinel_tree_cover_b200 import resegment as R
    ta, tf = R.make_tiles_right_neighb(np.array([0]), np.array([0, 138, 276, 412]), 670, 206)
    assert np.array_equal(ta, G["tiles_array"]) and np.array_equal(tf, G["tiles_folder"])


def test_check_if_artifact_matches_reference():
    from sentinel_tree_cover_b200 import resegment as R
    r = np.random.default_rng(3)
    got = []
    for k in range(6):
        tile = r.uniform(0, 100, (618, 40)).astype(np.float32); nb = tile[:, ::-1] + r.normal(0, [1, 4, 9, 15, 25, 2][k], tile.shape).astype(np.float32)
        if k == 5:
            nb += 8
        tile[r.random(tile.shape) < 0.02] = np.nan
        got.append(R.check_if_artifact(tile, nb))
    assert got == G["artifact"].tolist() and 0 in got and 1 in got
    p = r.uniform(0, 1, (40, 50)).astype(np.float32); ref = r.uniform(0.2, 0.9, (40, 50)).astype(np.float32); ref[3, 4] = np.nan
    np.testing.assert_allclose(R.adjust_predictions(np.copy(p), ref), G["adjust"], rtol=1e-6, atol=1e-7)


def test_balance_seam_predictions_properties():
    from sentinel_tree_cover_b200 import resegment as R
    r = np.random.default_rng(9)
    p = r.uniform(0.3, 0.5, (64, 80)).astype(np.float32)
    assert np.array_equal(R.balance_seam_predictions(p, 80), p)                  # no jump: untouched
    q = p.copy(); q[:, 40:] += 0.3; q[5, 3] = 0.01
    out = R.balance_seam_predictions(q, 80)
    assert abs(out[:, 36:40].mean() - out[:, 40:44].mean()) < 0.05 and out[5, 3] == np.float32(0.01)
    assert out.min() >= 0 and out.max() <= 1


@pytest.mark.gpu
def test_gpu_align_subtile_histograms_matches_reference(sess):
    from sentinel_tree_cover_b200 import resegment as R
    for i, (T, H, size, C, seed) in enumerate(MK.HIST_CASES):
        x = MK.hist_input(T, H, size, C, seed)
        y = R.align_subtile_histograms(np.copy(x), sess, size)
        want = G["hist_%d" % i]
        assert np.array_equal(np.isnan(y), np.isnan(want))
        np.testing.assert_allclose(y, want, rtol=1e-5, atol=1e-6, equal_nan=True)
    # two statistically identical halves: the transform cannot shrink the seam jump reliably; whatever the decision, the
    # result is either the input or a finite rescaling of it
    r = np.random.default_rng(1)
    z = r.uniform(0.1, 0.3, (2, 40, 94, 14)).astype(np.float32)
    out = R.align_subtile_histograms(np.copy(z), sess, 80)
    assert np.isfinite(out).all() and np.abs(out - z).max() < 0.05


@pytest.mark.gpu
def test_gpu_regularize_and_smooth_matches_reference(sess):
    from sentinel_tree_cover_b200 import resegment as R
    img, _ = cloud_ref.synth_cloudy_cube(9, 40, 44, 95)
    dates = np.array([12, 40, 75, 101, 140, 170, 220, 260, 320])
    y = R.regularize_and_smooth(img, dates, sess)
    assert y.shape == (12, 40, 44, 10)
    assert np.abs(y - G["regsmooth"]).max() < 1e-4


@pytest.mark.gpu
def test_gpu_preprocess_tile_matches_reference(sess):
    from sentinel_tree_cover_b200 import resegment as R
    for i, (T, H, W, seed, rseed, with_clm) in enumerate(MK.PRE_CASES):
        img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed)
        dts = (np.arange(T) * (330 // T) + 10).astype(np.int64)
        clm = None
        if with_clm:
            clm = np.zeros((T, H, W), np.float32); clm[2, 10:30, 20:50] = 1.
        random.seed(rseed)
        arr, interp2, d2 = R.preprocess_tile(np.copy(img), np.copy(dts), None, clm, "tile", np.copy(dem), None, sess)
        assert np.array_equal(d2, G["pre_dates_%d" % i])
        assert np.array_equal(np.asarray(interp2, np.float32).astype(np.float16), G["pre_interp_%d" % i])
        got = np.asarray(arr)[tuple(G["pre_idx_%d" % i].T)]
        np.testing.assert_allclose(got, G["pre_val_%d" % i], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(float(np.asarray(arr, np.float64).sum()), float(G["pre_arr_sum_%d" % i][0]), rtol=1e-6)
        assert random.random() == float(G["pre_next_random_%d" % i][0])
