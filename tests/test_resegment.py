"""Border re-segmentation pass (sentinel_tree_cover_b200/resegment.py) against outputs of the reference's own functions
(/root/reference/src/resegment_tiles_wide.py run through oracle/refshim.py by tools/make_golden_resegment.py).
CPU: the integer / scalar host logic (date alignment, border window table, seam-artifact test, prediction adjustment).
GPU: align_subtile_histograms (rtol 1e-5: float64 accumulation of the statistics), regularize_and_smooth (1e-4, the K1
tolerance), preprocess_tile (cloud pipeline: dates, feather weights and generator state exact, filled values rtol 1e-4),
the border process_subtiles loop against the reference function run with a stub session (tools/make_golden_seam.py: the
tensors it feeds to sess.run and the files it writes), and the rectangular 220 x 684 seam forward against the float32
restatement of the graph evaluated at the same size."""
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
GS = np.load(os.path.join(HERE, "golden", "seam.npz"))
spec = importlib.util.spec_from_file_location("mk_seam", os.path.join(HERE, "..", "tools", "make_golden_seam.py"))
MKS = importlib.util.module_from_spec(spec); spec.loader.exec_module(MKS)


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


def _reference_normalise(x):
    """resegment_tiles_wide.py:201-202 with the float32 constants of :1677-1685."""
    from sentinel_tree_cover_b200 import resegment as R
    mn, mx = np.float32(R.MIN_ALL), np.float32(R.MAX_ALL)
    mid, rng = ((mx + mn) / 2).astype(np.float32), (mx - mn).astype(np.float32)
    return (np.clip(x, mn, mx) - mid) / (rng / 2)


def test_border_window_assembly_pads_the_end_windows():
    """:455-489 on the host: the first window is mirrored upwards, the last one downwards, channel order
    [10 S2 | DEM | 2 S1 | 4 indices], the median frame last."""
    from sentinel_tree_cover_b200 import resegment as R
    r = np.random.default_rng(4)
    size, size_y = 20, 12
    for start_y, h in ((0, size_y + 7), (30, size_y + 7), (9, size_y + 14)):
        s2 = r.random((4, h, size + 14, 14)).astype(np.float32); s1 = r.random((4, h, size + 14, 2)).astype(np.float32)
        dem = r.random((1, h, size + 14)).astype(np.float32)
        m2 = r.random((1, h, size + 14, 14)).astype(np.float32); m1 = r.random((1, h, size + 14, 2)).astype(np.float32)
        x = R.assemble_border_subtile(s2, s1, dem, m2, m1, start_y, size, size_y)
        assert x.shape == (5, size_y + 14, size + 14, 17) and x.dtype == np.float32
        pad = ((7, 0) if start_y == 0 else (0, 7)) if h == size_y + 7 else (0, 0)
        want = np.pad(s2, ((0, 0), pad, (0, 0), (0, 0)), "reflect")
        assert np.array_equal(x[:4, ..., :10], want[..., :10]) and np.array_equal(x[:4, ..., 13:], want[..., 10:])
        assert np.array_equal(x[2, ..., 11:13], np.pad(s1, ((0, 0), pad, (0, 0), (0, 0)), "reflect")[2])
        assert np.array_equal(x[4, ..., :10], np.pad(m2, ((0, 0), pad, (0, 0), (0, 0)), "reflect")[0, ..., :10])
        assert all(np.array_equal(x[t, ..., 10], np.pad(dem, ((0, 0), pad, (0, 0)), "reflect")[0]) for t in range(5))


def test_seam_prediction_acceptance_rule():
    """:536-611: every finite prediction is written (the three range branches save the same array); a NaN mean next to
    finite reference maps is the one case that is skipped; no reference data at all -> written."""
    from sentinel_tree_cover_b200 import resegment as R
    r = np.random.default_rng(2)
    left, right = r.uniform(20, 60, (30, 16)), r.uniform(40, 80, (30, 16))
    p = r.uniform(0, 1, (30, 32)).astype(np.float32)
    assert R.seam_prediction_accepted(p, left, right, 32) and R.seam_prediction_accepted(p * 0, left, right, 32)
    assert R.seam_prediction_accepted(np.full((30, 32), 255), left, right, 32)
    q = p.copy(); q[3, 4] = np.nan
    assert R.seam_prediction_accepted(q, left, right, 32)                       # np.nanmean ignores it
    assert not R.seam_prediction_accepted(np.full((30, 32), np.nan, np.float32), left, right, 32)
    assert R.seam_prediction_accepted(np.full((30, 32), np.nan, np.float32), left * np.nan, right * np.nan, 32)


@pytest.mark.gpu
def test_gpu_border_process_subtiles_matches_reference(sess):
    """The reference's border process_subtiles (:360-616) with a stub session vs the mirror with the same stub as its
    forward: identical tensors reach the model (every third pixel + per-band sums over all pixels) and identical arrays
    come out of the balance / acceptance steps."""
    from sentinel_tree_cover_b200 import resegment as R
    for seed, hist_align in MKS.CASES:
        s2, dates, interp, s1, dem, left_all, right_all = MKS.seam_inputs(seed)
        gap_y = int(np.ceil((MKS.H - MKS.SIZE_Y) / 3))
        tfy = np.hstack([np.arange(0, MKS.H - MKS.SIZE_Y, gap_y), np.array(MKS.H - MKS.SIZE_Y)])
        ta, tf = R.make_tiles_right_neighb(np.array([0]), tfy, MKS.SIZE, MKS.SIZE_Y)
        assert np.array_equal(ta, GS["tiles_array_%d" % seed]) and np.array_equal(tf, GS["tiles_folder_%d" % seed])
        fed = []

        def forward(x):
            bx = _reference_normalise(x)[np.newaxis]
            fed.append(bx[0])
            return MKS.stub_forward(bx, len(fed) - 1)[0, ..., 0]
        out = R.process_subtiles(s2, dates, interp, s1, dem, sess, tf, ta, right_all, left_all, MKS.SIZE, MKS.SIZE_Y,
                                 hist_align=hist_align, forward=forward)
        fed = np.stack(fed)
        tol = dict(rtol=2e-5, atol=2e-5) if hist_align else dict(rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(fed[:, :, ::3, ::3, :], GS["fed_%d" % seed], **tol)
        np.testing.assert_allclose(fed.astype(np.float64).sum(axis=(2, 3)), GS["fed_sums_%d" % seed], rtol=1e-5, atol=1e-2)
        for t, f in enumerate(np.asarray(tf)):
            key = "preds_%d_%d" % (seed, t)
            assert (key in GS.files) == ((int(f[0]), int(f[1])) in out)
            if key in GS.files:
                np.testing.assert_allclose(out[(int(f[0]), int(f[1]))], GS[key], rtol=1e-5, atol=1e-5)
        assert len(out) == 4


@pytest.mark.gpu
def test_gpu_seam_forward_220x684(sess, predict_weights):
    """resegment.predict_subtile on the production border window [5, 220, 684, 17] -> [206, 670]: ONE rectangular forward.
    The reference's wide-border weight set is unreleased (:1605); the network is fully convolutional, so the released 172-px
    weights evaluated at 220 x 684 by the float32 restatement (pinned to the GraphDef interpreter at the square sizes) are
    the oracle.  Tolerance: the north_star 1e-3.  Also: uint16 storage input, the all-zero fill, the SIMT cross-check."""
    import torch
    from oracle import preproc_ref as P
    from oracle.model_ref import PredictRef
    from sentinel_tree_cover_b200 import resegment as R
    from sentinel_tree_cover_b200.api import StcSession
    raw = P.synth_model_input(1, 684, 77)[0, :, 100:320]                        # [5, 220, 684, 17], already in [-1, 1]
    mn, mx = np.float32(R.MIN_ALL), np.float32(R.MAX_ALL)
    x = ((raw * 0.5 + 0.5) * (mx - mn) + mn).astype(np.float32)                 # band units, so the clip / normalise step matters
    x[0, :5, :5, 3] = 2.0                                                       # clipped to max_all
    got = R.predict_subtile(x, sess)
    assert got.shape == (206, 670) and got.dtype == np.float32
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = PredictRef(predict_weights).forward(_reference_normalise(x)[np.newaxis].astype(np.float32))[0]
    err = np.abs(got - ref).max()
    print("seam forward 220x684 vs f32 restatement", err)
    assert err < 1e-3
    s1 = StcSession(0, predict_weights=predict_weights, conv_impl=1)
    d = np.abs(R.predict_subtile(x, s1) - got).max()
    s1.close()
    print("seam forward: tcgen05 vs simt", d)
    assert d < 2e-4
    u16 = np.clip(np.round(np.clip(x, 0, 1) * 65535), 2, 65535).astype(np.uint16)
    gu = R.predict_subtile(u16, sess)
    assert np.abs(gu - R.predict_subtile((u16 / 65535.).astype(np.float32), sess)).max() < 1e-6
    z = R.predict_subtile(np.zeros((5, 220, 684, 17), np.float32), sess, size=670)
    assert z.shape == (670, 670) and (z == 255).all()
