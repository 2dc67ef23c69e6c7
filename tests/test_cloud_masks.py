"""Multi-temporal cloud / shadow mask (identify_clouds_shadows, cloud_removal.py:1215-1677).
CPU: the NumPy restatement (oracle/cloud_ref.py) reproduces the reference's own outputs stored in
tests/golden/cloud_masks.npz (made by tools/make_golden_cloud.py from the reference function).
GPU: the CUDA pipeline equals the oracle at every stage tap and the golden outputs, bit for bit."""
import os
import numpy as np
import pytest
from oracle import cloud_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "cloud_masks.npz")


def _cases():
    g = np.load(GOLD)
    for i, (T, H, W, seed) in enumerate(g["cases"]):
        n = T * H * W
        clouds = np.unpackbits(g["clouds_%d" % i])[:n].reshape(T, H, W)
        fcps = np.unpackbits(g["fcps_%d" % i])[:n].reshape(T, H, W)
        yield int(T), int(H), int(W), int(seed), clouds, fcps


def test_oracle_matches_reference_golden():
    for T, H, W, seed, clouds, fcps in _cases():
        img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed)
        c, f = cloud_ref.identify_clouds_shadows(img, dem)
        assert np.array_equal(c > 0, clouds > 0), (T, H, W)
        assert np.array_equal(np.asarray(f) > 0, fcps > 0), (T, H, W)


def test_windows_small_T():
    # the window helpers must stay inside [0, T) (or use Python's negative wrap exactly like the reference lists)
    for T in range(1, 14):
        for t in range(T):
            w = cloud_ref.shadow_window(t, T)
            assert w and min(w) >= 0 and max(w) < T
            others, close = cloud_ref.cloud_windows(t, T)
            assert min(others) >= 0 and max(others) < T
            if T > 2:        # the T <= 2 branch never indexes with `close`
                assert all(-T <= c < T for c in close)


@pytest.mark.gpu
def test_gpu_cloud_masks_golden(sess):
    from sentinel_tree_cover_b200 import api
    for T, H, W, seed, clouds, fcps in _cases():
        img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed)
        c, f = api.identify_clouds_shadows(img, dem, None, sess)
        assert c.dtype == np.float32 and c.shape == (T, H, W)
        assert np.array_equal(c > 0, clouds > 0), ("clouds", T, H, W, int(np.sum((c > 0) != (clouds > 0))))
        assert np.array_equal(f, fcps > 0), ("fcps", T, H, W)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(9, 64, 72, 21), (12, 96, 80, 22), (4, 40, 44, 23), (7, 150, 130, 24)])
def test_gpu_cloud_masks_stages(sess, shape):
    T, H, W, seed = shape
    img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed)
    c0, f0, st = cloud_ref.identify_clouds_shadows(img, dem, stages=True)
    for name in sess.CLOUD_STAGES:
        c, f, tap = sess.cloud_masks(img, dem, stage=name)
        want = np.asarray(st[name]) > 0
        bad = int(np.sum((tap > 0) != want))
        assert bad == 0, (name, bad, shape)
    assert np.array_equal(c > 0, np.asarray(c0) > 0)
    assert np.array_equal(f, np.asarray(f0) > 0)


@pytest.mark.gpu
def test_gpu_cloud_masks_full_tile_properties(sess):
    """Reference-scale cube (T=24, 640x640): size-independent properties -- masks are 0/1, a date
    flagged as haze is entirely 1, clear dates of a clean cube stay (almost) clear, determinism."""
    img, dem = cloud_ref.synth_cloudy_cube(24, 640, 640, 31)
    c1, f1 = sess.cloud_masks(img, dem)
    c2, f2 = sess.cloud_masks(img, dem)
    assert np.array_equal(c1, c2) and np.array_equal(f1, f2)
    assert set(np.unique(c1).tolist()) <= {0.0, 1.0}
    clean = np.repeat(img[:1], 6, 0) * np.linspace(0.97, 1.03, 6, dtype=np.float32)[:, None, None, None]
    cc, _ = sess.cloud_masks(clean, dem)
    assert cc.mean() < 0.02


ANC = os.path.join(os.path.dirname(__file__), "golden", "cloud_masks_anc.npz")


def _anc_cases():
    """tests/golden/cloud_masks_anc.npz: the reference run with seeded forest / urban rasters (tools/make_golden_cloud_anc.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk_anc", os.path.join(os.path.dirname(__file__), "..", "tools", "make_golden_cloud_anc.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    g = np.load(ANC)
    for i, (T, H, W, seed, wf, wu) in enumerate(g["cases"]):
        T, H, W, seed = int(T), int(H), int(W), int(seed)
        n = T * H * W
        clouds = np.unpackbits(g["clouds_%d" % i])[:n].reshape(T, H, W)
        fcps = np.unpackbits(g["fcps_%d" % i])[:n].reshape(T, H, W)
        frst, urb = mk.synth_rasters(H, W, seed)
        yield T, H, W, seed, (frst if wf else None), (urb if wu else None), clouds, fcps


def test_raster_helpers_match_scipy():
    """api.ancillary_masks_from_rasters (NumPy only) == the SciPy dilation + order-0 zoom of the oracle."""
    from sentinel_tree_cover_b200 import api
    r = np.random.default_rng(5)
    for shape, rs in (((96, 104), (7, 8)), ((75, 81), (5, 6)), ((618, 618), (39, 40)), ((300, 280), (20, 18))):
        frst, urb = r.random(rs) < 0.3, r.random(rs) < 0.2
        f, (c, n) = api.ancillary_masks_from_rasters(frst, urb, shape)
        of, oc, on = cloud_ref.rasters_to_masks(frst, urb, shape)
        assert np.array_equal(f, of) and np.array_equal(c, oc) and np.array_equal(n, on)
    for n_in in range(2, 60):
        for n_out in (n_in - 1, n_in + 1, 2 * n_in, 16 * n_in + 3):
            if n_out > 0:
                assert np.array_equal(api._nn_index(n_out, n_in), cloud_ref.nn_resize_index(n_out, n_in)), (n_out, n_in)


def test_oracle_matches_reference_golden_with_forest_and_urban_rasters():
    for T, H, W, seed, frst, urb, clouds, fcps in _anc_cases():
        if T * H * W > 400000:
            continue                     # the 24 x 200 x 208 case is checked on the GPU only (CPU suite budget)
        img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed, urban=True)
        forest, core, near = cloud_ref.rasters_to_masks(frst, urb, (H, W))
        c, f = cloud_ref.identify_clouds_shadows(img, dem, forest=forest, urban=(core, near) if urb is not None else None)
        assert np.array_equal(c > 0, clouds > 0), (T, H, W)
        assert np.array_equal(np.asarray(f) > 0, fcps > 0), (T, H, W)


@pytest.mark.gpu
def test_gpu_cloud_masks_with_forest_and_urban_rasters_match_reference(sess):
    """P5 with the ancillary rasters: forest thresholds (:1412-1416, :1443-1447, :1549) and the Fmask-4 parallax test of
    detect_pfcp (:1109-1212), incl. an odd-sided tile (75 x 81: the order-0 resizes)."""
    from sentinel_tree_cover_b200 import api
    for T, H, W, seed, frst, urb, clouds, fcps in _anc_cases():
        img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed, urban=True)
        forest, urban = api.ancillary_masks_from_rasters(frst, urb, (H, W))
        c, f = api.identify_clouds_shadows(img, dem, None, sess, forest_mask=forest, urban_mask=urban)
        nbad = int(np.sum((c > 0) != (clouds > 0)))
        assert nbad == 0, ("clouds", T, H, W, nbad)
        assert np.array_equal(f, fcps > 0), ("fcps", T, H, W, int(np.sum(f != (fcps > 0))))
    # the masks do not leak into the next call
    img, dem = cloud_ref.synth_cloudy_cube(5, 40, 44, 23)
    c0, f0 = cloud_ref.identify_clouds_shadows(img, dem)
    c1, f1 = api.identify_clouds_shadows(img, dem, None, sess)
    assert np.array_equal(c1 > 0, np.asarray(c0) > 0) and np.array_equal(f1, np.asarray(f0) > 0)


@pytest.mark.gpu
def test_gpu_parallax_stage_matches_oracle(sess):
    """fcps before the false-positive removal (detect_pfcp's return value) against the oracle's tap."""
    from sentinel_tree_cover_b200 import api
    for T, H, W, seed, frst, urb, clouds, fcps in _anc_cases():
        if urb is None or T * H * W > 400000:
            continue
        img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed, urban=True)
        forest, urban = api.ancillary_masks_from_rasters(frst, urb, (H, W))
        _, _, st = cloud_ref.identify_clouds_shadows(img, dem, stages=True, forest=forest, urban=urban)
        sess.set_ancillary_masks(forest, urban, (H, W))
        try:
            _, _, tap = sess.cloud_masks(img, dem, stage="fcps0")
        finally:
            sess.set_ancillary_masks(None, None)
        assert np.array_equal(tap > 0, st["fcps0"] > 0), (T, H, W, int(np.sum((tap > 0) != (st["fcps0"] > 0))))


def test_bad_args_fail_loudly():
    from sentinel_tree_cover_b200 import api
    assert "stc_cloud_masks_host" in [s[0] for s in api.SYMBOLS]
