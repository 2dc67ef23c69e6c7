"""Subtile post-filters: bright-bare-surface ramp, no-image block vote, rounding."""
import numpy as np
import pytest
from oracle import morph_ref as M
from oracle import refshim


@pytest.mark.skipif(not refshim.available(), reason="reference tree absent")
def test_bright_surface_oracle_equals_reference(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    job = refshim.ref("download_and_predict_job")
    x, _ = M.synth_subtile_stack(0)
    assert np.array_equal(job.identify_bright_bare_surfaces(x), M.identify_bright_bare_surfaces(x))


@pytest.mark.gpu
def test_edt_capped_exact(sess):
    from scipy.ndimage import distance_transform_edt
    r = np.random.default_rng(3)
    t = r.uniform(0, 1, (2, 90, 77)) > 0.99
    for cap in (3, 5, 12):
        want = np.minimum(np.stack([distance_transform_edt(1 - m) for m in t]), cap)
        assert np.array_equal(sess.edt_capped(t, cap), want)


@pytest.mark.gpu
def test_postprocess_subtile_exact(sess):
    from sentinel_tree_cover_b200.api import postprocess_subtile, identify_bright_bare_surfaces
    for seed in (0, 1):
        x, clear = M.synth_subtile_stack(seed)
        preds = np.random.default_rng(seed).uniform(0, 1, (158, 158)).astype(np.float32)
        assert np.array_equal(identify_bright_bare_surfaces(x, sess), M.identify_bright_bare_surfaces(x))
        got = postprocess_subtile(preds, x, clear, sess)
        want = M.postprocess_subtile(preds, x, clear)
        assert got.dtype == np.float32 and np.array_equal(got, want)
        assert (got == 255).any() and (got < 1).any()
