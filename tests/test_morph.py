"""Feathering (EDT + grey closing) and binary dilation: oracle vs the reference function run
through the shim (CPU, when the reference tree exists) and CUDA vs oracle (GPU, exact)."""
import numpy as np
import pytest
from oracle import morph_ref as M
from oracle import refshim


@pytest.mark.skipif(not refshim.available(), reason="reference tree absent")
def test_feather_oracle_equals_reference_function(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    cr = refshim.ref("preprocessing.cloud_removal")
    m = M.synth_cloud_masks(5, 70, 83, 0)
    ref = cr.id_areas_to_interp(None, m.copy(), None, None, None)
    assert np.array_equal(M.feather(np.clip(m, 0, 1), 15), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("size", [15, 20])
def test_feather_gpu_exact(sess, size):
    for (n, H, W, seed) in ((5, 70, 83, 1), (3, 168, 168, 2), (2, 31, 29, 3)):
        m = M.synth_cloud_masks(n, H, W, seed)
        assert np.array_equal(sess.feather(m, size), M.feather(m, size)), (n, H, W, size)


@pytest.mark.gpu
def test_id_areas_to_interp_contract(sess):
    from sentinel_tree_cover_b200.api import id_areas_to_interp
    probs = M.synth_cloud_masks(4, 64, 64, 5) * 1.7 - 0.2       # values outside [0,1] get clipped (:784)
    out = id_areas_to_interp(None, probs, None, None, None, sess)
    assert out.dtype == np.float32 and np.array_equal(out, M.feather(np.clip(probs, 0, 1), 15))


@pytest.mark.gpu
def test_binary_dilation_gpu_exact(sess):
    r = np.random.default_rng(0)
    x = r.uniform(0, 1, (3, 57, 64)) > 0.985
    for conn in (1, 2):
        for k in (1, 2, 3, 5, 10):
            assert np.array_equal(sess.binary_dilation(x, k, conn), M.dilate(x, k, conn)), (conn, k)
    # erosion idiom of the reference: 1 - binary_dilation(x == 0, iterations=k)  
    y = r.uniform(0, 1, (40, 40)) > 0.2
    got = ~sess.binary_dilation(~y, 2, 1)
    assert np.array_equal(got, ~M.dilate(~y, 2, 1))
