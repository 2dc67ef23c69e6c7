"""GPU parity tests for the model path, through the C ABI (ctypes)."""
import numpy as np
import pytest
from conftest import golden
from oracle.model_ref import PredictRef, SuperresolveRef
from oracle import preproc_ref as P
from sentinel_tree_cover_b200.api import StcSession, predict_subtile, MIN_ALL, MAX_ALL
from sentinel_tree_cover_b200.weights import random_predict_weights

pytestmark = pytest.mark.gpu
TOL = 1e-3   # north_star: <= 1e-3 abs on the probability maps


def _taps_nhwc(t):
    return np.ascontiguousarray(t.numpy().transpose(0, 2, 3, 1))


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
def test_predict_small_random_weights_vs_oracle(impl):
    """H=44, random weights, both conv implementations (1 = SIMT verification kernel,
    0 = tcgen05), with intermediate taps to localise any mismatch."""
    w = random_predict_weights(3)
    s = StcSession(0, predict_weights=w, conv_impl=impl)
    x = P.synth_model_input(3, 44, 21)
    taps = {}
    ref = PredictRef(w).forward(x, taps=taps)
    y = s.predict(x, length=4)
    ccin = s.debug_read("ccin").reshape(3, 44, 44, 128)
    gru = _taps_nhwc(taps["gru"])
    med = _taps_nhwc(taps["conv_median"])
    e_gru = np.abs(ccin[..., :64] - gru).max()
    e_med = np.abs(ccin[..., 64:] - med).max()
    cat2 = s.debug_read("cat2").reshape(3, 32, 32, 128)
    e_up3 = np.abs(cat2[..., :64] - _taps_nhwc(taps["up3"])).max()
    e_cc = np.abs(cat2[..., 64:] - _taps_nhwc(taps["conv_concat"])[:, 6:-6, 6:-6]).max()
    err = np.abs(y - ref).max()
    print("impl", impl, "gru", e_gru, "median", e_med, "conv_concat", e_cc, "up3", e_up3, "out", err)
    assert e_gru < 5e-3 and e_med < 2e-2 and e_cc < 2e-2 and e_up3 < 3e-2
    assert y.shape == (3, 30, 30) and err < TOL
    s.close()


def test_umma_matches_simt_tightly():
    """Same fp16 operands, fp32 accumulation: the tensor-core path must agree with the
    CUDA-core kernel to accumulation-order noise."""
    w = random_predict_weights(4)
    x = P.synth_model_input(2, 60, 22)
    ys = []
    for impl in (1, 0):
        s = StcSession(0, predict_weights=w, conv_impl=impl)
        ys.append(s.predict(x))
        s.close()
    assert np.abs(ys[0] - ys[1]).max() < 2e-4


def test_predict_released_weights_golden_172(sess):
    g = golden("model_172.npz")
    for k in ("a", "b"):
        x = P.synth_model_input(1, 172, int(g["seed_" + k]))
        y = sess.predict(x, length=int(g["length_" + k]))[0]
        err = np.abs(y - g["y_" + k])
        print("golden", k, "max", err.max(), "mean", err.mean())
        assert err.max() < TOL


def test_predict_batch_chunks_and_independence(sess, predict_weights, monkeypatch):
    x = P.synth_model_input(5, 76, 23)
    monkeypatch.setenv("STC_CHUNK", "2")        # 3 chunks: 2+2+1
    y = sess.predict(x)
    monkeypatch.delenv("STC_CHUNK")
    ref = PredictRef(predict_weights).forward(x)
    assert np.abs(y - ref).max() < TOL
    y1 = sess.predict(x[3:4])
    assert np.abs(y1[0] - y[3]).max() < 1e-5


def test_predict_subtile_contract(sess):
    """src/download_and_predict_job.py:328-369: all-zero -> int 255 fill; uint16 input is
    rescaled; output centre-cropped to `size`."""
    z = predict_subtile(np.zeros((5, 76, 76, 17), np.float32), sess, None, 62)
    assert z.shape == (62, 62) and z.dtype.kind == "i" and (z == 255).all()
    x = P.synth_model_input(1, 76, 5)[0]
    p = predict_subtile(x, sess, None, 62)
    assert p.shape == (62, 62) and p.dtype == np.float32 and (p > 0).all() and (p < 1).all()
    p2 = predict_subtile(x, sess, None, 58)
    assert np.array_equal(p2, p[2:-2, 2:-2])


def test_normalize_fused_matches_host(sess, predict_weights):
    r = np.random.default_rng(9)
    raw = r.uniform(-0.2, 1.0, (2, 5, 44, 44, 17)).astype(np.float32)
    y = sess.predict(raw, normalize=True)
    ref = PredictRef(predict_weights).forward(P.normalize_subtile(raw, MIN_ALL, MAX_ALL))
    assert np.abs(y - ref).max() < TOL


def test_superresolve_vs_graph_golden(sess, sr_weights):
    g = golden("superresolve.npz")
    y = sess.superresolve(g["x"], g["x"][..., 4:])
    err = np.abs(y - g["y"]).max()
    print("superresolve err", err)
    assert err < 1e-3
    r = np.random.default_rng(2)
    x = r.uniform(0, 0.6, (3, 118, 118, 10)).astype(np.float32)
    y = sess.superresolve(x, x[..., 4:])
    assert np.abs(y - SuperresolveRef(sr_weights).forward(x, x[..., 4:])).max() < 1e-3


def test_full_size_properties(sess):
    """BASELINE size (168 -> 154), batch 8: size-independent properties."""
    m = P.synth_monthly(8, 168, 77)
    y = sess.predict_patches(m)
    assert y.shape == (8, 154, 154) and np.isfinite(y).all() and (y > 0).all() and (y < 1).all()
    y2 = sess.predict_patches(m[::-1].copy())
    assert np.abs(y2[::-1] - y).max() < 1e-5          # batch order / chunk position invariance
