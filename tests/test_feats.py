"""--gen_feats path: feature taps of the forward, the int16 x1000 codec and the feature mosaic
(load_mosaic_predictions depth > 1).  Codec and mosaic are pinned against the reference's own functions
(tests/golden/mosaic_feats.npz, tools/make_golden.py --mosaic-feats-only); the taps against the graph oracle."""
import os
import numpy as np
import pytest
from conftest import golden
from oracle import preproc_ref as P
from oracle.model_ref import PredictRef


def _layers():
    g = golden("mosaic_feats.npz")
    feats, xs, ys = P.synth_subtile_feats(250, 158, 16, seed=3)
    d = {(x, y): f for f, x, y in zip(feats, xs, ys)}
    order = g["order"]
    return [d[(int(x), int(y))] for x, y in order], [int(x) for x, _ in order], [int(y) for _, y in order], g["out"]


def test_float_to_int16_oracle_matches_reference_golden():
    g = golden("mosaic_feats.npz")
    got = P.float_to_int16(g["f2i_in"])
    assert got.dtype == np.int16 and np.array_equal(got, g["f2i_out"])


def test_mosaic_feats_oracle_matches_reference_golden():
    fl, xs, ys, want = _layers()
    got = P.mosaic_feats(fl, xs, ys, want.shape[1:], 16)
    assert got.dtype == np.int16 and np.array_equal(got, want)


@pytest.mark.gpu
def test_float_to_int16_gpu_bit_exact(sess):
    g = golden("mosaic_feats.npz")
    got = sess.float_to_int16(g["f2i_in"])
    assert got.dtype == np.int16 and np.array_equal(got, g["f2i_out"])
    r = np.random.default_rng(5)
    x = r.normal(0, 20, (7, 158, 9)).astype(np.float32)
    assert np.array_equal(sess.float_to_int16(x), P.float_to_int16(x))


@pytest.mark.gpu
def test_mosaic_feats_gpu_bit_exact(sess):
    fl, xs, ys, want = _layers()
    got = sess.mosaic_feats(fl, xs, ys, want.shape[1:])
    bad = int((got != want).sum())
    print("feature mosaic mismatching values", bad, "of", want.size)
    assert got.dtype == np.int16 and got.shape == want.shape and bad == 0


@pytest.mark.gpu
def test_load_mosaic_predictions_depth_from_folder(sess, tmp_path):
    from sentinel_tree_cover_b200.api import load_mosaic_predictions
    feats, xs, ys = P.synth_subtile_feats(250, 158, 16, seed=8)
    d = str(tmp_path) + "/"
    for f, x, y in zip(feats, xs, ys):
        os.makedirs(d + str(x), exist_ok=True)
        np.save(d + str(x) + "/" + str(y) + ".npy", f)
    out = load_mosaic_predictions(d, 8, sess)                # depth smaller than the stored stacks: first 8 channels
    order = [(int(x), int(y[:-4])) for x in os.listdir(d) for y in os.listdir(d + x + "/")]
    lut = {(x, y): f for f, x, y in zip(feats, xs, ys)}
    want = P.mosaic_feats([lut[o] for o in order], [o[0] for o in order], [o[1] for o in order], (250, 250), 8)
    assert out.shape == (8, 250, 250) and np.array_equal(out, want)


@pytest.mark.gpu
def test_feature_taps_vs_oracle(sess, predict_weights):
    """early = GRU output (fp16 on the device, values in (-1, 1)), late = last block's sSE output; both against the
    float32 graph oracle.  Bounds: fp16 storage / fp16 conv operands, same budget as the intermediate taps of
    test_predict_small_vs_oracle_with_taps (4e-3 of the tensor's scale + 2e-3)."""
    from sentinel_tree_cover_b200.api import predict_subtile, PREDICT_EARLYFEATS, PREDICT_LATEFEATS
    x = P.synth_model_input(2, 44, 21)
    taps = {}
    ref = PredictRef(predict_weights).forward(x, taps=taps)
    probs, early, late = sess.predict_feats(x, length=4)
    assert np.abs(probs - ref).max() < 1e-3
    want_e = taps["gru"].numpy().transpose(0, 2, 3, 1)[:, 7:-7, 7:-7]
    want_l = taps["out"].numpy().transpose(0, 2, 3, 1)
    for name, got, want in (("early", early, want_e), ("late", late, want_l)):
        e, scale = np.abs(got - want).max(), np.abs(want).max()
        print("feature tap", name, "err", e, "scale", scale)
        assert got.shape == want.shape and e < 4e-3 * scale + 2e-3
    # reference-signature access (one subtile, centre crop to `size`)
    e1 = predict_subtile(x[0], sess, PREDICT_EARLYFEATS, 26)
    l1 = predict_subtile(x[0], sess, PREDICT_LATEFEATS, 26)
    assert e1.shape == (26, 26, 64) and np.abs(e1 - early[0, 2:-2, 2:-2]).max() < 2e-3
    assert l1.shape == (26, 26, 64) and np.abs(l1 - late[0, 2:-2, 2:-2]).max() < 4e-3 * np.abs(late).max() + 2e-3
