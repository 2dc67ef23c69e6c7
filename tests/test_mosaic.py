"""Gaussian overlap-blend mosaic (load_mosaic_predictions, depth 1): oracle vs the golden
produced by the reference itself (CPU), and the CUDA path vs both (GPU, uint8 bit-exact)."""
import os
import numpy as np
import pytest
from conftest import golden
from oracle import preproc_ref as P

CASES = (("a", dict(seed=1)), ("b", dict(seed=2, all_nodata=(5,))))


def _layers(case, kw):
    g = golden("mosaic.npz")
    preds, xs, ys = P.synth_subtile_preds(618, 158, **kw)
    d = {(x, y): p for p, x, y in zip(preds, xs, ys)}
    order = g["order_" + case]                      # the reference's os.listdir walk order
    return [d[(int(x), int(y))] for x, y in order], [int(x) for x, _ in order], [int(y) for _, y in order], g["out_" + case]


@pytest.mark.parametrize("case,kw", CASES)
def test_mosaic_oracle_matches_reference_golden(case, kw):
    pl, xs, ys, want = _layers(case, kw)
    with np.errstate(all="ignore"):
        got = P.mosaic(pl, xs, ys, want.shape)
    assert got.dtype == np.uint8 and np.array_equal(got, want)


def test_fspecial_gauss_known_values():
    g = P.fspecial_gauss(158, 36)
    assert abs(g.min() - 0.0081025) < 1e-7 and abs(g[0, 0] - 0.0091459) < 1e-7 and g[78, 78] == 1.0
    assert np.array_equal(g, golden("preproc.npz")["gauss_158_36"])


@pytest.mark.gpu
@pytest.mark.parametrize("case,kw", CASES)
def test_mosaic_gpu_bit_exact(sess, case, kw):
    pl, xs, ys, want = _layers(case, kw)
    got = sess.mosaic(pl, xs, ys, want.shape)
    bad = int((got != want).sum())
    print("mosaic", case, "mismatching px", bad, "of", want.size)
    assert got.dtype == np.uint8 and bad == 0


@pytest.mark.gpu
def test_load_mosaic_predictions_from_folder(sess, tmp_path):
    from sentinel_tree_cover_b200.api import load_mosaic_predictions
    preds, xs, ys = P.synth_subtile_preds(618, 158, seed=4)
    d = str(tmp_path) + "/"
    for p, x, y in zip(preds, xs, ys):
        os.makedirs(d + str(x), exist_ok=True)
        np.save(d + str(x) + "/" + str(y) + ".npy", p)
    out = load_mosaic_predictions(d, 1, sess)
    order = [(int(x), int(y[:-4])) for x in os.listdir(d) for y in os.listdir(d + x + "/")]
    lut = {(x, y): p for p, x, y in zip(preds, xs, ys)}
    with np.errstate(all="ignore"):
        want = P.mosaic([lut[o] for o in order], [o[0] for o in order], [o[1] for o in order], (618, 618))
    assert out.shape == (618, 618) and np.array_equal(out, want)
