"""process_subtiles (src/download_and_predict_job.py:1125-1486): the batched mirror against the
reference function's own per-subtile .npy outputs (tests/golden/process_subtiles.npz; the reference ran
with its TensorFlow call replaced by the oracle graph restatement, tools/make_golden_subtiles.py)."""
import os
import numpy as np
import pytest
from oracle import subtiles_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "process_subtiles.npz")


def test_golden_layout():
    g = np.load(GOLD)
    assert len(g["names"]) == 36
    assert g["pred_0_0"].shape == (158, 158) and g["pred_0_0"].dtype == np.float32


@pytest.mark.gpu
@pytest.mark.parametrize("host_gather", [False, True])
def test_gpu_process_subtiles_matches_reference_golden(sess, tmp_path, monkeypatch, host_gather):
    """Both routes: windows gathered on the device (default) and the statement-by-statement host gather."""
    from sentinel_tree_cover_b200.tile import process_subtiles
    if host_gather:
        monkeypatch.setenv("STC_TILE_HOST_GATHER", "1")
    g = np.load(GOLD)
    seed, n, H, W = [int(v) for v in g["case"]]
    s2, dates, interp, s1, dem = subtiles_ref.synth_ard(seed, n, H, W)
    root = str(tmp_path) + "/"
    process_subtiles(3, 4, s2, dates, interp, s1, dem, sess, [0, 0, 1, 1], 158, None, local_path=root, length=4)
    path = root + "3/4/processed/"
    got_names = sorted((int(fy), int(f[:-4])) for fy in os.listdir(path) for f in os.listdir(path + fy))
    assert got_names == sorted(map(tuple, g["names"].tolist()))
    worst = 0.0
    for fy, fx in got_names:
        got = np.load(f"{path}{fy}/{fx}.npy")
        want = g["pred_%d_%d" % (fy, fx)]
        assert got.dtype == np.float32 and got.shape == (158, 158)
        assert np.array_equal(got == 255, want == 255), (fy, fx)          # no-data block votes: exact
        # attenuated no-data (255 * ramp) and probabilities rounded to 3 decimals
        m = want < 2
        worst = max(worst, float(np.abs(got[m] - want[m]).max()) if m.any() else 0.0)
        assert np.allclose(got[~m], want[~m], rtol=0, atol=1e-3), (fy, fx)
    assert worst <= 1.5e-3, worst      # 1e-3 model tolerance + half a rounding step
