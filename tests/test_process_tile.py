"""process_tile front half (src/download_and_predict_job.py:640-997): the mirror in
sentinel_tree_cover_b200/tile.py against golden outputs of the reference function itself
(tests/golden/process_tile.npz, tools/make_golden_tile.py) on seeded synthetic raw tiles."""
import os
import random
import numpy as np
import pytest
from oracle import tile_ref, refshim

GOLD = os.path.join(os.path.dirname(__file__), "golden", "process_tile.npz")


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_adjust_shape_matches_reference():
    from sentinel_tree_cover_b200.tile import adjust_shape
    job = refshim.ref("download_and_predict_job")
    r = np.random.default_rng(0)
    for shp in [(3, 20, 24, 2), (3, 21, 24, 2), (3, 19, 27, 2), (3, 26, 18, 4), (20, 24), (23, 25), (3, 20, 24), (3, 17, 30)]:
        a = r.random(shp).astype(np.float32)
        for (wd, ht) in [(20, 24), (22, 22)]:
            try:
                want = job.adjust_shape(a, wd, ht)
            except Exception as e:
                with pytest.raises(type(e)):
                    adjust_shape(a, wd, ht)
                continue
            assert np.array_equal(adjust_shape(a, wd, ht), want), (shp, wd, ht)


def test_fake_store_keys():
    raw = tile_ref.synth_raw_tile(5, n=6, h=40, w=52, with_clm=True)
    st = tile_ref.FakeStore(raw)
    assert st.exists("/t/1/2/raw/clouds/cloudmask_1X2Y.hkl") and not st.exists("/t/1/2/raw/clouds/shadows_1X2Y.hkl")
    assert st.load("/t/1/2/raw/s2_20/1X2Y.hkl").shape == (6, 40, 52, 6)
    assert st.load("/t/1/2/raw/misc/dem_1X2Y.hkl").dtype == np.float32


def test_process_tile_requires_session():
    from sentinel_tree_cover_b200.tile import process_tile
    with pytest.raises(RuntimeError):
        process_tile(1, 2, None, "/x/", [0, 0, 1, 1])


@pytest.mark.gpu
@pytest.mark.parametrize("idx", [0, 1])
def test_gpu_process_tile_matches_reference_golden(sess, idx):
    from sentinel_tree_cover_b200.tile import process_tile
    g = np.load(GOLD)
    seed, n, h, w, with_clm, ragged, rseed = [int(v) for v in g["case_%d" % idx]]
    store = tile_ref.FakeStore(tile_ref.synth_raw_tile(seed, n, h, w, bool(with_clm), bool(ragged)))
    random.seed(rseed)
    s2, dates, interp, s1, dem, cloudshad, snow = process_tile(1, 2, None, "/nonexistent/", [0, 0, 1, 1], make_shadow=True, sess=sess,
                                                               loader=store.load, exists=store.exists)
    assert np.array_equal(np.asarray(dates), g["dates_%d" % idx])
    assert np.array_equal(np.packbits(np.asarray(cloudshad) > 0), g["cloudshad_%d" % idx])
    assert np.array_equal(interp, g["interp_%d" % idx])
    assert np.array_equal(np.asarray(snow).astype(np.int8), g["snow_%d" % idx]) and np.asarray(snow).dtype == np.int64
    assert np.array_equal(dem, g["dem_%d" % idx])                                    # median filter + /90: exact
    np.testing.assert_allclose(s1[:, ::2, ::2], g["s1_sub_%d" % idx], rtol=0, atol=2e-6)      # log10 in the dB transform
    assert s2.dtype == np.float32 and s2.min() >= 0 and s2.max() <= 1
    np.testing.assert_allclose(s2[:, ::3, ::3], g["s2_sub_%d" % idx], rtol=1e-4, atol=1e-6)   # NNLS-filled pixels
    assert abs(float(np.sum(s2, dtype=np.float64)) / float(g["s2_sum_%d" % idx][0]) - 1) < 1e-6
    assert random.random() == float(g["next_random_%d" % idx][0])
