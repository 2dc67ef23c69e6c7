"""Cloud / shadow removal (remove_cloud_and_shadows, cloud_removal.py:888-973).
CPU: oracle/cloudfill_ref.py reproduces the reference's outputs stored in tests/golden/cloud_fill.npz
(tools/make_golden_cloudfill.py ran the reference with a pinned random.seed).
GPU: feathered weights, mosaic, removal list and the advanced `random` state are exact; blended pixel
values agree to rtol 1e-4 (float64 NNLS with a different summation order than LAPACK)."""
import os
import random
import numpy as np
import pytest
from oracle import cloud_ref, cloudfill_ref as F, refshim

GOLD = os.path.join(os.path.dirname(__file__), "golden", "cloud_fill.npz")


def _inputs(T, H, W, seed):
    img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed)
    clouds, fcps = cloud_ref.identify_clouds_shadows(img, dem)
    return img, clouds, fcps


def _golden():
    g = np.load(GOLD)
    for i, (T, H, W, seed, rseed) in enumerate(g["cases"]):
        yield (int(T), int(H), int(W), int(seed), int(rseed)), {k[:-2]: g[k] for k in g.files if k.endswith("_%d" % i)}


def _check_against_golden(img, tiles, areas, to_remove, gold, exact_values):
    assert np.array_equal(areas, gold["areas"])
    assert list(to_remove) == gold["to_remove"].tolist()
    changed = (tiles != img).any(-1)
    assert int(changed.sum()) == int(gold["changed_count"][0])
    got = tiles[tuple(gold["sample_idx"].T)]
    if exact_values:
        assert np.array_equal(got, gold["sample_val"])
    else:
        np.testing.assert_allclose(got, gold["sample_val"], rtol=1e-4, atol=1e-6)
    assert random.random() == float(gold["next_random"][0])       # generator left where the reference leaves it


def test_oracle_matches_reference_golden():
    for (T, H, W, seed, rseed), gold in _golden():
        img, clouds, fcps = _inputs(T, H, W, seed)
        random.seed(rseed)
        tiles, areas, to_remove = F.remove_cloud_and_shadows(np.copy(img), np.copy(clouds), np.copy(fcps))
        _check_against_golden(img, tiles, areas, to_remove, gold, exact_values=True)


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_oracle_matches_reference_large_clear_branch(tmp_path, monkeypatch):
    """> 40000 clear pixels: the fit uses the date itself only (:385-392)."""
    monkeypatch.chdir(tmp_path)                       # the reference writes debug .npy files
    cr = refshim.ref("preprocessing.cloud_removal")
    img, clouds, fcps = _inputs(4, 212, 208, 45)
    random.seed(5)
    r_tiles, r_areas, r_rm = cr.remove_cloud_and_shadows(np.copy(img), np.copy(clouds), np.copy(clouds), list(range(4)), np.copy(fcps),
                                                         np.zeros((212, 208, 2), np.float32))
    random.seed(5)
    o_tiles, o_areas, o_rm = F.remove_cloud_and_shadows(np.copy(img), np.copy(clouds), np.copy(fcps))
    assert np.array_equal(r_tiles, o_tiles) and np.array_equal(r_areas, o_areas) and list(r_rm) == list(o_rm)


@pytest.mark.gpu
def test_gpu_remove_clouds_golden(sess):
    from sentinel_tree_cover_b200 import api
    for (T, H, W, seed, rseed), gold in _golden():
        img, clouds, fcps = _inputs(T, H, W, seed)
        tiles = np.copy(img)
        random.seed(rseed)
        out, areas, to_remove = api.remove_cloud_and_shadows(tiles, clouds, clouds, list(range(T)), fcps, None, sess=sess)
        assert out is tiles
        _check_against_golden(img, tiles, areas, to_remove, gold, exact_values=False)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(6, 230, 220, 42, 11), (9, 120, 100, 46, 12), (3, 80, 80, 47, 13)])
def test_gpu_remove_clouds_vs_oracle(sess, case):
    T, H, W, seed, rseed = case
    img, clouds, fcps = _inputs(T, H, W, seed)
    taps = {}
    random.seed(rseed)
    o_tiles, o_areas, o_rm = F.remove_cloud_and_shadows(np.copy(img), np.copy(clouds), np.copy(fcps), taps)
    o_next = random.random()
    random.seed(rseed)
    state = np.array(random.getstate()[1], dtype=np.uint32)
    tiles = np.copy(img)
    areas, rm, mosaic = sess.remove_clouds(tiles, clouds, fcps, state, want_mosaic=True)
    assert np.array_equal(mosaic, taps["mosaic"], equal_nan=True), int(np.sum(mosaic != taps["mosaic"]))   # bit-exact mosaic
    assert np.array_equal(areas, o_areas)
    assert rm == list(o_rm)
    untouched = ~(o_tiles != img).any(-1)
    assert np.array_equal(tiles[untouched], img[untouched])
    np.testing.assert_allclose(tiles, o_tiles, rtol=1e-4, atol=1e-6)
    random.setstate((3, tuple(int(v) for v in state), None))
    assert random.random() == o_next


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(6, 230, 220, 42, 11), (9, 120, 100, 46, 12)])
def test_gpu_warp_parallel_nnls_equals_the_serial_solver(sess, case, monkeypatch):
    """k_nnls (one warp per band, the elimination spread over the lanes) against k_nnls_serial (one thread per band, the
    direct transcription of scipy.optimize.nnls' Lawson-Hanson loop): the same operations on the same operands, so the
    blended tiles are bit-identical."""
    T, H, W, seed, rseed = case
    img, clouds, fcps = _inputs(T, H, W, seed)
    outs = []
    for serial in ("1", "0"):
        monkeypatch.setenv("STC_NNLS_SERIAL", serial)
        random.seed(rseed)
        state = np.array(random.getstate()[1], dtype=np.uint32)
        tiles = np.copy(img)
        areas, rm = sess.remove_clouds(tiles, clouds, fcps, state)
        outs.append((tiles, areas, rm, state))
    assert (outs[0][0] != img).any()                                  # the fit did change pixels
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]) and outs[0][2] == outs[1][2]
    assert np.array_equal(outs[0][3], outs[1][3])


@pytest.mark.gpu
def test_gpu_remove_clouds_no_clouds_is_identity(sess):
    img, _ = cloud_ref.synth_cloudy_cube(5, 64, 64, 50)
    tiles = np.copy(img)
    state = np.array(random.Random(1).getstate()[1], dtype=np.uint32)
    before = state.copy()
    areas, rm = sess.remove_clouds(tiles, np.zeros((5, 64, 64), np.float32), np.zeros((5, 64, 64), np.uint8), state)
    assert np.array_equal(tiles, img) and not areas.any() and rm == [] and np.array_equal(state, before)


def test_mirror_requires_session():
    from sentinel_tree_cover_b200 import api
    with pytest.raises(RuntimeError):
        api.remove_cloud_and_shadows(np.zeros((2, 8, 8, 10), np.float32), np.zeros((2, 8, 8)), None, [0, 1], np.zeros((2, 8, 8)), None)
