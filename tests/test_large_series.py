"""Production-scale parity: 24-step series (BASELINE configs[2]) and 618 x 618 tiles.
tests/golden/large_t24.npz holds outputs of the REFERENCE's identify_clouds_shadows (cloud_removal.py:1215-1677) and
remove_cloud_and_shadows (:888-973) on seeded cubes with T in {24, 17, 12} and up to 618 x 618 px (tools/make_golden_large.py
ran them through oracle/refshim.py).  The date-window logic (:1266-1273, 1352-1363) depends on T; the small fixtures of
test_cloud_masks.py / test_cloud_fill.py stop at T = 12 and 230 px.
CPU: the oracle restatement equals the golden on the cases it finishes in seconds.
GPU: masks bit-identical on every case (incl. 24 x 618 x 618); removal: feather weights / removal list / generator state
exact, filled values rtol 1e-4."""
import os
import random
import numpy as np
import pytest
from oracle import cloud_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "large_t24.npz")


def _mask_case(g, i):
    T, H, W, seed = [int(v) for v in g["mask_cases"][i]]
    n = T * H * W
    clouds = np.unpackbits(g["masks_clouds_%d" % i])[:n].reshape(T, H, W)
    fcps = np.unpackbits(g["masks_fcps_%d" % i])[:n].reshape(T, H, W)
    return T, H, W, seed, clouds, fcps


@pytest.mark.parametrize("i", [0, 2])
def test_oracle_masks_match_reference_T24_T17(i):
    g = np.load(GOLD)
    T, H, W, seed, clouds, fcps = _mask_case(g, i)
    img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed)
    c, f = cloud_ref.identify_clouds_shadows(img, dem)
    assert np.array_equal(c > 0, clouds > 0)
    assert np.array_equal(np.asarray(f) > 0, fcps > 0)


@pytest.mark.gpu
@pytest.mark.parametrize("i", [0, 1, 2])
def test_gpu_masks_match_reference_large(sess, i):
    from sentinel_tree_cover_b200 import api
    g = np.load(GOLD)
    T, H, W, seed, clouds, fcps = _mask_case(g, i)
    img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed)
    c, f = api.identify_clouds_shadows(img, dem, None, sess)
    assert np.array_equal(c > 0, clouds > 0), ("clouds", T, H, W, int(np.sum((c > 0) != (clouds > 0))))
    assert np.array_equal(f, fcps > 0), ("fcps", T, H, W)


@pytest.mark.gpu
@pytest.mark.parametrize("i", [0, 1])
def test_gpu_remove_clouds_matches_reference_large(sess, i):
    from sentinel_tree_cover_b200 import api
    g = np.load(GOLD)
    T, H, W, seed, rseed = [int(v) for v in g["fill_cases"][i]]
    img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed)
    clouds, fcps = api.identify_clouds_shadows(img, dem, None, sess)       # bit-identical to the reference (test above)
    tiles = np.copy(img)
    random.seed(rseed)
    out, areas, to_remove = api.remove_cloud_and_shadows(tiles, clouds, clouds, list(range(T)), fcps, None, sess=sess)
    assert out is tiles
    areas = np.asarray(areas, np.float32)
    assert np.array_equal(areas[:, ::5, ::5].astype(np.float16), g["fill_areas_sub_%d" % i])
    s = g["fill_areas_sum_%d" % i]
    assert float(areas.sum(dtype=np.float64)) == s[0] and float((areas > 0).sum()) == s[1] and float((areas == 1).sum()) == s[2]
    assert list(to_remove) == g["fill_to_remove_%d" % i].tolist()
    changed = (tiles != img).any(-1)
    assert int(changed.sum()) == int(g["fill_changed_count_%d" % i][0])
    got = tiles[tuple(g["fill_sample_idx_%d" % i].T)]
    np.testing.assert_allclose(got, g["fill_sample_val_%d" % i], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(float(tiles.sum(dtype=np.float64)), float(g["fill_tiles_sum_%d" % i][0]), rtol=1e-6)
    assert random.random() == float(g["fill_next_random_%d" % i][0])
