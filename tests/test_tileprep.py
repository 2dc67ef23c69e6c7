"""Tile-prep kernels of process_tile's front half (csrc/stc_tileprep.cu) against NumPy / SciPy directly."""
import numpy as np
import pytest


@pytest.mark.gpu
def test_s1_fill_matches_numpy(sess):
    from sentinel_tree_cover_b200 import tile
    r = np.random.default_rng(0)
    raw = np.trunc(r.uniform(0.01, 0.6, (5, 40, 36, 2)) * 65535).astype(np.uint16)
    raw[r.random(raw.shape) < 0.01] = 65535
    raw[3] = np.minimum(raw[3], 65534)                                    # a date without saturated values
    want = np.float32(raw) / 65535
    for i in range(want.shape[0]):                                        # :702-705
        s = want[i]
        s[s == 1] = np.median(s[s < 65535], axis=0)
        want[i] = s
    got = tile.s1_fill(sess.to_float32(raw), sess)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_median_filter5_matches_scipy(sess):
    from scipy.ndimage import median_filter
    from sentinel_tree_cover_b200 import tile
    r = np.random.default_rng(1)
    for shp in [(40, 52), (7, 9), (5, 5), (3, 64)]:
        d = r.normal(100, 30, shp).astype(np.float32)
        assert np.array_equal(tile.median_filter5(d, sess), median_filter(d, size=5)), shp


@pytest.mark.gpu
def test_clm_pairs_matches_reference_loop(sess):
    from sentinel_tree_cover_b200 import tile
    r = np.random.default_rng(2)
    clm = (r.random((9, 30, 34)) < 0.35).astype(np.float32)
    want = clm.copy()
    for i in range(want.shape[0]):                                        # :688-695
        mins, maxs = np.maximum(i - 1, 0), np.minimum(i + 1, want.shape[0])
        sums = np.sum(want[mins:maxs], axis=0) == 2
        want[mins:maxs, sums] = 0.
    assert np.array_equal(tile.clm_pairs(clm, sess), want)


@pytest.mark.gpu
def test_snow_mask_counts_and_dilation(sess):
    from scipy.ndimage import binary_dilation
    from oracle.cloudfill_ref import snow_filter
    from sentinel_tree_cover_b200 import tile
    r = np.random.default_rng(3)
    s2 = r.uniform(0.02, 0.6, (6, 48, 44, 10)).astype(np.float32)
    s2[:, 10:30, 5:25, 8] *= 0.1                                           # low SWIR -> high NDSI -> snow
    s2[:, 10:30, 5:25, 0] += 0.2
    s2[:, 10:30, 5:25, 2] = s2[:, 10:30, 5:25, 0]                          # B2/B4 = 1
    s2[:, 10:30, 5:25, 3] = 0.4                                            # bright NIR
    ndsis = snow_filter(np.copy(s2)) > 0
    per_date, snow = tile.snow_mask(s2, sess)
    assert np.array_equal(per_date, ndsis.sum(axis=(1, 2)))
    want = 1 - binary_dilation(np.mean(ndsis, axis=0) < 0.7, iterations=2)
    assert snow.dtype == np.int64 and np.array_equal(snow, want)
    assert 0 < want.sum() < want.size


@pytest.mark.gpu
def test_counts_clip_divide_max(sess):
    from sentinel_tree_cover_b200 import tile
    r = np.random.default_rng(4)
    a = r.normal(0.3, 0.5, (5, 33, 31)).astype(np.float32)
    assert np.array_equal(tile.count_gt(a, 0.0, sess), (a > 0).reshape(5, -1).sum(1))
    assert np.array_equal(tile.count_lt_axis0(a, 0.33, sess), np.sum(a < 0.33, axis=0))
    b = a.copy(); b[0, 0, 0] = np.nan
    assert np.array_equal(tile.clip01(b, sess), np.clip(b, 0, 1), equal_nan=True)
    assert np.array_equal(tile.divide(a, 90, sess), a / 90)
    assert np.array_equal(tile.nan_to_zero(b, sess), np.nan_to_num(b, nan=0.0))
    m = (r.random(a.shape) < 0.4).astype(np.float32)
    z = r.random(a.shape) < 0.2
    want = m.copy(); want[z] = 0.
    assert np.array_equal(tile.max_masked(np.abs(a), m, z, sess), np.maximum(np.abs(a), want))
