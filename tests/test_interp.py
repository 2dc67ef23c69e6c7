"""P3 (20 m -> 10 m bilinear band stack, process_tile :743-782) and P4 (id_missing_px /
deal_w_missing_px) -- oracle pinned against the reference where it can run, CUDA vs oracle bit-exact."""
import numpy as np
import pytest
from oracle import upsample_ref as U, refshim


def _cube_with_sentinels(n, H, W, C, seed, p0=0.02, p1=0.01, kill_date=None):
    r = np.random.default_rng(seed)
    a = r.uniform(0.01, 0.9, (n, H, W, C)).astype(np.float32)
    a[r.random(a.shape) < p0] = 0.0
    a[r.random(a.shape) < p1] = 1.0
    if kill_date is not None:
        a[kill_date, : H // 2] = 0.0
    return a


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_oracle_missing_px_matches_reference():
    job = refshim.ref("download_and_predict_job")
    interp_mod = refshim.ref("preprocessing.interpolation")
    for seed, kill in [(1, None), (2, 3), (3, 0)]:
        a = _cube_with_sentinels(9, 24, 24, 10, seed, kill_date=kill)
        dates = np.arange(9) * 30
        interp = np.zeros((9, 24, 24), np.float32)
        assert np.array_equal(interp_mod.id_missing_px(a, 10), U.id_missing_px(a, 10))
        r_arr, r_dates, r_interp = job.deal_w_missing_px(np.copy(a), np.copy(dates), np.copy(interp))
        o_arr, o_dates, o_interp = U.deal_w_missing_px(np.copy(a), np.copy(dates), np.copy(interp))
        assert np.array_equal(r_arr, o_arr) and np.array_equal(r_dates, o_dates) and r_interp.shape == o_interp.shape


def test_oracle_resize_is_half_pixel_mirror_bilinear():
    x = np.array([1., 2., 4., 8., 16.], np.float32)
    assert np.array_equal(U.resize_bilinear(x, (10,)), np.array([1.25, 1.25, 1.75, 2.5, 3.5, 5., 7., 10., 14., 14.], np.float32))
    r = np.random.default_rng(0)
    a, b = r.random((2, 22, 26, 4)).astype(np.float32), r.random((2, 11, 13, 6)).astype(np.float32)
    o = U.build_sentinel2(a, b)
    assert np.array_equal(o[..., :4], a)
    assert np.array_equal(o[:, 0, :, 8], b[:, 0, :, 4].repeat(2, axis=1))     # odd rows: first 40 m row copied
    assert np.array_equal(o[:, :, 0, 9], b[:, :, 0, 5].repeat(2, axis=1))


@pytest.mark.gpu
@pytest.mark.parametrize("hw", [(10, 12), (11, 13), (11, 12), (10, 13), (84, 90), (3, 2)])
def test_gpu_build_sentinel2_bit_exact(sess, hw):
    from sentinel_tree_cover_b200 import api
    h, w = hw
    r = np.random.default_rng(h * 100 + w)
    a = (r.random((3, 2 * h, 2 * w, 4)) * r.choice([1e-3, 1.0], (3, 2 * h, 2 * w, 4))).astype(np.float32)
    b = (r.random((3, h, w, 6)) * r.choice([1e-3, 1.0, 50.0], (3, h, w, 6))).astype(np.float32)
    got = api.build_sentinel2(a, b, sess)
    want = U.build_sentinel2(a, b)
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got, want), int(np.sum(got != want))


@pytest.mark.gpu
def test_gpu_missing_px_and_median_fill(sess):
    from sentinel_tree_cover_b200 import api
    for seed, kill, n in [(1, None, 9), (2, 3, 9), (3, 0, 12), (4, None, 2), (5, None, 1), (6, 5, 24)]:
        a = _cube_with_sentinels(n, 40, 40, 10, seed, kill_date=kill if kill is None or kill < n else None)
        if seed == 3:
            a[2, 5, 5, 3] = np.nan                       # NaN date must be dropped after the fill
        dates = np.arange(n) * 15
        interp = np.zeros((n, 40, 40), np.float32)
        assert np.array_equal(api.id_missing_px(a, 10, sess), U.id_missing_px(a, 10))
        assert np.array_equal(api.id_missing_px(a, 2, sess), U.id_missing_px(a, 2))
        g_arr, g_dates, g_interp = api.deal_w_missing_px(np.copy(a), np.copy(dates), np.copy(interp), sess)
        o_arr, o_dates, o_interp = U.deal_w_missing_px(np.copy(a), np.copy(dates), np.copy(interp))
        assert np.array_equal(g_dates, o_dates)
        assert np.array_equal(g_arr, o_arr, equal_nan=True), int(np.sum(g_arr != o_arr))
        assert g_interp.shape == o_interp.shape


@pytest.mark.gpu
def test_gpu_median_fill_in_place_and_idempotent(sess):
    a = _cube_with_sentinels(7, 32, 32, 14, 11)
    b = np.copy(a)
    sess.median_fill(b)
    assert not np.any(b == 0.0) or np.any(np.median(b, axis=0) == 0.0)
    c = np.copy(b)
    sess.median_fill(c)
    assert np.array_equal(b, c)                              # nothing left to fill
    with pytest.raises(ValueError):
        sess.median_fill(a.astype(np.float64))


def test_bilinear_x2_oracle_interior_matches_opencv_and_pillow():
    """scikit-image is absent (P3 unpinned against skimage itself); two third-party bilinear resizers with the same
    pixel-centre alignment are here.  Away from the one-pixel border (where skimage's mode='reflect' mirrors and OpenCV /
    Pillow replicate the edge) the x2 upsampling of the oracle equals both to float32 rounding."""
    import cv2
    from PIL import Image
    from oracle import upsample_ref as U
    r = np.random.default_rng(0)
    for (h, w) in ((37, 41), (64, 30)):
        a = r.uniform(0, 1, (h, w)).astype(np.float32)
        want = U.resize_bilinear(a, (2 * h, 2 * w))
        got_cv = cv2.resize(a, (2 * w, 2 * h), interpolation=cv2.INTER_LINEAR)
        got_pil = np.array(Image.fromarray(a, mode="F").resize((2 * w, 2 * h), Image.BILINEAR))
        assert np.abs(got_cv[1:-1, 1:-1] - want[1:-1, 1:-1]).max() < 5e-7
        assert np.abs(got_pil[1:-1, 1:-1] - want[1:-1, 1:-1]).max() < 5e-7
        # the border follows ndimage 'mirror' (d c b | a b c d): 0.75 * a[0] + 0.25 * a[1] per axis, not the replicated edge
        row = 0.75 * a[0] + 0.25 * a[1]
        assert abs(want[0, 0] - (0.75 * row[0] + 0.25 * row[1])) < 1e-6
