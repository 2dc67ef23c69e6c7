"""Exact order statistics (csrc/stc_select.cu, stc_order_stats_host): the select under every np.median / np.percentile of the
path (cloud_removal.py:455-467, 598-677, 1458-1481; download_and_predict_job.py:702-705).  Oracle: np.sort."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def want_pairs(a, ks):
    s = np.sort(a, axis=0)                      # NaN last, like the kernel's keys
    out = np.empty((a.shape[1], 2), np.float32)
    for c, k in enumerate(ks):
        out[c, 0] = s[k, c]
        out[c, 1] = s[min(k + 1, a.shape[0] - 1), c]
    return out


def same_bits(x, y):
    return np.array_equal(np.asarray(x, np.float32).view(np.uint32), np.asarray(y, np.float32).view(np.uint32))


@pytest.mark.parametrize("rows,cols", [(1, 1), (2, 1), (7, 3), (1000, 10), (70001, 10), (300000, 1), (65537, 16), (250 * 256 + 3, 7)])
def test_ranks_match_sort(sess, rows, cols):
    r = np.random.default_rng(rows * 31 + cols)
    a = r.normal(0.1, 0.2, (rows, cols)).astype(np.float32)
    a[r.random((rows, cols)) < 0.2] = np.float32(0.125)          # many equal keys: rank k + 1 repeats the value
    if rows > 5:
        a[::5, 0] = -a[::5, 0]
        a[3, :] = 0.0
        a[4, :] = -0.0
    for ks in ([0] * cols, [rows - 1] * cols, [rows // 2] * cols, list(r.integers(0, rows, cols))):
        got = sess.order_stats(a, ks)
        want = want_pairs(a, ks)
        # -0.0 and 0.0 compare equal in np.sort; the kernel orders them by key (-0 first): compare values, and bits where no zero is involved
        assert np.array_equal(got, want), (rows, cols, ks)
        nz = want != 0
        assert same_bits(got[nz], want[nz])


def test_nan_sorts_last_and_strided_rows(sess):
    r = np.random.default_rng(5)
    full = r.uniform(0, 1, (40000, 13)).astype(np.float32)
    view = full[:, 2:12]                                           # 10 columns inside rows of 13 floats
    valid = 30000
    full[valid:, :] = np.nan                                       # producers write NaN for "not selected"
    ks = [valid - 1, 0, valid // 2, 17, valid - 2, 1, 2, 3, 4, 5]
    got = sess.order_stats(view, ks)
    want = want_pairs(np.ascontiguousarray(view), ks)
    assert np.array_equal(got[1:], want[1:]) and got[0, 0] == want[0, 0] and np.isnan(got[0, 1]) and np.isnan(want[0, 1])


def test_successor_outside_the_bucket_of_the_last_digit(sess):
    """x_(k+1) may lie anywhere above x_(k): right next to it (same 22-bit bucket), in another bucket, or not exist."""
    base = np.float32(0.3)
    nxt = np.nextafter(base, np.float32(1), dtype=np.float32)
    a = np.array([0.1, 0.2, base, nxt, 0.9, 7.0e8, -3.0, -1e-30], np.float32)[:, None]
    for k in range(len(a)):
        got = sess.order_stats(a, [k])
        assert same_bits(got, want_pairs(a, [k])), k
    b = np.full((5000, 2), 0.25, np.float32)                       # one key only: both statistics are that key
    b[:, 1] = np.arange(5000, dtype=np.float32) * np.float32(1e-3)
    got = sess.order_stats(b, [2499, 2499])
    assert same_bits(got, want_pairs(b, [2499, 2499]))


def test_median_and_percentile_from_pairs(sess):
    """NumPy's even-length median and linear-interpolation percentile from the pair, as the library forms them."""
    r = np.random.default_rng(9)
    x = r.gamma(2.0, 0.05, 123456).astype(np.float32)
    n = len(x)
    a, b = sess.order_stats(x, [n // 2 - 1])[0]
    assert np.float32((a + b) / np.float32(2)) == np.median(x)
    for pct in (2, 25, 99):
        vi = np.float32(n - 1) * (np.float32(pct) / np.float32(100))
        lo = int(np.floor(vi)); g = np.float32(vi - np.float32(lo))
        a, b = sess.order_stats(x, [lo])[0]
        got = np.float32(a + (b - a) * g) if g < 0.5 else np.float32(b - (b - a) * (np.float32(1) - g))
        assert abs(got - np.percentile(x, pct)) <= 1e-7 * max(1.0, abs(got))


def test_argument_checks(sess):
    a = np.zeros((10, 3), np.float32)
    with pytest.raises(Exception):
        sess.order_stats(a, [10, 0, 0])
    with pytest.raises(Exception):
        sess.order_stats(np.zeros((10, 17), np.float32), [0] * 17)
