"""Storage codecs, Sentinel-1 dB transform and monthly compositing (P1, P2, P12 of SURVEY 8a)."""
import numpy as np
import pytest
from oracle import refshim
from sentinel_tree_cover_b200 import regrid


def _ref_to_int16(a):            # src/tof/tof_downloading.py:51-61
    return np.trunc(np.clip(a, 0, 1) * 65535).astype(np.uint16)


def _ref_to_float32(a):          # :64-72
    return np.divide(np.float32(a), 65535.)


def _ref_db(x, min_db):          # src/download_and_predict_job.py:74-89
    x = 10 * np.log10(x + 1 / 65535)
    x[x < -min_db] = -min_db
    x = (x + min_db) / min_db
    return np.clip(x, 0, 1)


def _ref_s1(s1, dates):          # src/tof/tof_downloading.py:75-95 with the date logic of regrid.py
    from oracle import preproc_ref as P
    G, _ = regrid.regrid_matrix(dates)
    s24 = P.regrid_apply(G, s1)
    return np.median(s24.reshape((12, 2) + s1.shape[1:]), axis=1).astype(np.float32)


@pytest.mark.skipif(not refshim.available(), reason="reference tree absent")
def test_restatements_equal_reference_functions(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    tofd = refshim.ref("tof.tof_downloading")
    job = refshim.ref("download_and_predict_job")
    r = np.random.default_rng(0)
    a = r.uniform(0, 1, (4, 9, 9, 2)).astype(np.float32)
    u = r.integers(0, 65536, (5, 7, 3)).astype(np.uint16)
    assert np.array_equal(tofd.to_int16(a), _ref_to_int16(a))
    assert np.array_equal(tofd.to_float32(u), _ref_to_float32(u))
    assert np.array_equal(job.convert_to_db(a.copy(), 22), _ref_db(a.copy(), 22))
    dates = np.array([6, 18, 30, 42, 54, 90, 126, 162, 198, 234, 270, 306, 342])
    s1 = r.uniform(0, 1, (13, 6, 6, 2)).astype(np.float32)
    assert np.abs(tofd.process_sentinel_1_tile(s1, dates) - _ref_s1(s1, dates)).max() < 1e-6


@pytest.mark.gpu
def test_codecs_gpu_bit_exact(sess):
    from sentinel_tree_cover_b200.api import to_float32, to_int16
    r = np.random.default_rng(1)
    u = r.integers(0, 65536, (3, 33, 31, 10)).astype(np.uint16)
    u[0, 0, 0, 0] = 65535
    f = to_float32(u, sess)
    assert f.dtype == np.float32 and np.array_equal(f, _ref_to_float32(u))
    a = r.uniform(0, 1, (2, 45, 17, 4)).astype(np.float32)
    a.flat[:3] = (0.0, 1.0, 0.5)
    assert np.array_equal(to_int16(a, sess), _ref_to_int16(a))
    assert np.array_equal(to_int16(f, sess), u)                 # round trip of every representable value class


@pytest.mark.gpu
def test_convert_to_db_and_s1_monthly(sess):
    from sentinel_tree_cover_b200.api import convert_to_db, process_sentinel_1_tile
    r = np.random.default_rng(2)
    x = r.uniform(0, 1, (12, 40, 40, 2)).astype(np.float32)
    x.flat[:2] = (0.0, 1.0)
    got = convert_to_db(x, 22, sess)
    assert np.abs(got - _ref_db(x.copy(), 22)).max() < 2e-6     # log10f: <= 2 ulp
    dates = np.array([6, 18, 30, 42, 54, 90, 126, 162, 198, 234, 270, 306, 342])
    s1 = r.uniform(0, 1, (13, 24, 20, 2)).astype(np.float32)
    assert np.abs(process_sentinel_1_tile(s1, dates, sess) - _ref_s1(s1, dates)).max() < 2e-6
