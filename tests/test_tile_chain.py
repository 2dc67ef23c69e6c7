"""The whole-tile, device-resident chain (stc_tile_run_host, csrc/stc_tile.cu): the body of the reference's main loop
(src/download_and_predict_job.py:1995-2020) in one C call.

CPU: the C++ host logic of the chain (date regridding operator, window tables, adjust_shape plans) against the NumPy
mirrors regrid.py / windows.py -- which tests/test_oracle_preproc.py and tests/test_host_logic.py pin to the reference's
own functions -- and against the reference's adjust_shape itself where the reference tree is mounted.
GPU: the chain against the stage-by-stage mirrors (tile.process_tile -> api.superresolve_large_tile ->
tile.process_subtiles -> StcSession.mosaic), which are pinned to outputs of the reference's own functions
(tests/golden/process_tile.npz, process_subtiles.npz, mosaic.npz)."""
import ctypes as C
import os
import random
import numpy as np
import pytest
from oracle import tile_ref, refshim


def _lib():
    from sentinel_tree_cover_b200 import api
    return api.load_library()


def test_monthly_operator_plan_matches_numpy_mirror():
    from sentinel_tree_cover_b200 import api, regrid
    lib = _lib()
    rng = np.random.default_rng(0)
    n_ok = 0
    for trial in range(200):
        n = int(rng.integers(1, 25))
        if trial % 3 == 0:
            d = np.sort(rng.integers(-40, 400, n))                       # duplicates, previous / next year
        elif trial % 3 == 1:
            d = np.sort(rng.choice(np.arange(0, 365), n, replace=False))
        else:
            d = 15 + 30 * np.arange(n) + rng.integers(-5, 6, n)
        d = np.ascontiguousarray(d, np.int32)
        G = np.zeros((24, n), np.float32)
        M = np.zeros((12, n), np.float32)
        rc = lib.stc_monthly_operator_plan(api._dptr(d), n, api._dptr(G), api._dptr(M))
        try:
            Gp, _ = regrid.regrid_matrix(d)
            Mp, _ = regrid.monthly_operator(d)
        except Exception:
            assert rc != 0, d                                            # both refuse the same date sets
            continue
        assert rc == 0, d
        assert np.array_equal(G, Gp), d                                  # float32 weights: bit-exact
        assert np.abs(M - Mp).max() <= 1e-7, d                           # float64 Whittaker inverse: LAPACK vs Gauss-Jordan
        n_ok += 1
    assert n_ok > 100


def test_subtile_windows_plan_matches_numpy_mirror():
    from sentinel_tree_cover_b200 import api, windows
    lib = _lib()
    for Lx, Ly, size, rows in [(618, 618, 158, 6), (600, 618, 158, 6), (316, 316, 158, 6), (620, 600, 158, 6), (340, 330, 158, 6),
                               (618, 618, 142, 6), (700, 690, 222, 7)]:
        f = np.zeros((64, 4), np.int32)
        a = np.zeros((64, 4), np.int32)
        nt = lib.stc_subtile_windows_plan(Lx, Ly, size, rows, api._dptr(f), api._dptr(a), 64)
        fp, ap = windows.subtile_windows(Lx, Ly, size, rows)
        assert nt == len(fp)
        assert np.array_equal(f[:nt], fp) and np.array_equal(a[:nt], ap), (Lx, Ly, size)


def test_subtile_table_survey_constants():
    """window starts and array windows at L = 618, S = 158."""
    from sentinel_tree_cover_b200 import windows
    from sentinel_tree_cover_b200.tile import subtile_table
    folder, arr = windows.subtile_windows(618, 618, 158, 6)
    assert sorted(set(folder[:, 0].tolist())) == [0, 92, 184, 276, 368, 460]
    assert arr[1].tolist() == [0, 85, 165, 172] and arr[7].tolist() == [85, 85, 172, 172] and arr[-1].tolist() == [453, 453, 165, 165]
    t = subtile_table(arr, 618, 618, 158)
    assert t.shape == (36, 12)
    assert (t[:, 2] + t[:, 4] + t[:, 5] == 172).all() and (t[:, 3] + t[:, 6] + t[:, 7] == 172).all()
    assert t[0, 4:8].tolist() == [7, 0, 7, 0] and t[-1, 4:8].tolist() == [0, 7, 0, 7] and t[7, 4:8].tolist() == [0, 0, 0, 0]


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_adjust_shape_plan_matches_reference():
    """out[i] = in[clip(i + shift)] with the plan's length == the reference's adjust_shape on every small case, including
    the ones where the reference leaves an axis off by one or two pixels."""
    job = refshim.ref("download_and_predict_job")
    lib = _lib()
    for L in range(8, 22):
        for target in range(8, 22):
            a = np.arange(3 * L * 11, dtype=np.float32).reshape(3, L, 11, 1)
            want = job.adjust_shape(a, target, 11)
            shift, out_len = C.c_int32(0), C.c_int32(0)
            assert lib.stc_adjust_shape_plan(L, target, C.byref(shift), C.byref(out_len)) == 0
            got = np.take(a, np.clip(np.arange(out_len.value) + shift.value, 0, L - 1), axis=1).squeeze()
            assert got.shape == want.shape and np.array_equal(got, want), (L, target)
            b = np.ascontiguousarray(a.transpose(0, 2, 1, 3))
            want2 = job.adjust_shape(b, 11, target)
            got2 = np.take(b, np.clip(np.arange(out_len.value) + shift.value, 0, L - 1), axis=2).squeeze()
            assert np.array_equal(got2, want2), (L, target)


def _mirror_chain(sess, store, root, seed, forest=None, urban=None):
    from sentinel_tree_cover_b200 import api, tile
    random.seed(seed)
    s2, dates, interp, s1, dem, cloudshad, snow = tile.process_tile(1, 2, None, "/nonexistent/", [0, 0, 1, 1], make_shadow=True, sess=sess,
                                                                    loader=store.load, exists=store.exists, forest_mask=forest, urban_mask=urban)
    s2 = api.superresolve_large_tile(np.ascontiguousarray(s2), sess)
    tile.process_subtiles(1, 2, s2, dates, interp, s1, dem, sess, [0, 0, 1, 1], 158, None, local_path=root, length=4)
    path = root + "1/2/processed/"
    files = sorted((int(fy), int(f[:-4])) for fy in os.listdir(path) for f in os.listdir(path + fy))      # ascending (x, y) = the chain's layer order
    preds = [np.load(f"{path}{fy}/{fx}.npy") for fy, fx in files]
    xs = [f[0] for f in files]
    ys = [f[1] for f in files]
    out = sess.mosaic(preds, xs, ys, (max(xs) + 158, max(ys) + 158))
    return out, np.asarray(dates), dict(zip(files, preds)), random.getstate()


@pytest.mark.gpu
@pytest.mark.parametrize("case", [dict(seed=21, n=8, h=170, w=170, with_clm=False, ragged=False),
                                  dict(seed=22, n=7, h=172, w=166, with_clm=True, ragged=True),
                                  dict(seed=91, n=12, h=309, w=309, with_clm=False, ragged=False)],      # the 618 x 618 px production size
                         ids=["340px", "ragged", "618px"])
def test_gpu_tile_chain_matches_stage_mirrors(sess, tmp_path, case):
    from sentinel_tree_cover_b200 import windows
    raw = tile_ref.synth_raw_tile(case["seed"], n=case["n"], h=case["h"], w=case["w"], with_clm=case["with_clm"], ragged=case["ragged"])
    store = tile_ref.FakeStore(raw)
    want, want_dates, want_sub, want_state = _mirror_chain(sess, store, str(tmp_path) + "/", 4)
    random.seed(4)
    clm = (raw["cloudmask"] != 0) if case["with_clm"] else None
    got, kept, sub = sess.run_tile(raw["s2_10"], raw["s2_20"], raw["s1"], raw["dem"], raw["s2_dates"], clm=clm, return_subtiles=True)
    assert np.array_equal(kept, want_dates)
    assert random.getstate() == want_state                       # same draws from Python's generator as the stage-by-stage path
    folder, _ = windows.subtile_windows(2 * case["h"], 2 * case["w"], 158, 6)
    worst = 0.0
    for t in range(len(folder)):
        w = want_sub[(int(folder[t][1]), int(folder[t][0]))]
        assert np.array_equal(sub[t] == 255, w == 255), t
        worst = max(worst, float(np.abs(sub[t] - w).max()))
    assert worst == 0.0, worst                                   # same kernels, deterministic reductions: identical
    assert got.shape == want.shape and got.dtype == np.uint8
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_gpu_tile_chain_is_repeatable_and_pooled(sess):
    """Two runs of the same tile give identical bytes, and the second run takes every scratch buffer from the pool."""
    raw = tile_ref.synth_raw_tile(23, n=6, h=160, w=164)
    random.seed(9)
    a, ka = sess.run_tile(raw["s2_10"], raw["s2_20"], raw["s1"], raw["dem"], raw["s2_dates"])
    before = sess.pool_info()
    random.seed(9)
    b, kb = sess.run_tile(raw["s2_10"], raw["s2_20"], raw["s1"], raw["dem"], raw["s2_dates"])
    after = sess.pool_info()
    assert np.array_equal(a, b) and np.array_equal(ka, kb)
    assert after["misses"] == before["misses"], (before, after)
    assert 0 < (a <= 100).mean() <= 1


@pytest.mark.gpu
def test_gpu_tile_chain_with_ancillary_rasters_matches_stage_mirrors(sess, tmp_path):
    """The forest / urban rasters set on the session (stc_set_ancillary_masks_host) reach the cloud masks of the one-call chain
    exactly as they reach the stage mirror (process_tile(forest_mask=, urban_mask=)), and they change the result."""
    from sentinel_tree_cover_b200 import api
    raw = tile_ref.synth_raw_tile(25, n=7, h=170, w=170)
    store = tile_ref.FakeStore(raw)
    H = W = 340
    r = np.random.default_rng(8)
    frst = np.zeros((22, 22), bool); frst[:9] = True
    urb = np.zeros((22, 22), bool); urb[4:10, 10:18] = r.random((6, 8)) < 0.7
    forest, urban = api.ancillary_masks_from_rasters(frst, urb, (H, W))
    want, want_dates, _, want_state = _mirror_chain(sess, store, str(tmp_path) + "/a/", 4, forest, urban)
    plain, _, _, _ = _mirror_chain(sess, store, str(tmp_path) + "/b/", 4)
    sess.set_ancillary_masks(forest, urban, (H, W))
    try:
        random.seed(4)
        got, kept = sess.run_tile(raw["s2_10"], raw["s2_20"], raw["s1"], raw["dem"], raw["s2_dates"])
    finally:
        sess.set_ancillary_masks(None, None)
    assert np.array_equal(kept, want_dates) and random.getstate() == want_state
    assert np.array_equal(got, want)
    random.seed(4)
    again, _ = sess.run_tile(raw["s2_10"], raw["s2_20"], raw["s1"], raw["dem"], raw["s2_dates"])
    assert np.array_equal(again, plain)                              # masks cleared: back to the shipped configuration
