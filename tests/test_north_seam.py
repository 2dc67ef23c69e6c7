"""Host-only functions of the NORTH-seam variant of the border pass (`edge="up"` in sentinel_tree_cover_b200/resegment.py)
against outputs of /root/reference/src/resegment_tiles_north_wide.py (tools/make_golden_north.py through oracle/refshim.py):
the window table, the seam-artifact test with that file's thresholds, the re-mosaic with its feather exponent, load_tif's
flag.  The north file's array functions (regularize_and_smooth, align_dates, recreate_resegmented_tifs, adjust_*) are
line-identical to the east file's; its preprocess_tile is `edge="up"` too (host logic checked below against both files'
functions with the oracle standing in for the GPU calls); its process_subtiles differs in its rules and is not mirrored."""
import importlib.util
import os
import numpy as np
import pytest

HERE = os.path.dirname(__file__)
G = np.load(os.path.join(HERE, "golden", "north.npz"))
spec = importlib.util.spec_from_file_location("mk_north", os.path.join(HERE, "..", "tools", "make_golden_north.py"))
MK = importlib.util.module_from_spec(spec); spec.loader.exec_module(MK)


def test_north_window_table_and_artifact_test():
    from sentinel_tree_cover_b200 import resegment as R
    ta, tf = R.make_tiles_right_neighb(np.array([0, 138, 276, 412]), np.array([0]), 670, 206, edge="up")
    assert np.array_equal(ta, G["tiles_array"]) and np.array_equal(tf, G["tiles_folder"])
    got = [R.check_if_artifact(t, n, edge="up") for t, n in MK.artifact_inputs()]
    assert got == G["artifact"].tolist() and 0 in got and 1 in got


def test_north_recreate_and_load_tif(tmp_path):
    from sentinel_tree_cover_b200 import api, resegment as R
    case = MK.MKR.CASES[0]
    folder = str(tmp_path) + "/m/"
    os.makedirs(folder)
    MK.MKR.write_case(folder, case)
    preds, sums = R.recreate_resegmented_tifs(folder, case[2], size=case[1], edge="up")
    st = MK.MKR.SAMPLE[1]
    got, want = preds[::st[0], ::st[1]], G["recreate_preds_sample"]
    assert np.array_equal(got == 255, want == 255)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-4)                 # layer order follows os.listdir (see test_recreate.py)
    np.testing.assert_allclose(sums[::st[0], ::st[1]], G["recreate_sums_sample"], rtol=1e-6, atol=1e-9)
    east, _ = R.recreate_resegmented_tifs(folder, case[2], size=case[1])
    assert not np.array_equal(east, preds)                                     # the feather exponent matters on this case
    r = np.random.default_rng(3)
    root = str(tmp_path) + "/t"
    flags = []
    for case_i, names in enumerate(MK.TIF_SETS):
        d = os.path.join(root, str(case_i), "7") + "/"
        os.makedirs(d)
        for s in names:
            api.write_tif(r.integers(0, 101, (40, 44)).astype(np.uint8), [10.0, 5.0, 10.1, 5.1], case_i, 7, d, s)
        flags.append(R.load_tif((str(case_i), "7"), root, edge="up")[1])
    assert flags == G["load_tif_flags"].tolist()


def test_north_functions_against_live_reference(tmp_path):
    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference not mounted")
    import contextlib, io
    from sentinel_tree_cover_b200 import resegment as R
    m = refshim.ref("resegment_tiles_north_wide")
    m.x, m.y = 0, 0
    r = np.random.default_rng(99)
    n_art = 0
    for k in range(40):
        w = 2 * int(r.integers(150, 320))             # (an odd common width makes the reference's 10-bin reshape fail: see below)
        tile = r.uniform(0, 100, (30, w)).astype(np.float32)
        nb = tile[::-1, :w - 2 * int(r.integers(0, 5))] + r.normal(0, float(r.uniform(0.5, 30)), (30, 1)).astype(np.float32) * r.standard_normal((30, 1)).astype(np.float32)
        nb = nb + r.normal(0, float(r.uniform(0.5, 25)), nb.shape).astype(np.float32) + (r.uniform(-9, 9) if k % 4 == 0 else 0)
        tile[r.random(tile.shape) < 0.03] = np.nan
        with contextlib.redirect_stdout(io.StringIO()):
            want = m.check_if_artifact(tile, nb)
        assert R.check_if_artifact(tile, nb, edge="up") == want, k
        n_art += want
    assert 0 < n_art < 40
    odd = r.uniform(0, 100, (30, 301)).astype(np.float32)
    for fn in (lambda: m.check_if_artifact(odd, odd), lambda: R.check_if_artifact(odd, odd, edge="up")):
        with pytest.raises(ValueError), contextlib.redirect_stdout(io.StringIO()):
            fn()                                      # both fail alike on a width the padding cannot bring to a multiple of 10
    case = MK.MKR.CASES[0]
    folder = str(tmp_path) + "/"
    MK.MKR.write_case(folder, case)
    m.SIZE = case[1]
    with contextlib.redirect_stdout(io.StringIO()):
        want_p, want_s = m.recreate_resegmented_tifs(folder, case[2])
    got_p, got_s = R.recreate_resegmented_tifs(folder, case[2], size=case[1], edge="up")
    assert np.array_equal(got_p, want_p) and np.array_equal(got_s, want_s, equal_nan=True)


@pytest.mark.parametrize("edge", ["right", "up"])
def test_preprocess_tile_host_logic_against_live_reference(edge, tmp_path, monkeypatch):
    """The HOST logic of resegment.preprocess_tile (which dates are screened out, how the Sen2Cor mask is merged, the second
    mask pass after a removal) with the four GPU calls replaced by the oracle's NumPy restatements -- so the comparison with
    the reference function (east file :619-672, north file :573-625) isolates the glue and runs without a GPU: a date missing
    11 % of its pixels (dropped by the east file's threshold H^2 / 20, kept by the north file's H^2 / 5), one missing 40 %
    (dropped by both), a Sen2Cor mask,
    an almost fully clouded date (> 95 % interpolated: removed, masks recomputed).  Bit-identical, incl. Python's generator."""
    import contextlib, io, random
    from oracle import refshim, cloud_ref, cloudfill_ref, upsample_ref
    if not refshim.available():
        pytest.skip("reference not mounted")
    from sentinel_tree_cover_b200 import api, resegment as R
    m = refshim.ref("resegment_tiles_wide" if edge == "right" else "resegment_tiles_north_wide")
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(api, "id_missing_px", lambda arr, thresh, sess: upsample_ref.id_missing_px(arr, thresh))
    monkeypatch.setattr(api, "identify_clouds_shadows", lambda img, dem, bbx, sess, forest_mask=None, urban_mask=None:
                        cloud_ref.identify_clouds_shadows(np.copy(img), np.copy(dem))[:2])
    monkeypatch.setattr(api, "id_areas_to_interp", lambda tiles, probs, shadows, dates, pfcps, sess:
                        cloudfill_ref.feather(np.clip(np.copy(probs).astype(np.float32), 0, 1), 15))
    monkeypatch.setattr(api, "remove_cloud_and_shadows", lambda tiles, probs, shadows, dates, pfcps, s1, mosaic=None, sess=None:
                        cloudfill_ref.remove_cloud_and_shadows(tiles, probs, pfcps)[:3])
    T, H, W = 9, 64, 64
    img, dem = cloud_ref.synth_cloudy_cube(T, H, W, 321)
    img[3, :, :7, :10] = 0.0                               # 11 % of the pixels missing: >= H^2 / 20 (east drops it), < H^2 / 5 (north keeps it)
    img[6, :26, :, :10] = 0.0                              # 40 %: dropped by both
    img[1, :, :, :3] = 0.9; img[1, :, :, 3:10] = 0.8       # a date that is cloud everywhere
    dts = (np.arange(T) * 36 + 10).astype(np.int64)
    clm = np.zeros((T, H, W), np.float32); clm[2, 10:30, 20:50] = 1.; clm[4, 40:60, 5:25] = 1.
    random.seed(5)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()), np.errstate(all="ignore"):
        want_arr, want_interp, want_dates = m.preprocess_tile(np.copy(img), np.copy(dts), None, np.copy(clm), "tile", np.copy(dem), None)
    want_next = random.random()
    random.seed(5)
    with np.errstate(all="ignore"):
        arr, interp, dates = R.preprocess_tile(np.copy(img), np.copy(dts), None, np.copy(clm), "tile", np.copy(dem), None, sess=None, edge=edge)
    assert np.array_equal(dates, want_dates) and len(dates) < T
    assert (dts[3] in dates) == (edge == "up") and dts[6] not in dates and dts[1] not in dates
    assert np.array_equal(np.asarray(interp), np.asarray(want_interp)) and np.array_equal(np.asarray(arr), np.asarray(want_arr))
    assert random.random() == want_next
