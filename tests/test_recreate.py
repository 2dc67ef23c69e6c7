"""Re-mosaic of a tile after the border pass (`resegment.recreate_resegmented_tifs`, `mosaic_subtiles`, host logic) against
outputs of the reference's own functions (/root/reference/src/resegment_tiles_wide.py:1169-1547 run through
oracle/refshim.py by tools/make_golden_recreate.py) on seeded folders of subtile predictions: normal subtiles only, all four
kinds of border strips, the 618-px tile with the 206 x 670 strips the border pass writes; no-data subtiles / pixels.
Bit-exact against the reference function run live (same NumPy expressions on the same layer order: files are listed in
os.listdir order as the reference does); 1e-4 percent against the stored golden (listing order may differ between machines)."""
import importlib.util
import os
import numpy as np
import pytest

HERE = os.path.dirname(__file__)
G = np.load(os.path.join(HERE, "golden", "recreate.npz"))
spec = importlib.util.spec_from_file_location("mk_recreate", os.path.join(HERE, "..", "tools", "make_golden_recreate.py"))
MK = importlib.util.module_from_spec(spec); spec.loader.exec_module(MK)


@pytest.mark.parametrize("case", MK.CASES, ids=[c[0] for c in MK.CASES])
def test_recreate_matches_reference_golden(case, tmp_path):
    from sentinel_tree_cover_b200 import resegment as R
    name, size, shape = case[0], case[1], case[2]
    folder = str(tmp_path) + "/"
    MK.write_case(folder, case)
    preds, sums = R.recreate_resegmented_tifs(folder, shape, size=size)
    assert preds.shape == (shape[1], shape[0]) and preds.dtype == np.float64
    st = MK.SAMPLE[0] if preds.size > 200000 else MK.SAMPLE[1]
    # the layer order follows os.listdir, which may differ between the file system the golden was made on and this one: the
    # float32 layer sums then round differently in the last bits (the live comparison below is exact); no-data pixels are exact
    got, want = preds[::st[0], ::st[1]], G[name + "_preds_sample"]
    assert np.array_equal(got == 255, want == 255)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-4)
    np.testing.assert_allclose(sums[::st[0], ::st[1]], G[name + "_sums_sample"], rtol=1e-6, atol=1e-9)
    chk = G[name + "_check"]
    assert float((preds == 255).sum()) == chk[2] and abs(preds.sum() - chk[0]) < 1e-6 * chk[0] and abs(np.nansum(sums) - chk[1]) < 1e-6 * chk[1]
    valid = preds[preds < 255]
    assert valid.size > 0 and valid.min() >= 0 and valid.max() <= 100 + 1e-9


def test_recreate_against_reference_function_when_mounted(tmp_path):
    """Where /root/reference is mounted: the whole arrays, not the stored sample."""
    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference not mounted")
    from sentinel_tree_cover_b200 import resegment as R
    m = refshim.ref("resegment_tiles_wide")
    case = MK.CASES[0]
    folder = str(tmp_path) + "/"
    MK.write_case(folder, case)
    m.SIZE = case[1]
    want_p, want_s = m.recreate_resegmented_tifs(folder, case[2])
    got_p, got_s = R.recreate_resegmented_tifs(folder, case[2], size=case[1])
    assert np.array_equal(got_p, want_p) and np.array_equal(got_s, want_s, equal_nan=True)


def test_resize_linear_is_the_published_skimage_algorithm():
    """Enlarging / same size: plain grid-mode zoom; shrinking: the Gaussian pre-filter first (scikit-image >= 0.19 default)."""
    import scipy.ndimage as ndi
    from sentinel_tree_cover_b200 import resegment as R
    r = np.random.default_rng(4)
    a = r.uniform(0, 1, (40, 64))
    assert np.array_equal(R.resize_linear(a, (40, 64)), a)
    up = R.resize_linear(a, (80, 128))
    assert np.array_equal(up, ndi.zoom(a, 2, order=1, mode="mirror", grid_mode=True))
    down = R.resize_linear(a, (40, 16))
    want = ndi.zoom(ndi.gaussian_filter(a, (0, 1.5), mode="mirror"), (1, 0.25), order=1, mode="mirror", grid_mode=True)
    assert np.array_equal(down, want) and not np.array_equal(down, R.resize_linear(a, (40, 16), anti_aliasing=False))


def test_mosaic_subtiles_weights():
    """The blending weight of a border kind: zero on the far half, rising to the shared edge, feathered where flagged."""
    from sentinel_tree_cover_b200 import resegment as R
    X = Y = 420
    def run(kind, **flags):
        P = np.full((X, Y, 2), 50, np.float32); M = np.ones((X, Y, 2), np.float32)
        f = {k: (P if flags.get(k) else None) for k in ("left", "right", "up", "down")}
        return R.mosaic_subtiles(P, M, np.zeros((X, Y, 1)), kind, f["left"], f["right"], f["up"], f["down"], size=400)
    p, m = run("r")
    assert np.allclose(p, 50) and np.all(m[:X - 200] == 0) and np.all(np.diff(m[X - 200:, 7]) > 0) and m.max() < 1
    p, m = run("l")
    assert np.all(m[200:] == 0) and np.all(np.diff(m[:200, 7]) < 0)
    _, mu = run("u"); _, md = run("d")
    assert np.all(mu[:, 200:] == 0) and np.all(md[:, :Y - 200] == 0) and mu[:, :200].max() > 0.9 and md[:, Y - 200:].max() > 0.9
    _, mf = run("r", up=True)
    assert np.all(mf[X - 100:, 0] == 0) and np.array_equal(mf[:, 300:], run("r")[1][:, 300:])
    assert np.allclose(R.adjust_resegment(np.full((3, 3), 2.0), np.full((3, 3, 4), 0.75), 3), 4.5)


def test_seam_smooth_diff_and_write_out(tmp_path):
    """:1763-1816: the seam statistic of two re-mosaicked neighbours, the acceptance rule and the _SMOOTH_X / _SMOOTH_XY names."""
    from sentinel_tree_cover_b200 import resegment as R
    r = np.random.default_rng(21)
    left = r.uniform(0, 100, (618, 618)); right = r.uniform(0, 100, (618, 618))
    left[-8:, 5] = 255; right[:3, 9] = 255; left[-8:, 11] = 255; right[:8, 11] = 255
    l8, r8 = left[-8:].astype(np.float32), right[:8].astype(np.float32)
    l8[l8 == 255] = np.nan; r8[r8 == 255] = np.nan
    cols = [c for c in range(618) if c not in (5, 11)]
    want = np.mean([abs(np.nanmean(r8[:, c]) - np.nanmean(l8[:, c])) for c in cols])
    got = R.seam_smooth_diff(left, right)
    assert abs(got - want) < 1e-4
    assert np.isnan(R.seam_smooth_diff(np.full((20, 4), 255.), right[:, :4]))
    folder = str(tmp_path) + "/"
    box, nbox = [10.0, 5.0, 10.0 + 1 / 18, 5.0 + 1 / 18], [10.0 + 1 / 18, 5.0, 10.0 + 2 / 18, 5.0 + 1 / 18]
    assert R.write_smoothed_pair(left, right, box, nbox, "12", "34", folder, folder, smooth_diff=50.0, diff=10.0) is None
    files = R.write_smoothed_pair(left, right, box, nbox, "12", "34", folder, folder, smooth_diff=25.0, diff=10.0)
    assert [os.path.basename(f) for f in files] == ["12X34Y_SMOOTH_X.tif", "13X34Y_SMOOTH_X.tif"] and all(os.path.getsize(f) > 1000 for f in files)
    open(folder + "12X34Y_SMOOTH_Y.tif", "wb").close()
    files = R.write_smoothed_pair(left, right, box, nbox, "12", "34", folder, folder, smooth_diff=float("nan"), diff=float("nan"))
    assert [os.path.basename(f) for f in files] == ["12X34Y_SMOOTH_XY.tif", "13X34Y_SMOOTH_X.tif"]
    from PIL import Image
    im = np.array(Image.open(files[0]))
    assert np.array_equal(im, left.T.astype(np.uint8))


def test_load_tif_picks_the_product_like_the_reference(tmp_path):
    """:713-751: _SMOOTH_XY > _SMOOTH_X > _SMOOTH_Y > _FINAL > _POST; band 1 read by libstc's TIFF reader; the file the border
    pass itself wrote (write_smoothed_pair) comes back pixel for pixel."""
    from sentinel_tree_cover_b200 import api, resegment as R
    r = np.random.default_rng(2)
    root = str(tmp_path)
    folder = os.path.join(root, "12", "34") + "/"
    os.makedirs(folder)
    box = [10.0, 5.0, 10.0 + 1 / 18, 5.0 + 1 / 18]
    maps = {s: r.integers(0, 101, (60, 64)).astype(np.uint8) for s in ("_POST", "_FINAL", "_SMOOTH_Y", "_SMOOTH_X", "_SMOOTH_XY")}
    expect_flag = {"_POST": 0, "_FINAL": 0, "_SMOOTH_Y": 1, "_SMOOTH_X": 0, "_SMOOTH_XY": 1}
    with pytest.raises(IndexError):
        R.load_tif(("12", "34"), root)
    for s in ("_POST", "_FINAL", "_SMOOTH_Y", "_SMOOTH_X", "_SMOOTH_XY"):         # each new product outranks the ones before
        api.write_tif(maps[s], box, "12", "34", folder, s)
        got, flag = R.load_tif(("12", "34"), root)
        assert np.array_equal(got, maps[s].T) and flag == expect_flag[s], s
    a = np.zeros((3, 10, 30, 4)); b = np.zeros((3, 10, 24, 4))
    x, y = R.concatenate_s2_files(a, b)
    assert x.shape[2] == 24 and y.shape[2] == 24
    x, y = R.concatenate_s2_files(b, a)
    assert x.shape[2] == 24 and y.shape[2] == 24


def test_load_tif_against_reference_function_with_a_stub_rasterio(tmp_path):
    """Where the reference is mounted: its load_tif with `rasterio.open(f).read(1)` replaced by Pillow picks the same file."""
    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference not mounted")
    from PIL import Image
    from sentinel_tree_cover_b200 import api, resegment as R
    m = refshim.ref("resegment_tiles_wide")

    class _DS:
        def __init__(self, f): self.f = f
        def read(self, band): return np.array(Image.open(self.f))
    m.rasterio.open = lambda f: _DS(f)
    r = np.random.default_rng(3)
    root = str(tmp_path)
    box = [10.0, 5.0, 10.0 + 1 / 18, 5.0 + 1 / 18]
    for case, names in enumerate((["_POST"], ["_POST", "_FINAL"], ["_FINAL", "_SMOOTH_Y"], ["_FINAL", "_SMOOTH_X", "_SMOOTH_Y"],
                                  ["_FINAL", "_SMOOTH_X", "_SMOOTH_XY"])):
        folder = os.path.join(root, str(case), "7") + "/"
        os.makedirs(folder)
        for s in names:
            api.write_tif(r.integers(0, 101, (40, 44)).astype(np.uint8), box, case, 7, folder, s)
        want, want_flag = m.load_tif((str(case), "7"), root)
        got, flag = R.load_tif((str(case), "7"), root)
        assert np.array_equal(got, want) and flag == want_flag, names
