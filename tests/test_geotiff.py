"""write_tif (src/downloading/io.py:229-263): the uint8 LZW GeoTIFF product, written by libstc's own encoder
(csrc/stc_geotiff.cpp) and read back here by libtiff -- the decoder GDAL / rasterio use -- through Pillow and OpenCV.
rasterio is not installable in this image, so byte-identity with a GDAL-written file is NOT claimed (GDAL's files differ
between versions anyway); what is pinned: every pixel survives the LZW round trip through an independent decoder,
and the georeferencing tags decode to the transform / CRS `rasterio.transform.from_bounds` + EPSG:4326 describe.
CPU-only: the writer is host code and needs no session."""
import os
import struct
import numpy as np
import pytest

from sentinel_tree_cover_b200 import api

Image = pytest.importorskip("PIL.Image")


def _read_pil(path):
    with Image.open(path) as im:
        return np.array(im), dict(im.tag_v2)


def _cases():
    r = np.random.default_rng(0)
    yield "constant", np.full((618, 618), 7, np.uint8)
    yield "noise", r.integers(0, 256, (618, 640)).astype(np.uint8)            # incompressible: dictionary resets inside strips
    yield "tree_cover", np.repeat(np.repeat(r.integers(0, 101, (70, 70)), 9, 0), 9, 1).astype(np.uint8)
    yield "nodata", np.where(r.random((206, 670)) < 0.3, 255, r.integers(0, 101, (206, 670))).astype(np.uint8)
    yield "one_pixel", np.array([[5]], np.uint8)
    yield "wide_row", r.integers(0, 3, (3, 9000)).astype(np.uint8)            # a row longer than the 8 KiB strip target
    yield "binary", r.integers(0, 2, (700, 700)).astype(np.uint8)


@pytest.mark.parametrize("name,img", list(_cases()), ids=[c[0] for c in _cases()])
def test_write_tif_round_trips_through_libtiff(name, img, tmp_path):
    point = [10.25, -3.5, 10.25 + 0.061, -3.5 + 0.0607]                       # [west, south, east, north]
    f = api.write_tif(img.T, point, 1203, 977, str(tmp_path) + "/")           # the reference hands over the TRANSPOSED tile
    assert f == str(tmp_path) + "/1203X977Y_FINAL.tif" and os.path.exists(f) and not os.path.exists(f + ".part")
    got, tags = _read_pil(f)
    assert got.dtype == np.uint8 and np.array_equal(got, img)
    cv2 = pytest.importorskip("cv2")
    got2 = cv2.imread(f, cv2.IMREAD_UNCHANGED)
    assert got2 is not None and np.array_equal(got2.reshape(img.shape), img)
    rows, cols = img.shape
    assert tags[256] == cols and tags[257] == rows and tags[258] == (8,) and tags[259] == 5          # LZW
    assert tags[262] == 1 and tags[277] == 1 and tags[284] == 1 and tags[339] == (1,)
    assert tags[278] == max(1, min(rows, 8192 // cols))                                              # GDAL's strip geometry
    # rasterio.transform.from_bounds: a = (east - west) / width, e = -(north - south) / height, origin (west, north)
    np.testing.assert_allclose(tags[33550], ((point[2] - point[0]) / cols, (point[3] - point[1]) / rows, 0.0), rtol=1e-15)
    assert tags[33922] == (0.0, 0.0, 0.0, point[0], point[3], 0.0)
    keys = np.array(tags[34735]).reshape(-1, 4)
    assert tuple(keys[0]) == (1, 1, 0, len(keys) - 1)
    kd = {int(k[0]): tuple(int(v) for v in k[1:]) for k in keys[1:]}
    assert kd[1024] == (0, 1, 2) and kd[1025] == (0, 1, 1) and kd[2048] == (0, 1, 4326) and kd[2054] == (0, 1, 9102)
    assert tags[34736] == (6378137.0, 298.257223563) and tags[34737].startswith("WGS 84")
    assert api.geotiff_bytes(img, point) == open(f, "rb").read()


def test_lzw_code_width_steps_and_table_resets_at_every_length(tmp_path):
    """Strip lengths around the points where the 9 -> 10 -> 11 -> 12 bit steps and the dictionary reset (4094 codes)
    fall on the LAST code of a strip: the decoder steps its width one code after the encoder's table does."""
    r = np.random.default_rng(5)
    lib = api.load_library()
    for n in list(range(250, 262)) + list(range(760, 775)) + list(range(1785, 1800)) + list(range(3830, 3845)) + [1, 2, 3, 8191, 8192]:
        for kind in (0, 1):
            row = (r.integers(0, 256, n) if kind == 0 else (np.arange(n) * 7 % 251)).astype(np.uint8)[np.newaxis]
            f = str(tmp_path / ("w%d_%d.tif" % (n, kind)))
            assert lib.stc_write_geotiff_u8(os.fsencode(f), api._dptr(row), 1, n, 0.0, 0.0, 1.0, 1.0) == 0
            got, _ = _read_pil(f)
            assert np.array_equal(got.reshape(row.shape), row), (n, kind)


def test_geotiff_header_is_classic_little_endian_tiff_and_errors_are_reported(tmp_path):
    img = np.arange(12, dtype=np.uint8).reshape(3, 4)
    b = api.geotiff_bytes(img, [0, 0, 4e-4, 3e-4])
    assert b[:4] == b"II*\x00"
    ifd = struct.unpack("<I", b[4:8])[0]
    n = struct.unpack("<H", b[ifd:ifd + 2])[0]
    tags = [struct.unpack("<H", b[ifd + 2 + 12 * i: ifd + 4 + 12 * i])[0] for i in range(n)]
    assert tags == sorted(tags) and len(b) == ifd + 2 + 12 * n + 4                # ascending tags, IFD last, no trailing bytes
    lib = api.load_library()
    p = api._dptr(img)
    assert lib.stc_write_geotiff_u8(os.fsencode(str(tmp_path / "a.tif")), p, 3, 4, 1.0, 0.0, 1.0, 1.0) == -2    # west == east
    assert lib.stc_write_geotiff_u8(os.fsencode(str(tmp_path / "a.tif")), p, 0, 4, 0.0, 0.0, 1.0, 1.0) == -2
    assert lib.stc_write_geotiff_u8(os.fsencode(str(tmp_path / "no_such_dir" / "a.tif")), p, 3, 4, 0.0, 0.0, 1.0, 1.0) == -3
    with pytest.raises(RuntimeError):
        api.write_tif(img, [0, 0, 1, 1], 1, 2, str(tmp_path / "no_such_dir") + "/")
