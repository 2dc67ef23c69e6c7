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


# ---- the reader (api.read_tif = rasterio.open(f).read(1) of a tile product, src/resegment_tiles_wide.py:750) ----------------

@pytest.mark.parametrize("name,img", list(_cases()), ids=[c[0] for c in _cases()])
def test_read_tif_round_trips_own_writer(name, img, tmp_path):
    point = [10.25, -3.5, 10.25 + 0.061, -3.5 + 0.0607]
    f = api.write_tif(img.T, point, 7, 8, str(tmp_path) + "/")
    got, bounds = api.read_tif(f, return_bounds=True)
    assert got.dtype == np.uint8 and np.array_equal(got, img)
    assert np.allclose(bounds, point, rtol=0, atol=1e-12)


def _third_party_files(tmp_path, img):
    """The same raster written by libtiff through Pillow (LZW, PackBits, uncompressed, tiled LZW, LZW + horizontal predictor)
    and by OpenCV (its default: LZW with the horizontal predictor)."""
    out = []
    im = Image.fromarray(img)
    for name, kw in (("pil_lzw", dict(compression="tiff_lzw")), ("pil_raw", dict(compression=None)), ("pil_packbits", dict(compression="packbits")),
                     ("pil_lzw_pred2", dict(compression="tiff_lzw", tiffinfo={317: 2})),
                     ("pil_lzw_tiled", dict(compression="tiff_lzw", tiffinfo={322: 128, 323: 64}))):
        f = os.path.join(str(tmp_path), name + ".tif")
        try:
            im.save(f, **{k: v for k, v in kw.items() if v is not None})
        except Exception:
            continue
        out.append((name, f))
    try:
        import cv2
        f = os.path.join(str(tmp_path), "cv_lzw.tif")
        if cv2.imwrite(f, img):
            out.append(("cv_lzw", f))
    except ImportError:
        pass
    return out


@pytest.mark.parametrize("name,img", [c for c in _cases() if c[0] in ("noise", "tree_cover", "nodata", "binary", "one_pixel")],
                         ids=["noise", "tree_cover", "nodata", "one_pixel", "binary"])
def test_read_tif_decodes_files_written_by_libtiff_and_opencv(name, img, tmp_path):
    files = _third_party_files(tmp_path, img)
    assert len(files) >= 3
    seen = set()
    for kind, f in files:
        with Image.open(f) as im:
            tags = dict(im.tag_v2)
            want = np.array(im)
        seen.add((tags.get(259), tags.get(317, 1), 322 in tags))
        got, bounds = api.read_tif(f, return_bounds=True)
        assert np.array_equal(got, want) and np.array_equal(got, img), (kind, tags.get(259), tags.get(317))
        assert all(np.isnan(b) for b in bounds)                            # no GeoTIFF tags in these
    assert any(c == 5 for c, _, _ in seen)                                     # at least one third-party LZW stream was decoded


def test_read_tif_multiband_big_endian_and_refusals(tmp_path):
    r = np.random.default_rng(5)
    rgb = r.integers(0, 256, (50, 70, 3)).astype(np.uint8)
    f = os.path.join(str(tmp_path), "rgb.tif")
    Image.fromarray(rgb).save(f, compression="tiff_lzw")
    for b in range(3):
        assert np.array_equal(api.read_tif(f, band=b + 1), rgb[..., b])
    with pytest.raises(RuntimeError):
        api.read_tif(f, band=4)
    # a hand-made big-endian, uncompressed, single-strip file
    img = r.integers(0, 256, (5, 7)).astype(np.uint8)
    ent = [(256, 3, 1, 7), (257, 3, 1, 5), (258, 3, 1, 8), (259, 3, 1, 1), (262, 3, 1, 1), (273, 4, 1, 8), (277, 3, 1, 1), (278, 3, 1, 5), (279, 4, 1, 35)]
    body = img.tobytes() + b"\0"
    ifd = struct.pack(">H", len(ent)) + b"".join(struct.pack(">HHI", t, ty, c) + (struct.pack(">HH", v, 0) if ty == 3 else struct.pack(">I", v))
                                                 for t, ty, c, v in ent) + struct.pack(">I", 0)
    be = os.path.join(str(tmp_path), "be.tif")
    open(be, "wb").write(b"MM" + struct.pack(">HI", 42, 8 + len(body)) + body + ifd)
    assert np.array_equal(api.read_tif(be), img)
    with Image.open(be) as im:
        assert np.array_equal(np.array(im), img)                               # libtiff agrees that this is a valid file
    # refusals: 16-bit samples, deflate, truncated file, not a TIFF
    f16 = os.path.join(str(tmp_path), "u16.tif"); Image.fromarray(r.integers(0, 60000, (9, 9)).astype(np.uint16)).save(f16)
    fz = os.path.join(str(tmp_path), "z.tif"); Image.fromarray(img).save(fz, compression="tiff_adobe_deflate")
    ft = os.path.join(str(tmp_path), "trunc.tif"); open(ft, "wb").write(open(f, "rb").read()[:200])
    fn = os.path.join(str(tmp_path), "no.tif"); open(fn, "wb").write(b"not a tiff at all")
    for bad in (f16, fz, ft, fn, os.path.join(str(tmp_path), "missing.tif")):
        with pytest.raises(RuntimeError):
            api.read_tif(bad)


def test_read_tif_tiled_layout(tmp_path):
    """Pillow does not write tiles; GDAL does with TILED=YES.  A hand-made tiled file (16 x 16 tiles, image 40 x 50: partial
    tiles on both edges), uncompressed and with LZW tiles from libstc's own encoder, checked against libtiff's reading."""
    r = np.random.default_rng(8)
    img = r.integers(0, 256, (40, 50)).astype(np.uint8)
    th = tw = 16
    down, across = -(-40 // th), -(-50 // tw)
    tiles = []
    for by in range(down):
        for bx in range(across):
            t = np.zeros((th, tw), np.uint8)
            part = img[by * th:(by + 1) * th, bx * tw:(bx + 1) * tw]
            t[:part.shape[0], :part.shape[1]] = part
            tiles.append(t)
    for compression in (1, 5):
        blobs = []
        for t in tiles:
            if compression == 1:
                blobs.append(t.tobytes())
            else:                                         # one LZW stream per tile: a single-strip TIFF of the tile from the writer
                enc = api.geotiff_bytes(t, [0, 0, 1, 1])
                off, cnt = [struct.unpack_from("<I", enc, enc.index(struct.pack("<HHI", tag, 4, 1)) + 8)[0] for tag in (273, 279)]
                blobs.append(enc[off:off + cnt])
        body, offs = b"", []
        for b in blobs:
            offs.append(8 + len(body)); body += b + (b"\0" if len(b) & 1 else b"")
        n = len(blobs)
        off_arr, cnt_arr = 8 + len(body), 8 + len(body) + 4 * n
        body += struct.pack("<%dI" % n, *offs) + struct.pack("<%dI" % n, *[len(b) for b in blobs])
        ent = [(256, 3, 1, 50), (257, 3, 1, 40), (258, 3, 1, 8), (259, 3, 1, compression), (262, 3, 1, 1), (277, 3, 1, 1),
               (322, 3, 1, tw), (323, 3, 1, th), (324, 4, n, off_arr), (325, 4, n, cnt_arr)]
        ifd = struct.pack("<H", len(ent)) + b"".join(struct.pack("<HHI", t, ty, c) + (struct.pack("<HH", v, 0) if ty == 3 else struct.pack("<I", v))
                                                     for t, ty, c, v in ent) + struct.pack("<I", 0)
        f = os.path.join(str(tmp_path), "tiled_%d.tif" % compression)
        open(f, "wb").write(b"II" + struct.pack("<HI", 42, 8 + len(body)) + body + ifd)
        with Image.open(f) as im:
            assert np.array_equal(np.array(im), img)                           # libtiff reads the hand-made file
        assert np.array_equal(api.read_tif(f), img), compression


def test_tiff_codec_fuzz_against_libtiff(tmp_path):
    """60 random rasters (noise, binary, constant rows, long runs, sparse): libstc writer -> libstc reader, libstc writer ->
    libtiff, libtiff writer (LZW, predictor 1 / 2) -> libstc reader.  Exercises the code-width steps and table resets of both
    LZW directions on strips of every length."""
    r = np.random.default_rng(123)
    for k in range(60):
        h, w = int(r.integers(1, 300)), int(r.integers(1, 900))
        mode = k % 5
        if mode == 0:
            img = r.integers(0, 256, (h, w))
        elif mode == 1:
            img = r.integers(0, 2, (h, w)) * 255
        elif mode == 2:
            img = np.repeat(r.integers(0, 101, (h, 1)), w, 1)
        elif mode == 3:
            img = (np.arange(h * w).reshape(h, w) // int(r.integers(1, 50))) % int(r.integers(2, 256))
        else:
            img = np.where(r.random((h, w)) < 0.9, 0, r.integers(0, 256, (h, w)))
        img = img.astype(np.uint8)
        f = os.path.join(str(tmp_path), "own.tif")
        open(f, "wb").write(api.geotiff_bytes(img, [0, 0, 1, 1]))
        assert np.array_equal(api.read_tif(f), img), (k, h, w)
        with Image.open(f) as im:
            assert np.array_equal(np.array(im), img), (k, h, w)
        f2 = os.path.join(str(tmp_path), "pil.tif")
        Image.fromarray(img).save(f2, compression="tiff_lzw", tiffinfo={317: 1 + (k % 2)})
        assert np.array_equal(api.read_tif(f2), img), (k, h, w)
