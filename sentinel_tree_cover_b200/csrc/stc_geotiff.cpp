// Single-band uint8 GeoTIFF writer with LZW strips: the on-disk form of the reference's `write_tif`
// (/root/reference/src/downloading/io.py:229-263 -- rasterio `driver='GTiff', count=1, dtype='uint8', compress='lzw',
// crs='+proj=longlat +datum=WGS84 +no_defs', transform=from_bounds(west, south, east, north, width, height)`), called once per
// tile after load_mosaic_predictions (/root/reference/src/download_and_predict_job.py:2033-2036) and by the re-segmentation pass
// (/root/reference/src/resegment_tiles_wide.py:1240 ff).  Host-only code: no device work, no session.
//
// Layout written (classic little-endian TIFF 6.0, one IFD):
//   header | LZW strips (GDAL's default geometry: as many rows per strip as fit 8 KiB, at least one) | tag payloads | IFD
// Geo-referencing follows the GeoTIFF 1.0 encoding GDAL emits for a north-up EPSG:4326 raster: ModelPixelScale
// ((east-west)/width, (north-south)/height, 0), ModelTiepoint (0,0,0) -> (west, north, 0), GeoKeyDirectory {GTModelType =
// geographic, GTRasterType = PixelIsArea, GeographicType = 4326, GeogCitation "WGS 84", angular unit degree, WGS-84 ellipsoid}.
// The LZW stream is TIFF's variant: MSB-first codes of 9..12 bits, Clear = 256, EOI = 257, the code width grows ONE code early
// ("early change", what libtiff writes and every TIFF reader expects), a Clear code opens every strip and is re-issued when the
// table reaches 4094 entries.
#include "../../include/stc.h"
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <string>
#include <thread>
#include <vector>

namespace {

// MSB-first bit packer writing through a raw pointer (the caller reserves the worst case: 12 bits per input byte + resets).
struct BitSink {
  uint8_t* p;
  uint64_t acc = 0;
  int nbits = 0;
  explicit BitSink(uint8_t* dst) : p(dst) {}
  inline void put(uint32_t code, int width) {
    acc = (acc << width) | code;
    nbits += width;
    while (nbits >= 8) {
      *p++ = uint8_t(acc >> (nbits - 8));
      nbits -= 8;
    }
  }
  inline uint8_t* flush() {
    if (nbits > 0) *p++ = uint8_t(acc << (8 - nbits));
    nbits = 0;
    return p;
  }
};

// Dictionary: open-addressed hash of (prefix code, next byte) -> code.  At most 3836 entries in 8192 slots.  A slot is live
// when its generation stamp equals the table's, so the reset at every strip start and at every Clear code is one increment
// instead of a 32 KB memset (a 618-pixel-wide tile has a strip every 13 rows).
struct LzwTable {
  static constexpr int SLOTS = 8192;
  uint32_t tag[SLOTS];      // generation << 20 | prefix << 8 | byte
  uint16_t val[SLOTS];
  uint32_t gen = 0;
  void clear() {
    if (gen == 0 || gen == 0xfff) { memset(tag, 0, sizeof(tag)); gen = 0; }
    ++gen;
  }
  static inline uint32_t slot(uint32_t k) { return (k * 2654435761u) >> 19; }   // 13 bits
  inline int find(uint32_t k, uint32_t& s) const {
    const uint32_t want = (gen << 20) | k;
    s = slot(k);
    while ((tag[s] >> 20) == gen) {
      if (tag[s] == want) return val[s];
      s = (s + 1) & (SLOTS - 1);
    }
    return -1;
  }
  inline void insert(uint32_t s, uint32_t k, int code) { tag[s] = (gen << 20) | k; val[s] = uint16_t(code); }
};

// Worst-case encoded size of an n-byte strip: every byte its own 12-bit code, a Clear every 3836 codes, Clear + EOI, padding.
inline size_t lzw_bound(size_t n) { return n + n / 2 + (n / 3836 + 4) * 2 + 8; }

size_t lzw_encode_strip(const uint8_t* src, size_t n, uint8_t* dst, LzwTable& tab) {
  constexpr int CLEAR = 256, EOI = 257, FIRST = 258, LAST = 4094;   // the table is reset when code 4094 has been assigned
  BitSink bits(dst);
  int width = 9, next = FIRST;
  tab.clear();
  bits.put(CLEAR, width);
  if (n == 0) { bits.put(EOI, width); return size_t(bits.flush() - dst); }
  int prefix = src[0];
  for (size_t i = 1; i < n; ++i) {
    const uint32_t c = src[i];
    const uint32_t k = (uint32_t(prefix) << 8) | c;
    uint32_t s;
    const int hit = tab.find(k, s);
    if (hit >= 0) { prefix = hit; continue; }
    bits.put(uint32_t(prefix), width);
    tab.insert(s, k, next);
    ++next;
    // "early change": the decoder's table runs one entry behind the encoder's and steps its code width when ITS next free
    // code is 511 / 1023 / 2047 -- i.e. when the encoder's is 512 / 1024 / 2048 (libtiff's LZWEncode does the same).
    if (next == LAST) {
      bits.put(CLEAR, width);
      tab.clear();
      width = 9; next = FIRST;
    } else if (next == 512 || next == 1024 || next == 2048) {
      ++width;
    }
    prefix = int(c);
  }
  bits.put(uint32_t(prefix), width);
  // the decoder adds a table entry after this last code as well, so its width may step (or its table fill up) before EOI
  ++next;
  if (next == LAST) { bits.put(CLEAR, width); width = 9; }
  else if (next == 512 || next == 1024 || next == 2048) ++width;
  bits.put(EOI, width);
  return size_t(bits.flush() - dst);
}

struct Entry { uint16_t tag, type; uint32_t count; uint32_t value; };

template <class T> void append(std::vector<uint8_t>& buf, const T* p, size_t n) {
  const uint8_t* b = reinterpret_cast<const uint8_t*>(p);
  buf.insert(buf.end(), b, b + n * sizeof(T));
}
inline void pad_even(std::vector<uint8_t>& buf) { if (buf.size() & 1) buf.push_back(0); }

}  // namespace

// STC_OK; STC_ERR_ARG bad arguments; STC_ERR_NOMEM allocation failure.  *out_buf is malloc'ed (release with stc_geotiff_free).
extern "C" int stc_geotiff_encode_u8(const uint8_t* img, int rows, int cols, double west, double south, double east, double north,
                                     uint8_t** out_buf, int64_t* out_len) {
  if (!img || !out_buf || !out_len || rows < 1 || cols < 1) return STC_ERR_ARG;
  if (!(east > west) || !(north > south)) return STC_ERR_ARG;
  if (int64_t(rows) * cols > (int64_t(1) << 31)) return STC_ERR_ARG;          // classic TIFF: 32-bit offsets
  try {
    std::vector<uint8_t> f;
    f.reserve(size_t(rows) * cols / 2 + 65536);
    const uint8_t hdr[8] = {'I', 'I', 42, 0, 0, 0, 0, 0};             // IFD offset patched below
    append(f, hdr, 8);

    int rps = 8192 / cols;
    if (rps < 1) rps = 1;
    if (rps > rows) rps = rows;
    const int nstrips = (rows + rps - 1) / rps;
    std::vector<uint32_t> offs(nstrips), lens(nstrips);
    // Strips are independent LZW streams (each opens with a Clear code), and the encoder is a serial dependency chain through
    // its dictionary (~10 ns per byte): encode them on up to 8 host threads into per-strip slots, then pack the file in order.
    const size_t slot_bytes = lzw_bound(size_t(rps) * cols);
    std::vector<uint8_t> scratch(slot_bytes * size_t(nstrips));
    std::vector<size_t> enc_len(nstrips);
    auto encode_range = [&](std::atomic<int>& next_strip) {
      static thread_local LzwTable tab;
      for (int s = next_strip.fetch_add(1); s < nstrips; s = next_strip.fetch_add(1)) {
        const int r0 = s * rps, nr = (r0 + rps <= rows) ? rps : rows - r0;
        enc_len[s] = lzw_encode_strip(img + size_t(r0) * cols, size_t(nr) * cols, scratch.data() + slot_bytes * size_t(s), tab);
      }
    };
    {
      std::atomic<int> next_strip{0};
      unsigned nt = std::thread::hardware_concurrency();
      if (nt > 8) nt = 8;
      if (nt > unsigned(nstrips)) nt = unsigned(nstrips);
      if (size_t(rows) * cols < (size_t(1) << 16) || nt < 2) {
        encode_range(next_strip);
      } else {
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < nt; ++t) pool.emplace_back([&] { encode_range(next_strip); });
        encode_range(next_strip);
        for (auto& th : pool) th.join();
      }
    }
    for (int s = 0; s < nstrips; ++s) {
      if (f.size() + enc_len[s] > 0xfffffff0u) return STC_ERR_ARG;     // classic TIFF: 32-bit file offsets
      offs[s] = uint32_t(f.size());
      lens[s] = uint32_t(enc_len[s]);
      f.insert(f.end(), scratch.data() + slot_bytes * size_t(s), scratch.data() + slot_bytes * size_t(s) + enc_len[s]);
      pad_even(f);
    }

    // tag payloads that do not fit the 4-byte value field
    auto put_u32s = [&](const std::vector<uint32_t>& v) -> uint32_t {
      if (v.size() == 1) return v[0];
      pad_even(f); const uint32_t o = uint32_t(f.size()); append(f, v.data(), v.size()); return o; };
    const uint32_t off_offs = put_u32s(offs);
    const uint32_t off_lens = put_u32s(lens);
    const double scale[3] = {(east - west) / cols, (north - south) / rows, 0.0};
    const double tie[6] = {0, 0, 0, west, north, 0};
    pad_even(f); const uint32_t off_scale = uint32_t(f.size()); append(f, scale, 3);
    const uint32_t off_tie = uint32_t(f.size()); append(f, tie, 6);
    const char citation[] = "WGS 84|";                               // GeoASCII: '|' terminates each string
    const double dparams[2] = {6378137.0, 298.257223563};
    const uint16_t keys[] = {
        1, 1, 0, 7,                       // directory version 1, revision 1.0, 7 keys
        1024, 0, 1, 2,                    // GTModelTypeGeoKey = ModelTypeGeographic
        1025, 0, 1, 1,                    // GTRasterTypeGeoKey = RasterPixelIsArea
        2048, 0, 1, 4326,                 // GeographicTypeGeoKey = WGS 84
        2049, 34737, 7, 0,                // GeogCitationGeoKey -> GeoAsciiParams[0..7)
        2054, 0, 1, 9102,                 // GeogAngularUnitsGeoKey = degree
        2057, 34736, 1, 0,                // GeogSemiMajorAxisGeoKey -> GeoDoubleParams[0]
        2059, 34736, 1, 1,                // GeogInvFlatteningGeoKey -> GeoDoubleParams[1]
    };
    const uint32_t off_keys = uint32_t(f.size()); append(f, keys, sizeof(keys) / 2);
    const uint32_t off_dpar = uint32_t(f.size()); append(f, dparams, 2);
    const uint32_t off_cit = uint32_t(f.size()); append(f, citation, sizeof(citation));   // incl. the NUL
    pad_even(f);

    const Entry ifd[] = {
        {256, 4, 1, uint32_t(cols)},                  // ImageWidth
        {257, 4, 1, uint32_t(rows)},                  // ImageLength
        {258, 3, 1, 8},                               // BitsPerSample
        {259, 3, 1, 5},                               // Compression = LZW
        {262, 3, 1, 1},                               // Photometric = MinIsBlack
        {273, 4, uint32_t(nstrips), off_offs},        // StripOffsets
        {277, 3, 1, 1},                               // SamplesPerPixel
        {278, 4, 1, uint32_t(rps)},                   // RowsPerStrip
        {279, 4, uint32_t(nstrips), off_lens},        // StripByteCounts
        {284, 3, 1, 1},                               // PlanarConfiguration = contiguous
        {339, 3, 1, 1},                               // SampleFormat = unsigned integer
        {33550, 12, 3, off_scale},                    // ModelPixelScaleTag
        {33922, 12, 6, off_tie},                      // ModelTiepointTag
        {34735, 3, uint32_t(sizeof(keys) / 2), off_keys},   // GeoKeyDirectoryTag
        {34736, 12, 2, off_dpar},                     // GeoDoubleParamsTag
        {34737, 2, uint32_t(sizeof(citation)), off_cit},    // GeoAsciiParamsTag
    };
    const uint32_t ifd_off = uint32_t(f.size());
    const uint16_t n_entries = uint16_t(sizeof(ifd) / sizeof(ifd[0]));
    append(f, &n_entries, 1);
    for (const Entry& e : ifd) {
      append(f, &e.tag, 1); append(f, &e.type, 1); append(f, &e.count, 1);
      if (e.type == 3 && e.count == 1) { const uint16_t v[2] = {uint16_t(e.value), 0}; append(f, v, 2); }
      else append(f, &e.value, 1);
    }
    const uint32_t next_ifd = 0;
    append(f, &next_ifd, 1);
    memcpy(f.data() + 4, &ifd_off, 4);

    uint8_t* buf = static_cast<uint8_t*>(malloc(f.size()));
    if (!buf) return STC_ERR_NOMEM;
    memcpy(buf, f.data(), f.size());
    *out_buf = buf;
    *out_len = int64_t(f.size());
    return STC_OK;
  } catch (...) {
    return STC_ERR_NOMEM;
  }
}

extern "C" void stc_geotiff_free(uint8_t* buf) { free(buf); }

// Writes the file in one piece (temporary name + rename, so a reader never sees a truncated raster).  STC_ERR_STATE: I/O error.
extern "C" int stc_write_geotiff_u8(const char* path, const uint8_t* img, int rows, int cols, double west, double south, double east,
                                    double north) {
  if (!path) return STC_ERR_ARG;
  uint8_t* buf = nullptr; int64_t len = 0;
  const int rc = stc_geotiff_encode_u8(img, rows, cols, west, south, east, north, &buf, &len);
  if (rc) return rc;
  const std::string tmp = std::string(path) + ".part";
  FILE* fp = fopen(tmp.c_str(), "wb");
  if (!fp) { free(buf); return STC_ERR_STATE; }
  const bool ok = fwrite(buf, 1, size_t(len), fp) == size_t(len);
  const bool closed = fclose(fp) == 0;
  free(buf);
  if (!ok || !closed || rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); return STC_ERR_STATE; }
  return STC_OK;
}

// ---- reader: the tile products the border pass loads back (`load_tif`, /root/reference/src/resegment_tiles_wide.py:713-751:
// `rasterio.open(tif).read(1)` of a `_FINAL` / `_POST` / `_SMOOTH*` product) ------------------------------------------------------
// Classic TIFF (little- or big-endian), 8 bits per sample, strips or tiles, chunky or planar samples, compression none /
// LZW (TIFF variant, as above) / PackBits, predictor 1 or 2 (horizontal differencing).  That covers what this library, GDAL
// (`compress='lzw'`, striped or TILED=YES), libtiff (Pillow) and OpenCV write for a uint8 raster.  BigTIFF, Deflate / ZSTD and
// other sample widths are refused with STC_ERR_STATE (no silent garbage).
namespace {

struct ByteView {
  const uint8_t* p; size_t n; bool be;
  bool ok(uint64_t off, uint64_t len) const { return off <= n && len <= n - off; }
  uint16_t u16(uint64_t o) const { return be ? uint16_t(p[o] << 8 | p[o + 1]) : uint16_t(p[o] | p[o + 1] << 8); }
  uint32_t u32(uint64_t o) const {
    return be ? (uint32_t(p[o]) << 24 | uint32_t(p[o + 1]) << 16 | uint32_t(p[o + 2]) << 8 | p[o + 3])
              : (uint32_t(p[o]) | uint32_t(p[o + 1]) << 8 | uint32_t(p[o + 2]) << 16 | uint32_t(p[o + 3]) << 24);
  }
  double f64(uint64_t o) const {
    uint8_t b[8];
    for (int i = 0; i < 8; ++i) b[i] = be ? p[o + 7 - i] : p[o + i];
    double d; memcpy(&d, b, 8); return d;
  }
};

struct TiffTag { uint16_t type = 0; uint32_t count = 0; uint64_t at = 0; bool present = false; };   // `at`: file offset of the values

inline int type_size(uint16_t t) { return (t == 1 || t == 2 || t == 6 || t == 7) ? 1 : (t == 3 || t == 8) ? 2 : (t == 4 || t == 9 || t == 11) ? 4 : (t == 5 || t == 10 || t == 12) ? 8 : 0; }

// value i of an integer-typed tag (BYTE / SHORT / LONG)
inline bool tag_uint(const ByteView& v, const TiffTag& t, uint32_t i, uint32_t& out) {
  if (!t.present || i >= t.count) return false;
  if (t.type == 3) out = v.u16(t.at + 2ull * i);
  else if (t.type == 4) out = v.u32(t.at + 4ull * i);
  else if (t.type == 1) out = v.p[t.at + i];
  else return false;
  return true;
}

// TIFF LZW decode of one strip / tile into exactly `want` bytes (fewer decoded bytes -> the rest stays 0, like libtiff pads a
// short strip; more -> error).  Returns false on a corrupt stream.
bool lzw_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t want) {
  constexpr int CLEAR = 256, EOI = 257, FIRST = 258;
  static thread_local uint16_t prefix[4096];
  static thread_local uint8_t suffix[4096], first[4096];
  static thread_local uint16_t length[4096];
  for (int i = 0; i < 256; ++i) { prefix[i] = 0xffff; suffix[i] = uint8_t(i); first[i] = uint8_t(i); length[i] = 1; }
  uint64_t acc = 0; int nbits = 0; size_t ip = 0;
  int width = 9, next = FIRST, old = -1;
  size_t op = 0;
  for (;;) {
    while (nbits < width) {
      if (ip >= n) return true;                      // stream ended without EOI: accept what was decoded (libtiff does)
      acc = (acc << 8) | src[ip++]; nbits += 8;
    }
    const int code = int((acc >> (nbits - width)) & ((1u << width) - 1));
    nbits -= width;
    if (code == EOI) return true;
    if (code == CLEAR) { width = 9; next = FIRST; old = -1; continue; }
    int emit = code;
    if (old < 0) {
      if (code >= 256) return false;                 // the first code after a Clear must be a literal
    } else {
      if (code > next || (code == next && next >= 4096)) return false;
      if (next < 4096) {       // (a full table without a Clear: keep decoding with what is there, as libtiff does)
        // new entry = string(old) + first byte of string(code), where string(next) (the KwKwK case) starts like string(old)
        prefix[next] = uint16_t(old); length[next] = uint16_t(length[old] + 1); first[next] = first[old];
        suffix[next] = (code == next) ? first[old] : first[code];
        ++next;
        if (next + 1 == (1 << width) && width < 12) ++width;    // early change: the width steps one code before it must
      }
    }
    const size_t len = length[emit];
    if (op + len > want) return false;
    uint8_t* q = dst + op + len;
    for (int c = emit; c != 0xffff; c = prefix[c]) *--q = suffix[c];
    op += len;
    old = code;
  }
}

bool packbits_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t want) {
  size_t ip = 0, op = 0;
  while (ip < n && op < want) {
    const int8_t h = int8_t(src[ip++]);
    if (h >= 0) {
      const size_t len = size_t(h) + 1;
      if (ip + len > n || op + len > want) return false;
      memcpy(dst + op, src + ip, len); ip += len; op += len;
    } else if (h != -128) {
      const size_t len = size_t(1 - int(h));
      if (ip >= n || op + len > want) return false;
      memset(dst + op, src[ip++], len); op += len;
    }
  }
  return true;
}

}  // namespace

// STC_OK; STC_ERR_ARG: null arguments / band out of range; STC_ERR_STATE: not a TIFF this reader supports, or a corrupt one;
// STC_ERR_NOMEM.  *out_img: malloc'ed [rows][cols] uint8 of sample `band` (1-based, rasterio's read(band)); release with
// stc_geotiff_free.  bounds4 (optional): west, south, east, north from ModelPixelScale + ModelTiepoint, NaN when absent.
extern "C" int stc_geotiff_decode_u8(const uint8_t* file, int64_t len, int band, uint8_t** out_img, int* rows, int* cols, double* bounds4) {
  if (!file || !out_img || !rows || !cols || len < 8 || band < 1) return STC_ERR_ARG;
  ByteView v{file, size_t(len), false};
  if (file[0] == 'I' && file[1] == 'I') v.be = false;
  else if (file[0] == 'M' && file[1] == 'M') v.be = true;
  else return STC_ERR_STATE;
  if (v.u16(2) != 42) return STC_ERR_STATE;                        // 43 = BigTIFF: not supported
  const uint64_t ifd = v.u32(4);
  if (!v.ok(ifd, 2)) return STC_ERR_STATE;
  const uint32_t n_entries = v.u16(ifd);
  if (!v.ok(ifd + 2, 12ull * n_entries)) return STC_ERR_STATE;
  TiffTag width, height, bps, comp, strip_off, spp, rps, strip_cnt, planar, pred, tile_w, tile_h, tile_off, tile_cnt, scale, tie;
  for (uint32_t e = 0; e < n_entries; ++e) {
    const uint64_t at = ifd + 2 + 12ull * e;
    const uint16_t tag = v.u16(at);
    TiffTag t; t.type = v.u16(at + 2); t.count = v.u32(at + 4); t.present = true;
    const uint64_t bytes = uint64_t(type_size(t.type)) * t.count;
    if (bytes == 0) continue;
    t.at = bytes <= 4 ? at + 8 : v.u32(at + 8);
    if (!v.ok(t.at, bytes)) return STC_ERR_STATE;
    switch (tag) {
      case 256: width = t; break;      case 257: height = t; break;     case 258: bps = t; break;       case 259: comp = t; break;
      case 273: strip_off = t; break;  case 277: spp = t; break;        case 278: rps = t; break;       case 279: strip_cnt = t; break;
      case 284: planar = t; break;     case 317: pred = t; break;       case 322: tile_w = t; break;    case 323: tile_h = t; break;
      case 324: tile_off = t; break;   case 325: tile_cnt = t; break;   case 33550: scale = t; break;   case 33922: tie = t; break;
      default: break;
    }
  }
  uint32_t W = 0, H = 0, bits = 1, compression = 1, samples = 1, planar_cfg = 1, predictor = 1;
  if (!tag_uint(v, width, 0, W) || !tag_uint(v, height, 0, H) || W < 1 || H < 1) return STC_ERR_STATE;
  tag_uint(v, spp, 0, samples); tag_uint(v, comp, 0, compression); tag_uint(v, planar, 0, planar_cfg); tag_uint(v, pred, 0, predictor);
  if (samples < 1 || uint32_t(band) > samples) return STC_ERR_ARG;
  for (uint32_t i = 0; i < (bps.present ? bps.count : 0); ++i) { tag_uint(v, bps, i, bits); if (bits != 8) return STC_ERR_STATE; }
  if (!bps.present) return STC_ERR_STATE;                           // default is 1 bit per sample
  if (compression != 1 && compression != 5 && compression != 32773) return STC_ERR_STATE;
  if (predictor != 1 && predictor != 2) return STC_ERR_STATE;
  if (uint64_t(W) * H > (uint64_t(1) << 31)) return STC_ERR_STATE;
  const bool tiled = tile_off.present;
  uint32_t bw = W, bh = H;                                          // block (strip or tile) size in pixels
  if (tiled) { if (!tag_uint(v, tile_w, 0, bw) || !tag_uint(v, tile_h, 0, bh) || bw < 1 || bh < 1) return STC_ERR_STATE; }
  else { uint32_t r = H; if (tag_uint(v, rps, 0, r) && r >= 1 && r < H) bh = r; }
  if (uint64_t(bw) * bh * samples > (uint64_t(1) << 31)) return STC_ERR_STATE;      // a block this large is a corrupt tag, not a raster
  const TiffTag& offs = tiled ? tile_off : strip_off;
  const TiffTag& cnts = tiled ? tile_cnt : strip_cnt;
  if (!offs.present) return STC_ERR_STATE;
  const uint32_t across = (W + bw - 1) / bw, down = (H + bh - 1) / bh;
  const uint32_t per_plane = across * down;
  const uint32_t chunky = planar_cfg == 2 ? 1 : samples;            // samples interleaved inside a block
  const uint32_t plane = planar_cfg == 2 ? uint32_t(band - 1) : 0;
  if (uint64_t(per_plane) * (planar_cfg == 2 ? samples : 1) > offs.count) return STC_ERR_STATE;
  uint8_t* img = static_cast<uint8_t*>(malloc(size_t(W) * H));
  if (!img) return STC_ERR_NOMEM;
  try {
    std::vector<uint8_t> block;
    for (uint32_t by = 0; by < down; ++by) {
      for (uint32_t bx = 0; bx < across; ++bx) {
        const uint32_t idx = plane * per_plane + by * across + bx;
        uint32_t off = 0, cnt = 0;
        tag_uint(v, offs, idx, off);
        const uint32_t rows_here = tiled ? bh : std::min(bh, H - by * bh);     // tiles are always full size, the last strip is not
        const size_t want = size_t(rows_here) * bw * chunky;
        if (!tag_uint(v, cnts, idx, cnt)) cnt = (compression == 1) ? uint32_t(want) : uint32_t(v.n - std::min<size_t>(v.n, off));
        if (!v.ok(off, cnt)) { free(img); return STC_ERR_STATE; }
        block.assign(want, 0);
        bool good = true;
        if (compression == 1) { if (cnt < want) good = false; else memcpy(block.data(), file + off, want); }
        else if (compression == 5) good = lzw_decode(file + off, cnt, block.data(), want);
        else good = packbits_decode(file + off, cnt, block.data(), want);
        if (!good) { free(img); return STC_ERR_STATE; }
        if (predictor == 2)
          for (uint32_t r = 0; r < rows_here; ++r) {
            uint8_t* row = block.data() + size_t(r) * bw * chunky;
            for (size_t i = chunky; i < size_t(bw) * chunky; ++i) row[i] = uint8_t(row[i] + row[i - chunky]);
          }
        const uint32_t x0 = bx * bw, y0 = by * bh;
        const uint32_t cols_here = std::min(bw, W - x0), rows_copy = std::min(rows_here, H - y0);
        const uint32_t s = planar_cfg == 2 ? 0 : uint32_t(band - 1);
        for (uint32_t r = 0; r < rows_copy; ++r) {
          const uint8_t* src = block.data() + size_t(r) * bw * chunky + s;
          uint8_t* dst = img + size_t(y0 + r) * W + x0;
          if (chunky == 1) memcpy(dst, src, cols_here);
          else for (uint32_t c = 0; c < cols_here; ++c) dst[c] = src[size_t(c) * chunky];
        }
      }
    }
  } catch (...) { free(img); return STC_ERR_NOMEM; }
  if (bounds4) {
    const double nan = std::nan("");
    bounds4[0] = bounds4[1] = bounds4[2] = bounds4[3] = nan;
    if (scale.present && scale.type == 12 && scale.count >= 2 && tie.present && tie.type == 12 && tie.count >= 6) {
      const double sx = v.f64(scale.at), sy = v.f64(scale.at + 8);
      const double px = v.f64(tie.at), py = v.f64(tie.at + 8), gx = v.f64(tie.at + 24), gy = v.f64(tie.at + 32);
      bounds4[0] = gx - px * sx;               // west
      bounds4[3] = gy + py * sy;               // north
      bounds4[2] = bounds4[0] + sx * W;        // east
      bounds4[1] = bounds4[3] - sy * H;        // south
    }
  }
  *out_img = img; *rows = int(H); *cols = int(W);
  return STC_OK;
}

extern "C" int stc_read_geotiff_u8(const char* path, int band, uint8_t** out_img, int* rows, int* cols, double* bounds4) {
  if (!path) return STC_ERR_ARG;
  FILE* fp = fopen(path, "rb");
  if (!fp) return STC_ERR_STATE;
  std::vector<uint8_t> buf;
  try {
    if (fseek(fp, 0, SEEK_END) != 0) { fclose(fp); return STC_ERR_STATE; }
    const long sz = ftell(fp);
    if (sz < 8 || fseek(fp, 0, SEEK_SET) != 0) { fclose(fp); return STC_ERR_STATE; }
    buf.resize(size_t(sz));
    if (fread(buf.data(), 1, size_t(sz), fp) != size_t(sz)) { fclose(fp); return STC_ERR_STATE; }
  } catch (...) { fclose(fp); return STC_ERR_NOMEM; }
  fclose(fp);
  return stc_geotiff_decode_u8(buf.data(), int64_t(buf.size()), band, out_img, rows, cols, bounds4);
}
