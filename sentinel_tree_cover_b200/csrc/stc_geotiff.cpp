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
#include <atomic>
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
