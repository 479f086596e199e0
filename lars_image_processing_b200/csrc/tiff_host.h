// Host-side ingest: TIFF 6.0 / BigTIFF reader for the frame formats of the path -- chunky 8- or 16-bit
// unsigned samples, 1 / 3 / 4 samples per pixel, chunky or planar, either byte order; strips or tiles; uncompressed, LZW,
// Deflate or PackBits, with or without the horizontal-differencing predictor; whole frames or a
// rectangular region (the tiles / row bands one rank owns of an orthomosaic, BASELINE config 4).
//
// Why it exists: the reference loads frames with
// PIL.Image.open (process-images.py:183-193; backend-process.py:52; process-ndvi.py:18;
// process-rgn.py:18), and Pillow opens a 16-bit RGB TIFF as 8-bit -- 16-bit survey frames
// (BASELINE config 3) cannot enter the reference through files at all.  This reader moves the
// strips / tiles of a memory-mapped file straight into a caller-supplied (pinned) HWC buffer,
// byte-swapping big-endian 16-bit samples on the way; compressed chunks are independent streams and
// are decoded by a few host threads side by side.  The formats follow the TIFF 6.0 specification
// (sections 9 PackBits, 13 LZW, 14 differencing predictor, 15 tiles), Adobe's Deflate note and the
// BigTIFF layout (64-bit offsets, 20-byte entries); no third-party source is involved.  Deflate
// streams are inflated by the system zlib, looked up at run time (no link dependency: a box without
// libz reports Deflate files as unsupported and the caller decodes them with Pillow).
#pragma once
#include <dlfcn.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../../include/lars_b200.h"

namespace lars_host {

struct TiffCursor {
  const uint8_t* p;
  size_t n;
  bool be;
  bool ok(uint64_t off, uint64_t len) const { return off <= n && len <= n - off; }
  uint16_t u16(uint64_t off) const {
    return be ? (uint16_t)((p[off] << 8) | p[off + 1]) : (uint16_t)(p[off] | (p[off + 1] << 8));
  }
  uint32_t u32(uint64_t off) const {
    return be ? ((uint32_t)p[off] << 24) | ((uint32_t)p[off + 1] << 16) | ((uint32_t)p[off + 2] << 8) | p[off + 3]
              : ((uint32_t)p[off + 3] << 24) | ((uint32_t)p[off + 2] << 16) | ((uint32_t)p[off + 1] << 8) | p[off];
  }
  uint64_t u64(uint64_t off) const {
    return be ? ((uint64_t)u32(off) << 32) | u32(off + 4) : ((uint64_t)u32(off + 4) << 32) | u32(off);
  }
};

inline int tiff_type_size(int type) {
  switch (type) {
    case 1: case 2: case 6: case 7: return 1;    // BYTE ASCII SBYTE UNDEFINED
    case 3: case 8: return 2;                    // SHORT SSHORT
    case 4: case 9: case 11: case 13: return 4;  // LONG SLONG FLOAT IFD
    case 5: case 10: case 12: return 8;          // RATIONAL SRATIONAL DOUBLE
    case 16: case 17: case 18: return 8;         // LONG8 SLONG8 IFD8 (BigTIFF)
    default: return 0;
  }
}

// value i of an IFD entry of type SHORT, LONG or LONG8; `pos` is where the values live
inline uint64_t tiff_value(const TiffCursor& c, uint64_t pos, int type, uint64_t i) {
  return type == 3 ? c.u16(pos + 2ull * i) : type == 4 ? c.u32(pos + 4ull * i) : c.u64(pos + 8ull * i);
}

inline bool tiff_compressed(int compression) { return compression != 1; }

// ---- Deflate through the system zlib, bound at run time --------------------------------------
typedef int (*zlib_uncompress_fn)(unsigned char*, unsigned long*, const unsigned char*, unsigned long);
inline zlib_uncompress_fn zlib_uncompress() {
  static zlib_uncompress_fn fn = []() -> zlib_uncompress_fn {
    for (const char* name : {"libz.so.1", "libz.so"}) {
      if (void* h = dlopen(name, RTLD_NOW | RTLD_LOCAL))
        if (void* s = dlsym(h, "uncompress")) return reinterpret_cast<zlib_uncompress_fn>(s);
    }
    return nullptr;
  }();
  return fn;
}

// Each decoder writes at most `cap` bytes and returns how many it produced (never reads past `end`).
inline size_t inflate_chunk(const uint8_t* in, size_t n_in, uint8_t* out, size_t cap) {
  unsigned long produced = (unsigned long)cap;
  const int rc = zlib_uncompress()(out, &produced, in, (unsigned long)n_in);
  // 0 = Z_OK; -5 = Z_BUF_ERROR with the buffer full: the stream holds more than the chunk needs
  if (rc == 0 || (rc == -5 && produced == cap)) return produced;
  return 0;
}

inline size_t packbits_chunk(const uint8_t* in, size_t n_in, uint8_t* out, size_t cap) {
  const uint8_t* end = in + n_in;
  size_t op = 0;
  while (in < end && op < cap) {
    const int n = (int8_t)*in++;
    if (n >= 0) {                                  // n + 1 literal bytes
      size_t lit = (size_t)n + 1;
      if (lit > (size_t)(end - in)) lit = (size_t)(end - in);
      const size_t cnt = lit < cap - op ? lit : cap - op;
      memcpy(out + op, in, cnt);
      in += lit;
      op += cnt;
    } else if (n != -128) {                        // the next byte 1 - n times
      if (in >= end) break;
      size_t cnt = (size_t)(1 - n);
      if (cnt > cap - op) cnt = cap - op;
      memset(out + op, *in++, cnt);
      op += cnt;
    }
  }
  return op;
}

// TIFF 6.0 section 13: MSB-first codes of 9..12 bits, 256 = Clear, 257 = EndOfInformation, the width
// grows one code early (after entry 510 / 1022 / 2046 has been added).  Every string the table ever
// holds already stands in the output: the entry made when `code` follows `old` is old's string plus
// the next byte, i.e. the bytes written for old and the first one written for code.  The table therefore
// stores (position, length) into the output and a string is emitted by copying from there -- no
// prefix chains to walk.
inline size_t lzw_chunk(const uint8_t* in, size_t n_in, uint8_t* out, size_t cap) {
  struct Entry { uint32_t pos; uint32_t len; };    // 32 KB: the table stays in L1
  // Literals are table entries as well -- they point into a page holding the bytes 0..255 (bit 31 of pos) --
  // so that noisy data, where literals and strings alternate unpredictably, takes no branch on the kind of code.
  static const struct LiteralPage { uint8_t b[256 + 8]; LiteralPage() { for (int i = 0; i < 264; ++i) b[i] = (uint8_t)i; } } page;
  constexpr uint32_t kLiteral = 0x80000000u;
  Entry table[4096];
  if (cap > 0x7fffffffull) return 0;               // one strip / tile of 2 GB: not a frame of this path
  for (uint32_t i = 0; i < 256; ++i) table[i] = Entry{kLiteral | i, 1};
  table[256] = table[257] = Entry{kLiteral, 1};
  const uint8_t* end = in + n_in;
  uint64_t acc = 0;
  int have = 0, nbits = 9, next_code = 258, old = -1;
  size_t op = 0, old_pos = 0, old_len = 0;
  while (op < cap) {
    if (have < nbits) {                            // refill: several codes' worth of bits at a time
      while (have <= 56 && in < end) { acc = (acc << 8) | *in++; have += 8; }
      if (have < nbits) break;                     // ran out of input: treat as EndOfInformation
    }
    const int code = (int)((acc >> (have - nbits)) & ((1u << nbits) - 1));
    have -= nbits;
    if (code == 256) { nbits = 9; next_code = 258; old = -1; continue; }
    if (code == 257) break;
    const size_t at = op;
    size_t len;
    if (code < next_code) {                        // a literal or a string made earlier (it ends before `op`)
      if (old < 0 && code > 255) return 0;         // the first code after a Clear must be a literal
      const Entry e = table[code];
      len = e.len;
      const uint8_t* src = (e.pos & kLiteral) ? page.b + (e.pos & 0xFFu) : out + e.pos;
      if (len <= 8 && op + 8 <= cap) {             // most strings are short: one 8-byte move, the excess is
        uint64_t w;                                // scratch that later output overwrites (never past cap)
        memcpy(&w, src, 8);
        memcpy(out + op, &w, 8);
        op += len;
      } else {
        const size_t keep = len < cap - op ? len : cap - op;
        memcpy(out + op, src, keep);
        op += keep;
      }
    } else if (code == next_code && next_code < 4096 && old >= 0) {   // the string being defined: old + first(old);
      len = old_len + 1;                                               // it overlaps its own source, so byte by byte
      const size_t keep = len < cap - op ? len : cap - op;
      for (size_t i = 0; i < keep; ++i) out[op + i] = out[old_pos + i];
      op += keep;
    } else {
      return 0;                                    // a code the table cannot hold yet: corrupt stream
    }
    if (old >= 0) {
      if (next_code < 4096) table[next_code++] = Entry{(uint32_t)old_pos, (uint32_t)(old_len + 1)};
      if (next_code >= (1 << nbits) - 1 && nbits < 12) ++nbits;
    }
    old = code;
    old_pos = at;
    old_len = len;
  }
  return op;
}

// Returns NULL on success, else a static description of what is wrong / unsupported.
inline const char* tiff_probe(const void* file, size_t file_bytes, lars_tiff_info* info, bool* unsupported) {
  *unsupported = false;
  memset(info, 0, sizeof(*info));
  TiffCursor c{static_cast<const uint8_t*>(file), file_bytes, false};
  if (file_bytes < 8) return "file shorter than a TIFF header";
  if (c.p[0] == 'I' && c.p[1] == 'I') c.be = false;
  else if (c.p[0] == 'M' && c.p[1] == 'M') c.be = true;
  else return "not a TIFF file (byte-order mark)";
  const uint16_t magic = c.u16(2);
  if (magic != 42 && magic != 43) return "not a TIFF file (magic number)";
  const bool big = magic == 43;
  uint64_t ifd, n_entries;
  if (big) {
    if (file_bytes < 16 || c.u16(4) != 8 || c.u16(6) != 0) return "malformed BigTIFF header";
    ifd = c.u64(8);
    if (!c.ok(ifd, 8)) return "IFD offset outside the file";
    n_entries = c.u64(ifd);
  } else {
    ifd = c.u32(4);
    if (!c.ok(ifd, 2)) return "IFD offset outside the file";
    n_entries = c.u16(ifd);
  }
  const uint64_t entry_bytes = big ? 20 : 12, first_entry = ifd + (big ? 8 : 2), inline_cap = big ? 8 : 4;
  if (n_entries > 65535 || !c.ok(first_entry, entry_bytes * n_entries)) return "IFD runs past the end of the file";

  info->big_endian = c.be ? 1 : 0;
  info->bigtiff = big ? 1 : 0;
  info->samples_per_pixel = 1;
  info->bits_per_sample = 1;
  info->compression = 1;
  info->planar_config = 1;
  info->sample_format = 1;
  info->predictor = 1;
  info->rows_per_strip = -1;
  uint64_t strip_pos[2] = {0, 0}, tile_pos[2] = {0, 0}, strip_cnt[2] = {0, 0}, tile_cnt[2] = {0, 0};
  int strip_type[2] = {0, 0}, tile_type[2] = {0, 0};
  bool bits_mixed = false;
  for (uint64_t e = 0; e < n_entries; ++e) {
    const uint64_t at = first_entry + entry_bytes * e;
    const int tag = c.u16(at), type = c.u16(at + 2);
    const uint64_t count = big ? c.u64(at + 4) : c.u32(at + 4);
    const uint64_t value_at = at + (big ? 12 : 8);
    const int tsz = tiff_type_size(type);
    if (tsz == 0) continue;                                   // unknown field type: ignore the entry
    if (count > (1ull << 40)) return "an IFD entry has an implausible count";
    const uint64_t bytes = (uint64_t)tsz * count;
    const uint64_t pos = bytes <= inline_cap ? value_at : (big ? c.u64(value_at) : c.u32(value_at));
    if (!c.ok(pos, bytes)) return "an IFD entry points outside the file";
    const bool integral = (type == 3 || type == 4 || type == 16);
    const uint64_t v64 = (integral && count >= 1) ? tiff_value(c, pos, type, 0) : 0;
    const uint32_t v0 = v64 > 0xffffffffull ? 0xffffffffu : (uint32_t)v64;
    const int32_t v31 = v0 > 0x7fffffffu ? -1 : (int32_t)v0;
    switch (tag) {
      case 256: info->width = v31; break;
      case 257: info->height = v31; break;
      case 258:
        info->bits_per_sample = v31;
        for (uint64_t i = 1; i < count && integral; ++i)
          if (tiff_value(c, pos, type, i) != v64) bits_mixed = true;
        break;
      case 259: info->compression = v31; break;
      case 262: info->photometric = v31; break;
      case 266: if (v0 == 2) { *unsupported = true; return "FillOrder 2 (bit-reversed data): decode it with Pillow"; } break;
      case 273: strip_pos[0] = pos; strip_type[0] = type; strip_cnt[0] = count; break;
      case 277: info->samples_per_pixel = v31; break;
      case 278: info->rows_per_strip = v31; break;
      case 279: strip_pos[1] = pos; strip_type[1] = type; strip_cnt[1] = count; break;
      case 284: info->planar_config = v31; break;
      case 317: info->predictor = v31; break;
      case 322: info->tile_width = v31; break;
      case 323: info->tile_length = v31; break;
      case 324: tile_pos[0] = pos; tile_type[0] = type; tile_cnt[0] = count; break;
      case 325: tile_pos[1] = pos; tile_type[1] = type; tile_cnt[1] = count; break;
      case 339: info->sample_format = v31; break;
      default: break;
    }
  }
  if (info->width < 1 || info->height < 1) return "missing ImageWidth / ImageLength";
  if (info->width > (1 << 24) || info->height > (1 << 24)) return "implausible image dimensions";   // keeps all byte counts far below 2^64
  switch (info->compression) {
    case 1: case 5: case 32773: break;
    case 32946: info->compression = 8;   // the older Deflate code, same streams
      [[fallthrough]];
    case 8:
      if (!zlib_uncompress()) { *unsupported = true; return "Deflate TIFF, but no zlib on this system: decode it with Pillow"; }
      break;
    default: *unsupported = true; return "this TIFF compression scheme is not supported: decode it with Pillow";
  }
  // libtiff (and so Pillow) honours the Predictor tag only inside the LZW / Deflate codecs; uncompressed and
  // PackBits data are taken as they are
  if (info->compression != 5 && info->compression != 8) info->predictor = 1;
  if (info->predictor != 1 && info->predictor != 2) { *unsupported = true; return "only the horizontal-differencing TIFF predictor is supported"; }
  // unsigned 8- / 16-bit samples (the frames of the path), or 32-bit IEEE floats (calibrated reflectance stacks
  // and stored index maps: Pillow opens only single-band float files)
  const bool is_float = info->sample_format == 3;
  if (bits_mixed || !((info->sample_format == 1 && (info->bits_per_sample == 8 || info->bits_per_sample == 16)) ||
                      (is_float && info->bits_per_sample == 32))) {
    *unsupported = true;
    return "only unsigned 8- / 16-bit or 32-bit float samples are supported";
  }
  if (is_float && info->predictor != 1) { *unsupported = true; return "floating-point predictor: decode it with Pillow"; }
  if (info->samples_per_pixel != 1 && info->samples_per_pixel != 3 && info->samples_per_pixel != 4) {
    *unsupported = true;
    return "only 1, 3 or 4 samples per pixel are supported";
  }
  if (info->planar_config != 1 && info->planar_config != 2) return "unknown PlanarConfiguration";
  if (info->samples_per_pixel == 1) info->planar_config = 1;      // one sample per pixel: the two layouts coincide
  if (info->photometric == 0) { *unsupported = true; return "WhiteIsZero TIFF: decode it with Pillow (it inverts the samples)"; }
  // BlackIsZero, RGB, palette indices and CMYK are handed over as stored (as np.array(Image.open()) does);
  // YCbCr, CIELab, CFA ... are converted or re-interpreted by Pillow
  if (info->photometric != 1 && info->photometric != 2 && info->photometric != 3 && info->photometric != 5) {
    *unsupported = true;
    return "this PhotometricInterpretation is left to Pillow";
  }

  const uint64_t px_bytes = (uint64_t)info->samples_per_pixel * (info->bits_per_sample / 8);
  info->frame_bytes = (uint64_t)info->width * px_bytes * (uint64_t)info->height;
  const bool tiled = tile_pos[0] != 0 || info->tile_width > 0 || info->tile_length > 0;
  // PlanarConfiguration 2: every sample of the pixel has its own run of strips / tiles (plane after plane),
  // each chunk holding one sample per pixel
  const uint64_t planes = info->planar_config == 2 ? (uint64_t)info->samples_per_pixel : 1;
  const uint64_t file_px_bytes = px_bytes / planes;
  uint64_t n_chunks;
  if (tiled) {
    if (!tile_pos[0]) return "missing TileOffsets";
    if (info->tile_width < 1 || info->tile_length < 1 || info->tile_width > (1 << 24) || info->tile_length > (1 << 24))
      return "missing or implausible TileWidth / TileLength";
    info->tiles_across = (info->width + info->tile_width - 1) / info->tile_width;
    info->tiles_down = (info->height + info->tile_length - 1) / info->tile_length;
    n_chunks = (uint64_t)info->tiles_across * (uint64_t)info->tiles_down * planes;
    info->rows_per_strip = info->tile_length;
    info->strip_offsets_pos = tile_pos[0]; info->strip_offsets_type = tile_type[0];
    info->strip_counts_pos = tile_pos[1]; info->strip_counts_type = tile_type[1];
    if (tile_cnt[0] != n_chunks) return "TileOffsets count does not match the tile grid";
    if (!tile_pos[1]) return "missing TileByteCounts";
    if (tile_cnt[1] != n_chunks) return "TileByteCounts count does not match TileOffsets";
  } else {
    if (!strip_pos[0]) return "missing StripOffsets";
    if (info->rows_per_strip < 1 || info->rows_per_strip > info->height) info->rows_per_strip = info->height;
    n_chunks = (uint64_t)((info->height + info->rows_per_strip - 1) / info->rows_per_strip) * planes;
    info->strip_offsets_pos = strip_pos[0]; info->strip_offsets_type = strip_type[0];
    info->strip_counts_pos = strip_pos[1]; info->strip_counts_type = strip_type[1];
    if (strip_cnt[0] != n_chunks) return "StripOffsets count does not match the image height";
    if (strip_pos[1] && strip_cnt[1] != n_chunks) return "StripByteCounts count does not match StripOffsets";
    if (!strip_pos[1] && tiff_compressed(info->compression)) return "compressed TIFF without StripByteCounts";
  }
  if (n_chunks > (1ull << 26)) return "implausible number of strips / tiles";
  info->n_strips = (int32_t)n_chunks;
  auto index_type_ok = [](int t) { return t == 3 || t == 4 || t == 16; };
  if (!index_type_ok(info->strip_offsets_type)) return "strip / tile offsets must be SHORT, LONG or LONG8";
  if (info->strip_counts_pos && !index_type_ok(info->strip_counts_type)) return "strip / tile byte counts must be SHORT, LONG or LONG8";

  // every chunk must lie inside the file; uncompressed ones must hold their rows
  const uint64_t chunk_row_bytes = (uint64_t)(tiled ? info->tile_width : info->width) * file_px_bytes;
  const uint64_t per_plane = n_chunks / planes;
  for (uint64_t s = 0; s < n_chunks; ++s) {
    const uint64_t off = tiff_value(c, info->strip_offsets_pos, info->strip_offsets_type, s);
    const uint64_t in_plane = s % per_plane;
    const uint64_t cy = tiled ? in_plane / (uint64_t)info->tiles_across : in_plane;
    const uint64_t left = (uint64_t)info->height - cy * (uint64_t)info->rows_per_strip;
    const uint64_t rows = tiled ? (uint64_t)info->tile_length : (left < (uint64_t)info->rows_per_strip ? left : (uint64_t)info->rows_per_strip);
    const uint64_t need = chunk_row_bytes * rows;
    if (tiff_compressed(info->compression)) {
      const uint64_t cnt = tiff_value(c, info->strip_counts_pos, info->strip_counts_type, s);
      if (!c.ok(off, cnt)) return tiled ? "a tile runs past the end of the file" : "a strip runs past the end of the file";
    } else {
      if (!c.ok(off, need)) return tiled ? "a tile runs past the end of the file" : "a strip runs past the end of the file";
      if (info->strip_counts_pos && tiff_value(c, info->strip_counts_pos, info->strip_counts_type, s) < need)
        return tiled ? "a tile is shorter than its rows" : "a strip is shorter than its rows";
    }
  }
  return nullptr;
}

// TIFF 6.0 section 14: every sample holds the difference to the same sample of the pixel on its left.
inline void undo_predictor_row(uint8_t* row, uint64_t width_px, int spp, int sample_bytes) {
  const uint64_t n = width_px * (uint64_t)spp;
  if (sample_bytes == 1) {
    for (uint64_t i = (uint64_t)spp; i < n; ++i) row[i] = (uint8_t)(row[i] + row[i - spp]);
  } else {
    uint16_t* r = reinterpret_cast<uint16_t*>(row);      // native little-endian samples by now
    for (uint64_t i = (uint64_t)spp; i < n; ++i) r[i] = (uint16_t)(r[i] + r[i - spp]);
  }
}

inline void swap_inplace(uint8_t* p, uint64_t bytes, int sb) {
  if (sb == 2) {
    for (uint64_t i = 0; i + 1 < bytes; i += 2) { const uint8_t t = p[i]; p[i] = p[i + 1]; p[i + 1] = t; }
  } else {
    for (uint64_t i = 0; i + 3 < bytes; i += 4) {
      const uint8_t a = p[i], b = p[i + 1];
      p[i] = p[i + 3]; p[i + 1] = p[i + 2]; p[i + 2] = b; p[i + 3] = a;
    }
  }
}

inline void swap16_inplace(uint8_t* p, uint64_t bytes) {
  for (uint64_t i = 0; i + 1 < bytes; i += 2) { const uint8_t t = p[i]; p[i] = p[i + 1]; p[i + 1] = t; }
}

inline void copy_samples(uint8_t* dst, const uint8_t* src, uint64_t bytes, bool swap, int sb = 2) {
  if (!swap) { memcpy(dst, src, bytes); return; }
  if (sb == 4) {
    for (uint64_t i = 0; i + 3 < bytes; i += 4) { dst[i] = src[i + 3]; dst[i + 1] = src[i + 2]; dst[i + 2] = src[i + 1]; dst[i + 3] = src[i]; }
    return;
  }
  for (uint64_t i = 0; i + 1 < bytes; i += 2) { dst[i] = src[i + 1]; dst[i + 1] = src[i]; }
}

// Rows [row0, row1) x columns [col0, col1) of the image into dst as one contiguous block with
// little-endian samples.  Only the strips / tiles that touch the region are read; n_threads host
// threads take chunks from a shared counter (compressed chunks are independent streams).
inline const char* tiff_read_region(const void* file, size_t file_bytes, const lars_tiff_info* info, int32_t row0,
                                    int32_t row1, int32_t col0, int32_t col1, void* dst, size_t dst_bytes,
                                    int n_threads) {
  if (row0 < 0 || col0 < 0 || row1 > info->height || col1 > info->width || row0 >= row1 || col0 >= col1)
    return "region outside the image";
  const int spp = info->samples_per_pixel, sb = info->bits_per_sample / 8;
  const uint64_t px_bytes = (uint64_t)spp * sb;                 // of the destination (always interleaved)
  const int planes = info->planar_config == 2 ? spp : 1;
  const int file_spp = spp / planes;                            // samples per pixel inside one chunk
  const uint64_t file_px_bytes = (uint64_t)file_spp * sb;
  const uint64_t out_row_bytes = (uint64_t)(col1 - col0) * px_bytes;
  if ((uint64_t)dst_bytes < out_row_bytes * (uint64_t)(row1 - row0)) return "destination buffer smaller than the region";
  const bool tiled = info->tile_width > 0;
  const int64_t chunk_w = tiled ? info->tile_width : info->width, chunk_h = info->rows_per_strip;
  const int64_t across = tiled ? info->tiles_across : 1;
  const int64_t per_plane = info->n_strips / planes;
  const uint64_t chunk_row_bytes = (uint64_t)chunk_w * file_px_bytes;
  const int64_t cy0 = row0 / chunk_h, cy1 = (row1 - 1) / chunk_h, cx0 = col0 / chunk_w, cx1 = (col1 - 1) / chunk_w;
  const int64_t n_cx = cx1 - cx0 + 1, n_grid = (cy1 - cy0 + 1) * n_cx, n_work = n_grid * planes;
  const bool compressed = tiff_compressed(info->compression);
  const bool direct = !compressed && info->predictor == 1;     // no scratch: copy out of the file
  const bool file_swap = sb > 1 && info->big_endian;
  TiffCursor c{static_cast<const uint8_t*>(file), file_bytes, info->big_endian != 0};
  uint8_t* out = static_cast<uint8_t*>(dst);
  std::atomic<int64_t> next{0};
  std::atomic<const char*> error{nullptr};

  auto worker = [&]() {
    uint8_t* scratch = nullptr;
    size_t scratch_bytes = 0;
    for (;;) {
      const int64_t k = next.fetch_add(1);
      if (k >= n_work || error.load() != nullptr) break;
      const int64_t plane = k / n_grid, kg = k % n_grid;
      const int64_t cy = cy0 + kg / n_cx, cx = cx0 + kg % n_cx;
      const uint64_t idx = (uint64_t)(plane * per_plane + cy * across + cx);
      if (idx >= (uint64_t)info->n_strips) { error = "corrupt info block"; break; }
      const int64_t valid_rows = (info->height - cy * chunk_h < chunk_h) ? info->height - cy * chunk_h : chunk_h;
      const uint64_t need = chunk_row_bytes * (uint64_t)valid_rows;                       // bytes the image uses
      const uint64_t cap = chunk_row_bytes * (uint64_t)(tiled ? chunk_h : valid_rows);    // bytes a chunk may hold
      const uint64_t off = tiff_value(c, info->strip_offsets_pos, info->strip_offsets_type, idx);
      const uint8_t* base;
      bool swap = file_swap;
      if (direct) {
        if (!c.ok(off, need)) { error = "a strip / tile runs past the end of the file"; break; }
        base = c.p + off;
      } else {
        if (scratch_bytes < cap) {
          free(scratch);
          scratch = static_cast<uint8_t*>(malloc(cap));
          scratch_bytes = scratch ? cap : 0;
          if (!scratch) { error = "out of host memory for a decode buffer"; break; }
        }
        uint64_t produced;
        if (!compressed) {
          if (!c.ok(off, need)) { error = "a strip / tile runs past the end of the file"; break; }
          memcpy(scratch, c.p + off, need);
          produced = need;
        } else {
          const uint64_t cnt = tiff_value(c, info->strip_counts_pos, info->strip_counts_type, idx);
          if (!c.ok(off, cnt)) { error = "a strip / tile runs past the end of the file"; break; }
          switch (info->compression) {
            case 5: produced = lzw_chunk(c.p + off, cnt, scratch, cap); break;
            case 8: produced = zlib_uncompress() ? inflate_chunk(c.p + off, cnt, scratch, cap) : 0; break;
            default: produced = packbits_chunk(c.p + off, cnt, scratch, cap); break;
          }
        }
        if (produced < need) { error = "a compressed strip / tile is corrupt or shorter than its rows"; break; }
        if (swap) { swap_inplace(scratch, need, sb); swap = false; }
        if (info->predictor == 2)
          for (int64_t r = 0; r < valid_rows; ++r) undo_predictor_row(scratch + (uint64_t)r * chunk_row_bytes, (uint64_t)chunk_w, file_spp, sb);
        base = scratch;
      }
      // the part of this chunk that lies inside the region
      const int64_t r_lo = row0 > cy * chunk_h ? row0 : cy * chunk_h;
      const int64_t r_hi = row1 < cy * chunk_h + valid_rows ? row1 : cy * chunk_h + valid_rows;
      const int64_t c_lo = col0 > cx * chunk_w ? col0 : cx * chunk_w;
      int64_t c_hi = (cx + 1) * chunk_w < info->width ? (cx + 1) * chunk_w : info->width;
      if (col1 < c_hi) c_hi = col1;
      const uint64_t run = (uint64_t)(c_hi - c_lo) * file_px_bytes;
      const uint8_t* src = base + (uint64_t)(r_lo - cy * chunk_h) * chunk_row_bytes + (uint64_t)(c_lo - cx * chunk_w) * file_px_bytes;
      uint8_t* o = out + (uint64_t)(r_lo - row0) * out_row_bytes + (uint64_t)(c_lo - col0) * px_bytes;
      if (planes > 1) {                                          // one sample of every pixel: interleave on the way out
        o += (uint64_t)plane * sb;
        for (int64_t r = r_lo; r < r_hi; ++r, src += chunk_row_bytes, o += out_row_bytes) {
          const int64_t n_px = c_hi - c_lo;
          if (sb == 1) {
            for (int64_t x = 0; x < n_px; ++x) o[(uint64_t)x * px_bytes] = src[x];
          } else if (sb == 4) {
            for (int64_t x = 0; x < n_px; ++x)
              for (int b = 0; b < 4; ++b) o[(uint64_t)x * px_bytes + b] = src[4 * x + (swap ? 3 - b : b)];
          } else {
            for (int64_t x = 0; x < n_px; ++x) {
              o[(uint64_t)x * px_bytes] = src[2 * x + (swap ? 1 : 0)];
              o[(uint64_t)x * px_bytes + 1] = src[2 * x + (swap ? 0 : 1)];
            }
          }
        }
      } else if (run == chunk_row_bytes && run == out_row_bytes) {   // full-width rows on both sides: one run
        copy_samples(o, src, run * (uint64_t)(r_hi - r_lo), swap, sb);
      } else {
        for (int64_t r = r_lo; r < r_hi; ++r, src += chunk_row_bytes, o += out_row_bytes) copy_samples(o, src, run, swap, sb);
      }
    }
    free(scratch);
  };

  int64_t n_thr = n_threads < 1 ? 1 : n_threads;
  if (n_thr > n_work) n_thr = n_work;
  if (n_thr > 64) n_thr = 64;
  std::vector<std::thread> pool;
  try {
    for (int64_t t = 1; t < n_thr; ++t) pool.emplace_back(worker);
  } catch (...) {
    // could not start (all of) the helpers: the calling thread drains the counter on its own
  }
  worker();
  for (auto& th : pool) th.join();
  return error.load();
}

}  // namespace lars_host
