// Host-side ingest: baseline TIFF 6.0 reader for the frame formats of the path (uncompressed,
// chunky, strips, 8- or 16-bit unsigned samples, 1 / 3 / 4 samples per pixel, either byte order).
//
// Why it exists: the reference loads frames with
// PIL.Image.open (process-images.py:183-193; backend-process.py:52; process-ndvi.py:18;
// process-rgn.py:18), and Pillow opens a 16-bit RGB TIFF as 8-bit -- 16-bit survey frames
// (BASELINE config 3) cannot enter the reference through files at all.  This reader copies the
// strips of a memory-mapped file straight into a caller-supplied (pinned) HWC buffer, byte-swapping
// big-endian 16-bit samples on the way, so file bytes reach the H2D copy without a decode step.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

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
};

inline int tiff_type_size(int type) {
  switch (type) {
    case 1: case 2: case 6: case 7: return 1;   // BYTE ASCII SBYTE UNDEFINED
    case 3: case 8: return 2;                   // SHORT SSHORT
    case 4: case 9: case 11: return 4;          // LONG SLONG FLOAT
    case 5: case 10: case 12: return 8;         // RATIONAL SRATIONAL DOUBLE
    default: return 0;
  }
}

// value i of an IFD entry of type SHORT or LONG; `pos` is where the values live
inline uint32_t tiff_value(const TiffCursor& c, uint64_t pos, int type, uint32_t i) {
  return type == 3 ? c.u16(pos + 2ull * i) : c.u32(pos + 4ull * i);
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
  if (magic == 43) { *unsupported = true; return "BigTIFF is not supported"; }
  if (magic != 42) return "not a TIFF file (magic number)";
  const uint64_t ifd = c.u32(4);
  if (!c.ok(ifd, 2)) return "IFD offset outside the file";
  const uint32_t n_entries = c.u16(ifd);
  if (!c.ok(ifd + 2, 12ull * n_entries)) return "IFD runs past the end of the file";

  info->big_endian = c.be ? 1 : 0;
  info->samples_per_pixel = 1;
  info->bits_per_sample = 1;
  info->compression = 1;
  info->planar_config = 1;
  info->sample_format = 1;
  info->rows_per_strip = -1;
  uint32_t offsets_count = 0, counts_count = 0;
  bool bits_mixed = false;
  for (uint32_t e = 0; e < n_entries; ++e) {
    const uint64_t at = ifd + 2 + 12ull * e;
    const int tag = c.u16(at), type = c.u16(at + 2);
    const uint32_t count = c.u32(at + 4);
    const int tsz = tiff_type_size(type);
    if (tsz == 0) continue;                                   // unknown field type: ignore the entry
    const uint64_t bytes = (uint64_t)tsz * count;
    const uint64_t pos = bytes <= 4 ? at + 8 : c.u32(at + 8); // values inline or at an offset
    if (!c.ok(pos, bytes)) return "an IFD entry points outside the file";
    const bool integral = (type == 3 || type == 4);
    const uint32_t v0 = (integral && count >= 1) ? tiff_value(c, pos, type, 0) : 0;
    switch (tag) {
      case 256: info->width = (int32_t)v0; break;
      case 257: info->height = (int32_t)v0; break;
      case 258:
        info->bits_per_sample = (int32_t)v0;
        for (uint32_t i = 1; i < count && integral; ++i)
          if (tiff_value(c, pos, type, i) != v0) bits_mixed = true;
        break;
      case 259: info->compression = (int32_t)v0; break;
      case 262: info->photometric = (int32_t)v0; break;
      case 273: info->strip_offsets_pos = pos; info->strip_offsets_type = type; offsets_count = count; break;
      case 277: info->samples_per_pixel = (int32_t)v0; break;
      case 278: info->rows_per_strip = (v0 > 0x7fffffffu) ? -1 : (int32_t)v0; break;
      case 279: info->strip_counts_pos = pos; info->strip_counts_type = type; counts_count = count; break;
      case 284: info->planar_config = (int32_t)v0; break;
      case 317: if (v0 != 1) { *unsupported = true; return "TIFF predictor is not supported"; } break;
      case 322: case 323: case 324: case 325: *unsupported = true; return "tiled TIFF is not supported";
      case 339: info->sample_format = (int32_t)v0; break;
      default: break;
    }
  }
  if (info->width < 1 || info->height < 1) return "missing ImageWidth / ImageLength";
  if (info->width > (1 << 24) || info->height > (1 << 24)) return "implausible image dimensions";   // keeps all byte counts far below 2^64
  if (info->compression != 1) { *unsupported = true; return "compressed TIFF: decode it with Pillow"; }
  if (bits_mixed || (info->bits_per_sample != 8 && info->bits_per_sample != 16)) {
    *unsupported = true;
    return "only 8- or 16-bit samples are supported";
  }
  if (info->sample_format != 1) { *unsupported = true; return "only unsigned integer samples are supported"; }
  if (info->samples_per_pixel != 1 && info->samples_per_pixel != 3 && info->samples_per_pixel != 4) {
    *unsupported = true;
    return "only 1, 3 or 4 samples per pixel are supported";
  }
  if (info->planar_config != 1 && info->samples_per_pixel != 1) {
    *unsupported = true;
    return "planar (separate-plane) TIFF is not supported";
  }
  if (!info->strip_offsets_pos) return "missing StripOffsets";
  if (info->strip_offsets_type != 3 && info->strip_offsets_type != 4) return "StripOffsets must be SHORT or LONG";
  if (info->rows_per_strip < 1 || info->rows_per_strip > info->height) info->rows_per_strip = info->height;
  info->n_strips = (info->height + info->rows_per_strip - 1) / info->rows_per_strip;
  if (offsets_count != (uint32_t)info->n_strips) return "StripOffsets count does not match the image height";
  if (info->strip_counts_pos) {
    if (info->strip_counts_type != 3 && info->strip_counts_type != 4) return "StripByteCounts must be SHORT or LONG";
    if (counts_count != offsets_count) return "StripByteCounts count does not match StripOffsets";
  }
  const uint64_t row_bytes = (uint64_t)info->width * info->samples_per_pixel * (info->bits_per_sample / 8);
  info->frame_bytes = row_bytes * (uint64_t)info->height;
  // every strip must lie inside the file and hold its rows
  for (int s = 0; s < info->n_strips; ++s) {
    const uint64_t off = tiff_value(c, info->strip_offsets_pos, info->strip_offsets_type, (uint32_t)s);
    const int rows = (s == info->n_strips - 1) ? info->height - s * info->rows_per_strip : info->rows_per_strip;
    const uint64_t need = row_bytes * (uint64_t)rows;
    if (!c.ok(off, need)) return "a strip runs past the end of the file";
    if (info->strip_counts_pos &&
        tiff_value(c, info->strip_counts_pos, info->strip_counts_type, (uint32_t)s) < need)
      return "a strip is shorter than its rows";
  }
  return nullptr;
}

// Copies the strips into dst as one contiguous HWC frame with little-endian samples.
inline const char* tiff_read(const void* file, size_t file_bytes, const lars_tiff_info* info, void* dst,
                             size_t dst_bytes) {
  if (dst_bytes < info->frame_bytes) return "destination buffer smaller than the frame";
  TiffCursor c{static_cast<const uint8_t*>(file), file_bytes, info->big_endian != 0};
  const uint64_t row_bytes = (uint64_t)info->width * info->samples_per_pixel * (info->bits_per_sample / 8);
  uint8_t* out = static_cast<uint8_t*>(dst);
  const bool swap = info->bits_per_sample == 16 && info->big_endian;
  for (int s = 0; s < info->n_strips; ++s) {
    if (!c.ok(info->strip_offsets_pos, 4)) return "corrupt info block";
    const uint64_t off = tiff_value(c, info->strip_offsets_pos, info->strip_offsets_type, (uint32_t)s);
    const int rows = (s == info->n_strips - 1) ? info->height - s * info->rows_per_strip : info->rows_per_strip;
    const uint64_t bytes = row_bytes * (uint64_t)rows;
    if (!c.ok(off, bytes)) return "a strip runs past the end of the file";
    uint8_t* o = out + (uint64_t)s * info->rows_per_strip * row_bytes;
    if (!swap) {
      memcpy(o, c.p + off, bytes);
    } else {
      const uint8_t* in = c.p + off;
      for (uint64_t i = 0; i + 1 < bytes; i += 2) { o[i] = in[i + 1]; o[i + 1] = in[i]; }
    }
  }
  return nullptr;
}

}  // namespace lars_host
