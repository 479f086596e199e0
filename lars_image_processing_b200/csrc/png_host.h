// Host-side ingest: PNG reader for the frame formats of the path -- non-interlaced grayscale, RGB or
// RGBA with 8- or 16-bit samples (BASELINE config 1 is a 1280x960 uint8 RGNir PNG).
//
// Why it exists: the reference loads frames with PIL.Image.open
// (process-images.py:183-193; backend-process.py:52; process-ndvi.py:18; process-rgn.py:18).  In the
// many-frame survey pipeline Pillow's PNG decode is the bottleneck (it alternates between Python and C
// per 64 KB of file, so decode threads queue on the interpreter lock); this reader is one C call per
// frame -- chunk walk, one zlib inflate of the concatenated IDAT data, the five PNG row filters --
// straight into the caller's (pinned) HWC buffer.  16-bit samples come out as native little-endian
// uint16 (Pillow reduces 16-bit RGB to 8 bits).  Palette, gray + alpha, sub-byte depths and Adam7
// interlace return LARS_ERR_UNSUPPORTED and the caller decodes with Pillow.  Written from the PNG
// specification (ISO/IEC 15948 sections 5, 9, 10, 11); zlib is bound at run time like in tiff_host.h.
// Chunk CRCs are not verified: the zlib stream's own Adler-32 guards the pixel data.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/lars_b200.h"
#include "tiff_host.h"   // zlib_uncompress()

namespace lars_host {

inline uint32_t png_be32(const uint8_t* p) {
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

inline int png_channels(int color_type) { return color_type == 0 ? 1 : color_type == 2 ? 3 : color_type == 6 ? 4 : 0; }

// Returns NULL on success, else a static description of what is wrong / unsupported.
inline const char* png_probe(const void* file, size_t file_bytes, lars_png_info* info, bool* unsupported) {
  static const uint8_t kSignature[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  *unsupported = false;
  memset(info, 0, sizeof(*info));
  const uint8_t* p = static_cast<const uint8_t*>(file);
  if (file_bytes < 8 || memcmp(p, kSignature, 8) != 0) return "not a PNG file (signature)";
  uint64_t at = 8;
  bool have_ihdr = false, ended = false;
  while (!ended) {
    if (file_bytes - at < 12) return have_ihdr ? "PNG ends without IEND" : "PNG without IHDR";
    const uint64_t len = png_be32(p + at);
    const uint8_t* type = p + at + 4;
    if (len > 0x7fffffffull || len + 12 > file_bytes - at) return "a PNG chunk runs past the end of the file";
    const uint8_t* data = p + at + 8;
    if (!have_ihdr) {
      if (memcmp(type, "IHDR", 4) != 0 || len != 13) return "PNG does not start with IHDR";
      have_ihdr = true;
      const uint32_t w = png_be32(data), h = png_be32(data + 4);
      if (w < 1 || h < 1 || w > (1u << 24) || h > (1u << 24)) return "implausible PNG dimensions";
      info->width = (int32_t)w;
      info->height = (int32_t)h;
      info->bit_depth = data[8];
      info->color_type = data[9];
      info->interlace = data[12];
      if (data[10] != 0 || data[11] != 0) return "unknown PNG compression / filter method";
      info->channels = png_channels(info->color_type);
      if (info->channels == 0) { *unsupported = true; return "palette and gray + alpha PNGs: decode with Pillow"; }
      if (info->bit_depth != 8 && info->bit_depth != 16) { *unsupported = true; return "only 8- or 16-bit PNG samples are supported"; }
      if (info->interlace != 0) { *unsupported = true; return "interlaced PNG: decode with Pillow"; }
    } else if (memcmp(type, "IDAT", 4) == 0) {
      info->n_idat += 1;
      info->idat_bytes += len;
    } else if (memcmp(type, "IEND", 4) == 0) {
      ended = true;
    } else if (memcmp(type, "tRNS", 4) == 0) {
      *unsupported = true;                    // Pillow keeps the array but some writers rely on it: leave it to Pillow
      return "PNG with a tRNS chunk: decode with Pillow";
    }
    at += 12 + len;
  }
  if (info->n_idat == 0) return "PNG without image data";
  if (!zlib_uncompress()) { *unsupported = true; return "no zlib on this system: decode the PNG with Pillow"; }
  info->frame_bytes = (uint64_t)info->width * (uint64_t)info->channels * (uint64_t)(info->bit_depth / 8) * (uint64_t)info->height;
  return nullptr;
}

// Paeth rows: the predictor is whichever of left (a), up (b), upper-left (c) is closest to a + b - c, ties in
// that order; |p - a| = |b - c|, |p - b| = |a - c|, |p - c| = |(b - c) + (a - c)|.  The left and upper-left
// pixels stay in registers (no store-to-load forwarding on the dependency chain) and the choice is made
// with selects, not branches: on noisy frames the branches mispredict about every other byte (2.6x slower).
template <int BPP>
inline void png_paeth_row(const uint8_t* cur, const uint8_t* up, uint8_t* dst, size_t n) {
  int a[BPP], c[BPP];
  for (int k = 0; k < BPP; ++k) a[k] = c[k] = 0;   // left of the first pixel: zeros, so its predictor is `up`
  for (size_t i = 0; i + BPP <= n; i += BPP) {
    for (int k = 0; k < BPP; ++k) {
      const int b = up[i + k];
      const int bc = b - c[k], ac = a[k] - c[k], s = bc + ac;
      int pa = bc < 0 ? -bc : bc;
      const int pb = ac < 0 ? -ac : ac, pc = s < 0 ? -s : s;
      int pred = a[k];
      pred = pb < pa ? b : pred;
      pa = pb < pa ? pb : pa;
      pred = pc < pa ? c[k] : pred;
      a[k] = (cur[i + k] + pred) & 255;
      dst[i + k] = (uint8_t)a[k];
      c[k] = b;
    }
  }
}

// One row: `cur` (filtered bytes) -> `dst`; `up` is the reconstructed previous row or NULL for row 0
// (an all-zero row above: Up becomes None, Paeth becomes Sub, Average halves the left neighbour).
inline bool png_unfilter_row(int filter, const uint8_t* cur, const uint8_t* up, uint8_t* dst, size_t n, size_t bpp) {
  const size_t head = bpp < n ? bpp : n;
  switch (filter) {
    case 0:
      memcpy(dst, cur, n);
      return true;
    case 1:
      memcpy(dst, cur, head);
      for (size_t i = bpp; i < n; ++i) dst[i] = (uint8_t)(cur[i] + dst[i - bpp]);
      return true;
    case 2:
      if (!up) { memcpy(dst, cur, n); return true; }
      for (size_t i = 0; i < n; ++i) dst[i] = (uint8_t)(cur[i] + up[i]);
      return true;
    case 3:
      if (!up) {
        memcpy(dst, cur, head);
        for (size_t i = bpp; i < n; ++i) dst[i] = (uint8_t)(cur[i] + (dst[i - bpp] >> 1));
        return true;
      }
      for (size_t i = 0; i < head; ++i) dst[i] = (uint8_t)(cur[i] + (up[i] >> 1));
      for (size_t i = bpp; i < n; ++i) dst[i] = (uint8_t)(cur[i] + ((dst[i - bpp] + up[i]) >> 1));
      return true;
    case 4:
      if (!up) {
        memcpy(dst, cur, head);
        for (size_t i = bpp; i < n; ++i) dst[i] = (uint8_t)(cur[i] + dst[i - bpp]);
        return true;
      }
      switch (bpp) {                                 // bytes per pixel: 1 / 2 gray, 3 / 6 RGB, 4 / 8 RGBA
        case 1: png_paeth_row<1>(cur, up, dst, n); break;
        case 2: png_paeth_row<2>(cur, up, dst, n); break;
        case 3: png_paeth_row<3>(cur, up, dst, n); break;
        case 4: png_paeth_row<4>(cur, up, dst, n); break;
        case 6: png_paeth_row<6>(cur, up, dst, n); break;
        default: png_paeth_row<8>(cur, up, dst, n); break;
      }
      return true;
    default:
      return false;
  }
}

// dst receives height x width x channels, little-endian samples, rows contiguous.
inline const char* png_read(const void* file, size_t file_bytes, const lars_png_info* info, void* dst, size_t dst_bytes) {
  if ((uint64_t)dst_bytes < info->frame_bytes) return "destination buffer smaller than the frame";
  const uint8_t* p = static_cast<const uint8_t*>(file);
  const size_t bpp = (size_t)info->channels * (size_t)(info->bit_depth / 8);
  const size_t row_bytes = (size_t)info->width * bpp;
  const size_t raw_bytes = (row_bytes + 1) * (size_t)info->height;
  if (raw_bytes > 0xffffffffffull) return "implausible PNG size";
  // gather the IDAT payloads (one zlib stream cut into chunks) unless there is only one
  const uint8_t* z = nullptr;
  uint8_t* joined = nullptr;
  size_t z_bytes = 0;
  {
    uint64_t at = 8;
    size_t filled = 0;
    for (;;) {
      if (file_bytes - at < 12) break;
      const uint64_t len = png_be32(p + at);
      if (len + 12 > file_bytes - at) break;
      const uint8_t* type = p + at + 4;
      if (memcmp(type, "IDAT", 4) == 0) {
        if (info->n_idat == 1) { z = p + at + 8; z_bytes = (size_t)len; break; }
        if (!joined) {
          joined = static_cast<uint8_t*>(malloc(info->idat_bytes ? info->idat_bytes : 1));
          if (!joined) return "out of host memory for the PNG data";
        }
        if (filled + len > info->idat_bytes) { free(joined); return "corrupt info block"; }
        memcpy(joined + filled, p + at + 8, (size_t)len);
        filled += (size_t)len;
      } else if (memcmp(type, "IEND", 4) == 0) {
        break;
      }
      at += 12 + len;
    }
    if (info->n_idat != 1) { z = joined; z_bytes = filled; }
  }
  if (!z) { free(joined); return "PNG without image data"; }
  uint8_t* raw = static_cast<uint8_t*>(malloc(raw_bytes));
  if (!raw) { free(joined); return "out of host memory for the PNG rows"; }
  unsigned long produced = (unsigned long)raw_bytes;
  const int rc = zlib_uncompress()(raw, &produced, z, (unsigned long)z_bytes);
  free(joined);
  if (rc != 0 || produced != raw_bytes) {        // strict: the stream must end, checksum verified, exactly at the last row
    free(raw);
    return "the PNG image data are corrupt or shorter than the image";
  }
  uint8_t* out = static_cast<uint8_t*>(dst);
  for (int32_t r = 0; r < info->height; ++r) {
    const uint8_t* cur = raw + (size_t)r * (row_bytes + 1);
    uint8_t* o = out + (size_t)r * row_bytes;
    if (!png_unfilter_row(cur[0], cur + 1, r ? o - row_bytes : nullptr, o, row_bytes, bpp)) {
      free(raw);
      return "unknown PNG row filter";
    }
  }
  free(raw);
  if (info->bit_depth == 16) swap16_inplace(out, (uint64_t)row_bytes * (uint64_t)info->height);   // network order -> little-endian
  return nullptr;
}

}  // namespace lars_host
