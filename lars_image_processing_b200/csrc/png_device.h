// PNG row filters undone on the device (experimental, with inflate_warp.h: not yet run on hardware).  After the
// image data of a batch of PNG frames has been inflated into [image][row][1 + row_bytes] scratch, every byte lane k of
// the pixel (k = 0 .. bytes_per_pixel - 1) is independent of the others for all five filters -- left, up and
// upper-left neighbours of a byte are bytes of the same lane -- so one thread owns one lane of one image and walks
// rows and columns in order, writing the frame slot directly (16-bit samples leave in little-endian order).
// The same source is compiled for the host by tests/hostcheck and compared with the host PNG reader.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LARS_PNG_HD __host__ __device__ inline
#else
#define LARS_PNG_HD static inline
#endif

// raw: h rows of (1 filter byte + row_bytes); dst: h rows of row_bytes (HWC).  Returns false for an unknown filter.
LARS_PNG_HD bool lars_png_unfilter_lane(const uint8_t* raw, uint8_t* dst, int h, long long row_bytes, int bpp, int k, int swap16) {
  const int ko = swap16 ? (k ^ 1) : k;                     // where this lane's byte goes inside the pixel
  for (int r = 0; r < h; ++r) {
    const uint8_t* cur = raw + (long long)r * (row_bytes + 1);
    const int filter = cur[0];
    if (filter > 4) return false;
    uint8_t* o = dst + (long long)r * row_bytes;
    const uint8_t* up = r ? o - row_bytes : nullptr;
    int a = 0, c = 0;                                      // left and upper-left of the current byte
    for (long long x = 0; x + bpp <= row_bytes; x += bpp) {
      const int b = up ? up[x + ko] : 0;
      int pred;
      switch (filter) {
        case 0: pred = 0; break;
        case 1: pred = a; break;
        case 2: pred = b; break;
        case 3: pred = (a + b) >> 1; break;
        default: {
          const int bc = b - c, ac = a - c, s = bc + ac;
          int pa = bc < 0 ? -bc : bc;
          const int pb = ac < 0 ? -ac : ac, pc = s < 0 ? -s : s;
          pred = a;
          pred = pb < pa ? b : pred;
          pa = pb < pa ? pb : pa;
          pred = pc < pa ? c : pred;
        }
      }
      a = (cur[1 + x + k] + pred) & 255;
      o[x + ko] = (uint8_t)a;
      c = b;
    }
  }
  return true;
}
