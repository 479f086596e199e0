// Device-side decode of LZW-compressed TIFF strips (GPU-side decode).
//
// The compressed file bytes are uploaded as they are; every strip is an independent LZW stream, so a batch
// of frames offers thousands of them: one warp per strip (lzw_warp.h), table in shared memory, output written
// straight into the padded HWC frame slots the analysis passes read.  A second small kernel undoes the
// horizontal-differencing predictor and turns big-endian 16-bit samples into native ones, one thread per
// (row, sample of the pixel) running along the row.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lars_b200.h"
#include "lzw_warp.h"
#include "inflate_warp.h"

namespace lars {

constexpr int LZW_WARPS = 4;                                  // per CTA
constexpr int LZW_SMEM_BYTES = LZW_WARPS * 4096 * 4;          // 64 KB: three CTAs (12 warps) per SM

struct LzwParams {
  const uint8_t* src;              // compressed bytes of every file of the batch
  const lars_lzw_chunk* chunks;    // [n_chunks]
  uint8_t* dst;                    // frame batch
  uint32_t* status;                // [1]: number of corrupt / short strips
  unsigned int* next;              // [1]: work counter (zeroed by the caller)
  int n_chunks;
};

__global__ void __launch_bounds__(LZW_WARPS * 32) lzw_decode_kernel(const LzwParams p) {
  extern __shared__ __align__(16) uint32_t lzw_tables[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* table = lzw_tables + warp * 4096;
  for (;;) {
    unsigned int k = 0;
    if (lane == 0) k = atomicAdd(p.next, 1u);                 // strips differ in length: dynamic hand-out
    k = __shfl_sync(0xffffffffu, k, 0);
    if (k >= (unsigned int)p.n_chunks) break;
    const lars_lzw_chunk c = p.chunks[k];
    const uint32_t produced = lars_lzw_decode_warp(p.src + c.src_offset, c.src_bytes, p.dst + c.dst_offset,
                                                   c.dst_bytes, table);
    if (lane == 0 && produced < c.dst_bytes) atomicAdd(p.status, 1u);
    __syncwarp();
  }
}

// Deflate strips (inflate_warp.h): one warp per zlib stream, 5 warps per CTA, one CTA per SM.
constexpr int INF_WARPS = 5;
constexpr int INF_SMEM_BYTES = INF_WARPS * (int)sizeof(LarsInflateSmem);     // 194 KB

__global__ void __launch_bounds__(INF_WARPS * 32, 1) inflate_decode_kernel(const LzwParams p) {
  extern __shared__ __align__(16) uint32_t lzw_tables[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  LarsInflateSmem* sm = reinterpret_cast<LarsInflateSmem*>(lzw_tables) + warp;
  for (;;) {
    unsigned int k = 0;
    if (lane == 0) k = atomicAdd(p.next, 1u);
    k = __shfl_sync(0xffffffffu, k, 0);
    if (k >= (unsigned int)p.n_chunks) break;
    const lars_lzw_chunk c = p.chunks[k];
    const uint32_t produced = lars_inflate_warp(p.src + c.src_offset, c.src_bytes, p.dst + c.dst_offset, c.dst_bytes, sm);
    if (lane == 0 && produced < c.dst_bytes) atomicAdd(p.status, 1u);
    __syncwarp();
  }
}

// Strips / tiles decoded into scratch slots -> their place inside one frame (a row band of a mosaic): per chunk a
// row range of the slot goes to a row range of the frame at a byte column.  One CTA per (chunk, row stripe).
struct UntileParams {
  const uint8_t* scratch;
  const long long* table;          // [n_chunks][5]: first slot row, rows, first frame row, byte column in the frame, bytes per row
  uint8_t* dst;
  long long slot_bytes, slot_row_bytes, dst_row_bytes;
  int n_chunks;
};

__global__ void __launch_bounds__(256) untile_kernel(const UntileParams p) {
  const long long* t = p.table + 5ll * blockIdx.x;
  const long long src_row0 = t[0], rows = t[1], dst_row0 = t[2], dst_col = t[3], n = t[4];
  const uint8_t* src = p.scratch + (long long)blockIdx.x * p.slot_bytes + src_row0 * p.slot_row_bytes;
  uint8_t* dst = p.dst + dst_row0 * p.dst_row_bytes + dst_col;
  for (long long r = blockIdx.y; r < rows; r += gridDim.y) {
    const uint8_t* s = src + r * p.slot_row_bytes;
    uint8_t* d = dst + r * p.dst_row_bytes;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
  }
}

struct TiffPostParams {
  uint8_t* dst;
  long long frame_stride;          // bytes between frames
  long long row_bytes;
  int n_frames, rows, width, spp, sample_bytes, predictor, swap16;
};

// thread = (frame, row, sample k of the pixel): walks its row, undoing byte order and differencing
__global__ void __launch_bounds__(256) tiff_post_kernel(const TiffPostParams p) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_frame = (long long)p.rows * p.spp;
  if (t >= per_frame * p.n_frames) return;
  const long long f = t / per_frame, r = (t % per_frame) / p.spp;
  const int k = (int)(t % p.spp);
  uint8_t* row = p.dst + f * p.frame_stride + r * p.row_bytes;
  if (p.sample_bytes == 1) {
    uint32_t acc = 0;
    for (int x = 0; x < p.width; ++x) {
      uint8_t* q = row + (long long)x * p.spp + k;
      acc = (acc + *q) & 0xFFu;
      *q = (uint8_t)acc;
    }
  } else {
    uint32_t acc = 0;
    for (int x = 0; x < p.width; ++x) {
      uint16_t* q = reinterpret_cast<uint16_t*>(row) + (long long)x * p.spp + k;
      uint32_t v = *q;
      if (p.swap16) v = ((v & 0xFFu) << 8) | (v >> 8);
      if (p.predictor == 2) { acc = (acc + v) & 0xFFFFu; v = acc; }
      *q = (uint16_t)v;
    }
  }
}

}  // namespace lars
