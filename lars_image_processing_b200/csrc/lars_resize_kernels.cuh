// K10: Lanczos down/up-scaling of interleaved uint8 frames, bit-identical to Pillow's
// Image.resize(..., Image.Resampling.LANCZOS) -- the step in front of the analysis path
// (preprocess_large_image, process-images.py:398-422).
//
// Pillow's algorithm (src/libImaging/Resample.c): separable two-pass convolution, horizontal pass
// first into a uint8 intermediate image that only holds the source rows the vertical pass uses,
// then the vertical pass; coefficients are 22-bit fixed point, each output sample is
// clip8((2^21 + sum(sample * coef)) >> 22) in 32-bit integer arithmetic.  The coefficient tables
// are computed on the HOST (lars_resize_tables_lanczos: double precision + libm sin, the same
// evaluation Pillow performs) and handed to the kernels as a device block.
//
//   K10h (horizontal): a warp owns one output column at a time and its 32 lanes are 32 image ROWS, so
//         window position, tap count and coefficients are warp-uniform.  A CTA stages a
//         [32 rows] x [input span of its output columns] tile in shared memory TRANSPOSED
//         (word column major, 33-word pitch): the lanes of a warp then hit 32 different banks for
//         every tap.  Results go through a second shared tile so the global stores are row-contiguous.
//   K10v (vertical): lanes run along the row (4 consecutive bytes per thread); every tap is one
//         coalesced 32-bit load of the intermediate image (L2-resident: it is 1/scale of the input).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lars {

constexpr int RS_PRECISION_BITS = 32 - 8 - 2;   // Resample.c PRECISION_BITS for 8-bit channels
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROWS = 32;                     // rows per K10h tile (= lanes)
constexpr int RS_IN_PITCH = 33;                 // words between word columns of the transposed tile

__device__ __forceinline__ uint32_t rs_clip8(int acc) {
  const int v = acc >> RS_PRECISION_BITS;       // arithmetic shift, as Resample.c clip8()
  return (uint32_t)min(max(v, 0), 255);
}

struct ResizeHParams {
  const uint8_t* src;        // [frame][in_h][in_w * C]
  uint8_t* dst;              // [frame][row_count][out_w * C]
  const int* bounds;         // [out_w][2]  (first source column, tap count)
  const int* kk;             // [out_w][ksize]
  long long src_frame_stride, dst_frame_stride;
  int in_w, out_w, ksize;
  int row_first, row_count;  // source rows to produce
  int xo_tile;               // output columns per CTA
  int span_words;            // shared words per row of the input tile (max over tiles)
  int out_pitch;             // bytes per row of the output tile (odd number of words)
};

template <int C>
__global__ void __launch_bounds__(RS_THREADS) resize_h_kernel(const ResizeHParams p) {
  extern __shared__ __align__(16) uint8_t rs_smem[];
  uint8_t* in_s = rs_smem;                                                  // [span_words][33] words
  uint8_t* out_s = rs_smem + (size_t)p.span_words * RS_IN_PITCH * 4;        // [32][out_pitch] bytes
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int xo0 = blockIdx.x * p.xo_tile;
  const int nxo = min(p.xo_tile, p.out_w - xo0);
  const int r0 = blockIdx.y * RS_ROWS;                                      // relative to row_first
  const int nrows = min(RS_ROWS, p.row_count - r0);
  const long long in_pitch = (long long)p.in_w * C;
  const uint8_t* src = p.src + blockIdx.z * p.src_frame_stride + (long long)(p.row_first + r0) * in_pitch;
  uint8_t* dst = p.dst + blockIdx.z * p.dst_frame_stride + (long long)r0 * p.out_w * C;

  const int b0 = p.bounds[2 * xo0] * C;                                     // first input byte of the tile
  const int b1 = (p.bounds[2 * (xo0 + nxo - 1)] + p.bounds[2 * (xo0 + nxo - 1) + 1]) * C;
  const int span = b1 - b0;

  // ---- stage: global rows (lanes along the row) -> transposed shared tile ----
  for (int r = warp; r < nrows; r += RS_WARPS) {
    const uint8_t* row = src + (long long)r * in_pitch + b0;
    for (int j = lane; j < span; j += 32) in_s[((j >> 2) * RS_IN_PITCH + r) * 4 + (j & 3)] = row[j];
  }
  __syncthreads();

  // ---- convolve: warp = output column, lane = row ----
  const uint8_t* my_in = in_s + lane * 4;
  for (int xl = warp; xl < nxo; xl += RS_WARPS) {
    const int xo = xo0 + xl;
    const int first = p.bounds[2 * xo], cnt = p.bounds[2 * xo + 1];
    const int* kk = p.kk + (long long)xo * p.ksize;
    int acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 1 << (RS_PRECISION_BITS - 1);
    int j = first * C - b0;
    for (int k = 0; k < cnt; ++k) {
      const int coef = __ldg(kk + k);
#pragma unroll
      for (int c = 0; c < C; ++c, ++j) acc[c] += (int)my_in[(j >> 2) * (RS_IN_PITCH * 4) + (j & 3)] * coef;
    }
    uint8_t* o = out_s + lane * p.out_pitch + xl * C;
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = (uint8_t)rs_clip8(acc[c]);
  }
  __syncthreads();

  // ---- write: row-contiguous stores ----
  const int out_bytes = nxo * C;
  for (int r = warp; r < nrows; r += RS_WARPS) {
    uint8_t* row = dst + (long long)r * p.out_w * C + (long long)xo0 * C;
    const uint8_t* s = out_s + r * p.out_pitch;
    for (int j = lane; j < out_bytes; j += 32) row[j] = s[j];
  }
}

struct ResizeVParams {
  const uint8_t* src;        // [frame][rows][row_bytes]  (intermediate image, or the source when no horizontal pass)
  uint8_t* dst;              // [frame][out_h][row_bytes]
  const int* bounds;         // [out_h][2]  (first row relative to the intermediate image, tap count)
  const int* kk;             // [out_h][ksize]
  long long src_frame_stride, dst_frame_stride;
  int row_bytes, out_h, ksize;
};

// VEC = 4: row_bytes and both frame strides are multiples of 4; VEC = 1: anything.
template <int VEC>
__global__ void __launch_bounds__(RS_THREADS) resize_v_kernel(const ResizeVParams p) {
  const int yo = blockIdx.y;
  const int xb = (blockIdx.x * RS_THREADS + threadIdx.x) * VEC;
  if (xb >= p.row_bytes) return;
  const int first = p.bounds[2 * yo], cnt = p.bounds[2 * yo + 1];
  const int* kk = p.kk + (long long)yo * p.ksize;
  const uint8_t* src = p.src + blockIdx.z * p.src_frame_stride + (long long)first * p.row_bytes + xb;
  int acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 1 << (RS_PRECISION_BITS - 1);
  for (int k = 0; k < cnt; ++k) {
    const int coef = __ldg(kk + k);
    if (VEC == 4) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(src + (long long)k * p.row_bytes);
      acc[0] += (int)(w & 255u) * coef;
      acc[1 % VEC] += (int)((w >> 8) & 255u) * coef;
      acc[2 % VEC] += (int)((w >> 16) & 255u) * coef;
      acc[3 % VEC] += (int)(w >> 24) * coef;
    } else {
      acc[0] += (int)src[(long long)k * p.row_bytes] * coef;
    }
  }
  uint8_t* dst = p.dst + blockIdx.z * p.dst_frame_stride + (long long)yo * p.row_bytes + xb;
  if (VEC == 4) {
    *reinterpret_cast<uint32_t*>(dst) = rs_clip8(acc[0]) | (rs_clip8(acc[1 % VEC]) << 8) |
                                        (rs_clip8(acc[2 % VEC]) << 16) | (rs_clip8(acc[3 % VEC]) << 24);
  } else {
    dst[0] = (uint8_t)rs_clip8(acc[0]);
  }
}

}  // namespace lars
