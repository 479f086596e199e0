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
//         [32 rows] x [input span of its output columns] tile in shared memory DE-INTERLEAVED and
//         TRANSPOSED (channel plane, word column major, 33-word pitch): one conflict-free 32-bit
//         load brings four taps of one channel, DP4A applies four coefficients at once.  Results go
//         through a second shared tile so the global stores are row-contiguous.
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

// dp4a with unsigned bytes in `a` and signed bytes in `b`
__device__ __forceinline__ int rs_dp4a_us(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

struct ResizeHParams {
  const uint8_t* src;        // [frame][in_h][in_w * C]
  uint8_t* dst;              // [frame][row_count][out_w * C]
  const int* bounds;         // [out_w][2]  (first source column, tap count)
  const uint32_t* kk;        // [out_w][groups][3]: byte planes of 4 coefficients (bits 0-7, 8-15, 16-23 signed)
  long long src_frame_stride, dst_frame_stride;
  int in_w, out_w, groups;
  int row_first, row_count;  // source rows to produce
  int xo_tile;               // output columns per CTA
  int plane_words;           // shared words per row per channel plane of the input tile (max over tiles, + slack)
  int out_pitch;             // bytes per row of the output tile (odd number of words)
};

// The 22-bit coefficients are split into three byte planes k = k0 + 2^8 k1 + 2^16 k2 (k0, k1 unsigned,
// k2 signed), so four taps of one channel cost three DP4A on one gathered word instead of four
// byte extractions + four IMAD; the three partial sums recombine exactly modulo 2^32 and the true
// value fits in int32 (Resample.c accumulates in int as well).
// FAST: C == 3 and every row of every frame starts on a 4-byte boundary (in_w * 3 and the frame
// stride are multiples of 4, the base pointer is 4-byte aligned).
template <int C, bool FAST>
__global__ void __launch_bounds__(RS_THREADS) resize_h_kernel(const ResizeHParams p) {
  extern __shared__ __align__(16) uint8_t rs_smem[];
  const int PW = p.plane_words;
  uint32_t* in_w32 = reinterpret_cast<uint32_t*>(rs_smem);                  // [C][PW][33] words (planar, transposed)
  uint8_t* in_s = rs_smem;
  uint8_t* out_s = rs_smem + (size_t)C * PW * RS_IN_PITCH * 4;              // [32][out_pitch] bytes
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int xo0 = blockIdx.x * p.xo_tile;
  const int nxo = min(p.xo_tile, p.out_w - xo0);
  const int r0 = blockIdx.y * RS_ROWS;                                      // relative to row_first
  const int nrows = min(RS_ROWS, p.row_count - r0);
  const long long in_pitch = (long long)p.in_w * C;
  const uint8_t* src = p.src + blockIdx.z * p.src_frame_stride + (long long)(p.row_first + r0) * in_pitch;
  uint8_t* dst = p.dst + blockIdx.z * p.dst_frame_stride + (long long)r0 * p.out_w * C;

  // tile origin: the first input pixel of the tile, rounded down to a 4-pixel group when the fast
  // staging path is used (12 bytes = 3 aligned words per group)
  const int px_first = p.bounds[2 * xo0];
  const int px0 = FAST ? (px_first & ~3) : px_first;
  const int px1 = p.bounds[2 * (xo0 + nxo - 1)] + p.bounds[2 * (xo0 + nxo - 1) + 1];

  // ---- stage: global rows (lanes along the row) -> de-interleaved, transposed shared tile ----
  if (FAST) {
    // C == 3, rows 4-byte aligned: a lane turns three aligned words (4 RGB pixels) into one word per
    // channel plane with six PRMT; groups past the end of the row read inside the frame slot's padding
    // or the next row and only feed taps whose coefficients are zero / never read
    const int ngroups = (px1 - px0 + 3) >> 2;
    const long long row_words_max = ((long long)p.in_w * 3 + 3) >> 2;       // words of one row
    for (int r = warp; r < nrows; r += RS_WARPS) {
      const uint32_t* row = reinterpret_cast<const uint32_t*>(src + (long long)r * in_pitch) + (px0 >> 2) * 3;
      const long long avail = row_words_max - (long long)(px0 >> 2) * 3;    // words left in this row
      for (int u = lane; u < ngroups; u += 32) {
        uint32_t w0 = 0, w1 = 0, w2 = 0;
        if (3 * u + 0 < avail) w0 = __ldg(row + 3 * u);
        if (3 * u + 1 < avail) w1 = __ldg(row + 3 * u + 1);
        if (3 * u + 2 < avail) w2 = __ldg(row + 3 * u + 2);
        // bytes: w0 = R0 G0 B0 R1, w1 = G1 B1 R2 G2, w2 = B2 R3 G3 B3
        const uint32_t pr = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);   // R0 R1 R2 R3
        const uint32_t pg = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);   // G0 G1 G2 G3
        const uint32_t pb = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);   // B0 B1 B2 B3
        in_w32[(0 * PW + u) * RS_IN_PITCH + r] = pr;
        in_w32[(1 * PW + u) * RS_IN_PITCH + r] = pg;
        in_w32[(2 * PW + u) * RS_IN_PITCH + r] = pb;
      }
    }
  } else {
    const int span = (px1 - px0) * C;
    for (int r = warp; r < nrows; r += RS_WARPS) {
      const uint8_t* row = src + (long long)r * in_pitch + (long long)px0 * C;
      for (int j = lane; j < span; j += 32) {
        const int x = j / C, c = j - x * C;
        in_s[((c * PW + (x >> 2)) * RS_IN_PITCH + r) * 4 + (x & 3)] = row[j];
      }
    }
  }
  __syncthreads();

  // ---- convolve: warp = output column, lane = row ----
  for (int xl = warp; xl < nxo; xl += RS_WARPS) {
    const int xo = xo0 + xl;
    const int xs = p.bounds[2 * xo] - px0, cnt = p.bounds[2 * xo + 1];
    const int ng = (cnt + 3) >> 2;
    const uint32_t* kk = p.kk + (long long)xo * p.groups * 3;
    const uint32_t sel = 0x3210u + 0x1111u * (uint32_t)(xs & 3);            // realign 4 bytes out of 2 words
    const uint32_t* base = in_w32 + (xs >> 2) * RS_IN_PITCH + lane;
    uint32_t a0[C], a1[C];
    int a2[C];
    uint32_t prev[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      a0[c] = 0u; a1[c] = 0u; a2[c] = 0;
      prev[c] = base[c * PW * RS_IN_PITCH];
    }
    for (int g = 0; g < ng; ++g) {
      const uint32_t k0 = __ldg(kk + 3 * g), k1 = __ldg(kk + 3 * g + 1), k2 = __ldg(kk + 3 * g + 2);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const uint32_t next = base[(c * PW + g + 1) * RS_IN_PITCH];
        const uint32_t v = __byte_perm(prev[c], next, sel);
        prev[c] = next;
        a0[c] = __dp4a(v, k0, a0[c]);
        a1[c] = __dp4a(v, k1, a1[c]);
        a2[c] = rs_dp4a_us(v, k2, a2[c]);
      }
    }
    uint8_t* o = out_s + lane * p.out_pitch + xl * C;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const uint32_t acc = (1u << (RS_PRECISION_BITS - 1)) + a0[c] + (a1[c] << 8) + ((uint32_t)a2[c] << 16);
      o[c] = (uint8_t)rs_clip8((int)acc);
    }
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

// ------------------------------------------------------------------------------------------
// K10h on the tensor cores.  The horizontal pass is a banded matrix product
//   out[row][xo] = sum_x in[row][x] * K[x][xo]
// -- the one contraction in this code base.  It is exact in integers: the 22-bit coefficients are split
// into three byte planes (k = k0 + 2^8 k1 + 2^16 k2, k0 / k1 unsigned, k2 signed), each plane is one
// int8 MMA (mma.sync.m16n8k32, u8 x u8 / u8 x s8 -> s32) and the three accumulators recombine modulo
// 2^32 to Pillow's 32-bit sum.  A = 16 rows x 32 input pixels of one channel plane (the de-interleaved
// shared tile, XOR-swizzled so that staging stores and fragment loads are both conflict-free),
// B = 32 input pixels x 8 output columns of coefficient bytes, prepared on the host in fragment order.
// One warp owns one block of 8 output columns: it loads the B fragments of a k-step once and applies
// them to 3 channels x 2 row blocks (6 MMA tiles x 3 planes).
// ------------------------------------------------------------------------------------------
struct ResizeHMmaParams {
  const uint8_t* src;
  uint8_t* dst;
  const int* bounds;         // [out_w][2] (tile extents)
  const int* kstart;         // [n_blocks] first input pixel (multiple of 4) of each 8-column block
  const uint2* bfrag;        // [n_blocks][ksteps][3 planes][32 lanes] B fragments
  long long src_frame_stride, dst_frame_stride;
  int in_w, out_w, ksteps;
  int row_first, row_count;
  int plane_words;           // shared words per row per channel plane
  int out_pitch;             // bytes per row of the output tile
};

constexpr int RS_MMA_COLS = 64;   // output columns per CTA = 8 warps x 8
constexpr int RS_MMA_OUT_PITCH = RS_MMA_COLS + 1;   // words per row of the output tile (one packed pixel per word)

__device__ __forceinline__ int rs_swz(int xw) { return ((xw & 3) << 3) | ((xw >> 2) & 7); }

__device__ __forceinline__ void rs_mma_u8u8(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void rs_mma_u8s8(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// KS > 0: the number of k-steps is the compile-time KS and the warp's B fragments (KS x 3 planes) are
// fetched into registers BEFORE the tile is staged, so their latency hides behind the staging loads and
// they serve all three channels; KS == 0: any k-step count, fragments re-read per channel.
#ifndef LARS_RS_MMA_CTAS
#define LARS_RS_MMA_CTAS 3
#endif
template <int KS>
__global__ void __launch_bounds__(RS_THREADS, LARS_RS_MMA_CTAS) resize_h_mma_kernel(const ResizeHMmaParams p) {
  extern __shared__ __align__(16) uint8_t rs_smem[];
  const int PW = p.plane_words;
  uint32_t* in_w32 = reinterpret_cast<uint32_t*>(rs_smem);                  // [3][PW][32] words, row index swizzled
  uint32_t* out_w32 = in_w32 + (size_t)3 * PW * 32;                         // [32][RS_MMA_OUT_PITCH] words: R | G << 8 | B << 16
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int xo0 = blockIdx.x * RS_MMA_COLS;
  const int nxo = min(RS_MMA_COLS, p.out_w - xo0);
  const int r0 = blockIdx.y * RS_ROWS;
  const int nrows = min(RS_ROWS, p.row_count - r0);
  const long long in_pitch = (long long)p.in_w * 3;
  const uint8_t* src = p.src + blockIdx.z * p.src_frame_stride + (long long)(p.row_first + r0) * in_pitch;
  uint8_t* dst = p.dst + blockIdx.z * p.dst_frame_stride + (long long)r0 * p.out_w * 3;
  const int nb0 = xo0 >> 3, nbt = (nxo + 7) >> 3;
  const int ksteps = KS > 0 ? KS : p.ksteps;
  const int px0 = p.kstart[nb0];                                             // tile origin, multiple of 4
  const int px1 = p.kstart[nb0 + nbt - 1] + ksteps * 32;                     // everything a fragment can touch
  const int my_nb = nb0 + (warp < nbt ? warp : 0);
  const int my_kstart = p.kstart[my_nb];
  uint2 breg[KS > 0 ? KS * 3 : 1];
  if (KS > 0) {
    const uint2* bf = p.bfrag + ((long long)my_nb * KS * 3) * 32 + lane;
#pragma unroll
    for (int i = 0; i < KS * 3; ++i) breg[i] = __ldg(bf + i * 32);
  }

  // ---- stage ----
  {
    const int ngroups = (px1 - px0) >> 2;
    const long long row_words = in_pitch >> 2;                               // rows are whole words (fast path)
    const long long first_word = (long long)(px0 >> 2) * 3;
    // a warp stages its four rows (warp, warp + 8, +16, +24) together: 12 independent loads per thread
    // in flight (one row at a time left ~12 KB in flight per SM and the pass latency-bound)
    const long long avail_full = row_words - first_word;
    for (int u = lane; u < ngroups; u += 32) {
      uint32_t w[4][3];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = warp + q * RS_WARPS;
        const uint32_t* row = reinterpret_cast<const uint32_t*>(src + (long long)r * in_pitch) + first_word;
        const long long avail = (r < nrows) ? avail_full : 0;                  // rows past the image: zeros
#pragma unroll
        for (int k = 0; k < 3; ++k) w[q][k] = (3 * u + k < avail) ? __ldg(row + 3 * u + k) : 0u;
      }
      const int sw = rs_swz(u);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int sr = (warp + q * RS_WARPS) ^ sw;
        in_w32[(0 * PW + u) * 32 + sr] = __byte_perm(__byte_perm(w[q][0], w[q][1], 0x0630), w[q][2], 0x5210);
        in_w32[(1 * PW + u) * 32 + sr] = __byte_perm(__byte_perm(w[q][0], w[q][1], 0x0741), w[q][2], 0x6210);
        in_w32[(2 * PW + u) * 32 + sr] = __byte_perm(__byte_perm(w[q][0], w[q][1], 0x0052), w[q][2], 0x7410);
      }
    }
  }
  __syncthreads();

  // ---- contract: warp = block of 8 output columns ----
  if (warp < nbt) {
    const int nb = my_nb;
    const int g = lane >> 2, t = lane & 3;
    const int xw_base = (my_kstart - px0) >> 2;
    const uint2* bf = p.bfrag + ((long long)nb * ksteps * 3) * 32 + lane;
    // one channel at a time keeps the accumulators at 24 registers (the B fragments are re-read from L1
    // per channel; all three channels at once needed 123 registers and halved the occupancy)
    uint32_t pix[2][4];                                                      // packed R | G << 8 | B << 16 of this thread's 8 outputs
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
      for (int i = 0; i < 4; ++i) pix[mb][i] = 0u;
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      int acc[2][3][4];                                                      // [row block][plane][fragment]
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int pl = 0; pl < 3; ++pl)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[mb][pl][i] = 0;
#pragma unroll
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint2 b_lo = KS > 0 ? breg[(KS > 0 ? ks : 0) * 3 + 0] : __ldg(bf + (ks * 3 + 0) * 32);
        const uint2 b_mid = KS > 0 ? breg[(KS > 0 ? ks : 0) * 3 + 1] : __ldg(bf + (ks * 3 + 1) * 32);
        const uint2 b_hi = KS > 0 ? breg[(KS > 0 ? ks : 0) * 3 + 2] : __ldg(bf + (ks * 3 + 2) * 32);
        const int xa = xw_base + ks * 8 + t, xb = xa + 4;
        const int sa = rs_swz(xa), sb = rs_swz(xb);
        const uint32_t* pa = in_w32 + (c * PW + xa) * 32;
        const uint32_t* pb = in_w32 + (c * PW + xb) * 32;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          const int row = mb * 16 + g;
          const uint32_t a[4] = {pa[row ^ sa], pa[(row + 8) ^ sa], pb[row ^ sb], pb[(row + 8) ^ sb]};
          rs_mma_u8u8(acc[mb][0], a, b_lo.x, b_lo.y);
          rs_mma_u8u8(acc[mb][1], a, b_mid.x, b_mid.y);
          rs_mma_u8s8(acc[mb][2], a, b_hi.x, b_hi.y);
        }
      }
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t v = (1u << (RS_PRECISION_BITS - 1)) + (uint32_t)acc[mb][0][i] +
                             ((uint32_t)acc[mb][1][i] << 8) + ((uint32_t)acc[mb][2][i] << 16);
          pix[mb][i] |= rs_clip8((int)v) << (8 * c);
        }
    }
    // thread holds rows g, g + 8 and columns 2 t, 2 t + 1 of every 16 x 8 tile: one word per pixel
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = mb * 16 + g + ((i >> 1) << 3);
        const int col = warp * 8 + 2 * t + (i & 1);
        out_w32[row * RS_MMA_OUT_PITCH + col] = pix[mb][i];
      }
  }
  __syncthreads();

  // ---- write: four packed pixels -> three output words, row-contiguous 32-bit stores ----
  const bool words_ok = (p.out_w & 3) == 0 && (p.dst_frame_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dst) & 3u) == 0;
  const int nquads = words_ok ? (nxo >> 2) : 0;                              // whole 4-pixel groups of this tile
  for (int r = warp * 2 + (lane >> 4); r < nrows; r += RS_WARPS * 2) {       // half a warp per row
    const int q = lane & 15;
    if (q < nquads) {
      const uint32_t* s4 = out_w32 + r * RS_MMA_OUT_PITCH + 4 * q;
      const uint32_t p0 = s4[0], p1 = s4[1], p2 = s4[2], p3 = s4[3];
      uint32_t* o = reinterpret_cast<uint32_t*>(dst + (long long)r * p.out_w * 3 + (long long)xo0 * 3) + 3 * q;
      o[0] = __byte_perm(p0, p1, 0x4210);   // R0 G0 B0 R1
      o[1] = __byte_perm(p1, p2, 0x5421);   // G1 B1 R2 G2
      o[2] = __byte_perm(p2, p3, 0x6542);   // B2 R3 G3 B3
    }
  }
  for (int r = warp; r < nrows; r += RS_WARPS) {                             // leftover pixels (or everything, unaligned)
    uint8_t* row = dst + (long long)r * p.out_w * 3 + (long long)xo0 * 3;
    for (int x = 4 * nquads + lane; x < nxo; x += 32) {
      const uint32_t px = out_w32[r * RS_MMA_OUT_PITCH + x];
      row[3 * x] = (uint8_t)px; row[3 * x + 1] = (uint8_t)(px >> 8); row[3 * x + 2] = (uint8_t)(px >> 16);
    }
  }
}

struct ResizeVParams {
  const uint8_t* src;        // [frame][rows][row_bytes]  (intermediate image, or the source when no horizontal pass)
  uint8_t* dst;              // [frame][out_h][row_bytes]
  const int* bounds;         // [out_h][2]  (first row relative to the intermediate image, tap count)
  const uint32_t* kk;        // [out_h][groups][3] byte planes of 4 coefficients, as in ResizeHParams
  long long src_frame_stride, dst_frame_stride;
  int row_bytes, out_h, groups, src_rows;
};

// VEC = 4: row_bytes and both frame strides are multiples of 4; VEC = 1: anything.
// Four taps at a time: four coalesced loads (rows k .. k+3 of the same 4 columns), a 4x4 byte
// transpose (8 PRMT) and three DP4A per column against the byte planes of the coefficients.
template <int VEC>
__global__ void __launch_bounds__(RS_THREADS) resize_v_kernel(const ResizeVParams p) {
  const int yo = blockIdx.y;
  const int xb = (blockIdx.x * RS_THREADS + threadIdx.x) * VEC;
  if (xb >= p.row_bytes) return;
  const int first = p.bounds[2 * yo], cnt = p.bounds[2 * yo + 1];
  const int ng = (cnt + 3) >> 2;
  const uint32_t* kk = p.kk + (long long)yo * p.groups * 3;
  const uint8_t* src = p.src + blockIdx.z * p.src_frame_stride + xb;
  const int last_row = p.src_rows - 1;          // rows past the window carry zero coefficients: clamp the address
  uint32_t a0[VEC], a1[VEC];
  int a2[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) { a0[i] = 0u; a1[i] = 0u; a2[i] = 0; }
  for (int g = 0; g < ng; ++g) {
    const uint32_t k0 = __ldg(kk + 3 * g), k1 = __ldg(kk + 3 * g + 1), k2 = __ldg(kk + 3 * g + 2);
    const int r = first + 4 * g;
    if (VEC == 4) {
      const uint32_t w0 = *reinterpret_cast<const uint32_t*>(src + (long long)min(r, last_row) * p.row_bytes);
      const uint32_t w1 = *reinterpret_cast<const uint32_t*>(src + (long long)min(r + 1, last_row) * p.row_bytes);
      const uint32_t w2 = *reinterpret_cast<const uint32_t*>(src + (long long)min(r + 2, last_row) * p.row_bytes);
      const uint32_t w3 = *reinterpret_cast<const uint32_t*>(src + (long long)min(r + 3, last_row) * p.row_bytes);
      const uint32_t t0 = __byte_perm(w0, w1, 0x5140), t1 = __byte_perm(w2, w3, 0x5140);
      const uint32_t t2 = __byte_perm(w0, w1, 0x7362), t3 = __byte_perm(w2, w3, 0x7362);
      const uint32_t col[4] = {__byte_perm(t0, t1, 0x5410), __byte_perm(t0, t1, 0x7632),
                               __byte_perm(t2, t3, 0x5410), __byte_perm(t2, t3, 0x7632)};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a0[i % VEC] = __dp4a(col[i], k0, a0[i % VEC]);
        a1[i % VEC] = __dp4a(col[i], k1, a1[i % VEC]);
        a2[i % VEC] = rs_dp4a_us(col[i], k2, a2[i % VEC]);
      }
    } else {
      const uint32_t b0 = src[(long long)min(r, last_row) * p.row_bytes];
      const uint32_t b1 = src[(long long)min(r + 1, last_row) * p.row_bytes];
      const uint32_t b2 = src[(long long)min(r + 2, last_row) * p.row_bytes];
      const uint32_t b3 = src[(long long)min(r + 3, last_row) * p.row_bytes];
      const uint32_t v = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
      a0[0] = __dp4a(v, k0, a0[0]);
      a1[0] = __dp4a(v, k1, a1[0]);
      a2[0] = rs_dp4a_us(v, k2, a2[0]);
    }
  }
  uint32_t out[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const uint32_t acc = (1u << (RS_PRECISION_BITS - 1)) + a0[i] + (a1[i] << 8) + ((uint32_t)a2[i] << 16);
    out[i] = rs_clip8((int)acc);
  }
  uint8_t* dst = p.dst + blockIdx.z * p.dst_frame_stride + (long long)yo * p.row_bytes + xb;
  if (VEC == 4) *reinterpret_cast<uint32_t*>(dst) = out[0] | (out[1 % VEC] << 8) | (out[2 % VEC] << 16) | (out[3 % VEC] << 24);
  else dst[0] = (uint8_t)out[0];
}

// ------------------------------------------------------------------------------------------
// RGBA frames: Pillow resizes them with premultiplied alpha (Image.resize: convert("RGBa") -> resize ->
// convert("RGBA")), so preprocess_large_image (process-images.py:398-422) on np.array(Image.open(upload)) of an RGBA PNG
// does too.  The two conversions, in place, one thread per pixel (libImaging/Convert.c: rgbA2rgba, rgba2rgbA):
//   premultiply    c' = MULDIV255(c, a) = ((t = c * a + 128) + (t >> 8)) >> 8
//   unpremultiply  a in {0, 255}: c unchanged; else c' = min(255, 255 * c / a)   (integer division)
// ------------------------------------------------------------------------------------------
template <bool PREMULTIPLY>
__global__ void __launch_bounds__(256) rgba_alpha_kernel(uint8_t* data, long long frame_stride, long long n_pixels) {
  uint8_t* frame = data + (long long)blockIdx.y * frame_stride;
  uint32_t* px = reinterpret_cast<uint32_t*>(frame);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t v = px[i];
    const uint32_t a = v >> 24;
    uint32_t c[3] = {v & 0xFFu, (v >> 8) & 0xFFu, (v >> 16) & 0xFFu};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (PREMULTIPLY) {
        const uint32_t t = c[k] * a + 128u;
        c[k] = ((t >> 8) + t) >> 8;
      } else if (a != 0u && a != 255u) {
        const uint32_t q = (255u * c[k]) / a;
        c[k] = q > 255u ? 255u : q;
      }
    }
    px[i] = c[0] | (c[1] << 8) | (c[2] << 16) | (a << 24);
  }
}

}  // namespace lars
