// K2: the fused Pass 2 of the RGNir analysis path, and K2f, its fixed-order statistics merge.
//
// One read of the raw uint8 frame produces, per pixel: the white-balanced bytes (stretch LUT,
// process-images.py:438-441), NDVI / GNDVI / NDWI as fp32 (process-images.py:449-490), three
// colormapped RGB pixels (process-images.py:689-695) and the per-index statistics + histogram
// (process-images.py:492-513, process-ndvi.py:65, :97).
//
// CTA = 8 consumer warps + 1 TMA-load warp + 1 TMA-store warp, persistent over a balanced
// contiguous range of pipeline tiles (1024 pixels for uint8 frames, 2048 for uint16):
//   load warp  : cp.async.bulk global->shared into a 3-stage ring (mbarrier full / empty)
//   consumers  : one (uint8) or two (uint16) 4-pixel groups per thread per tile; fp32 maps leave as coalesced 128-bit streaming
//                stores; byte outputs (WB, 3 x RGB) are staged in a 2-stage shared ring
//   store warp : cp.async.bulk shared->global of the staged bytes (mbarrier full / empty), so
//                no CTA-wide barrier sits on the consumers' critical path
// Statistics: per-thread registers, lane-private shared histograms (bank == lane, conflict
// free), one partial record per (CTA, frame) merged in fixed order by K2f -- deterministic.
// Per-pixel arithmetic avoids the quarter-rate XU pipe: magic-number int<->float conversions
// and a branch-free correctly-rounded division (pixel_math.h).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/lars_b200.h"
#include "pixel_math.h"
#include "ptx_sm100.cuh"

namespace lars {

#ifndef LARS_K2_STORE_EVICT_FIRST
#define LARS_K2_STORE_EVICT_FIRST 1   /* outputs carry an L2 evict-first policy (never re-read on the GPU) */
#endif
#ifndef LARS_K2_PREFETCH_TILES
#define LARS_K2_PREFETCH_TILES 0    /* >0: the load warp pulls this many tiles into L2 in one burst */
#endif
#ifndef LARS_K2_GROUPS
#define LARS_K2_GROUPS 1          /* 4-pixel groups per consumer thread per tile, uint8 frames: 1024-pixel tiles
                                     (measured: 0.876 of the HBM peak against 0.829 with 2048-pixel tiles) */
#endif
#ifndef LARS_K2_GROUPS_U16
#define LARS_K2_GROUPS_U16 2      /* uint16 frames keep 2048-pixel tiles (1024: 0.72 against 0.78) */
#endif
#ifndef LARS_K2_IN_STAGES
#define LARS_K2_IN_STAGES 3
#endif
#ifndef LARS_K2_U16_STAGES
#define LARS_K2_U16_STAGES 2      /* input stages of the uint16 variant (tiles are twice as large) */
#endif
#ifndef LARS_K2_OUT_STAGES
#define LARS_K2_OUT_STAGES 2
#endif
#ifndef LARS_K2_CTAS_PER_SM
#define LARS_K2_CTAS_PER_SM 2
#endif
constexpr int K2_CONSUMERS = 256;
constexpr int K2_CONSUMER_WARPS = K2_CONSUMERS / 32;
__host__ __device__ constexpr int k2_groups(int bps) { return bps == 2 ? LARS_K2_GROUPS_U16 : LARS_K2_GROUPS; }
constexpr int K2_GROUP_PX = 4 * K2_CONSUMERS;            // pixels one pass of the consumers covers
__host__ __device__ constexpr bool k2_derive_sum(int bps) { return bps == 2; }   // see the flush of fused_index_kernel
__host__ __device__ constexpr int k2_tile_px(int bps) { return K2_GROUP_PX * k2_groups(bps); }   // pixels per pipeline tile
constexpr int K2_THREADS = K2_CONSUMERS + 64;            // + TMA load warp + TMA store warp
constexpr int K2_IN_STAGES = LARS_K2_IN_STAGES;
constexpr int K2_OUT_STAGES = LARS_K2_OUT_STAGES;
constexpr int K2_CTAS_PER_SM = LARS_K2_CTAS_PER_SM;
constexpr int K2_BINS_PAD = LARS_MAX_BINS;
constexpr int K2_HIST_ROWS = LARS_MAX_BINS + 1;  // + the row that catches x == 1.0
constexpr int K2_CMAP_SLOTS = 257;               // + the slot that catches x == 1.0
constexpr int K2_MAX_GRID = 148 * 4;

// Per-CTA, per-frame partial record; merged in fixed order by fused_finalize_kernel.
struct __align__(16) K2Partial {
  double sx[2];       // sum x            (NDVI, GNDVI)
  double sd[2];       // sum (x - K)
  double sdd[2];      // sum (x - K)^2
  float k[2];         // the shift K (index value of the frame's first pixel)
  float mn[2], mx[2];
  uint32_t above[3];  // NDVI > t0, GNDVI > t1, NDWI > t2
  uint32_t count;
  uint32_t hist[3][K2_BINS_PAD];
  uint32_t pad_[2];
};
static_assert(sizeof(K2Partial) % 16 == 0, "partial record must stay 16-byte aligned");

struct K2Params {
  const uint8_t* src;
  const uint8_t* wb_lut;
  uint8_t* wb_out;
  float* maps[3];
  uint8_t* rgb[3];
  K2Partial* partials;      // [frame][slots_per_frame]
  const uint32_t* cmaps;    // [3 cmaps][256] packed R | G<<8 | B<<16 (device global)
  long long n_pixels;
  long long src_frame_stride, lut_frame_stride, wb_frame_stride, map_frame_stride, rgb_frame_stride;
  long long tiles_per_frame, total_tiles;
  int n_frames, bins, slots_per_frame;
  int cmap_id[3];
  float thresholds[3];
};

template <int C, int BPS>
struct K2Smem {
  static constexpr int GROUPS = k2_groups(BPS);
  static constexpr int TILE_PX = k2_tile_px(BPS);
  static constexpr int IN_BYTES = TILE_PX * C * BPS;
  static constexpr int WB_BYTES = TILE_PX * C;
  static constexpr int RGB_BYTES = TILE_PX * 3;
  static constexpr int OUT_BYTES = WB_BYTES + 3 * RGB_BYTES;     // one output stage
  // uint8: 3 x 256 B stretch tables, each 256-aligned (1 KB reserved to absorb any base alignment)
  // uint16: 3 x lars_stretch_u16 (guess parameters + 256 threshold pairs = 2064 B each)
  static constexpr int OFF_LUT = 0;
  static constexpr int OFF_CMAP = (BPS == 1) ? 1024 : 6208;      // 3 x 257 words
  static constexpr int OFF_HIST = OFF_CMAP + 3104;               // 3 x 65 rows x 32 lanes x 4 B
  static constexpr int OFF_IN = OFF_HIST + 3 * K2_HIST_ROWS * 128;
  // uint16 tiles are twice as large: two input stages keep two CTAs per SM resident
  static constexpr int IN_STAGES = (BPS == 2) ? LARS_K2_U16_STAGES : K2_IN_STAGES;
  static_assert(IN_STAGES <= K2_IN_STAGES, "the barrier block is sized for K2_IN_STAGES");
  static constexpr int OFF_OUT = OFF_IN + IN_STAGES * IN_BYTES;
  static constexpr int OFF_RED = OFF_OUT + K2_OUT_STAGES * OUT_BYTES;
  static constexpr int RED_BYTES = K2_CONSUMER_WARPS * 16 * 8;
  static constexpr int OFF_BAR = OFF_RED + RED_BYTES;
  static constexpr int N_BARS = 2 * K2_IN_STAGES + 2 * K2_OUT_STAGES;
  static constexpr int TOTAL = OFF_BAR + N_BARS * 8;
  static_assert(3 * K2_CMAP_SLOTS * 4 <= 3104 && OFF_HIST % 16 == 0 && OFF_IN % 16 == 0 && OFF_OUT % 16 == 0 &&
                    OFF_RED % 8 == 0 && OFF_BAR % 8 == 0,
                "shared layout alignment");
};

// CTA b owns the global tile range [b * T / G, (b + 1) * T / G): balanced to +-1 tile.
__device__ __forceinline__ long long k2_range_begin(long long b, long long total, long long grid) {
  return (b * total) / grid;
}
// the CTA whose range contains global tile t
__host__ __device__ inline long long k2_owner_of_tile(long long t, long long total, long long grid) {
  return ((t + 1) * grid - 1) / total;
}

struct K2ThreadStats {
  float mn[2], mx[2];
  double sx[2], sd[2], sdd[2];   // sum x (uint8 variant only), sum (x - K), sum (x - K)^2
  uint32_t above[3];
};

struct K2ThreadConst {
  uint32_t lut_addr[3];    // shared addresses of the three 256-byte stretch tables (256-aligned)
  uint32_t hist_cst[3];    // hist_base[i] + 4 lane - (MAGIC_U << 7)   (mod 2^32)
  uint32_t cmap_cst[3];    // cmap_base[i]          - (MAGIC_U << 2)   (mod 2^32)
  float half_bins, half_bins_bias_m05;
  float kshift[2];
  bool ndwi_by_sign;       // thresholds[2] == 0: count NDWI > 0 from the sign bit of GNDVI
  uint64_t store_policy;   // L2 evict-first policy for the output streams
};

__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void red_shared_inc(uint32_t addr) {
  asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}
// One PRMT builds the LUT address: low byte = byte `K` of the raw word, upper bytes from the
// 256-aligned table address.
template <int K>
__device__ __forceinline__ uint32_t lut_gather(uint32_t raw_word, uint32_t table_addr) {
  return lds_u8(prmt(raw_word, table_addr, 0x7650u | (uint32_t)K));
}
// three-input min / max (FMNMX3 on sm_100a)
__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ uint32_t pack4(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
  return prmt(prmt(b0, b1, 0x0040), prmt(b2, b3, 0x0040), 0x5410);
}

// uint16 sample -> white-balanced uint8, branch-free: a float guess of the stretch that is provably
// within +-1 of the reference's value (the difference v - p_lo is formed from an exact integer part,
// so its error is relative to itself; p_hi - p_lo >= 0.02 whenever it is not zero), then one 8-byte
// load of the bracketing thresholds (thr[g], thr[g + 1]) and a +-1 correction.  Exact because the
// thresholds come from the reference's fp64 -> fp32 -> uint8 chain evaluated on all 65,536 values.
struct U16Guess {
  int lo_int;
  float lo_frac, scale;
};
__device__ __forceinline__ U16Guess load_guess(const lars_stretch_u16* st) {
  const uint4 h = *reinterpret_cast<const uint4*>(st);   // lo_int, lo_frac, scale, reserved
  U16Guess g;
  g.lo_int = (int)h.x; g.lo_frac = __uint_as_float(h.y); g.scale = __uint_as_float(h.z);
  return g;
}
__device__ __forceinline__ uint32_t stretch_u16(uint32_t v, const U16Guess& st, uint32_t pairs_addr) {
  const float d = lars_small_int_to_float((int)v - st.lo_int);
  float t = fmaf(d - st.lo_frac, st.scale, -0.5f);
  t = fminf(fmaxf(t, -0.5f), 254.5f);
  const uint32_t gb = lars_f2u(t + LARS_MAGIC_F);              // MAGIC_U + g0, g0 in [0, 255]
  uint32_t lo, hi;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(gb * 8u + pairs_addr));
  return (gb - LARS_MAGIC_U) + (v >= hi ? 1u : 0u) - (v < lo ? 1u : 0u);
}

// raw words of one 4-pixel group of one thread
template <int C, int BPS>
struct K2Raw {
  uint32_t w[C * BPS];
};

template <int C, int BPS>
__device__ __forceinline__ void k2_load_group(const uint8_t* in_group, int tid, K2Raw<C, BPS>& r) {
  constexpr int NW = C * BPS;                      // 4 pixels * C samples * BPS bytes / 4
  if (NW == 3) {
    const uint32_t* in = reinterpret_cast<const uint32_t*>(in_group) + 3 * tid;
    r.w[0] = in[0]; r.w[1] = in[1]; r.w[2] = in[2];
  } else if (NW == 4) {
    const uint4 v = *(reinterpret_cast<const uint4*>(in_group) + tid);
    r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
  } else if (NW == 6) {
    const uint2* in = reinterpret_cast<const uint2*>(in_group) + 3 * tid;
    const uint2 a = in[0], b = in[1], c = in[2];
    r.w[0] = a.x; r.w[1] = a.y; r.w[2] = b.x; r.w[3] = b.y; r.w[4] = c.x; r.w[5] = c.y;
  } else {
    const uint4* in = reinterpret_cast<const uint4*>(in_group) + 2 * tid;
    const uint4 a = in[0], b = in[1];
    r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w; r.w[4] = b.x; r.w[5] = b.y; r.w[6] = b.z; r.w[7] = b.w;
  }
}

// fp32 partial sums of one tile of one thread (float64 accumulation happens once per tile)
struct K2TileSums {
  float gx[2], gd[2], gdd[2];
};

// One 4-pixel group of one consumer thread: stretch, indices, fp32 maps, staged bytes, statistics.
//   out_wb / out_rgb : this thread's slots in the staging buffers of the group
//   map_dst[i]       : this thread's 4 floats in index map i (nullptr = not requested)
template <int C, int BPS, bool FULL>
__device__ __forceinline__ void k2_process_group(const K2Params& p, const uint8_t* smem, const K2Raw<C, BPS>& raw,
                                                 uint32_t* out_wb, uint32_t* out_rgb, int rgb_stride_words,
                                                 float* const map_dst[3], int first, int nvalid,
                                                 const K2ThreadConst& tc, bool stage_bytes, K2ThreadStats& st,
                                                 K2TileSums& ts) {
  using L = K2Smem<C, BPS>;

  // ---- white balance: one stretch lookup per sample (process-images.py:438-441) ----
  uint32_t wb[3][4];
  if (BPS == 2) {
    const lars_stretch_u16* st16 = reinterpret_cast<const lars_stretch_u16*>(smem + L::OFF_LUT);
    const U16Guess g16[3] = {load_guess(st16), load_guess(st16 + 1), load_guess(st16 + 2)};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int sidx = j * C + c;                      // sample index within the 4-pixel group
        const uint32_t word = raw.w[sidx >> 1];
        wb[c][j] = stretch_u16((sidx & 1) ? (word >> 16) : (word & 0xFFFFu), g16[c], tc.lut_addr[c]);
      }
  } else if (C == 3) {
    const uint32_t w0 = raw.w[0], w1 = raw.w[1], w2 = raw.w[2];
    wb[0][0] = lut_gather<0>(w0, tc.lut_addr[0]); wb[1][0] = lut_gather<1>(w0, tc.lut_addr[1]);
    wb[2][0] = lut_gather<2>(w0, tc.lut_addr[2]); wb[0][1] = lut_gather<3>(w0, tc.lut_addr[0]);
    wb[1][1] = lut_gather<0>(w1, tc.lut_addr[1]); wb[2][1] = lut_gather<1>(w1, tc.lut_addr[2]);
    wb[0][2] = lut_gather<2>(w1, tc.lut_addr[0]); wb[1][2] = lut_gather<3>(w1, tc.lut_addr[1]);
    wb[2][2] = lut_gather<0>(w2, tc.lut_addr[2]); wb[0][3] = lut_gather<1>(w2, tc.lut_addr[0]);
    wb[1][3] = lut_gather<2>(w2, tc.lut_addr[1]); wb[2][3] = lut_gather<3>(w2, tc.lut_addr[2]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      wb[0][j] = lut_gather<0>(raw.w[j], tc.lut_addr[0]);
      wb[1][j] = lut_gather<1>(raw.w[j], tc.lut_addr[1]);
      wb[2][j] = lut_gather<2>(raw.w[j], tc.lut_addr[2]);
    }
  }
  if (stage_bytes && p.wb_out) {
    uint32_t* o = out_wb;
    if (C == 3) {
      o[0] = pack4(wb[0][0], wb[1][0], wb[2][0], wb[0][1]);
      o[1] = pack4(wb[1][1], wb[2][1], wb[0][2], wb[1][2]);
      o[2] = pack4(wb[2][2], wb[0][3], wb[1][3], wb[2][3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = pack4(wb[0][j], wb[1][j], wb[2][j], 0u);  // alpha = 0
    }
  }

  // ---- indices (process-images.py:449-490) ----
  float ndvi[4], gndvi[4], ndwi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    ndvi[j] = lars_ratio_pair_u8((int)wb[2][j], (int)wb[0][j]);
    gndvi[j] = lars_ratio_pair_u8((int)wb[2][j], (int)wb[1][j]);
    ndwi[j] = lars_negate_index(gndvi[j]);
  }

  // ---- fp32 maps: one coalesced 128-bit streaming store per index ----
  const bool any_valid = FULL || first < nvalid;
  const bool all_valid = FULL || first + 4 <= nvalid;
  {
    const float* vals[3] = {ndvi, gndvi, ndwi};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float* dst = map_dst[i];
      if (dst && any_valid) {
        if (all_valid) {
#if LARS_K2_STORE_EVICT_FIRST
          stg_v4_policy(dst, vals[i][0], vals[i][1], vals[i][2], vals[i][3], tc.store_policy);
#else
          stg_stream_v4(dst, vals[i][0], vals[i][1], vals[i][2], vals[i][3]);
#endif
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (first + j < nvalid) dst[j] = vals[i][j];
        }
      }
    }
  }

  // ---- colormap (process-images.py:689-695): slot -> packed RGB -> staged bytes ----
  if (stage_bytes) {
    const float* vals[3] = {ndvi, gndvi, ndwi};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (p.rgb[i]) {
        uint32_t c[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) c[j] = lds_u32(lars_cmap_slot_bits(vals[i][j]) * 4u + tc.cmap_cst[i]);
        uint32_t* o = out_rgb + i * rgb_stride_words;
        o[0] = prmt(c[0], c[1], 0x4210);  // R0 G0 B0 R1
        o[1] = prmt(c[1], c[2], 0x5421);  // G1 B1 R2 G2
        o[2] = prmt(c[2], c[3], 0x6542);  // B2 R3 G3 B3
      }
    }
  }

  // ---- statistics + histograms ----
  if (p.partials) {
    float* gx = ts.gx; float* gd = ts.gd; float* gdd = ts.gdd;
    if (FULL) {
      st.mn[0] = fmin3(fmin3(st.mn[0], ndvi[0], ndvi[1]), ndvi[2], ndvi[3]);
      st.mx[0] = fmax3(fmax3(st.mx[0], ndvi[0], ndvi[1]), ndvi[2], ndvi[3]);
      st.mn[1] = fmin3(fmin3(st.mn[1], gndvi[0], gndvi[1]), gndvi[2], gndvi[3]);
      st.mx[1] = fmax3(fmax3(st.mx[1], gndvi[0], gndvi[1]), gndvi[2], gndvi[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (FULL || first + j < nvalid) {
        const float x0 = ndvi[j], x1 = gndvi[j], x2 = ndwi[j];
        if (!FULL) {
          st.mn[0] = fminf(st.mn[0], x0); st.mx[0] = fmaxf(st.mx[0], x0);
          st.mn[1] = fminf(st.mn[1], x1); st.mx[1] = fmaxf(st.mx[1], x1);
        }
        st.above[0] += (x0 > p.thresholds[0]) ? 1u : 0u;   // float32 compare (0.2f)
        st.above[1] += (x1 > p.thresholds[1]) ? 1u : 0u;
        if (tc.ndwi_by_sign) st.above[2] += __float_as_uint(x1) >> 31;   // NDWI > 0  <=>  GNDVI < 0
        else st.above[2] += (x2 > p.thresholds[2]) ? 1u : 0u;
        const float d0 = LARS_FSUB(x0, tc.kshift[0]), d1 = LARS_FSUB(x1, tc.kshift[1]);
        if (!k2_derive_sum(BPS)) { gx[0] += x0; gx[1] += x1; }
        gd[0] += d0; gdd[0] = fmaf(d0, d0, gdd[0]);
        gd[1] += d1; gdd[1] = fmaf(d1, d1, gdd[1]);
        red_shared_inc(lars_hist_row_bits(x0, tc.half_bins, tc.half_bins_bias_m05) * 128u + tc.hist_cst[0]);
        red_shared_inc(lars_hist_row_bits(x1, tc.half_bins, tc.half_bins_bias_m05) * 128u + tc.hist_cst[1]);
        red_shared_inc(lars_hist_row_bits(x2, tc.half_bins, tc.half_bins_bias_m05) * 128u + tc.hist_cst[2]);
      }
    }
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}

template <int C, int BPS>
__global__ void __launch_bounds__(K2_THREADS, K2_CTAS_PER_SM) fused_index_kernel(const K2Params p) {
  using L = K2Smem<C, BPS>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + L::OFF_BAR;
  // barrier slots: in_full[0..4) in_empty[4..8) out_full[8..11) out_empty[11..14)
  const uint32_t bar_out_full = bar_base + 8u * (2 * K2_IN_STAGES);
  const uint32_t bar_out_empty = bar_out_full + 8u * K2_OUT_STAGES;

  const long long grid = gridDim.x;
  const long long t_begin = k2_range_begin(blockIdx.x, p.total_tiles, grid);
  const long long t_end = k2_range_begin(blockIdx.x + 1, p.total_tiles, grid);
  const bool stage_bytes = (p.wb_out != nullptr) || p.rgb[0] || p.rgb[1] || p.rgb[2];

  if (tid == 0) {
    for (int s = 0; s < K2_IN_STAGES; ++s) {
      mbar_init(bar_base + 8u * s, 1);                                   // load warp's expect_tx arrive
      mbar_init(bar_base + 8u * (K2_IN_STAGES + s), K2_CONSUMER_WARPS);  // one arrive per consumer warp
    }
    for (int s = 0; s < K2_OUT_STAGES; ++s) {
      mbar_init(bar_out_full + 8u * s, K2_CONSUMER_WARPS);
      mbar_init(bar_out_empty + 8u * s, 1);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= K2_CONSUMERS + 32) {
    // ===================== TMA store warp (one elected lane) =====================
    if (lane == 0 && stage_bytes && t_begin < t_end) {
      long long frame = t_begin / p.tiles_per_frame;          // one division, then incremental
      long long tile = t_begin - frame * p.tiles_per_frame;
      uint32_t so = 0, phase = 0;
      const uint64_t spol = l2_policy_evict_first();
      (void)spol;
      for (long long t = t_begin; t < t_end; ++t) {
        const long long px0 = tile * L::TILE_PX;
        const long long rem = p.n_pixels - px0;
        const uint32_t nvalid = rem < L::TILE_PX ? (uint32_t)rem : (uint32_t)L::TILE_PX;
        mbar_wait_relaxed(bar_out_full + 8u * so, phase);
        const uint32_t stage_addr = smem_base + L::OFF_OUT + so * L::OUT_BYTES;
        const uint32_t wb_bytes = (nvalid * C + 15u) & ~15u;
        const uint32_t rgb_bytes = (nvalid * 3u + 15u) & ~15u;
#if LARS_K2_STORE_EVICT_FIRST
        if (p.wb_out) tma_store_1d_hint(p.wb_out + frame * p.wb_frame_stride + px0 * C, stage_addr, wb_bytes, spol);
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (p.rgb[i])
            tma_store_1d_hint(p.rgb[i] + frame * p.rgb_frame_stride + px0 * 3,
                              stage_addr + L::WB_BYTES + i * L::RGB_BYTES, rgb_bytes, spol);
#else
        if (p.wb_out) tma_store_1d(p.wb_out + frame * p.wb_frame_stride + px0 * C, stage_addr, wb_bytes);
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (p.rgb[i])
            tma_store_1d(p.rgb[i] + frame * p.rgb_frame_stride + px0 * 3,
                         stage_addr + L::WB_BYTES + i * L::RGB_BYTES, rgb_bytes);
#endif
        tma_store_commit();
        // this warp has nothing else to do: wait until the bulk engine has read the staging buffer
        // (~1 us) and hand it straight back, so two buffers keep the consumers a full tile ahead
        tma_store_wait_read_all();
        mbar_arrive(bar_out_empty + 8u * so);
        if (++so == K2_OUT_STAGES) { so = 0; phase ^= 1u; }
        if (++tile == p.tiles_per_frame) { tile = 0; ++frame; }
      }
      tma_store_wait_all();  // global writes performed before the CTA (and its smem) goes away
    }
    return;
  }
  if (tid >= K2_CONSUMERS) {
    // ===================== TMA load warp (one elected lane) =====================
    if (lane == 0 && t_begin < t_end) {
      const uint64_t pol = l2_policy_evict_first();  // raw bytes are dead after this pass
#if LARS_K2_PREFETCH_TILES > 0
      const uint64_t pol_last = l2_policy_evict_last();
      int until_prefetch = 0;
#endif
      long long frame = t_begin / p.tiles_per_frame;
      long long tile = t_begin - frame * p.tiles_per_frame;
      uint32_t s = 0, phase = 0;
      for (long long t = t_begin; t < t_end; ++t) {
        const long long px0 = tile * L::TILE_PX;
        const long long rem = p.n_pixels - px0;
        const uint32_t nvalid = rem < L::TILE_PX ? (uint32_t)rem : (uint32_t)L::TILE_PX;
        const uint32_t bytes = (nvalid * (C * BPS) + 15u) & ~15u;
#if LARS_K2_PREFETCH_TILES > 0
        if (until_prefetch == 0) {
          // one large DRAM read burst for the next tiles of this frame: fewer read / write
          // turnarounds than a trickle of tile-sized reads between the output writes
          long long ntl = p.tiles_per_frame - tile;
          if (ntl > LARS_K2_PREFETCH_TILES) ntl = LARS_K2_PREFETCH_TILES;
          if (ntl > t_end - t) ntl = t_end - t;
          long long pbytes = ntl * (long long)(L::TILE_PX * C * BPS);
          const long long left = (p.n_pixels - px0) * (C * BPS);
          if (pbytes > left) pbytes = left;
          l2_prefetch_bulk(p.src + frame * p.src_frame_stride + px0 * (C * BPS), (uint32_t)((pbytes + 15) & ~15ll), pol_last);
          until_prefetch = (int)ntl;
        }
        --until_prefetch;
#endif
        mbar_wait_relaxed(bar_base + 8u * (K2_IN_STAGES + s), phase ^ 1u);
        mbar_arrive_expect_tx(bar_base + 8u * s, bytes);
        tma_load_1d_hint(smem_base + L::OFF_IN + s * L::IN_BYTES,
                         p.src + frame * p.src_frame_stride + px0 * (C * BPS), bytes, bar_base + 8u * s, pol);
        if (++s == L::IN_STAGES) { s = 0; phase ^= 1u; }
        if (++tile == p.tiles_per_frame) { tile = 0; ++frame; }
      }
    }
    return;
  }

  // ============================== consumer warps ==============================
  const int warp = tid >> 5;
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem + L::OFF_HIST);
  // the stretch tables must start on a 256-byte boundary of the shared window (lut_gather);
  // the layout reserves 1 KB for 768 B of tables so any base alignment can be absorbed
  const uint32_t lut_shift = (256u - ((smem_base + L::OFF_LUT) & 255u)) & 255u;
  uint8_t* lut_s = smem + L::OFF_LUT + lut_shift;
  uint32_t* cm_s = reinterpret_cast<uint32_t*>(smem + L::OFF_CMAP);
  double* red = reinterpret_cast<double*>(smem + L::OFF_RED);

  for (int i = tid; i < 3 * K2_CMAP_SLOTS; i += K2_CONSUMERS) {
    const int t = i / K2_CMAP_SLOTS, k = i - t * K2_CMAP_SLOTS;
    cm_s[i] = p.cmaps[p.cmap_id[t] * 256 + (k < 256 ? k : 255)];
  }

  K2ThreadConst tc;
  tc.half_bins = 0.5f * (float)p.bins;
  tc.half_bins_bias_m05 = tc.half_bins + LARS_HIST_BIAS - 0.5f;
  tc.ndwi_by_sign = (p.thresholds[2] == 0.0f);
  tc.store_policy = l2_policy_evict_first();
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    tc.lut_addr[i] = (BPS == 2)
                         ? smem_base + L::OFF_LUT + (uint32_t)(i * sizeof(lars_stretch_u16) + offsetof(lars_stretch_u16, pairs)) -
                               (LARS_MAGIC_U << 3)
                         : smem_base + L::OFF_LUT + lut_shift + 256u * i;
    tc.hist_cst[i] = smem_base + L::OFF_HIST + (uint32_t)(i * K2_HIST_ROWS * 128 + 4 * lane) - (LARS_MAGIC_U << 7);
    tc.cmap_cst[i] = smem_base + L::OFF_CMAP + (uint32_t)(i * K2_CMAP_SLOTS * 4) - (LARS_MAGIC_U << 2);
  }

  uint32_t s_in = 0, ph_in = 0, s_out = 0, ph_out = 0;
  long long t = t_begin;
  while (t < t_end) {
    // ---------------- one span: the part of frame `frame` owned by this CTA ----------------
    const long long frame = t / p.tiles_per_frame;
    const long long frame_t0 = frame * p.tiles_per_frame;
    const long long span_end = (frame_t0 + p.tiles_per_frame < t_end) ? frame_t0 + p.tiles_per_frame : t_end;

    named_bar_sync(1, K2_CONSUMERS);  // previous span's flush has finished reading hist / red / LUT users
    if (BPS == 2) {
      const uint32_t* gl = reinterpret_cast<const uint32_t*>(p.wb_lut + frame * p.lut_frame_stride);
      uint32_t* dst = reinterpret_cast<uint32_t*>(smem + L::OFF_LUT);
      for (int i = tid; i < 3 * (int)(sizeof(lars_stretch_u16) / 4); i += K2_CONSUMERS) dst[i] = gl[i];
    } else if (p.wb_lut) {
      const uint8_t* gl = p.wb_lut + frame * p.lut_frame_stride;
      for (int i = tid; i < 3 * 256; i += K2_CONSUMERS) lut_s[i] = gl[i];
    } else {
      for (int i = tid; i < 3 * 256; i += K2_CONSUMERS) lut_s[i] = (uint8_t)(i & 255);
    }
    for (int i = tid; i < 3 * K2_HIST_ROWS * 32; i += K2_CONSUMERS) hist[i] = 0u;
    named_bar_sync(1, K2_CONSUMERS);

    // shift K = index value of the frame's first pixel (same for every CTA of the frame)
    {
      int r, g, n;
      if (BPS == 2) {
        const uint16_t* f0 = reinterpret_cast<const uint16_t*>(p.src + frame * p.src_frame_stride);
        const lars_stretch_u16* st16 = reinterpret_cast<const lars_stretch_u16*>(smem + L::OFF_LUT);
        r = (int)stretch_u16(f0[0], load_guess(st16), tc.lut_addr[0]);
        g = (int)stretch_u16(f0[1], load_guess(st16 + 1), tc.lut_addr[1]);
        n = (int)stretch_u16(f0[2], load_guess(st16 + 2), tc.lut_addr[2]);
      } else {
        const uint8_t* f0 = p.src + frame * p.src_frame_stride;
        r = lut_s[f0[0]]; g = lut_s[256 + f0[1]]; n = lut_s[512 + f0[2]];
      }
      tc.kshift[0] = lars_ratio_pair_u8(n, r);
      tc.kshift[1] = lars_ratio_pair_u8(n, g);
    }
    K2ThreadStats st;
    st.mn[0] = st.mn[1] = INFINITY;
    st.mx[0] = st.mx[1] = -INFINITY;
    st.sx[0] = st.sx[1] = st.sd[0] = st.sd[1] = st.sdd[0] = st.sdd[1] = 0.0;
    st.above[0] = st.above[1] = st.above[2] = 0u;

    for (; t < span_end; ++t) {
      const long long px0 = (t - frame_t0) * L::TILE_PX;
      const long long rem = p.n_pixels - px0;
      const int nvalid = rem < L::TILE_PX ? (int)rem : L::TILE_PX;
      if (stage_bytes) mbar_wait(bar_out_empty + 8u * s_out, ph_out ^ 1u);  // staging buffer drained
      mbar_wait(bar_base + 8u * s_in, ph_in);                               // tile landed
      // all of this thread's raw words of the tile, then release the input stage
      K2Raw<C, BPS> raw[L::GROUPS];
      const uint8_t* in_tile = smem + L::OFF_IN + s_in * L::IN_BYTES;
#pragma unroll
      for (int g = 0; g < L::GROUPS; ++g) k2_load_group<C, BPS>(in_tile + g * (K2_GROUP_PX * C * BPS), tid, raw[g]);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_base + 8u * (K2_IN_STAGES + s_in));
      uint8_t* out_tile = smem + L::OFF_OUT + s_out * L::OUT_BYTES;
      K2TileSums ts;
      ts.gx[0] = ts.gx[1] = ts.gd[0] = ts.gd[1] = ts.gdd[0] = ts.gdd[1] = 0.f;
      float* map_base[3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
        map_base[i] = p.maps[i] ? p.maps[i] + frame * p.map_frame_stride + px0 + 4 * tid : nullptr;
#pragma unroll
      for (int g = 0; g < L::GROUPS; ++g) {
        uint32_t* out_wb = reinterpret_cast<uint32_t*>(out_tile + g * (K2_GROUP_PX * C)) + C * tid;
        uint32_t* out_rgb = reinterpret_cast<uint32_t*>(out_tile + L::WB_BYTES + g * (K2_GROUP_PX * 3)) + 3 * tid;
        float* map_dst[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) map_dst[i] = map_base[i] ? map_base[i] + g * K2_GROUP_PX : nullptr;
        const int first = g * K2_GROUP_PX + 4 * tid;
        if (nvalid == L::TILE_PX)
          k2_process_group<C, BPS, true>(p, smem, raw[g], out_wb, out_rgb, L::RGB_BYTES / 4, map_dst, first, nvalid,
                                         tc, stage_bytes, st, ts);
        else
          k2_process_group<C, BPS, false>(p, smem, raw[g], out_wb, out_rgb, L::RGB_BYTES / 4, map_dst, first, nvalid,
                                          tc, stage_bytes, st, ts);
      }
      if (p.partials) {  // float32 within the tile, float64 across tiles (error analysis in DESIGN.md)
        if (!k2_derive_sum(BPS)) { st.sx[0] += (double)ts.gx[0]; st.sx[1] += (double)ts.gx[1]; }
        st.sd[0] += (double)ts.gd[0]; st.sdd[0] += (double)ts.gdd[0];
        st.sd[1] += (double)ts.gd[1]; st.sdd[1] += (double)ts.gdd[1];
      }
      if (stage_bytes) {
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the bulk-copy engine
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_out_full + 8u * s_out);
        if (++s_out == K2_OUT_STAGES) { s_out = 0; ph_out ^= 1u; }
      }
      if (++s_in == L::IN_STAGES) { s_in = 0; ph_in ^= 1u; }
    }

    // ---------------- flush this span into its partial record ----------------
    if (p.partials) {
      double v[6] = {st.sd[0], st.sd[1], st.sdd[0], st.sdd[1], st.sx[0], st.sx[1]};
#pragma unroll
      for (int k = 0; k < (k2_derive_sum(BPS) ? 4 : 6); ++k) v[k] = warp_sum(v[k]);
      const float mn0 = warp_min(st.mn[0]), mn1 = warp_min(st.mn[1]);
      const float mx0 = warp_max(st.mx[0]), mx1 = warp_max(st.mx[1]);
      const uint32_t a0 = warp_sum(st.above[0]), a1 = warp_sum(st.above[1]), a2 = warp_sum(st.above[2]);
      if (lane == 0) {
        double* r = red + warp * 16;
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = v[k];
        r[4] = (double)mn0; r[5] = (double)mn1; r[6] = (double)mx0; r[7] = (double)mx1;
        r[8] = (double)a0; r[9] = (double)a1; r[10] = (double)a2;   // exact: < 2^32
        r[11] = v[4]; r[12] = v[5];
      }
      named_bar_sync(1, K2_CONSUMERS);
      const long long owner0 = k2_owner_of_tile(frame_t0, p.total_tiles, grid);
      K2Partial* rec = p.partials + frame * p.slots_per_frame + ((long long)blockIdx.x - owner0);
      if (tid < 3 * K2_BINS_PAD) {
        const int idx = tid / K2_BINS_PAD, bin = tid % K2_BINS_PAD;
        const uint32_t* row = hist + (idx * K2_HIST_ROWS + bin) * 32;
        uint32_t sum = 0;
#pragma unroll 8
        for (int l = 0; l < 32; ++l) sum += row[(l + tid) & 31];
        if (bin == p.bins - 1) {  // fold the x == 1.0 row into the last bin
#pragma unroll 8
          for (int l = 0; l < 32; ++l) sum += row[32 + ((l + tid) & 31)];
        }
        rec->hist[idx][bin] = (bin < p.bins) ? sum : 0u;
      }
      if (tid == 0) {
        double acc[13];
#pragma unroll
        for (int k = 0; k < 13; ++k) acc[k] = red[k];
        for (int w = 1; w < K2_CONSUMER_WARPS; ++w) {
          const double* r = red + w * 16;
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[k] += r[k];
          acc[4] = fmin(acc[4], r[4]); acc[5] = fmin(acc[5], r[5]);
          acc[6] = fmax(acc[6], r[6]); acc[7] = fmax(acc[7], r[7]);
          acc[8] += r[8]; acc[9] += r[9]; acc[10] += r[10];
          acc[11] += r[11]; acc[12] += r[12];
        }
        const long long first_px = (t_begin > frame_t0 ? t_begin - frame_t0 : 0) * L::TILE_PX;
        long long last_px = (span_end - frame_t0) * L::TILE_PX;
        if (last_px > p.n_pixels) last_px = p.n_pixels;
        const double cnt = (double)(last_px - first_px);
        // uint16 variant: sum x = sum (x - K) + count K, so its hot loop accumulates one sum less per index
        // (measured +3.4 % there; the same change costs the uint8 variant 2.5 %, which keeps its own sum)
        rec->sx[0] = k2_derive_sum(BPS) ? acc[0] + cnt * (double)tc.kshift[0] : acc[11];
        rec->sx[1] = k2_derive_sum(BPS) ? acc[1] + cnt * (double)tc.kshift[1] : acc[12];
        rec->sd[0] = acc[0]; rec->sd[1] = acc[1];
        rec->sdd[0] = acc[2]; rec->sdd[1] = acc[3];
        rec->k[0] = tc.kshift[0]; rec->k[1] = tc.kshift[1];
        rec->mn[0] = (float)acc[4]; rec->mn[1] = (float)acc[5];
        rec->mx[0] = (float)acc[6]; rec->mx[1] = (float)acc[7];
        rec->above[0] = (uint32_t)acc[8]; rec->above[1] = (uint32_t)acc[9]; rec->above[2] = (uint32_t)acc[10];
        rec->count = (uint32_t)(last_px - first_px);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// K2f: merge the partial records of each frame in slot order -> lars_index_stats[frame][3]
// ------------------------------------------------------------------------------------------
struct K2fParams {
  const K2Partial* partials;
  lars_index_stats* stats;
  int slots_per_frame, bins;
  float thresholds[3];
};

constexpr int K2F_SPLIT = 4;   // slot groups per histogram bin (a single frame per launch has ~300 slots to fold)

__global__ void __launch_bounds__(3 * K2_BINS_PAD * K2F_SPLIT) fused_finalize_kernel(const K2fParams p) {
  __shared__ unsigned long long hpart[K2F_SPLIT][3 * K2_BINS_PAD];
  const int frame = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const K2Partial* recs = p.partials + (long long)frame * p.slots_per_frame;
  lars_index_stats* out = p.stats + (long long)frame * 3;
  {  // histograms: thread (slot group, index, bin); integer sums, any order is exact
    const int grp = tid / (3 * K2_BINS_PAD), rem = tid % (3 * K2_BINS_PAD);
    const int idx = rem / K2_BINS_PAD, bin = rem % K2_BINS_PAD;
    unsigned long long h = 0;
#pragma unroll 16
    for (int s = grp; s < p.slots_per_frame; s += K2F_SPLIT) h += recs[s].hist[idx][bin];   // unwritten slots are zero (memset)
    hpart[grp][rem] = h;
  }
  __syncthreads();
  if (tid < 3 * K2_BINS_PAD) {
    unsigned long long h = 0;
#pragma unroll
    for (int g = 0; g < K2F_SPLIT; ++g) h += hpart[g][tid];
    out[tid / K2_BINS_PAD].hist[tid % K2_BINS_PAD] = h;
  }
  if (warp < 3) {
    // one warp per index: lane l folds slots l, l + 32, ... in order, then a fixed butterfly --
    // the summation tree depends only on the launch geometry, so results are reproducible
    const int i = warp;
    const int g = (i == 0) ? 0 : 1;  // NDWI statistics derive from GNDVI's (x -> 0 - x)
    double sx = 0.0, sd = 0.0, sdd = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    unsigned long long cnt = 0, above = 0;
#pragma unroll 4
    for (int s = lane; s < p.slots_per_frame; s += 32) {
      const K2Partial& r = recs[s];
      const bool used = r.count != 0;              // unwritten slots are all zero: sums take them as they are,
      sx += r.sx[g]; sd += r.sd[g]; sdd += r.sdd[g];
      mn = used ? fminf(mn, r.mn[g]) : mn;         // only min / max must skip them (select, no branch: loads pipeline)
      mx = used ? fmaxf(mx, r.mx[g]) : mx;
      cnt += r.count; above += r.above[i];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, d);
      sd += __shfl_xor_sync(0xffffffffu, sd, d);
      sdd += __shfl_xor_sync(0xffffffffu, sdd, d);
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
      cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
      above += __shfl_xor_sync(0xffffffffu, above, d);
    }
    if (lane == 0) {
      lars_index_stats& o = out[i];
      if (i == 2) {
        const float t = mn;
        mn = 0.0f - mx; mx = 0.0f - t;
        sx = 0.0 - sx; sd = 0.0 - sd;
      }
      const double n = (double)cnt;
      const double mean = cnt ? sx / n : 0.0;
      const double md = cnt ? sd / n : 0.0;
      double var = cnt ? sdd / n - md * md : 0.0;
      var = var > 0.0 ? var : 0.0;
      o.count = cnt;
      o.count_above = above;
      o.sum = sx;
      o.sumsq = cnt ? (var + mean * mean) * n : 0.0;
      o.mean = mean;
      o.std = sqrt(var);
      o.min = cnt ? mn : 0.f;
      o.max = cnt ? mx : 0.f;
      o.threshold = p.thresholds[i];
      o.bins = (uint32_t)p.bins;
    }
  }
}

}  // namespace lars
