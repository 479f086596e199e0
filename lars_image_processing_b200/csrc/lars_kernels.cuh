// sm_100a kernels of the RGNir analysis path (u8 frames).
//
//   K1  wb_hist_u8_kernel      Pass 1: per-frame, per-channel 256-bin value histogram
//   K1b wb_lut_build_kernel    cumulative histogram -> NumPy "linear" percentiles -> stretch LUT
//   K2  fused_index_u8_kernel  Pass 2: WB LUT + NDVI/GNDVI/NDWI + fp32 maps + colormap RGB +
//                              per-index statistics and histograms, one read of the raw frame
//   K2f fused_finalize_kernel  fixed-order merge of the per-CTA partial records
//
// Nothing here is a contraction, so no tensor cores: the path is HBM-bound byte/float
// streaming.  Design rules followed: 1-D TMA bulk copies staged through shared memory with
// mbarrier pipelines, 128-bit global accesses, lane-private (bank = lane) shared-memory
// histograms so that no shared atomic ever conflicts, persistent grids sized from the SM
// count, and deterministic fixed-order reductions (no floating-point atomics).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lars_b200.h"
#include "pixel_math.h"
#include "ptx_sm100.cuh"

namespace lars {

// ------------------------------------------------------------------------------------------
// K1: white-balance histogram (replaces the sort inside np.percentile, process-images.py:437)
// ------------------------------------------------------------------------------------------
// Shared layout hist[channel][bin][lane] (u32): a lane only ever touches bank == lane, so
// every shared atomic of a warp is conflict-free whatever the image content (natural images
// have long runs of equal values, the worst case for a plain 256-bin shared histogram).
constexpr int K1_THREADS = 512;
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int K1_SMEM_BYTES = 3 * 256 * 32 * 4;  // 96 KB -> 2 CTAs / SM
constexpr int K1_WARP_BYTES = 3 * 512;           // one warp iteration: 3 coalesced 512-byte rows
constexpr int K1_CTA_BYTES = K1_WARPS * K1_WARP_BYTES;

struct K1Params {
  const uint8_t* src;
  unsigned long long* hist;  // [frame][3][256]
  long long n_pixels;
  long long frame_stride;    // bytes
  long long units_per_frame; // work unit = K1_CTA_BYTES (C == 3) / 16 * K1_THREADS bytes (C == 4)
  long long total_units;
  int n_frames;
};

template <int SHIFT>
__device__ __forceinline__ void k1_count_byte(uint32_t word, uint32_t lane_base) {
  // address = lane_base + byte * 128  (bin stride = 32 lanes * 4 bytes)
  const uint32_t off = (SHIFT == 0) ? ((word << 7) & 0x7F80u) : ((word >> (SHIFT - 7)) & 0x7F80u);
  asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(lane_base + off) : "memory");
}

// One 16-byte vector whose byte k belongs to channel slot (k + ROT) % 3; c0/c1/c2 are the
// lane's shared addresses for slots 0 / 1 / 2.
template <int ROT>
__device__ __forceinline__ void k1_count_vec3(const uint4& v, uint32_t c0, uint32_t c1, uint32_t c2) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k0 = 4 * i + ROT;
    k1_count_byte<0>(w[i], (k0 % 3 == 0) ? c0 : ((k0 % 3 == 1) ? c1 : c2));
    k1_count_byte<8>(w[i], ((k0 + 1) % 3 == 0) ? c0 : (((k0 + 1) % 3 == 1) ? c1 : c2));
    k1_count_byte<16>(w[i], ((k0 + 2) % 3 == 0) ? c0 : (((k0 + 2) % 3 == 1) ? c1 : c2));
    k1_count_byte<24>(w[i], ((k0 + 3) % 3 == 0) ? c0 : (((k0 + 3) % 3 == 1) ? c1 : c2));
  }
}

__device__ __forceinline__ void k1_count_vec4(const uint4& v, uint32_t c0, uint32_t c1, uint32_t c2) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {  // one RGNA pixel per word, alpha ignored (process-images.py:435)
    k1_count_byte<0>(w[i], c0);
    k1_count_byte<8>(w[i], c1);
    k1_count_byte<16>(w[i], c2);
  }
}

template <int C>
struct K1Unit {
  static constexpr long long BYTES = (C == 3) ? (long long)K1_CTA_BYTES : 16ll * K1_THREADS;
};

// Persistent CTAs; CTA b owns the contiguous range [b T / G, (b + 1) T / G) of the global
// sequence of work units (frame-major), so the load is balanced to +-1 unit for any batch
// shape; a range that crosses a frame boundary flushes its shared histogram in between.
template <int C>
__global__ void __launch_bounds__(K1_THREADS, 2) wb_hist_u8_kernel(const K1Params p) {
  extern __shared__ __align__(16) uint32_t k1_hist[];  // [3][256][32]
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const long long frame_bytes = p.n_pixels * C;
  const uint32_t a0 = smem_u32(k1_hist) + 4u * lane, a1 = a0 + 32768u, a2 = a0 + 65536u;
  const long long G = gridDim.x;
  long long u = ((long long)blockIdx.x * p.total_units) / G;
  const long long u_end = ((long long)(blockIdx.x + 1) * p.total_units) / G;

  while (u < u_end) {
    const long long frame = u / p.units_per_frame;
    const long long fu0 = frame * p.units_per_frame;
    const long long span_end = (fu0 + p.units_per_frame < u_end) ? fu0 + p.units_per_frame : u_end;
    const long long b0 = (u - fu0) * K1Unit<C>::BYTES;
    long long b1 = (span_end - fu0) * K1Unit<C>::BYTES;
    if (b1 > frame_bytes) b1 = frame_bytes;
    const uint8_t* fsrc = p.src + frame * p.frame_stride;
    u = span_end;

    for (int i = tid; i < 3 * 256 * 32; i += K1_THREADS) k1_hist[i] = 0u;
    __syncthreads();

    if (C == 3) {
      // channel of the byte at frame offset o is o % 3; a warp iteration starts at a multiple
      // of 1536, row j at +512 j, lane at +16 lane  ->  slot (2 j + lane + k) % 3.
      const int ph = lane % 3;
      const uint32_t c0 = ph == 0 ? a0 : (ph == 1 ? a1 : a2);
      const uint32_t c1 = ph == 0 ? a1 : (ph == 1 ? a2 : a0);
      const uint32_t c2 = ph == 0 ? a2 : (ph == 1 ? a0 : a1);
      const long long vec_end = b0 + ((b1 - b0) / K1_WARP_BYTES) * K1_WARP_BYTES;
      long long off = b0 + (long long)warp * K1_WARP_BYTES;
      // two warp-iterations in flight per loop trip: 6 independent 128-bit loads per thread
      for (; off + (long long)K1_CTA_BYTES + K1_WARP_BYTES <= vec_end; off += 2ll * K1_CTA_BYTES) {
        const uint8_t* q0 = fsrc + off + 16 * lane;
        const uint8_t* q1 = q0 + K1_CTA_BYTES;
        const uint4 x0 = ldg_stream_v4(q0), x1 = ldg_stream_v4(q0 + 512), x2 = ldg_stream_v4(q0 + 1024);
        const uint4 y0 = ldg_stream_v4(q1), y1 = ldg_stream_v4(q1 + 512), y2 = ldg_stream_v4(q1 + 1024);
        k1_count_vec3<0>(x0, c0, c1, c2);
        k1_count_vec3<2>(x1, c0, c1, c2);
        k1_count_vec3<1>(x2, c0, c1, c2);
        k1_count_vec3<0>(y0, c0, c1, c2);
        k1_count_vec3<2>(y1, c0, c1, c2);
        k1_count_vec3<1>(y2, c0, c1, c2);
      }
      for (; off + K1_WARP_BYTES <= vec_end; off += K1_CTA_BYTES) {
        const uint8_t* q0 = fsrc + off + 16 * lane;
        const uint4 x0 = ldg_stream_v4(q0), x1 = ldg_stream_v4(q0 + 512), x2 = ldg_stream_v4(q0 + 1024);
        k1_count_vec3<0>(x0, c0, c1, c2);
        k1_count_vec3<2>(x1, c0, c1, c2);
        k1_count_vec3<1>(x2, c0, c1, c2);
      }
      // scalar tail (< 1536 bytes, only at the end of a frame)
      for (long long o = vec_end + tid; o < b1; o += K1_THREADS)
        atomicAdd(&k1_hist[((int)(o % 3) * 256 + (int)fsrc[o]) * 32 + lane], 1u);
    } else {
      const long long vec_end = b0 + ((b1 - b0) / 16) * 16;  // frame_bytes is a multiple of 4
      for (long long off = b0 + 16ll * tid; off + 16 <= vec_end; off += 16ll * K1_THREADS)
        k1_count_vec4(ldg_stream_v4(fsrc + off), a0, a1, a2);
      for (long long o = vec_end + tid; o < b1; o += K1_THREADS) {
        const int ch = (int)(o & 3);
        if (ch < 3) atomicAdd(&k1_hist[(ch * 256 + (int)fsrc[o]) * 32 + lane], 1u);
      }
    }
    __syncthreads();

    // fold the 32 lane copies; rotate the start so that concurrent threads hit distinct banks
    for (int b = tid; b < 3 * 256; b += K1_THREADS) {
      uint32_t sum = 0;
#pragma unroll 8
      for (int l = 0; l < 32; ++l) sum += k1_hist[b * 32 + ((l + tid) & 31)];
      if (sum) atomicAdd(&p.hist[frame * 768 + b], (unsigned long long)sum);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// K1b: percentiles + stretch LUT (process-images.py:437-441), one CTA per (set, channel)
// ------------------------------------------------------------------------------------------
struct K1bParams {
  const unsigned long long* hist;  // [set][3][256]
  uint8_t* lut;                    // [set][3][256]
  double* pct;                     // [set][3][2] or nullptr
  double q_lo, q_hi;
};

__global__ void __launch_bounds__(256) wb_lut_build_u8_kernel(const K1bParams p) {
  __shared__ unsigned long long cum[256];
  __shared__ unsigned long long warp_tot[8];
  __shared__ int order_stat[4];
  __shared__ double pcts[2];
  const int v = threadIdx.x;
  const int lane = v & 31, warp = v >> 5;
  const long long base = (long long)blockIdx.x * 256;  // blockIdx.x = set * 3 + channel

  // inclusive scan of the 256 bins
  unsigned long long x = p.hist[base + v];
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) warp_tot[warp] = x;
  __syncthreads();
  unsigned long long add = 0;
  for (int w = 0; w < warp; ++w) add += warp_tot[w];
  x += add;
  cum[v] = x;
  __syncthreads();
  const unsigned long long n = cum[255];
  const unsigned long long below = (v == 0) ? 0ull : cum[v - 1];

  // virtual index (n - 1) * q, neighbours floor / floor + 1, clamped like _get_indexes
  const double nm1 = (double)(n > 0 ? n - 1 : 0);
  unsigned long long ranks[4];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double vi = LARS_DMUL(nm1, k == 0 ? p.q_lo : p.q_hi);
    unsigned long long lo = (unsigned long long)floor(vi);
    unsigned long long hi = lo + 1;
    if (vi >= nm1) { lo = (n > 0 ? n - 1 : 0); hi = lo; }
    ranks[2 * k] = lo;
    ranks[2 * k + 1] = hi;
  }
  // value at rank r = the bin v with cum[v-1] <= r < cum[v]
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (below <= ranks[k] && ranks[k] < x) order_stat[k] = v;
  if (n == 0 && v < 4) order_stat[v] = 0;
  __syncthreads();
  if (v < 2) {
    const double vi = LARS_DMUL(nm1, v == 0 ? p.q_lo : p.q_hi);
    const double gamma = LARS_DSUB(vi, floor(vi));
    const double a = (double)order_stat[2 * v], b = (double)order_stat[2 * v + 1];
    const double r = (vi >= nm1) ? a : lars_percentile_lerp(a, b, gamma);
    pcts[v] = r;
    if (p.pct) p.pct[(long long)blockIdx.x * 2 + v] = r;
  }
  __syncthreads();
  p.lut[base + v] = lars_wb_lut_entry((double)v, pcts[0], pcts[1]);
}

// ------------------------------------------------------------------------------------------
// K2: fused Pass 2
// ------------------------------------------------------------------------------------------
constexpr int K2_TILE_PX = 1024;          // pixels per pipeline tile
constexpr int K2_CONSUMERS = 256;         // 4 pixels per consumer thread per tile
constexpr int K2_THREADS = K2_CONSUMERS + 32;  // + one TMA producer warp
constexpr int K2_IN_STAGES = 4;
constexpr int K2_OUT_STAGES = 2;
constexpr int K2_BINS_PAD = LARS_MAX_BINS;
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

template <int C>
struct K2Smem {
  static constexpr int IN_BYTES = K2_TILE_PX * C;
  static constexpr int WB_BYTES = K2_TILE_PX * C;
  static constexpr int RGB_BYTES = K2_TILE_PX * 3;
  static constexpr int OFF_IN = 0;
  static constexpr int OFF_WB = OFF_IN + K2_IN_STAGES * IN_BYTES;
  static constexpr int OFF_RGB = OFF_WB + K2_OUT_STAGES * WB_BYTES;
  static constexpr int OFF_HIST = OFF_RGB + K2_OUT_STAGES * 3 * RGB_BYTES;
  static constexpr int OFF_CMAP = OFF_HIST + 3 * K2_BINS_PAD * 32 * 4;
  static constexpr int OFF_LUT = OFF_CMAP + 3 * 256 * 4;
  static constexpr int OFF_RED = OFF_LUT + 3 * 256;
  static constexpr int RED_BYTES = 8 * 16 * 8;  // 8 warps x 16 slots x 8 bytes
  static constexpr int OFF_BAR = OFF_RED + RED_BYTES;
  static constexpr int TOTAL = OFF_BAR + 2 * K2_IN_STAGES * 8;
};

// CTA b owns the global tile range [b * T / G, (b + 1) * T / G): balanced to +-1 tile.
__device__ __forceinline__ long long k2_range_begin(long long b, long long total, long long grid) {
  return (b * total) / grid;
}
// first CTA whose range contains global tile t
__host__ __device__ inline long long k2_owner_of_tile(long long t, long long total, long long grid) {
  return ((t + 1) * grid - 1) / total;
}

struct K2ThreadStats {
  float mn[2], mx[2];
  double sx[2], sd[2], sdd[2];
  uint32_t above[3];
};

template <int C, bool FULL>
__device__ __forceinline__ void k2_process_tile(const K2Params& p, const uint8_t* smem, int in_stage,
                                                int out_stage, int tid, int lane, long long frame,
                                                long long px0, int nvalid, const float kshift[2],
                                                float half_bins, float half_bins_bias, int last_bin,
                                                bool stage_bytes, K2ThreadStats& st) {
  using L = K2Smem<C>;
  const uint8_t* lut = smem + L::OFF_LUT;
  const uint32_t* cm = reinterpret_cast<const uint32_t*>(smem + L::OFF_CMAP);
  uint32_t* hist = const_cast<uint32_t*>(reinterpret_cast<const uint32_t*>(smem + L::OFF_HIST));

  // ---- 4 raw pixels of this thread ----
  uint32_t raw[3][4];
  if (C == 3) {
    const uint32_t* in = reinterpret_cast<const uint32_t*>(smem + L::OFF_IN + in_stage * L::IN_BYTES) + 3 * tid;
    const uint32_t w0 = in[0], w1 = in[1], w2 = in[2];
    raw[0][0] = w0 & 0xFF;         raw[1][0] = (w0 >> 8) & 0xFF;  raw[2][0] = (w0 >> 16) & 0xFF;
    raw[0][1] = w0 >> 24;          raw[1][1] = w1 & 0xFF;         raw[2][1] = (w1 >> 8) & 0xFF;
    raw[0][2] = (w1 >> 16) & 0xFF; raw[1][2] = w1 >> 24;          raw[2][2] = w2 & 0xFF;
    raw[0][3] = (w2 >> 8) & 0xFF;  raw[1][3] = (w2 >> 16) & 0xFF; raw[2][3] = w2 >> 24;
  } else {
    const uint4 v = *(reinterpret_cast<const uint4*>(smem + L::OFF_IN + in_stage * L::IN_BYTES) + tid);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      raw[0][j] = w[j] & 0xFF; raw[1][j] = (w[j] >> 8) & 0xFF; raw[2][j] = (w[j] >> 16) & 0xFF;
    }
  }
  // release the input stage as early as possible (values are in registers now)
  __syncwarp();
  if (lane == 0) mbar_arrive(smem_u32(smem + L::OFF_BAR) + 8u * (K2_IN_STAGES + in_stage));

  // ---- white balance: one byte-LUT gather per sample (process-images.py:438-441) ----
  uint32_t wb[3][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    wb[0][j] = lut[raw[0][j]];
    wb[1][j] = lut[256 + raw[1][j]];
    wb[2][j] = lut[512 + raw[2][j]];
  }
  if (stage_bytes && p.wb_out) {
    uint32_t* o = const_cast<uint32_t*>(reinterpret_cast<const uint32_t*>(
                      smem + L::OFF_WB + out_stage * L::WB_BYTES)) + C * tid;
    if (C == 3) {
      o[0] = wb[0][0] | (wb[1][0] << 8) | (wb[2][0] << 16) | (wb[0][1] << 24);
      o[1] = wb[1][1] | (wb[2][1] << 8) | (wb[0][2] << 16) | (wb[1][2] << 24);
      o[2] = wb[2][2] | (wb[0][3] << 8) | (wb[1][3] << 16) | (wb[2][3] << 24);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = wb[0][j] | (wb[1][j] << 8) | (wb[2][j] << 16);  // alpha = 0
    }
  }

  // ---- indices (process-images.py:449-490) ----
  float ndvi[4], gndvi[4], ndwi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float r = (float)wb[0][j], g = (float)wb[1][j], n = (float)wb[2][j];
    ndvi[j] = lars_ratio_f32(n, r);
    gndvi[j] = lars_ratio_f32(n, g);
    ndwi[j] = lars_negate_index(gndvi[j]);
  }

  // ---- fp32 maps: one coalesced 128-bit streaming store per index ----
  const int first = 4 * tid;
  const bool any_valid = FULL || first < nvalid;
  const bool all_valid = FULL || first + 4 <= nvalid;
  {
    const float* vals[3] = {ndvi, gndvi, ndwi};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float* mp = p.maps[i];
      if (mp && any_valid) {
        float* dst = mp + frame * p.map_frame_stride + px0 + first;
        if (all_valid) {
          stg_stream_v4(dst, vals[i][0], vals[i][1], vals[i][2], vals[i][3]);
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
        for (int j = 0; j < 4; ++j) c[j] = cm[i * 256 + lars_cmap_index(vals[i][j])];
        uint32_t* o = const_cast<uint32_t*>(reinterpret_cast<const uint32_t*>(
                          smem + L::OFF_RGB + (out_stage * 3 + i) * L::RGB_BYTES)) + 3 * tid;
        o[0] = prmt(c[0], c[1], 0x4210);  // R0 G0 B0 R1
        o[1] = prmt(c[1], c[2], 0x5421);  // G1 B1 R2 G2
        o[2] = prmt(c[2], c[3], 0x6542);  // B2 R3 G3 B3
      }
    }
  }

  // ---- statistics + histograms ----
  if (p.partials) {
    float gx[2] = {0.f, 0.f}, gd[2] = {0.f, 0.f}, gdd[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (FULL || first + j < nvalid) {
        const float x0 = ndvi[j], x1 = gndvi[j], x2 = ndwi[j];
        st.mn[0] = fminf(st.mn[0], x0); st.mx[0] = fmaxf(st.mx[0], x0);
        st.mn[1] = fminf(st.mn[1], x1); st.mx[1] = fmaxf(st.mx[1], x1);
        st.above[0] += (x0 > p.thresholds[0]) ? 1u : 0u;   // float32 compare (0.2f)
        st.above[1] += (x1 > p.thresholds[1]) ? 1u : 0u;
        st.above[2] += (x2 > p.thresholds[2]) ? 1u : 0u;
        const float d0 = LARS_FSUB(x0, kshift[0]), d1 = LARS_FSUB(x1, kshift[1]);
        gx[0] += x0; gd[0] += d0; gdd[0] = fmaf(d0, d0, gdd[0]);
        gx[1] += x1; gd[1] += d1; gdd[1] = fmaf(d1, d1, gdd[1]);
        const int b0 = lars_hist_bin_pair(x0, half_bins, half_bins_bias, last_bin);
        const int b1 = lars_hist_bin_pair(x1, half_bins, half_bins_bias, last_bin);
        const int b2 = lars_hist_bin_pair(x2, half_bins, half_bins_bias, last_bin);
        atomicAdd(&hist[(0 * K2_BINS_PAD + b0) * 32 + lane], 1u);
        atomicAdd(&hist[(1 * K2_BINS_PAD + b1) * 32 + lane], 1u);
        atomicAdd(&hist[(2 * K2_BINS_PAD + b2) * 32 + lane], 1u);
      }
    }
    // float32 over 4 pixels, float64 across tiles (error analysis in DESIGN.md)
    st.sx[0] += (double)gx[0]; st.sd[0] += (double)gd[0]; st.sdd[0] += (double)gdd[0];
    st.sx[1] += (double)gx[1]; st.sd[1] += (double)gd[1]; st.sdd[1] += (double)gdd[1];
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

template <int C>
__global__ void __launch_bounds__(K2_THREADS, 2) fused_index_u8_kernel(const K2Params p) {
  using L = K2Smem<C>;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const uint32_t bar_base = smem_u32(smem + L::OFF_BAR);  // full[0..S), empty[S..2S)

  const long long grid = gridDim.x;
  const long long t_begin = k2_range_begin(blockIdx.x, p.total_tiles, grid);
  const long long t_end = k2_range_begin(blockIdx.x + 1, p.total_tiles, grid);

  if (tid == 0) {
    for (int s = 0; s < K2_IN_STAGES; ++s) {
      mbar_init(bar_base + 8u * s, 1);                              // producer's expect_tx arrive
      mbar_init(bar_base + 8u * (K2_IN_STAGES + s), K2_CONSUMERS / 32);  // one arrive per consumer warp
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= K2_CONSUMERS) {
    // ===================== TMA producer warp (one elected lane) =====================
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_first();  // raw bytes are dead after this pass
      uint32_t it = 0;
      for (long long t = t_begin; t < t_end; ++t, ++it) {
        const long long frame = t / p.tiles_per_frame;
        const long long tile = t - frame * p.tiles_per_frame;
        const long long px0 = tile * K2_TILE_PX;
        const long long rem = p.n_pixels - px0;
        const uint32_t npx = rem < K2_TILE_PX ? (uint32_t)rem : (uint32_t)K2_TILE_PX;
        const uint32_t bytes = (npx * C + 15u) & ~15u;
        const int s = it % K2_IN_STAGES;
        const uint32_t phase = (it / K2_IN_STAGES) & 1u;
        mbar_wait(bar_base + 8u * (K2_IN_STAGES + s), phase ^ 1u);
        mbar_arrive_expect_tx(bar_base + 8u * s, bytes);
        tma_load_1d_hint(smem_u32(smem + L::OFF_IN + s * L::IN_BYTES),
                         p.src + frame * p.src_frame_stride + px0 * C, bytes, bar_base + 8u * s, pol);
      }
    }
    return;
  }

  // ============================== consumer warps ==============================
  const int warp = tid >> 5;
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem + L::OFF_HIST);
  uint8_t* lut_s = smem + L::OFF_LUT;
  uint32_t* cm_s = reinterpret_cast<uint32_t*>(smem + L::OFF_CMAP);
  double* red = reinterpret_cast<double*>(smem + L::OFF_RED);

  for (int i = tid; i < 3 * 256; i += K2_CONSUMERS) cm_s[i] = p.cmaps[p.cmap_id[i >> 8] * 256 + (i & 255)];

  const bool stage_bytes = (p.wb_out != nullptr) || p.rgb[0] || p.rgb[1] || p.rgb[2];
  const float half_bins = 0.5f * (float)p.bins;
  const float half_bins_bias = half_bins + LARS_HIST_BIAS;
  const int last_bin = p.bins - 1;

  uint32_t it = 0;
  long long t = t_begin;
  while (t < t_end) {
    // ---------------- one span: the part of frame `frame` owned by this CTA ----------------
    const long long frame = t / p.tiles_per_frame;
    const long long frame_t0 = frame * p.tiles_per_frame;
    const long long span_end = (frame_t0 + p.tiles_per_frame < t_end) ? frame_t0 + p.tiles_per_frame : t_end;

    named_bar_sync(1, K2_CONSUMERS);  // previous span's flush has finished reading hist / red
    if (p.wb_lut) {
      const uint8_t* gl = p.wb_lut + frame * p.lut_frame_stride;
      for (int i = tid; i < 3 * 256; i += K2_CONSUMERS) lut_s[i] = gl[i];
    } else {
      for (int i = tid; i < 3 * 256; i += K2_CONSUMERS) lut_s[i] = (uint8_t)(i & 255);
    }
    for (int i = tid; i < 3 * K2_BINS_PAD * 32; i += K2_CONSUMERS) hist[i] = 0u;
    named_bar_sync(1, K2_CONSUMERS);

    // shift K = index value of the frame's first pixel (same for every CTA of the frame)
    float kshift[2];
    {
      const uint8_t* f0 = p.src + frame * p.src_frame_stride;
      const float r = (float)lut_s[f0[0]], g = (float)lut_s[256 + f0[1]], n = (float)lut_s[512 + f0[2]];
      kshift[0] = lars_ratio_f32(n, r);
      kshift[1] = lars_ratio_f32(n, g);
    }
    K2ThreadStats st;
    st.mn[0] = st.mn[1] = INFINITY;
    st.mx[0] = st.mx[1] = -INFINITY;
    st.sx[0] = st.sx[1] = st.sd[0] = st.sd[1] = st.sdd[0] = st.sdd[1] = 0.0;
    st.above[0] = st.above[1] = st.above[2] = 0u;

    for (; t < span_end; ++t, ++it) {
      const long long tile = t - frame_t0;
      const long long px0 = tile * K2_TILE_PX;
      const long long rem = p.n_pixels - px0;
      const int nvalid = rem < K2_TILE_PX ? (int)rem : K2_TILE_PX;
      const int s = it % K2_IN_STAGES;
      const uint32_t phase = (it / K2_IN_STAGES) & 1u;
      const int so = it & (K2_OUT_STAGES - 1);
      mbar_wait(bar_base + 8u * s, phase);
      if (nvalid == K2_TILE_PX)
        k2_process_tile<C, true>(p, smem, s, so, tid, lane, frame, px0, nvalid, kshift, half_bins,
                                 half_bins_bias, last_bin, stage_bytes, st);
      else
        k2_process_tile<C, false>(p, smem, s, so, tid, lane, frame, px0, nvalid, kshift, half_bins,
                                  half_bins_bias, last_bin, stage_bytes, st);
      if (stage_bytes) {
        // Byte outputs leave through the async proxy: fence own writes, make sure the other
        // staging buffer has been drained (its stores were issued one tile ago), then one
        // thread issues this tile's bulk stores.
        fence_proxy_async_smem();
        if (tid == 0) tma_store_wait_read_all();
        named_bar_sync(1, K2_CONSUMERS);
        if (tid == 0) {
          const uint32_t wb_bytes = ((uint32_t)nvalid * C + 15u) & ~15u;
          const uint32_t rgb_bytes = ((uint32_t)nvalid * 3u + 15u) & ~15u;
          if (p.wb_out)
            tma_store_1d(p.wb_out + frame * p.wb_frame_stride + px0 * C,
                         smem_u32(smem + L::OFF_WB + so * L::WB_BYTES), wb_bytes);
#pragma unroll
          for (int i = 0; i < 3; ++i)
            if (p.rgb[i])
              tma_store_1d(p.rgb[i] + frame * p.rgb_frame_stride + px0 * 3,
                           smem_u32(smem + L::OFF_RGB + (so * 3 + i) * L::RGB_BYTES), rgb_bytes);
          tma_store_commit();
        }
      }
    }

    // ---------------- flush this span into its partial record ----------------
    if (p.partials) {
      double v[6] = {st.sx[0], st.sx[1], st.sd[0], st.sd[1], st.sdd[0], st.sdd[1]};
#pragma unroll
      for (int k = 0; k < 6; ++k) v[k] = warp_sum(v[k]);
      float mn0 = warp_min(st.mn[0]), mn1 = warp_min(st.mn[1]);
      float mx0 = warp_max(st.mx[0]), mx1 = warp_max(st.mx[1]);
      uint32_t a0 = warp_sum(st.above[0]), a1 = warp_sum(st.above[1]), a2 = warp_sum(st.above[2]);
      if (lane == 0) {
        double* r = red + warp * 16;
#pragma unroll
        for (int k = 0; k < 6; ++k) r[k] = v[k];
        r[6] = (double)mn0; r[7] = (double)mn1; r[8] = (double)mx0; r[9] = (double)mx1;
        r[10] = (double)a0; r[11] = (double)a1; r[12] = (double)a2;   // exact: < 2^32
      }
      named_bar_sync(1, K2_CONSUMERS);
      const long long owner0 = k2_owner_of_tile(frame_t0, p.total_tiles, grid);
      K2Partial* rec = p.partials + frame * p.slots_per_frame + ((long long)blockIdx.x - owner0);
      if (tid < 3 * K2_BINS_PAD) {
        uint32_t sum = 0;
#pragma unroll 8
        for (int l = 0; l < 32; ++l) sum += hist[tid * 32 + ((l + tid) & 31)];
        rec->hist[tid / K2_BINS_PAD][tid % K2_BINS_PAD] = sum;
      }
      if (tid == 0) {
        double acc[13];
#pragma unroll
        for (int k = 0; k < 13; ++k) acc[k] = red[k];
        for (int w = 1; w < K2_CONSUMERS / 32; ++w) {
          const double* r = red + w * 16;
#pragma unroll
          for (int k = 0; k < 6; ++k) acc[k] += r[k];
          acc[6] = fmin(acc[6], r[6]); acc[7] = fmin(acc[7], r[7]);
          acc[8] = fmax(acc[8], r[8]); acc[9] = fmax(acc[9], r[9]);
          acc[10] += r[10]; acc[11] += r[11]; acc[12] += r[12];
        }
        rec->sx[0] = acc[0]; rec->sx[1] = acc[1];
        rec->sd[0] = acc[2]; rec->sd[1] = acc[3];
        rec->sdd[0] = acc[4]; rec->sdd[1] = acc[5];
        rec->k[0] = kshift[0]; rec->k[1] = kshift[1];
        rec->mn[0] = (float)acc[6]; rec->mn[1] = (float)acc[7];
        rec->mx[0] = (float)acc[8]; rec->mx[1] = (float)acc[9];
        rec->above[0] = (uint32_t)acc[10]; rec->above[1] = (uint32_t)acc[11]; rec->above[2] = (uint32_t)acc[12];
        // pixels of this span
        const long long first_px = (t_begin > frame_t0 ? t_begin - frame_t0 : 0) * K2_TILE_PX;
        long long last_px = (span_end - frame_t0) * K2_TILE_PX;
        if (last_px > p.n_pixels) last_px = p.n_pixels;
        rec->count = (uint32_t)(last_px - first_px);
      }
    }
  }
  // all bulk stores must have completed before the CTA (and its shared memory) goes away
  if (stage_bytes && tid == 0) tma_store_wait_all();
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

__global__ void __launch_bounds__(3 * K2_BINS_PAD) fused_finalize_kernel(const K2fParams p) {
  const int frame = blockIdx.x;
  const int tid = threadIdx.x;
  const K2Partial* recs = p.partials + (long long)frame * p.slots_per_frame;
  lars_index_stats* out = p.stats + (long long)frame * 3;
  {  // histograms: thread (index, bin)
    const int idx = tid / K2_BINS_PAD, bin = tid % K2_BINS_PAD;
    unsigned long long h = 0;
    for (int s = 0; s < p.slots_per_frame; ++s)
      if (recs[s].count) h += recs[s].hist[idx][bin];
    out[idx].hist[bin] = h;
  }
  if (tid < 3) {
    const int i = tid;
    const int g = (i == 0) ? 0 : 1;  // NDWI statistics derive from GNDVI's (x -> 0 - x)
    double sx = 0.0, sd = 0.0, sdd = 0.0;
    float mn = INFINITY, mx = -INFINITY, k = 0.f;
    unsigned long long cnt = 0, above = 0;
    for (int s = 0; s < p.slots_per_frame; ++s) {
      const K2Partial& r = recs[s];
      if (!r.count) continue;
      sx += r.sx[g]; sd += r.sd[g]; sdd += r.sdd[g];
      mn = fminf(mn, r.mn[g]); mx = fmaxf(mx, r.mx[g]);
      k = r.k[g];
      cnt += r.count; above += r.above[i];
    }
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
    (void)k;
  }
}

}  // namespace lars
