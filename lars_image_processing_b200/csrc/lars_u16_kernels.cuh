// Pass 1 for uint16 frames (BASELINE config 3: 16-bit RGNir TIFF frames).
//
// A full 65,536-bin histogram per channel (768 KB of counters) is neither shared-memory
// resident nor affordable as L2 atomics, and only four order statistics per channel are
// needed (ranks floor((n-1)q), +1 for q = 2 %, 98 %).  So the percentiles come from a
// two-level radix histogram:
//   level A  wb_hist_u16_hi_kernel   histogram of the HIGH byte (256 bins, lane-private shared
//                                    layout as for uint8)                        -- reads 6 B/px
//   select   wb_u16_select_kernel    per (frame, channel): the high-byte buckets that hold the
//                                    four ranks (<= 4 distinct, typically 2) + rank residuals
//   level B  wb_hist_u16_lo_kernel   histogram of the LOW byte of the samples whose high byte is
//                                    one of two selected buckets per channel     -- reads 6 B/px
//                                    (a second launch covers buckets 3-4 and exits at once when
//                                    no channel needs them)
//   build    wb_stretch_build_u16_kernel  exact values at the four ranks -> NumPy "linear"
//                                    percentiles (fp64) -> the stretch as 256 thresholds:
//                                    thr[k] = smallest v with LUT(v) >= k, LUT evaluated with the
//                                    reference's fp64 -> fp32 -> uint8 chain on all 65,536 values
// The stretch is monotone, so Pass 2 maps a sample with a float guess plus a +-1 correction
// against the thresholds (1 KB per channel in shared memory instead of a 64 KB table).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lars_b200.h"
#include "lars_kernels.cuh"
#include "pixel_math.h"
#include "ptx_sm100.cuh"

namespace lars {

constexpr int U16_MAX_BUCKETS = 4;  // distinct high-byte buckets holding the 4 ranks of a channel

// per (frame, channel) selection state
struct __align__(16) U16Select {
  unsigned long long residual[4];   // rank of order statistic r inside its bucket
  int32_t bucket_of_rank[4];        // index into buckets[] for order statistic r
  int32_t buckets[U16_MAX_BUCKETS]; // distinct high bytes, ascending; -1 = unused
  int32_t n_buckets;
  int32_t pad_[3];
};

struct U16HistParams {
  const uint8_t* src;               // uint16 samples, little endian
  unsigned long long* hist_hi;      // [set][3][256]
  unsigned long long* hist_lo;      // [set][3][U16_MAX_BUCKETS][256]
  const U16Select* select;          // [set][3]
  long long n_pixels;
  long long frame_stride;           // bytes
  long long units_per_frame, total_units;
  long long set_stride;             // 1 = one set per frame, 0 = shared set (tiles of one image)
  int n_frames;
  int lo_pass;                      // level B: handles buckets [2 lo_pass, 2 lo_pass + 1]
};

template <int C>
struct U16Unit {  // bytes per work unit: every lane owns 48 (C == 3) or 16 (C == 4) contiguous bytes
  static constexpr long long BYTES = (C == 3) ? (long long)K1_CTA_BYTES : 16ll * K1_THREADS;
};

// Visit every sample of a span.  C == 3: lane reads 48 contiguous bytes = 24 samples = 8 pixels,
// so sample s belongs to channel s % 3 statically.  C == 4: 16 bytes = 2 RGNA pixels.
template <int C, class Visit>
__device__ __forceinline__ void u16_visit_span(const uint8_t* fsrc, long long b0, long long b1, int tid, Visit visit) {
  if (C == 3) {
    const long long vec_end = b0 + ((b1 - b0) / 48) * 48;
    for (long long off = b0 + 48ll * tid; off + 48 <= vec_end; off += 48ll * K1_THREADS) {
      const uint4* q = reinterpret_cast<const uint4*>(fsrc + off);
      const uint4 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2);
      const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        visit((2 * i) % 3, w[i] & 0xFFFFu);
        visit((2 * i + 1) % 3, w[i] >> 16);
      }
    }
    const uint16_t* s16 = reinterpret_cast<const uint16_t*>(fsrc);
    for (long long s = vec_end / 2 + tid; s < b1 / 2; s += K1_THREADS) visit((int)(s % 3), (uint32_t)s16[s]);
  } else {
    const long long vec_end = b0 + ((b1 - b0) / 16) * 16;  // frame bytes are a multiple of 8
    for (long long off = b0 + 16ll * tid; off + 16 <= vec_end; off += 16ll * K1_THREADS) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(fsrc + off));
      visit(0, v.x & 0xFFFFu); visit(1, v.x >> 16); visit(2, v.y & 0xFFFFu);
      visit(0, v.z & 0xFFFFu); visit(1, v.z >> 16); visit(2, v.w & 0xFFFFu);
    }
    const uint16_t* s16 = reinterpret_cast<const uint16_t*>(fsrc);
    for (long long s = vec_end / 2 + tid; s < b1 / 2; s += K1_THREADS)
      if ((s & 3) < 3) visit((int)(s & 3), (uint32_t)s16[s]);
  }
}

// ---- level A ------------------------------------------------------------------------------
constexpr int U16_HI_SMEM_BYTES = 3 * 256 * 32 * 4;   // [3][256][32] lane-private, 96 KB
template <int C>
__global__ void __launch_bounds__(K1_THREADS, 2) wb_hist_u16_hi_kernel(const U16HistParams p) {
  extern __shared__ __align__(16) uint32_t u16_hist[];  // [3][256][32] lane-private
  const int tid = threadIdx.x, lane = tid & 31;
  const long long frame_bytes = p.n_pixels * C * 2;
  const uint32_t base = smem_u32(u16_hist) + 4u * lane;
  const long long G = gridDim.x;
  long long u = ((long long)blockIdx.x * p.total_units) / G;
  const long long u_end = ((long long)(blockIdx.x + 1) * p.total_units) / G;
  while (u < u_end) {
    const long long frame = u / p.units_per_frame;
    const long long fu0 = frame * p.units_per_frame;
    const long long span_end = (fu0 + p.units_per_frame < u_end) ? fu0 + p.units_per_frame : u_end;
    const long long b0 = (u - fu0) * U16Unit<C>::BYTES;
    long long b1 = (span_end - fu0) * U16Unit<C>::BYTES;
    if (b1 > frame_bytes) b1 = frame_bytes;
    const uint8_t* fsrc = p.src + frame * p.frame_stride;
    u = span_end;
    for (int i = tid; i < 3 * 256 * 32; i += K1_THREADS) u16_hist[i] = 0u;
    __syncthreads();
    u16_visit_span<C>(fsrc, b0, b1, tid, [&](int ch, uint32_t v) {
      asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + (uint32_t)ch * 32768u + ((v >> 8) << 7)) : "memory");
    });
    __syncthreads();
    for (int b = tid; b < 3 * 256; b += K1_THREADS) {
      uint32_t sum = 0;
#pragma unroll 8
      for (int l = 0; l < 32; ++l) sum += u16_hist[b * 32 + ((l + tid) & 31)];
      if (sum) atomicAdd(&p.hist_hi[frame * p.set_stride * 768 + b], (unsigned long long)sum);
    }
    __syncthreads();
  }
}

// ---- select -------------------------------------------------------------------------------
struct U16SelectParams {
  const unsigned long long* hist_hi;  // [set][3][256]
  U16Select* select;                  // [set][3]
  double q_lo, q_hi;
};

__device__ __forceinline__ void percentile_ranks(unsigned long long n, double q, unsigned long long& lo,
                                                 unsigned long long& hi, double& gamma, bool& at_end) {
  const double nm1 = (double)(n > 0 ? n - 1 : 0);
  const double vi = LARS_DMUL(nm1, q);
  lo = (unsigned long long)floor(vi);
  hi = lo + 1;
  at_end = vi >= nm1;
  if (at_end) { lo = (n > 0 ? n - 1 : 0); hi = lo; }
  gamma = LARS_DSUB(vi, floor(vi));
}

__global__ void __launch_bounds__(256) wb_u16_select_kernel(const U16SelectParams p) {
  __shared__ unsigned long long cum[256];
  __shared__ unsigned long long warp_tot[8];
  __shared__ int rank_bucket[4];
  __shared__ unsigned long long rank_resid[4];
  const int v = threadIdx.x, lane = v & 31, warp = v >> 5;
  const long long base = (long long)blockIdx.x * 256;
  unsigned long long x = p.hist_hi[base + v];
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
  if (v < 4) { rank_bucket[v] = 0; rank_resid[v] = 0; }
  __syncthreads();
  const unsigned long long n = cum[255];
  const unsigned long long below = v ? cum[v - 1] : 0ull;
  unsigned long long ranks[4];
  double g; bool e;
  percentile_ranks(n, p.q_lo, ranks[0], ranks[1], g, e);
  percentile_ranks(n, p.q_hi, ranks[2], ranks[3], g, e);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (below <= ranks[k] && ranks[k] < x) { rank_bucket[k] = v; rank_resid[k] = ranks[k] - below; }
  __syncthreads();
  if (v == 0) {
    U16Select s;
    s.n_buckets = 0;
    for (int b = 0; b < U16_MAX_BUCKETS; ++b) s.buckets[b] = -1;
    for (int k = 0; k < 4; ++k) {       // ranks ascend, so buckets ascend
      int slot = -1;
      for (int b = 0; b < s.n_buckets; ++b) if (s.buckets[b] == rank_bucket[k]) slot = b;
      if (slot < 0) { slot = s.n_buckets; s.buckets[s.n_buckets++] = rank_bucket[k]; }
      s.bucket_of_rank[k] = slot;
      s.residual[k] = rank_resid[k];
    }
    s.pad_[0] = s.pad_[1] = s.pad_[2] = 0;
    p.select[blockIdx.x] = s;
  }
}

// ---- level B ------------------------------------------------------------------------------
// shared layout lo[channel][slot(2)][low byte][16]: lanes l and l + 16 share a counter, so any
// shared atomic is at most 2-way serialised even for a constant image.
constexpr int U16_LO_SMEM_BYTES = 3 * 2 * 256 * 16 * 4;  // 96 KB

template <int C>
__global__ void __launch_bounds__(K1_THREADS, 2) wb_hist_u16_lo_kernel(const U16HistParams p) {
  extern __shared__ __align__(16) uint32_t u16_lo[];  // [3][2][256][16]
  const int tid = threadIdx.x, lane = tid & 31;
  const long long frame_bytes = p.n_pixels * C * 2;
  const uint32_t base = smem_u32(u16_lo) + 4u * (lane & 15);
  const long long G = gridDim.x;
  long long u = ((long long)blockIdx.x * p.total_units) / G;
  const long long u_end = ((long long)(blockIdx.x + 1) * p.total_units) / G;
  while (u < u_end) {
    const long long frame = u / p.units_per_frame;
    const long long fu0 = frame * p.units_per_frame;
    const long long span_end = (fu0 + p.units_per_frame < u_end) ? fu0 + p.units_per_frame : u_end;
    const long long b0 = (u - fu0) * U16Unit<C>::BYTES;
    long long b1 = (span_end - fu0) * U16Unit<C>::BYTES;
    if (b1 > frame_bytes) b1 = frame_bytes;
    const uint8_t* fsrc = p.src + frame * p.frame_stride;
    u = span_end;
    const U16Select* sel = p.select + frame * p.set_stride * 3;
    // six scalars, not an array: the scalar tail indexes by a run-time channel, and a dynamically
    // indexed local array would put the selection in local memory for the hot loop as well
    const int w00 = sel[0].buckets[2 * p.lo_pass], w01 = sel[0].buckets[2 * p.lo_pass + 1];
    const int w10 = sel[1].buckets[2 * p.lo_pass], w11 = sel[1].buckets[2 * p.lo_pass + 1];
    const int w20 = sel[2].buckets[2 * p.lo_pass], w21 = sel[2].buckets[2 * p.lo_pass + 1];
    const bool any = (w00 >= 0) || (w10 >= 0) || (w20 >= 0);
    if (!any) continue;  // uniform per CTA: this pass has nothing to count for this frame
    for (int i = tid; i < 3 * 2 * 256 * 16; i += K1_THREADS) u16_lo[i] = 0u;
    __syncthreads();
    // (a branch-free form -- two predicated REDs per sample -- was measured 17 % slower: ~1 % of the
    // samples match, and the extra address arithmetic for the other 99 % costs more than the branches)
    u16_visit_span<C>(fsrc, b0, b1, tid, [&](int ch, uint32_t v) {
      const int hi = (int)(v >> 8);
      const uint32_t lo = v & 0xFFu;
      const int want0 = ch == 0 ? w00 : (ch == 1 ? w10 : w20);
      const int want1 = ch == 0 ? w01 : (ch == 1 ? w11 : w21);
      if (hi == want0)
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + (uint32_t)((ch * 2 + 0) * 256 + lo) * 64u) : "memory");
      else if (hi == want1)
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + (uint32_t)((ch * 2 + 1) * 256 + lo) * 64u) : "memory");
    });
    __syncthreads();
    for (int b = tid; b < 3 * 2 * 256; b += K1_THREADS) {
      uint32_t sum = 0;
#pragma unroll
      for (int l = 0; l < 16; ++l) sum += u16_lo[b * 16 + ((l + tid) & 15)];
      if (sum) {
        const int ch = b / 512, slot = (b >> 8) & 1, lo = b & 255;
        atomicAdd(&p.hist_lo[((frame * p.set_stride * 3 + ch) * U16_MAX_BUCKETS + 2 * p.lo_pass + slot) * 256 + lo],
                  (unsigned long long)sum);
      }
    }
    __syncthreads();
  }
}

// ---- build: percentiles + thresholds ------------------------------------------------------
struct U16BuildParams {
  const unsigned long long* hist_hi;  // [set][3][256]   (for n)
  const unsigned long long* hist_lo;  // [set][3][4][256]
  const U16Select* select;            // [set][3]
  lars_stretch_u16* stretch;          // [set][3]
  double* pct;                        // [set][3][2] or nullptr
  double q_lo, q_hi;
};

__global__ void __launch_bounds__(256) wb_stretch_build_u16_kernel(const U16BuildParams p) {
  __shared__ unsigned long long cum[256];
  __shared__ unsigned long long warp_tot[8];
  __shared__ int value_at[4];
  __shared__ double pcts[2];
  __shared__ unsigned long long n_total;
  const int v = threadIdx.x, lane = v & 31, warp = v >> 5;
  const long long sc = blockIdx.x;  // set * 3 + channel
  const U16Select sel = p.select[sc];
  {
    unsigned long long x = p.hist_hi[sc * 256 + v];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    if (lane == 0) warp_tot[warp] = x;
    __syncthreads();
    if (v == 0) { unsigned long long t = 0; for (int w = 0; w < 8; ++w) t += warp_tot[w]; n_total = t; }
    __syncthreads();
  }
  // exact 16-bit value at each of the four ranks: scan the low-byte histogram of its bucket
  for (int b = 0; b < sel.n_buckets; ++b) {
    unsigned long long x = p.hist_lo[(sc * U16_MAX_BUCKETS + b) * 256 + v];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    __syncthreads();
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    unsigned long long add = 0;
    for (int w = 0; w < warp; ++w) add += warp_tot[w];
    x += add;
    cum[v] = x;
    __syncthreads();
    const unsigned long long below = v ? cum[v - 1] : 0ull;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (sel.bucket_of_rank[k] == b && below <= sel.residual[k] && sel.residual[k] < x)
        value_at[k] = sel.buckets[b] * 256 + v;
    __syncthreads();
  }
  if (n_total == 0 && v < 4) value_at[v] = 0;
  __syncthreads();
  if (v < 2) {
    unsigned long long lo, hi; double gamma; bool at_end;
    percentile_ranks(n_total, v == 0 ? p.q_lo : p.q_hi, lo, hi, gamma, at_end);
    const double a = (double)value_at[2 * v], b = (double)value_at[2 * v + 1];
    const double r = at_end ? a : lars_percentile_lerp(a, b, gamma);
    pcts[v] = r;
    if (p.pct) p.pct[sc * 2 + v] = r;
  }
  __syncthreads();
  const double plo = pcts[0], phi = pcts[1];
  lars_stretch_u16& st = p.stretch[sc];
  // thr[k] = smallest v with LUT(v) >= k; the LUT is monotone, so thresholds sit where it steps
  __shared__ uint32_t thr[258];
  for (int k = v; k < 258; k += 256) thr[k] = (k == 0) ? 0u : 65536u;
  __syncthreads();
  // the LUT is monotone non-decreasing in val, so thread k finds thr[k] by bisection over [0, 65536]
  // (17 evaluations of the reference chain instead of a sweep over all 65,536 values)
  if (v >= 1) {
    int lo_v = 0, hi_v = 65536;                     // invariant: LUT(x) < k for x < lo_v, LUT(x) >= k for x >= hi_v
    while (lo_v < hi_v) {
      const int mid = (lo_v + hi_v) >> 1;
      if ((int)lars_wb_lut_entry((double)mid, plo, phi) >= v) hi_v = mid;
      else lo_v = mid + 1;
    }
    thr[v] = (uint32_t)lo_v;
  }
  __syncthreads();
  st.pairs[v][0] = thr[v];
  st.pairs[v][1] = thr[v + 1];
  if (v == 0) {
    const double span = LARS_DSUB(phi, plo);
    if (span > 0.0) {
      const double fl = floor(plo);
      st.lo_int = (int32_t)fl;
      st.lo_frac = (float)LARS_DSUB(plo, fl);
      st.scale = (float)(255.0 / span);
    } else {  // p_hi == p_lo: one step from 0 to 255 at thr[1]; a steep ramp centred just below it
      st.lo_int = (int32_t)thr[1] - 1;
      st.lo_frac = 0.5f;
      st.scale = 1048576.0f;
    }
    st.reserved = 0u;
  }
}

}  // namespace lars
