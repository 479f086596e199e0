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
//
// Round 2 -- the guided single pass (per-frame statistics only; tiles of one image keep the two-level form because
// their counters are all-reduced between the levels).  Level B re-read the whole frame although ~1 % of the samples
// matter, and it was instruction-bound on its per-sample compare-and-branch (231 us per 8 x 20 MP against 150 us
// for level A).  Now:
//   sample   wb_hist_u16_sample_kernel   high-byte histogram of every 16th work unit (6 % of a large frame)
//   guess    wb_u16_candidates_kernel    per (frame, channel): the bucket the sample puts each percentile in plus
//                                        the neighbour on the nearer side -> a 256-entry class table (<= 4 slots)
//   guided   wb_hist_u16_guided_kernel   ONE read of the frame: exact high-byte histogram (as level A) and, for the
//                                        samples whose high byte has a class (one byte load from shared memory),
//                                        the low-byte histogram of that slot                      -- reads 6 B/px
//   select   (as before, on the EXACT high-byte histogram) additionally checks that every bucket it needs was a
//            candidate; then the slot histograms are the level-B result and level B has nothing to do for that
//            frame.  A wrong guess costs that frame the old level-B pass -- never exactness.
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
  int32_t done;                     // 1: the guided pass already counted every bucket of this channel (level B skips it)
  int32_t pad_[2];
};

constexpr int U16_CAND_SLOTS = 4;   // candidate high-byte buckets per channel in the guided pass
constexpr int U16_SAMPLE_STEP = 16; // large frames: the guess is made from every 16th work unit (4,096 pixels each)
constexpr int U16_SAMPLE_MIN_UNITS = 64;  // ... but from at least 64 units (262,144 pixels per channel): with fewer samples
                                          // the sampling noise of the 2 % / 98 % positions exceeds half a bucket
__host__ __device__ inline int u16_sample_step(long long units_per_frame) {
  const long long s = units_per_frame / U16_SAMPLE_MIN_UNITS;
  return s < 1 ? 1 : (s > U16_SAMPLE_STEP ? U16_SAMPLE_STEP : (int)s);
}

// per (frame, channel): which high bytes the guided pass also counts by low byte
struct __align__(16) U16Candidates {
  uint8_t cls[256];                 // 0 = not a candidate, k + 1 = slot k
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
  // guided single pass
  unsigned long long* hist_sample;  // [set][3][256]  high-byte histogram of the sampled units
  const U16Candidates* cand;        // [set][3]
  unsigned long long* cand_lo;      // [set][3][U16_CAND_SLOTS][256]
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

// The same walk handing out the 32-bit word and which half holds the sample (a compile-time constant after
// unrolling), so that a visitor can pick bytes with one PRMT instead of first isolating the 16-bit sample.
template <int C, class Visit>
__device__ __forceinline__ void u16_visit_span_words(const uint8_t* fsrc, long long b0, long long b1, int tid, Visit visit) {
  if (C == 3) {
    const long long vec_end = b0 + ((b1 - b0) / 48) * 48;
    for (long long off = b0 + 48ll * tid; off + 48 <= vec_end; off += 48ll * K1_THREADS) {
      const uint4* q = reinterpret_cast<const uint4*>(fsrc + off);
      const uint4 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2);
      const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        visit((2 * i) % 3, w[i], 0);
        visit((2 * i + 1) % 3, w[i], 1);
      }
    }
    const uint16_t* s16 = reinterpret_cast<const uint16_t*>(fsrc);
    for (long long s = vec_end / 2 + tid; s < b1 / 2; s += K1_THREADS) visit((int)(s % 3), (uint32_t)s16[s], 0);
  } else {
    const long long vec_end = b0 + ((b1 - b0) / 16) * 16;
    for (long long off = b0 + 16ll * tid; off + 16 <= vec_end; off += 16ll * K1_THREADS) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(fsrc + off));
      visit(0, v.x, 0); visit(1, v.x, 1); visit(2, v.y, 0);
      visit(0, v.z, 0); visit(1, v.z, 1); visit(2, v.w, 0);
    }
    const uint16_t* s16 = reinterpret_cast<const uint16_t*>(fsrc);
    for (long long s = vec_end / 2 + tid; s < b1 / 2; s += K1_THREADS)
      if ((s & 3) < 3) visit((int)(s & 3), (uint32_t)s16[s], 0);
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


// ---- sample: level A over every U16_SAMPLE_STEP-th work unit ------------------------------------
template <int C>
__global__ void __launch_bounds__(K1_THREADS, 2) wb_hist_u16_sample_kernel(const U16HistParams p) {
  extern __shared__ __align__(16) uint32_t u16_hist[];  // [3][256][32] lane-private
  const int tid = threadIdx.x, lane = tid & 31;
  const long long frame_bytes = p.n_pixels * C * 2;
  const uint32_t base = smem_u32(u16_hist) + 4u * lane;
  const int step = u16_sample_step(p.units_per_frame);
  const long long per_frame = (p.units_per_frame + step - 1) / step;
  const long long total = per_frame * p.n_frames;
  const long long G = gridDim.x;
  long long u = ((long long)blockIdx.x * total) / G;
  const long long u_end = ((long long)(blockIdx.x + 1) * total) / G;
  while (u < u_end) {
    const long long frame = u / per_frame;
    const long long fu0 = frame * per_frame;
    const long long span_end = (fu0 + per_frame < u_end) ? fu0 + per_frame : u_end;
    const uint8_t* fsrc = p.src + frame * p.frame_stride;
    for (int i = tid; i < 3 * 256 * 32; i += K1_THREADS) u16_hist[i] = 0u;
    __syncthreads();
    for (; u < span_end; ++u) {
      const long long b0 = (u - fu0) * step * U16Unit<C>::BYTES;
      long long b1 = b0 + U16Unit<C>::BYTES;
      if (b1 > frame_bytes) b1 = frame_bytes;
      u16_visit_span<C>(fsrc, b0, b1, tid, [&](int ch, uint32_t v) {
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + (uint32_t)ch * 32768u + ((v >> 8) << 7)) : "memory");
      });
    }
    __syncthreads();
    for (int b = tid; b < 3 * 256; b += K1_THREADS) {
      uint32_t sum = 0;
#pragma unroll 8
      for (int l = 0; l < 32; ++l) sum += u16_hist[b * 32 + ((l + tid) & 31)];
      if (sum) atomicAdd(&p.hist_sample[frame * 768 + b], (unsigned long long)sum);
    }
    __syncthreads();
  }
}

// ---- guess: candidate buckets from the sampled histogram ----------------------------------------
struct U16CandParams {
  const unsigned long long* hist_sample;  // [set][3][256]
  U16Candidates* cand;                    // [set][3]
  double q_lo, q_hi;
};

__global__ void __launch_bounds__(256) wb_u16_candidates_kernel(const U16CandParams p) {
  __shared__ unsigned long long cum[256];
  __shared__ unsigned long long warp_tot[8];
  __shared__ int picked[U16_CAND_SLOTS];
  const int v = threadIdx.x, lane = v & 31, warp = v >> 5;
  const long long base = (long long)blockIdx.x * 256;
  const unsigned long long own = p.hist_sample[base + v];
  unsigned long long x = own;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) warp_tot[warp] = x;
  if (v < U16_CAND_SLOTS) picked[v] = -1;
  __syncthreads();
  unsigned long long add = 0;
  for (int w = 0; w < warp; ++w) add += warp_tot[w];
  x += add;
  cum[v] = x;
  __syncthreads();
  const unsigned long long m = cum[255];
  __shared__ uint8_t nonempty[256];
  __shared__ int bucket_of[2];
  nonempty[v] = own ? 1 : 0;
  // the sample's own percentile positions: every thread tests its bucket for rank floor((m - 1) q)
  const unsigned long long below_v = v ? cum[v - 1] : 0ull;
  if (m > 0) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const unsigned long long r = (unsigned long long)floor((double)(m - 1) * (k == 0 ? p.q_lo : p.q_hi));
      if (below_v <= r && r < x) bucket_of[k] = v;
    }
  }
  __syncthreads();
  if (v < 2 && m > 0) {
    // ... and the non-empty neighbour on the side the rank sits closer to (below the middle of the bucket's mass ->
    // previous, else next)
    const int k = v, b = bucket_of[k];
    const unsigned long long r = (unsigned long long)floor((double)(m - 1) * (k == 0 ? p.q_lo : p.q_hi));
    const unsigned long long below = b ? cum[b - 1] : 0ull;
    const unsigned long long cnt = cum[b] - below;
    int nb = -1;
    if (2 * (r - below) < cnt) { for (int t = b - 1; t >= 0 && nb < 0; --t) if (nonempty[t]) nb = t; }
    else { for (int t = b + 1; t < 256 && nb < 0; ++t) if (nonempty[t]) nb = t; }
    picked[2 * k] = b;
    picked[2 * k + 1] = nb;
  }
  __syncthreads();
  uint8_t c = 0;
#pragma unroll
  for (int k = U16_CAND_SLOTS - 1; k >= 0; --k)     // a bucket picked twice keeps its lowest slot
    if (picked[k] == v) c = (uint8_t)(k + 1);
  p.cand[blockIdx.x].cls[v] = c;
}

// ---- guided: exact high-byte histogram + low-byte histograms of the candidate buckets, one read --
// One CTA of 1,024 threads per SM.  Shared memory: hi[3][256][32] lane-private as in level A (96 KB),
// lo[3][slots][256][8] (96 KB: lanes congruent modulo 8 share a counter, so a constant channel or a saturated
// region -- every sample a hit on one address -- is serialised 4-fold at most) and the class bytes (768 B).
// Per sample: PRMT (address of the class byte), LDS.U8, IMAD + RED (high-byte counter), ISETP and a branch around
// the low-byte counter's three instructions.  History (ncu, 8 x 20 MP): a branch per sample with the
// class load in front of it 287 us (54 % issue-active, short_scoreboard); class loads hoisted eight at a time
// 271 us (190 M warp instructions: the ~1 % hits made 27 % of the warp-level sample slots take a 17-instruction
// divergent path with a warp match; the assembler turns a predicated shared atomic back into a branch, so the
// branch stays and what it guards shrank to PRMT + IMAD + RED); this form: see profiles/.
constexpr int U16_GUIDED_THREADS = 1024;
constexpr int U16_GUIDED_LO_COPIES = 8;
constexpr int U16_GUIDED_LO_BINS = 3 * U16_CAND_SLOTS * 256;
constexpr int U16_GUIDED_LO_WORDS = U16_GUIDED_LO_BINS * U16_GUIDED_LO_COPIES;
constexpr int U16_GUIDED_SMEM_BYTES = U16_HI_SMEM_BYTES + U16_GUIDED_LO_WORDS * 4 + 3 * 256;   // 96 KB + 96 KB + 768 B
template <int C>
__global__ void __launch_bounds__(U16_GUIDED_THREADS, 1) wb_hist_u16_guided_kernel(const U16HistParams p) {
  extern __shared__ __align__(16) uint32_t u16_hist[];  // hi[3][256][32], then lo[3][slots][256][8], then cls[3][256]
  uint32_t* lo_hist = u16_hist + 3 * 256 * 32;
  uint8_t* cls_s = reinterpret_cast<uint8_t*>(lo_hist + U16_GUIDED_LO_WORDS);
  const int tid = threadIdx.x, lane = tid & 31;
  const long long frame_bytes = p.n_pixels * C * 2;
  const uint32_t base = smem_u32(u16_hist) + 4u * lane;
  // low-byte counter of (channel ch, class c, low byte lo): lo_lane + ((ch * slots + c - 1) * 256 + lo) * 32
  const uint32_t lo_lane = smem_u32(lo_hist) + 4u * (lane & (U16_GUIDED_LO_COPIES - 1)) - 256u * U16_GUIDED_LO_COPIES * 4u;
  const uint32_t cls_base = smem_u32(cls_s);
  const long long G = gridDim.x;
  long long u = ((long long)blockIdx.x * p.total_units) / G;
  const long long u_end = ((long long)(blockIdx.x + 1) * p.total_units) / G;
  // Per-channel constants so that a sample costs PRMT, LDS.U8, IMAD, RED, ISETP and a branch:
  //   the class tables are 256-byte aligned in the shared window, so ONE byte permute forms the class byte's address
  //   (the table's address with the sample's high byte as its low byte), and the high-byte counter's address is that
  //   address times 128 plus a constant (counters are 128 bytes apart: hi[ch][byte][32 lanes]).
  // (scalars selected by ternaries, not arrays: the scalar tails index by a run-time channel, which would put arrays
  // in local memory for the hot loop as well)
  const uint32_t lo_ch = U16_CAND_SLOTS * 256u * U16_GUIDED_LO_COPIES * 4u;
  uint32_t cls0 = cls_base, cls1 = cls_base + 256u, cls2 = cls_base + 512u;
  asm("" : "+r"(cls0), "+r"(cls1), "+r"(cls2));     // opaque: keep them in registers instead of re-adding per sample
  const uint32_t red0 = base - cls0 * 128u, red1 = base + 32768u - cls1 * 128u, red2 = base + 65536u - cls2 * 128u;
  const uint32_t lo0 = lo_lane, lo1 = lo_lane + lo_ch, lo2 = lo_lane + 2u * lo_ch;
  auto one = [&](int ch, uint32_t w, int half) {
    const uint32_t cls_at = ch == 0 ? cls0 : (ch == 1 ? cls1 : cls2);
    const uint32_t red_k = ch == 0 ? red0 : (ch == 1 ? red1 : red2);
    const uint32_t la = __byte_perm(w, cls_at, half ? 0x7653 : 0x7651);
    uint32_t c;
    asm("ld.shared.u8 %0, [%1];" : "=r"(c) : "r"(la));
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(la * 128u + red_k));
    if (c) {       // ~1 % of the samples
      const uint32_t lob = __byte_perm(w, 0u, half ? 0x4442 : 0x4440);
      asm volatile("red.shared.add.u32 [%0], 1;" ::"r"((ch == 0 ? lo0 : (ch == 1 ? lo1 : lo2)) + c * (256u * U16_GUIDED_LO_COPIES * 4u) + lob * (U16_GUIDED_LO_COPIES * 4u)));
    }
  };
  while (u < u_end) {
    const long long frame = u / p.units_per_frame;
    const long long fu0 = frame * p.units_per_frame;
    const long long span_end = (fu0 + p.units_per_frame < u_end) ? fu0 + p.units_per_frame : u_end;
    const long long b0 = (u - fu0) * U16Unit<C>::BYTES;
    long long b1 = (span_end - fu0) * U16Unit<C>::BYTES;
    if (b1 > frame_bytes) b1 = frame_bytes;
    const uint8_t* fsrc = p.src + frame * p.frame_stride;
    u = span_end;
    for (int i = tid; i < 3 * 256 * 32 + U16_GUIDED_LO_WORDS; i += U16_GUIDED_THREADS) u16_hist[i] = 0u;
    {
      const uint32_t* src_cls = reinterpret_cast<const uint32_t*>(p.cand + frame * 3);
      uint32_t* dst_cls = reinterpret_cast<uint32_t*>(cls_s);
      for (int i = tid; i < 3 * 64; i += U16_GUIDED_THREADS) dst_cls[i] = src_cls[i];
    }
    __syncthreads();
    const uint16_t* s16 = reinterpret_cast<const uint16_t*>(fsrc);
    if (C == 3) {
      // 48 contiguous bytes per lane = 24 samples, sample s of channel s % 3
      const long long vec_end = b0 + ((b1 - b0) / 48) * 48;
      auto chunk = [&](const uint4& v0, const uint4& v1, const uint4& v2) {
        const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          one((2 * i) % 3, w[i], 0);
          one((2 * i + 1) % 3, w[i], 1);
        }
      };
      // two 48-byte chunks per trip, all six 128-bit loads issued before the first sample is counted: with one chunk
      // per trip the kernel waited on global loads (ncu: long_scoreboard the top stall, 48 % of the DRAM peak)
      long long off = b0 + 48ll * tid;
      for (; off + 48ll * U16_GUIDED_THREADS + 48 <= vec_end; off += 96ll * U16_GUIDED_THREADS) {
        const uint4* q = reinterpret_cast<const uint4*>(fsrc + off);
        const uint4* r = reinterpret_cast<const uint4*>(fsrc + off + 48ll * U16_GUIDED_THREADS);
        const uint4 a0 = __ldg(q), a1 = __ldg(q + 1), a2 = __ldg(q + 2);
        const uint4 c0 = __ldg(r), c1 = __ldg(r + 1), c2 = __ldg(r + 2);
        chunk(a0, a1, a2);
        chunk(c0, c1, c2);
      }
      for (; off + 48 <= vec_end; off += 48ll * U16_GUIDED_THREADS) {
        const uint4* q = reinterpret_cast<const uint4*>(fsrc + off);
        const uint4 a0 = __ldg(q), a1 = __ldg(q + 1), a2 = __ldg(q + 2);
        chunk(a0, a1, a2);
      }
      for (long long sidx = vec_end / 2 + tid; sidx < b1 / 2; sidx += U16_GUIDED_THREADS) one((int)(sidx % 3), (uint32_t)s16[sidx], 0);
    } else {
      const long long vec_end = b0 + ((b1 - b0) / 16) * 16;
      for (long long off = b0 + 16ll * tid; off + 16 <= vec_end; off += 16ll * U16_GUIDED_THREADS) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(fsrc + off));
        one(0, v.x, 0); one(1, v.x, 1); one(2, v.y, 0);
        one(0, v.z, 0); one(1, v.z, 1); one(2, v.w, 0);
      }
      for (long long sidx = vec_end / 2 + tid; sidx < b1 / 2; sidx += U16_GUIDED_THREADS)
        if ((sidx & 3) < 3) one((int)(sidx & 3), (uint32_t)s16[sidx], 0);
    }
    __syncthreads();
    for (int b = tid; b < 3 * 256; b += U16_GUIDED_THREADS) {
      uint32_t sum = 0;
#pragma unroll 8
      for (int l = 0; l < 32; ++l) sum += u16_hist[b * 32 + ((l + tid) & 31)];
      if (sum) atomicAdd(&p.hist_hi[frame * 768 + b], (unsigned long long)sum);
    }
    for (int b = tid; b < U16_GUIDED_LO_BINS; b += U16_GUIDED_THREADS) {
      uint32_t sum = 0;
#pragma unroll
      for (int l = 0; l < U16_GUIDED_LO_COPIES; ++l) sum += lo_hist[b * U16_GUIDED_LO_COPIES + ((l + tid) & (U16_GUIDED_LO_COPIES - 1))];
      if (sum) atomicAdd(&p.cand_lo[frame * U16_GUIDED_LO_BINS + b], (unsigned long long)sum);
    }
    __syncthreads();
  }
}

// ---- select -------------------------------------------------------------------------------
struct U16SelectParams {
  const unsigned long long* hist_hi;  // [set][3][256]
  U16Select* select;                  // [set][3]
  double q_lo, q_hi;
  // guided pass (all three nullptr otherwise): candidate classes, their low-byte histograms, and where level B's go
  const U16Candidates* cand;          // [set][3]
  const unsigned long long* cand_lo;  // [set][3][U16_CAND_SLOTS][256]
  unsigned long long* hist_lo;        // [set][3][U16_MAX_BUCKETS][256]
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
  __shared__ U16Select chosen;
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
    s.pad_[0] = s.pad_[1] = 0;
    // guided pass: usable iff every bucket that holds a rank was one of the candidates
    s.done = 0;
    if (p.cand) {
      s.done = 1;
      for (int b = 0; b < s.n_buckets; ++b)
        if (p.cand[blockIdx.x].cls[s.buckets[b]] == 0) s.done = 0;
    }
    chosen = s;
    p.select[blockIdx.x] = s;
  }
  __syncthreads();
  if (p.cand && chosen.done) {           // the candidates' low-byte histograms ARE level B's result for this channel
    for (int b = 0; b < chosen.n_buckets; ++b) {
      const int slot = (int)p.cand[blockIdx.x].cls[chosen.buckets[b]] - 1;
      p.hist_lo[((long long)blockIdx.x * U16_MAX_BUCKETS + b) * 256 + v] =
          p.cand_lo[((long long)blockIdx.x * U16_CAND_SLOTS + slot) * 256 + v];
    }
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
    // a channel the guided pass has already counted (done) wants nothing from this pass
    const int w00 = sel[0].done ? -1 : sel[0].buckets[2 * p.lo_pass], w01 = sel[0].done ? -1 : sel[0].buckets[2 * p.lo_pass + 1];
    const int w10 = sel[1].done ? -1 : sel[1].buckets[2 * p.lo_pass], w11 = sel[1].done ? -1 : sel[1].buckets[2 * p.lo_pass + 1];
    const int w20 = sel[2].done ? -1 : sel[2].buckets[2 * p.lo_pass], w21 = sel[2].done ? -1 : sel[2].buckets[2 * p.lo_pass + 1];
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
