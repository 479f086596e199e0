// Kernels over float maps / raw frames that sit next to the fused pass:
//
//   K3  select_*            exact order statistics (median) of a float32 map: 3-pass (11 + 11 + 10 bit)
//                           radix select on order-preserving keys (np.median,
//                           process-images.py:508, :654; process-ndvi.py:62)
//   K4  map_stats_f32       statistics + np.histogram of an arbitrary float32 map
//                           (analyze_index process-images.py:492-513 called on a host array;
//                           analyze_ndvi_statistics process-ndvi.py:50-73; plt.hist :97)
//   K5  colormap_f32        Normalize(vmin, vmax) + colormap LUT of a float32 map
//                           (process-images.py:689-695, :949-956)
//   K6  ndvi_f64_u8         float64 NDVI of a raw uint8 frame (calculate_ndvi, process-ndvi.py:18-31)
//   K7  index_planes_f32    (hi - lo) / (hi + lo + eps) clipped, separate float32 planes
//                           (backend-process.py:28-38)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lars_b200.h"
#include "pixel_math.h"
#include "ptx_sm100.cuh"

namespace lars {

constexpr int MAP_THREADS = 256;
constexpr int MAP_HIST_ROWS = LARS_MAX_BINS;  // bins supported by the generic map kernels

// ------------------------------------------------------------------------------------------
// K4: statistics of a float32 map
// ------------------------------------------------------------------------------------------
struct __align__(16) MapPartial {
  double sx, sd, sdd;
  float mn, mx;
  uint32_t above;
  uint32_t has_nan;
  unsigned long long count;
  uint32_t hist[MAP_HIST_ROWS];
};

struct MapStatsParams {
  const float* data;       // [n_maps][stride]
  long long n;             // elements per map
  long long stride;        // elements
  MapPartial* partials;    // [n_maps][gridDim.x]
  const uint8_t* subbin;   // [MAP_SUBBIN_BYTES] built by map_subbin_table_kernel for this bin count
  float threshold;
  int bins;
};

constexpr int MAP_SUBBIN_BYTES = 8192;   // 4097 entries; the rest is padding so that ANY 13-bit index is a valid load

// The sub-bin table of one bin count, built once per call (every CTA of the statistics kernel copies it).
__global__ void __launch_bounds__(256) map_subbin_table_kernel(uint8_t* table, int bins) {
  __shared__ float edges_s[LARS_MAX_BINS + 1];
  const double step = LARS_DDIV(2.0, (double)bins);
  for (int i = threadIdx.x; i <= bins; i += blockDim.x)
    edges_s[i] = (i == bins) ? 1.0f : (float)LARS_DADD(LARS_DMUL((double)i, step), -1.0);
  __syncthreads();
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < MAP_SUBBIN_BYTES; k += gridDim.x * blockDim.x)
    table[k] = k < LARS_SUBBIN_COUNT ? lars_hist_subbin_entry(k, edges_s, bins) : (uint8_t)LARS_SUBBIN_AMBIGUOUS;
}

// One element.  Branch-free except for the ~1 % of elements in a sub-bin that straddles an edge:
//   * sub-bin index without a float->int conversion (the XU pipe was 30 % busy with them): y = fma(x, 2048, 2047.5),
//     then adding 1.5 * 2^23 rounds y to an integer that sits in the low mantissa bits.  Round-to-nearest instead of
//     truncation moves the index by at most one sub-bin at a boundary, which the table's 2^-20 margin covers
//     (pixel_math.h);
//   * out-of-range values and NaN (np.histogram drops them) go to a trash row instead of around a branch;
//   * NaN is detected from the sum afterwards, not per element.
__device__ __forceinline__ void map_stats_accum(float x, float k, float thr, const float* edges_s, const uint8_t* subbin_s,
                                                int bins, uint32_t* hist_lane, float& mn, float& mx, float& gx,
                                                float& gd, float& gdd, uint32_t& above) {
  mn = fminf(mn, x);
  mx = fmaxf(mx, x);
  above += (x > thr) ? 1u : 0u;
  const float d = LARS_FSUB(x, k);
  gx += x;
  gd += d;
  gdd = fmaf(d, d, gdd);
  const bool in_range = fabsf(x) <= 1.0f;
  const uint32_t idx = lars_hist_subbin_index_rn(x);
  int b = subbin_s[idx];
  if (b == LARS_SUBBIN_AMBIGUOUS && in_range) b = lars_hist_bin_edges(x, edges_s, bins);
  b = in_range ? b : bins;                                   // row `bins` is the trash row
  atomicAdd(hist_lane + b * 32, 1u);
}

__global__ void __launch_bounds__(MAP_THREADS) map_stats_f32_kernel(const MapStatsParams p) {
  extern __shared__ __align__(16) uint8_t ms_smem[];
  uint8_t* subbin_s = ms_smem;                                                                  // [MAP_SUBBIN_BYTES]
  uint32_t* hist = reinterpret_cast<uint32_t*>(ms_smem + MAP_SUBBIN_BYTES);                     // [bins + 1][32]
  float* edges_s = reinterpret_cast<float*>(ms_smem + MAP_SUBBIN_BYTES + (size_t)(p.bins + 1) * 32 * 4);  // [bins + 1]
  double* red = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(edges_s) + ((p.bins + 1 + 3) / 4) * 16);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int map = blockIdx.y;
  const float* x = p.data + (long long)map * p.stride;
  for (int i = tid; i < (p.bins + 1) * 32; i += MAP_THREADS) hist[i] = 0u;
  {  // np.linspace(-1, 1, bins + 1) in float64, rounded to float32, last edge exact
    const double step = LARS_DDIV(2.0, (double)p.bins);
    for (int i = tid; i <= p.bins; i += MAP_THREADS)
      edges_s[i] = (i == p.bins) ? 1.0f : (float)LARS_DADD(LARS_DMUL((double)i, step), -1.0);
  }
  for (int i = tid; i < MAP_SUBBIN_BYTES / 16; i += MAP_THREADS)
    reinterpret_cast<uint4*>(subbin_s)[i] = __ldg(reinterpret_cast<const uint4*>(p.subbin) + i);
  __syncthreads();

  const float k = x[0];
  const float thr = p.threshold;
  float mn = INFINITY, mx = -INFINITY;
  double sx = 0.0, sd = 0.0, sdd = 0.0;
  uint32_t above = 0;
  unsigned long long count = 0;
  uint32_t* hist_lane = hist + lane;

  // contiguous chunk per CTA, 16-byte vectors inside (map bases are 16-byte aligned)
  const long long nvec = p.n / 4;
  const long long per = (nvec + gridDim.x - 1) / gridDim.x;
  const long long v0 = (long long)blockIdx.x * per;
  const long long v1 = (v0 + per < nvec) ? v0 + per : nvec;
  const float4* xv = reinterpret_cast<const float4*>(x);
  // four independent 128-bit loads in flight per thread (one per trip is latency-bound: measured 2.0 TB/s);
  // the moments of the 16 elements are summed in float32 and promoted to float64 once
  long long v = v0 + tid;
  for (; v + 3ll * MAP_THREADS < v1; v += 4ll * MAP_THREADS) {
    float4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = __ldg(xv + v + (long long)u * MAP_THREADS);
    float gx = 0.f, gd = 0.f, gdd = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      map_stats_accum(q[u].x, k, thr, edges_s, subbin_s, p.bins, hist_lane, mn, mx, gx, gd, gdd, above);
      map_stats_accum(q[u].y, k, thr, edges_s, subbin_s, p.bins, hist_lane, mn, mx, gx, gd, gdd, above);
      map_stats_accum(q[u].z, k, thr, edges_s, subbin_s, p.bins, hist_lane, mn, mx, gx, gd, gdd, above);
      map_stats_accum(q[u].w, k, thr, edges_s, subbin_s, p.bins, hist_lane, mn, mx, gx, gd, gdd, above);
    }
    sx += (double)gx; sd += (double)gd; sdd += (double)gdd;
    count += 16;
  }
  for (; v < v1; v += MAP_THREADS) {
    const float4 q = __ldg(xv + v);
    float gx = 0.f, gd = 0.f, gdd = 0.f;
    map_stats_accum(q.x, k, thr, edges_s, subbin_s, p.bins, hist_lane, mn, mx, gx, gd, gdd, above);
    map_stats_accum(q.y, k, thr, edges_s, subbin_s, p.bins, hist_lane, mn, mx, gx, gd, gdd, above);
    map_stats_accum(q.z, k, thr, edges_s, subbin_s, p.bins, hist_lane, mn, mx, gx, gd, gdd, above);
    map_stats_accum(q.w, k, thr, edges_s, subbin_s, p.bins, hist_lane, mn, mx, gx, gd, gdd, above);
    sx += (double)gx; sd += (double)gd; sdd += (double)gdd;
    count += 4;
  }
  if (blockIdx.x == gridDim.x - 1) {  // scalar tail (n % 4 elements)
    for (long long i = nvec * 4 + tid; i < p.n; i += MAP_THREADS) {
      float gx = 0.f, gd = 0.f, gdd = 0.f;
      map_stats_accum(x[i], k, thr, edges_s, subbin_s, p.bins, hist_lane, mn, mx, gx, gd, gdd, above);
      sx += (double)gx; sd += (double)gd; sdd += (double)gdd;
      count += 1;
    }
  }
  // block reduction
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, d);
    sd += __shfl_xor_sync(0xffffffffu, sd, d);
    sdd += __shfl_xor_sync(0xffffffffu, sdd, d);
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    above += __shfl_xor_sync(0xffffffffu, above, d);
    count += __shfl_xor_sync(0xffffffffu, count, d);
  }
  if (lane == 0) {
    double* r = red + warp * 8;
    r[0] = sx; r[1] = sd; r[2] = sdd; r[3] = (double)mn; r[4] = (double)mx;
    r[5] = (double)above; r[6] = 0.0; r[7] = (double)count;
  }
  __syncthreads();
  MapPartial* rec = p.partials + (long long)map * gridDim.x + blockIdx.x;
  for (int b = tid; b < p.bins; b += MAP_THREADS) {
    uint32_t s = 0;
    for (int l = 0; l < 32; ++l) s += hist[b * 32 + ((l + tid) & 31)];
    rec->hist[b] = s;
  }
  if (tid == 0) {
    double a[8];
    for (int q = 0; q < 8; ++q) a[q] = red[q];
    for (int w = 1; w < MAP_THREADS / 32; ++w) {
      const double* r = red + w * 8;
      a[0] += r[0]; a[1] += r[1]; a[2] += r[2];
      a[3] = fmin(a[3], r[3]); a[4] = fmax(a[4], r[4]);
      a[5] += r[5]; a[7] += r[7];
    }
    rec->sx = a[0]; rec->sd = a[1]; rec->sdd = a[2];
    rec->mn = (float)a[3]; rec->mx = (float)a[4];
    rec->above = (uint32_t)a[5];
    rec->has_nan = (a[0] != a[0]) ? 1u : 0u;     // a NaN anywhere poisons the sum (fminf / fmaxf skip NaN)
    rec->count = (unsigned long long)a[7];
  }
}

struct MapFinalizeParams {
  const MapPartial* partials;  // [n_maps][n_parts]
  const float* data;
  long long stride;
  lars_index_stats* stats;     // [n_maps]
  int n_parts, bins;
  float threshold;
};

constexpr int MAP_FIN_SPLIT = 8;   // partials are folded by 8 thread groups per bin, then combined in order

__global__ void __launch_bounds__(MAP_HIST_ROWS * MAP_FIN_SPLIT) map_stats_finalize_kernel(const MapFinalizeParams p) {
  __shared__ unsigned long long hpart[MAP_FIN_SPLIT][MAP_HIST_ROWS];
  const int map = blockIdx.x, tid = threadIdx.x;
  const MapPartial* recs = p.partials + (long long)map * p.n_parts;
  lars_index_stats& o = p.stats[map];
  {
    const int bin = tid % MAP_HIST_ROWS, grp = tid / MAP_HIST_ROWS;
    unsigned long long h = 0;
    if (bin < p.bins)
      for (int s = grp; s < p.n_parts; s += MAP_FIN_SPLIT) h += recs[s].hist[bin];
    hpart[grp][bin] = h;
  }
  __syncthreads();
  if (tid < MAP_HIST_ROWS) {
    unsigned long long h = 0;
#pragma unroll
    for (int g = 0; g < MAP_FIN_SPLIT; ++g) h += hpart[g][tid];   // integer sums: any order is exact
    o.hist[tid] = h;
  }
  // moments: lane l of warp 0 folds partials l, l + 32, ... in order, then a fixed butterfly -- the
  // summation tree depends only on n_parts, so results are reproducible (one thread walking all
  // partials took 119 us for 592 of them)
  if (tid < 32) {
    double sx = 0.0, sd = 0.0, sdd = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    unsigned long long cnt = 0, above = 0;
    uint32_t nan = 0;
    for (int s = tid; s < p.n_parts; s += 32) {
      const MapPartial& r = recs[s];
      if (!r.count) continue;
      sx += r.sx; sd += r.sd; sdd += r.sdd;
      mn = fminf(mn, r.mn); mx = fmaxf(mx, r.mx);
      cnt += r.count; above += r.above; nan |= r.has_nan;
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
      nan |= __shfl_xor_sync(0xffffffffu, nan, d);
    }
    if (tid == 0) {
      const double n = (double)cnt;
      const double mean = cnt ? sx / n : 0.0;
      const double md = cnt ? sd / n : 0.0;
      double var = cnt ? sdd / n - md * md : 0.0;
      var = var > 0.0 ? var : 0.0;
      o.count = cnt; o.count_above = above;
      o.sum = sx; o.sumsq = cnt ? (var + mean * mean) * n : 0.0;
      o.mean = nan ? NAN : mean; o.std = nan ? NAN : sqrt(var);
      o.min = nan ? NAN : mn; o.max = nan ? NAN : mx;
      o.threshold = p.threshold; o.bins = (uint32_t)p.bins;
    }
  }
}

// ------------------------------------------------------------------------------------------
// K3 / K3d: exact order statistics by radix select, ONE persistent kernel for all passes
// ------------------------------------------------------------------------------------------
// float32: 11 + 11 + 10 key bits (3 passes); float64: 11 + 11 + 11 + 11 + 11 + 9 (6 passes).  A pass streams the map,
// counts the current digit of the elements that match the prefix chosen so far in a shared histogram (2,048 bins, 4
// lane copies, one set per rank while the two ranks still share a prefix), folds it into the global histogram of that
// pass, meets the other CTAs at a grid barrier, and then EVERY CTA scans the global histogram for itself and extends
// the prefix -- no separate scan launch, no second barrier, nothing to clear between passes (each pass has its own
// global histogram, zeroed by one memset in front of the launch).
// Round 1 / early round 2 launched one kernel per pass: 59 us per 12 MP float32 map = 3 x ~20 us, of which ~7 us is
// streaming (ncu: 8 M warp instructions and 48 MB per pass, long_scoreboard on top) and the rest launch, ramp-up of
// 296 CTAs x 512 threads for ~80 elements per thread, and the last CTA's scan.  A cheaper pass (no shared histogram
// for the sparse passes) measured no gain; removing the launches and ramps is what this form does: 59.4 -> 53.6 us,
// 50.1 with one scan for both ranks while they share a prefix, 47.4 with one 1,024-thread CTA per SM instead of two
// of 512 (half the barrier arrivals and fold atomics); float64 155 -> 122 us.  The launch is cooperative (all CTAs
// must be co-resident for the barrier).
constexpr int SEL_BINS = 2048;                 // bins of the widest digit
#ifndef LARS_SEL_LANES
#define LARS_SEL_LANES 4
#endif
constexpr int SEL_LANES = LARS_SEL_LANES;      // lane copies of each counter (lanes congruent modulo this share one)
constexpr int SEL_THREADS = 1024;
constexpr int SEL_SMEM_BYTES = 2 * SEL_BINS * SEL_LANES * 4;   // 64 KB: [2][2048][4]

template <typename T> struct SelectTraits;
template <> struct SelectTraits<float> {
  typedef uint32_t Key;
  typedef float4 Vec;
  static constexpr int PASSES = 3, PER_VEC = 4;
  __host__ __device__ static constexpr int bits(int pass) { return pass == 2 ? 10 : 11; }
  __host__ __device__ static constexpr int shift(int pass) { return pass == 0 ? 21 : (pass == 1 ? 10 : 0); }
  __device__ static __forceinline__ Key key(float x) {
    const uint32_t b = __float_as_uint(x);
    return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);          // order-preserving
  }
  __device__ static __forceinline__ float value(Key k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k); }
  __device__ static __forceinline__ float mean2(float a, float b) { return LARS_FMUL(LARS_FADD(a, b), 0.5f); }   // np.mean of the two middle values
};
template <> struct SelectTraits<double> {
  typedef unsigned long long Key;
  typedef double2 Vec;
  static constexpr int PASSES = 6, PER_VEC = 2;
  __host__ __device__ static constexpr int bits(int pass) { return pass == 5 ? 9 : 11; }
  __host__ __device__ static constexpr int shift(int pass) { return pass == 5 ? 0 : 53 - 11 * pass; }
  __device__ static __forceinline__ Key key(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return b ^ ((unsigned long long)((long long)b >> 63) | 0x8000000000000000ull);
  }
  __device__ static __forceinline__ double value(Key k) {
    return __longlong_as_double((long long)((k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k));
  }
  __device__ static __forceinline__ double mean2(double a, double b) { return LARS_DMUL(LARS_DADD(a, b), 0.5); }
};

template <typename T>
struct SelectState {
  unsigned int barrier;         // arrivals at the grid barrier (zeroed with the histograms)
  unsigned int pad_[3];
  T value[2];                   // the two order statistics
  T median;                     // their mean in T (np.median for even n)
  T pad2_;
  unsigned long long hist[SelectTraits<T>::PASSES][2][SEL_BINS];
};

template <typename T>
__global__ void __launch_bounds__(SEL_THREADS) select_kernel(const T* __restrict__ data, long long n, SelectState<T>* st,
                                                             unsigned long long rank_lo, unsigned long long rank_hi) {
  typedef SelectTraits<T> Tr;
  typedef typename Tr::Key Key;
  extern __shared__ __align__(16) uint32_t sel_hist[];  // [2][SEL_BINS][SEL_LANES]
  __shared__ unsigned long long wtot[SEL_THREADS / 32];
  __shared__ unsigned long long sh_rank[2];
  __shared__ Key sh_prefix[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { sh_rank[0] = rank_lo; sh_rank[1] = rank_hi; sh_prefix[0] = 0; sh_prefix[1] = 0; }
  __syncthreads();
  const uint32_t a0 = smem_u32(sel_hist) + 4u * (lane & (SEL_LANES - 1)), a1 = a0 + (uint32_t)SEL_BINS * SEL_LANES * 4u;
  const long long nvec = n / Tr::PER_VEC;
  const typename Tr::Vec* xv = reinterpret_cast<const typename Tr::Vec*>(data);
  const long long stride = (long long)gridDim.x * SEL_THREADS;

#pragma unroll 1
  for (int pass = 0; pass < Tr::PASSES; ++pass) {
    const int bits = Tr::bits(pass), dg_shift = Tr::shift(pass);
    const int nb = 1 << bits;
    const Key p0 = sh_prefix[0], p1 = sh_prefix[1];
    const bool same = (pass == 0) || (p0 == p1);
    for (int i = tid; i < (same ? 1 : 2) * SEL_BINS * SEL_LANES; i += SEL_THREADS) sel_hist[i] = 0u;
    __syncthreads();
    // the prefix test is one masked compare: (key ^ prefix_bits) & hi_mask == 0, with hi_mask = 0 in pass 0
    const int hi_shift = dg_shift + bits;                  // the key's width in pass 0
    const Key hi_mask = pass == 0 ? (Key)0 : (Key)(~(Key)0 << hi_shift);
    const Key want0 = pass == 0 ? (Key)0 : (Key)(p0 << hi_shift);
    const Key want1 = pass == 0 ? (Key)0 : (Key)(p1 << hi_shift);
    const uint32_t dg_mask = (uint32_t)nb - 1u;
    auto visit = [&](T x) {
      const Key key = Tr::key(x);
      const uint32_t off = ((uint32_t)(key >> dg_shift) & dg_mask) * (SEL_LANES * 4u);
      if (((key ^ want0) & hi_mask) == (Key)0) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a0 + off) : "memory");
      if (!same && ((key ^ want1) & hi_mask) == (Key)0) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a1 + off) : "memory");
    };
    auto visit_vec = [&](const typename Tr::Vec& q) {
      if constexpr (Tr::PER_VEC == 4) { visit(q.x); visit(q.y); visit(q.z); visit(q.w); }
      else { visit(q.x); visit(q.y); }
    };
    long long v = (long long)blockIdx.x * SEL_THREADS + tid;
    for (; v + 3 * stride < nvec; v += 4 * stride) {     // four independent 128-bit loads in flight
      const typename Tr::Vec q0 = __ldg(xv + v), q1 = __ldg(xv + v + stride), q2 = __ldg(xv + v + 2 * stride), q3 = __ldg(xv + v + 3 * stride);
      visit_vec(q0); visit_vec(q1); visit_vec(q2); visit_vec(q3);
    }
    for (; v < nvec; v += stride) visit_vec(__ldg(xv + v));
    if (blockIdx.x == 0)
      for (long long i = nvec * Tr::PER_VEC + tid; i < n; i += SEL_THREADS) visit(data[i]);
    __syncthreads();
    for (int b = tid; b < (same ? 1 : 2) * nb; b += SEL_THREADS) {
      const int set = b >= nb ? 1 : 0, bin = b - set * nb;
      uint32_t s = 0;
#pragma unroll
      for (int l = 0; l < SEL_LANES; ++l) s += sel_hist[(set * SEL_BINS + bin) * SEL_LANES + ((l + tid) & (SEL_LANES - 1))];
      if (s) atomicAdd(&st->hist[pass][set][bin], (unsigned long long)s);
    }
    // ---- grid barrier: every CTA's counts of this pass are in the global histogram
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      atomicAdd(&st->barrier, 1u);
      const unsigned int target = gridDim.x * (unsigned int)(pass + 1);
      unsigned int seen;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&st->barrier) : "memory");
        if (seen < target) __nanosleep(40);
      } while (seen < target);
    }
    __syncthreads();
    // ---- digit selection, by every CTA for itself; thread t owns bins 4 t .. 4 t + 3.  While both ranks share a
    // prefix they read the same histogram: one load + scan serves both
#pragma unroll 1
    for (int r = 0; r < (same ? 1 : 2); ++r) {
      unsigned long long c[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) c[j] = (4 * tid + j < nb) ? __ldcg(&st->hist[pass][r][4 * tid + j]) : 0ull;
      unsigned long long x = c[0] + c[1] + c[2] + c[3];       // inclusive scan over threads of the 4-bin sums
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
      }
      if (lane == 31) wtot[warp] = x;
      __syncthreads();
      unsigned long long add = 0;
      for (int w = 0; w < warp; ++w) add += wtot[w];
      x += add;
      const unsigned long long rk0 = sh_rank[same ? 0 : r], rk1 = sh_rank[1];
      const Key pref0 = sh_prefix[same ? 0 : r], pref1 = sh_prefix[1];
      __syncthreads();
      unsigned long long below = x - (c[0] + c[1] + c[2] + c[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (below <= rk0 && rk0 < below + c[j]) {
          sh_prefix[same ? 0 : r] = (Key)(pref0 << bits) | (Key)(4 * tid + j);
          sh_rank[same ? 0 : r] = rk0 - below;
        }
        if (same && below <= rk1 && rk1 < below + c[j]) {
          sh_prefix[1] = (Key)(pref1 << bits) | (Key)(4 * tid + j);
          sh_rank[1] = rk1 - below;
        }
        below += c[j];
      }
      __syncthreads();
    }
  }
  if (blockIdx.x == 0 && tid == 0) {
    const T a = Tr::value(sh_prefix[0]), b = Tr::value(sh_prefix[1]);
    st->value[0] = a;
    st->value[1] = b;
    st->median = Tr::mean2(a, b);
  }
}

// ------------------------------------------------------------------------------------------
// K5: colormap a float32 map
// ------------------------------------------------------------------------------------------
struct ColormapParams {
  const float* data;
  uint8_t* rgb;            // [n][3]
  const uint32_t* cmap;    // [256] packed R | G<<8 | B<<16
  long long n;
  float vmin, vmax;
  int unit_range;          // vmin == -1 && vmax == 1 -> single-FMA slot formula
};

__global__ void __launch_bounds__(256) colormap_f32_kernel(const ColormapParams p) {
  __shared__ uint32_t cm[256];
  __shared__ uint32_t stage[8][2][96];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  cm[tid] = p.cmap[tid];
  __syncthreads();
  // each warp iteration: two groups of 128 pixels (2 x 32 float4 in flight), 2 x 384 bytes out
  const long long ngroups = (p.n + 127) / 128;
  uint32_t* out32 = reinterpret_cast<uint32_t*>(p.rgb);
  const long long total_bytes = p.n * 3;
  const long long stride = (long long)gridDim.x * 8;
  for (long long g0 = (long long)blockIdx.x * 8 + warp; g0 < ngroups; g0 += 2 * stride) {
    float v[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long px = (g0 + u * stride) * 128 + 4 * lane;
      if (px + 4 <= p.n) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p.data + px));
        v[u][0] = q.x; v[u][1] = q.y; v[u][2] = q.z; v[u][3] = q.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[u][j] = (px + j < p.n) ? p.data[px + j] : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      uint32_t c[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float x = v[u][j];
        // NaN -> matplotlib "bad" colour (transparent black); out-of-range values saturate to the
        // first / last slot (matplotlib under / over colours)
        const int k = p.unit_range ? lars_cmap_index(x) : lars_cmap_index_range(x, p.vmin, p.vmax);
        c[j] = (x != x) ? 0u : cm[k];
      }
      stage[warp][u][3 * lane + 0] = prmt(c[0], c[1], 0x4210);
      stage[warp][u][3 * lane + 1] = prmt(c[1], c[2], 0x5421);
      stage[warp][u][3 * lane + 2] = prmt(c[2], c[3], 0x6542);
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long g = g0 + u * stride;
      if (g >= ngroups) break;
      const long long wbase = g * 96;               // output word index of this group
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const long long w = wbase + r * 32 + lane;
        const uint32_t val = stage[warp][u][r * 32 + lane];
        if (w * 4 + 4 <= total_bytes) {
          out32[w] = val;
        } else {
          for (int b = 0; b < 4; ++b)
            if (w * 4 + b < total_bytes) p.rgb[w * 4 + b] = (uint8_t)(val >> (8 * b));
        }
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// K6: float64 NDVI of a raw uint8 frame (process-ndvi.py:18-31)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ndvi_f64_u8_kernel(const uint8_t* __restrict__ src, double* __restrict__ out,
                                                          long long n, int channels) {
  long long done = 0;
  if (channels == 3 && (reinterpret_cast<uintptr_t>(src) & 3u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
    // 4 pixels per thread: three aligned words in, two 16-byte stores out
    const long long ngroups = n / 4;
    const uint32_t* in = reinterpret_cast<const uint32_t*>(src);
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += (long long)gridDim.x * blockDim.x) {
      const uint32_t w0 = __ldg(in + 3 * g), w1 = __ldg(in + 3 * g + 1), w2 = __ldg(in + 3 * g + 2);
      // bytes: w0 = R0 G0 N0 R1, w1 = G1 N1 R2 G2, w2 = N2 R3 G3 N3
      const double x0 = lars_ratio_clip_f64((double)((w0 >> 16) & 255u), (double)(w0 & 255u));
      const double x1 = lars_ratio_clip_f64((double)((w1 >> 8) & 255u), (double)(w0 >> 24));
      const double x2 = lars_ratio_clip_f64((double)(w2 & 255u), (double)((w1 >> 16) & 255u));
      const double x3 = lars_ratio_clip_f64((double)(w2 >> 24), (double)((w2 >> 8) & 255u));
      double2* o = reinterpret_cast<double2*>(out + 4 * g);
      o[0] = make_double2(x0, x1);
      o[1] = make_double2(x2, x3);
    }
    done = ngroups * 4;
  }
  for (long long i = done + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const uint8_t* px = src + i * channels;
    out[i] = lars_ratio_clip_f64((double)px[2], (double)px[0]);
  }
}

// ------------------------------------------------------------------------------------------
// K7: normalized difference of two float32 planes (backend-process.py:28-38)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) index_planes_f32_kernel(const float* __restrict__ hi, const float* __restrict__ lo,
                                                               float* __restrict__ out, long long n) {
  const long long nvec = n / 4;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec;
       v += (long long)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(hi) + v);
    const float4 b = __ldg(reinterpret_cast<const float4*>(lo) + v);
    float4 r;
    r.x = lars_ratio_clip_f32(a.x, b.x); r.y = lars_ratio_clip_f32(a.y, b.y);
    r.z = lars_ratio_clip_f32(a.z, b.z); r.w = lars_ratio_clip_f32(a.w, b.w);
    reinterpret_cast<float4*>(out)[v] = r;
  }
  if (blockIdx.x == 0)
    for (long long i = nvec * 4 + threadIdx.x; i < n; i += blockDim.x) out[i] = lars_ratio_clip_f32(hi[i], lo[i]);
}

// ------------------------------------------------------------------------------------------
// Dataset-wide statistics: merge n_sets x 3 per-frame (or per-rank) records into 3, in order.
// This is the local half of the multi-GPU exchange: every rank merges its
// own frames, the packed records are all-gathered over NCCL, and the same kernel merges the
// per-rank records in rank order -- deterministic, no floating-point atomics.
// `out` may be one of the input sets (the running fold of a survey merges {fold, new} into fold): no
// __restrict__ on either pointer, every thread reads all it needs of a bin / field before that bin / field
// is written (own histogram bin per thread; the scalars are written by lane 0 after the warp's shuffles).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LARS_MAX_BINS) stats_merge_kernel(const lars_index_stats* in, int n_sets,
                                                                    lars_index_stats* out) {
  const int idx = blockIdx.x;   // index 0..2
  const int tid = threadIdx.x;  // histogram bin
  unsigned long long h = 0;
#pragma unroll 4
  for (int s = 0; s < n_sets; ++s) h += in[(long long)s * 3 + idx].hist[tid];
  out[idx].hist[tid] = h;
  if (tid < 32) {
    // lane l folds sets l, l + 32, ... in order, then a fixed butterfly: the summation tree depends only on
    // n_sets (reproducible), and the loads of a lane are independent (one thread walking all sets was a
    // chain of dependent loads)
    unsigned long long cnt = 0, above = 0;
    double sum = 0.0, sumsq = 0.0;
    float mn = INFINITY, mx = -INFINITY, thr = 0.f;
    uint32_t bins = 0;
    for (int s = tid; s < n_sets; s += 32) {
      const lars_index_stats& r = in[(long long)s * 3 + idx];
      const bool used = r.count != 0;
      cnt += r.count; above += r.count_above;
      sum += used ? r.sum : 0.0; sumsq += used ? r.sumsq : 0.0;
      mn = used ? fminf(mn, r.min) : mn; mx = used ? fmaxf(mx, r.max) : mx;
      thr = used ? r.threshold : thr; bins = used ? r.bins : bins;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
      above += __shfl_xor_sync(0xffffffffu, above, d);
      sum += __shfl_xor_sync(0xffffffffu, sum, d);
      sumsq += __shfl_xor_sync(0xffffffffu, sumsq, d);
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
      const float othr = __shfl_xor_sync(0xffffffffu, thr, d);
      const uint32_t obins = __shfl_xor_sync(0xffffffffu, bins, d);
      if (bins == 0) { thr = othr; bins = obins; }        // any used record carries the same threshold / bins
    }
    if (tid == 0) {
      const double n = (double)cnt;
      const double mean = cnt ? sum / n : 0.0;
      double var = cnt ? sumsq / n - mean * mean : 0.0;
      var = var > 0.0 ? var : 0.0;
      lars_index_stats& o = out[idx];
      o.count = cnt; o.count_above = above; o.sum = sum; o.sumsq = sumsq;
      o.mean = mean; o.std = sqrt(var);
      o.min = cnt ? mn : 0.f; o.max = cnt ? mx : 0.f;
      o.threshold = thr; o.bins = bins;
    }
  }
}

// ------------------------------------------------------------------------------------------
// K8: index map of an interleaved frame of any sample type (calculate_index on a frame that is
// not uint8, process-images.py:456-490: astype(float32), ratio, clip).  hi_c / lo_c select the
// channels: NDVI (2,0), GNDVI (2,1), NDWI (1,2).
// ------------------------------------------------------------------------------------------
template <typename T, int HI, int LO>
__global__ void __launch_bounds__(256) index_hwc_kernel(const T* __restrict__ src, float* __restrict__ out, long long n,
                                                        int channels) {
  constexpr int hi_c = HI, lo_c = LO;   // compile-time channels: sample positions resolve to fixed shifts
  long long done = 0;
  if (sizeof(T) == 2 && channels == 3 && (reinterpret_cast<uintptr_t>(src) & 7u) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
    // uint16 RGN: 4 pixels = 24 bytes = three 8-byte loads per thread, one float4 out
    const long long ngroups = n / 4;
    const uint2* in = reinterpret_cast<const uint2*>(src);
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += (long long)gridDim.x * blockDim.x) {
      const uint2 a = __ldg(in + 3 * g), b = __ldg(in + 3 * g + 1), c = __ldg(in + 3 * g + 2);
      const uint32_t w[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
      float x[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int sh = 3 * j + hi_c, sl = 3 * j + lo_c;           // sample indices 0..11
        const float hi = (float)((w[sh >> 1] >> (16 * (sh & 1))) & 0xFFFFu);
        const float lo = (float)((w[sl >> 1] >> (16 * (sl & 1))) & 0xFFFFu);
        x[j] = lars_ratio_clip_f32(hi, lo);
      }
      reinterpret_cast<float4*>(out)[g] = make_float4(x[0], x[1], x[2], x[3]);
    }
    done = ngroups * 4;
  }
  for (long long i = done + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const T* px = src + i * channels;
    out[i] = lars_ratio_clip_f32((float)px[hi_c], (float)px[lo_c]);
  }
}

// ------------------------------------------------------------------------------------------
// K9: change detection (process-images.py:908-923, :956): index of two white-balanced uint8
// frames, diff = late - early in float32, 'bwr' colormap over [-0.5, 0.5].
// ------------------------------------------------------------------------------------------
struct ChangeParams {
  const uint8_t* early;
  const uint8_t* late;
  float* early_map;     // optional
  float* late_map;      // optional
  float* diff;          // required
  uint8_t* rgb;         // optional [n][3]
  const uint32_t* cmap; // bwr, packed
  long long n;
  int channels, hi_c, lo_c;
  float vmin, vmax;
};

__device__ __forceinline__ uint32_t change_byte(const uint32_t w[3], int idx) {   // byte idx (0..11) of 3 words
  return (w[idx >> 2] >> (8 * (idx & 3))) & 255u;
}

template <int HI, int LO>
__global__ void __launch_bounds__(256) index_change_u8_kernel(const ChangeParams p) {
  __shared__ uint32_t cm[256];
  cm[threadIdx.x] = p.cmap[threadIdx.x];
  __syncthreads();
  long long done = 0;
  const bool aligned = ((reinterpret_cast<uintptr_t>(p.early) | reinterpret_cast<uintptr_t>(p.late) |
                         reinterpret_cast<uintptr_t>(p.rgb)) & 3u) == 0 &&
                       ((reinterpret_cast<uintptr_t>(p.diff) | reinterpret_cast<uintptr_t>(p.early_map) |
                         reinterpret_cast<uintptr_t>(p.late_map)) & 15u) == 0;
  if (p.channels == 3 && aligned) {
    // 4 pixels per thread: three aligned words per frame in, one float4 per map and three words of RGB out
    const long long ngroups = p.n / 4;
    const uint32_t* ein = reinterpret_cast<const uint32_t*>(p.early);
    const uint32_t* lin = reinterpret_cast<const uint32_t*>(p.late);
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += (long long)gridDim.x * blockDim.x) {
      const uint32_t e[3] = {__ldg(ein + 3 * g), __ldg(ein + 3 * g + 1), __ldg(ein + 3 * g + 2)};
      const uint32_t l[3] = {__ldg(lin + 3 * g), __ldg(lin + 3 * g + 1), __ldg(lin + 3 * g + 2)};
      float xe[4], xl[4], d[4];
      uint32_t c[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // HI / LO are compile-time, so the byte positions 3 j + channel are fixed shifts
        xe[j] = lars_ratio_pair_u8((int)change_byte(e, 3 * j + HI), (int)change_byte(e, 3 * j + LO));
        xl[j] = lars_ratio_pair_u8((int)change_byte(l, 3 * j + HI), (int)change_byte(l, 3 * j + LO));
        d[j] = LARS_FSUB(xl[j], xe[j]);
      }
      if (p.early_map) reinterpret_cast<float4*>(p.early_map)[g] = make_float4(xe[0], xe[1], xe[2], xe[3]);
      if (p.late_map) reinterpret_cast<float4*>(p.late_map)[g] = make_float4(xl[0], xl[1], xl[2], xl[3]);
      reinterpret_cast<float4*>(p.diff)[g] = make_float4(d[0], d[1], d[2], d[3]);
      if (p.rgb) {
#pragma unroll
        for (int j = 0; j < 4; ++j) c[j] = cm[lars_cmap_index_range(d[j], p.vmin, p.vmax)];
        uint32_t* o = reinterpret_cast<uint32_t*>(p.rgb) + 3 * g;
        o[0] = prmt(c[0], c[1], 0x4210);  // R0 G0 B0 R1
        o[1] = prmt(c[1], c[2], 0x5421);  // G1 B1 R2 G2
        o[2] = prmt(c[2], c[3], 0x6542);  // B2 R3 G3 B3
      }
    }
    done = ngroups * 4;
  }
  for (long long i = done + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
    const uint8_t* e = p.early + i * p.channels;
    const uint8_t* l = p.late + i * p.channels;
    const float xe = lars_ratio_pair_u8((int)e[HI], (int)e[LO]);
    const float xl = lars_ratio_pair_u8((int)l[HI], (int)l[LO]);
    const float d = LARS_FSUB(xl, xe);
    if (p.early_map) p.early_map[i] = xe;
    if (p.late_map) p.late_map[i] = xl;
    p.diff[i] = d;
    if (p.rgb) {
      const uint32_t c = cm[lars_cmap_index_range(d, p.vmin, p.vmax)];
      p.rgb[3 * i + 0] = (uint8_t)c; p.rgb[3 * i + 1] = (uint8_t)(c >> 8); p.rgb[3 * i + 2] = (uint8_t)(c >> 16);
    }
  }
}

}  // namespace lars
