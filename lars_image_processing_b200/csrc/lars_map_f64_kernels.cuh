// Statistics and exact order statistics of a FLOAT64 map -- the flavour process-ndvi.py works in:
// calculate_ndvi returns float64 (process-ndvi.py:18-31) and analyze_ndvi_statistics / generate_ndvi_report
// reduce that array as it is (:60-71, :97): min / max / median are float64 values, `ndvi > 0.2` compares in
// float64, and plt.hist bins against float64 edges.  Rounding the map to float32 first (round 1) moved the
// coverage count, the extremes and ~0.9 % of the histogram; these kernels keep the reference's arithmetic.
//
//   K4d  map_stats_f64     min / max / count / count > thr / sum / sum of squares (about the first element) /
//                          np.histogram with float64 linspace edges and the +-1 edge correction
//   K3d  select_f64        exact x[rank] by radix select on order-preserving 64-bit keys: 6 passes over
//                          11 + 11 + 11 + 11 + 11 + 9 bits (select_kernel<double>, lars_map_kernels.cuh)
//
// Both are HBM-bound streaming reads of 8 B per element; the float64 path is the report path of one script, so
// they follow the float32 kernels' structure without their tuning.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lars_b200.h"
#include "lars_map_kernels.cuh"
#include "pixel_math.h"

namespace lars {

struct __align__(16) MapPartialF64 {
  double sx, sd, sdd, mn, mx;
  unsigned long long count, above;
  uint32_t has_nan, pad_;
  uint32_t hist[MAP_HIST_ROWS];
};

struct MapStatsF64Params {
  const double* data;
  long long n;
  MapPartialF64* partials;   // [gridDim.x]
  double threshold;
  int bins;
};

__global__ void __launch_bounds__(MAP_THREADS) map_stats_f64_kernel(const MapStatsF64Params p) {
  extern __shared__ __align__(16) uint8_t msd_smem[];
  uint32_t* hist = reinterpret_cast<uint32_t*>(msd_smem);                                  // [bins][32]
  double* edges_s = reinterpret_cast<double*>(msd_smem + (size_t)p.bins * 32 * 4);        // [bins + 1]
  double* red = edges_s + ((p.bins + 1 + 1) & ~1);                                         // [8 warps][8]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < p.bins * 32; i += MAP_THREADS) hist[i] = 0u;
  {  // np.linspace(-1, 1, bins + 1): arange * step + start, last edge exact
    const double step = LARS_DDIV(2.0, (double)p.bins);
    for (int i = tid; i <= p.bins; i += MAP_THREADS)
      edges_s[i] = (i == p.bins) ? 1.0 : LARS_DADD(LARS_DMUL((double)i, step), -1.0);
  }
  __syncthreads();
  const double k = p.data[0];
  const double thr = p.threshold;
  double mn = INFINITY, mx = -INFINITY, sx = 0.0, sd = 0.0, sdd = 0.0;
  unsigned long long count = 0, above = 0;
  uint32_t nan = 0;
  uint32_t* hist_lane = hist + lane;
  auto visit = [&](double x) {
    mn = fmin(mn, x);
    mx = fmax(mx, x);
    above += (x > thr) ? 1ull : 0ull;
    nan |= (x != x) ? 1u : 0u;
    const double d = LARS_DSUB(x, k);
    sx += x;
    sd += d;
    sdd = fma(d, d, sdd);
    if (x >= -1.0 && x <= 1.0) atomicAdd(hist_lane + lars_hist_bin_edges_f64(x, edges_s, p.bins) * 32, 1u);
  };
  const long long nvec = p.n / 2;
  const long long per = (nvec + gridDim.x - 1) / gridDim.x;
  const long long v0 = (long long)blockIdx.x * per;
  const long long v1 = (v0 + per < nvec) ? v0 + per : nvec;
  const double2* xv = reinterpret_cast<const double2*>(p.data);
  long long v = v0 + tid;
  for (; v + 3ll * MAP_THREADS < v1; v += 4ll * MAP_THREADS) {
    double2 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = __ldg(xv + v + (long long)u * MAP_THREADS);
#pragma unroll
    for (int u = 0; u < 4; ++u) { visit(q[u].x); visit(q[u].y); }
    count += 8;
  }
  for (; v < v1; v += MAP_THREADS) {
    const double2 q = __ldg(xv + v);
    visit(q.x); visit(q.y);
    count += 2;
  }
  if (blockIdx.x == gridDim.x - 1 && (p.n & 1) && tid == 0) { visit(p.data[p.n - 1]); count += 1; }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, d);
    sd += __shfl_xor_sync(0xffffffffu, sd, d);
    sdd += __shfl_xor_sync(0xffffffffu, sdd, d);
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, d));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    above += __shfl_xor_sync(0xffffffffu, above, d);
    nan |= __shfl_xor_sync(0xffffffffu, nan, d);
    count += __shfl_xor_sync(0xffffffffu, count, d);
  }
  if (lane == 0) {
    double* r = red + warp * 8;
    r[0] = sx; r[1] = sd; r[2] = sdd; r[3] = mn; r[4] = mx;
    r[5] = (double)above; r[6] = (double)nan; r[7] = (double)count;     // counts < 2^53: exact
  }
  __syncthreads();
  MapPartialF64* rec = p.partials + blockIdx.x;
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
      a[5] += r[5]; a[6] += r[6]; a[7] += r[7];
    }
    rec->sx = a[0]; rec->sd = a[1]; rec->sdd = a[2]; rec->mn = a[3]; rec->mx = a[4];
    rec->above = (unsigned long long)a[5]; rec->has_nan = a[6] != 0.0 ? 1u : 0u; rec->pad_ = 0u;
    rec->count = (unsigned long long)a[7];
  }
}

// one CTA: fold the partials in a fixed order (reproducible), thread b owns histogram bin b
__global__ void __launch_bounds__(MAP_HIST_ROWS) map_stats_f64_finalize_kernel(const MapPartialF64* recs, int n_parts, int bins,
                                                                                double threshold, lars_map_record_f64* out) {
  const int tid = threadIdx.x;
  unsigned long long h = 0;
  if (tid < bins)
    for (int s = 0; s < n_parts; ++s) h += recs[s].hist[tid];
  out->hist[tid] = h;
  if (tid < 32) {
    double sx = 0.0, sd = 0.0, sdd = 0.0, mn = INFINITY, mx = -INFINITY;
    unsigned long long cnt = 0, above = 0;
    uint32_t nan = 0;
    for (int s = tid; s < n_parts; s += 32) {
      const MapPartialF64& r = recs[s];
      if (!r.count) continue;
      sx += r.sx; sd += r.sd; sdd += r.sdd;
      mn = fmin(mn, r.mn); mx = fmax(mx, r.mx);
      cnt += r.count; above += r.above; nan |= r.has_nan;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, d);
      sd += __shfl_xor_sync(0xffffffffu, sd, d);
      sdd += __shfl_xor_sync(0xffffffffu, sdd, d);
      mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, d));
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
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
      out->count = cnt; out->count_above = above;
      out->sum = sx; out->sumsq = cnt ? (var + mean * mean) * n : 0.0;
      out->mean = nan ? NAN : mean; out->std = nan ? NAN : sqrt(var);
      out->min = nan ? NAN : mn; out->max = nan ? NAN : mx;
      out->threshold = threshold; out->bins = (uint32_t)bins; out->has_nan = nan;
    }
  }
}

// K3d (exact order statistics of a float64 map) is select_kernel<double> in lars_map_kernels.cuh.

}  // namespace lars
