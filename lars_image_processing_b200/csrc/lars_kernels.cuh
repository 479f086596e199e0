// sm_100a kernels of the RGNir analysis path (u8 frames).
//
//   K1  wb_hist_u8_kernel      Pass 1: per-frame, per-channel 256-bin value histogram
//   K1b wb_lut_build_kernel    cumulative histogram -> NumPy "linear" percentiles -> stretch LUT
//   (K2, the fused Pass 2, lives in lars_fused_kernel.cuh)
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
#ifndef LARS_K1_LANES
#define LARS_K1_LANES 32          /* lane copies of every counter (32: bank == lane, never a conflict) */
#endif
#ifndef LARS_K1_CTAS
#define LARS_K1_CTAS 2
#endif
constexpr int K1_THREADS = 512;
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int K1_LANES = LARS_K1_LANES;
constexpr int K1_CTAS_PER_SM = LARS_K1_CTAS;
constexpr int K1_BIN_SHIFT = (K1_LANES == 32) ? 7 : (K1_LANES == 16 ? 6 : 5);   // log2(bytes per bin)
constexpr int K1_SMEM_BYTES = 3 * 256 * K1_LANES * 4;  // 96 KB -> 2 CTAs / SM
constexpr int K1_WARP_BYTES = 3 * 512;           // one warp iteration: 3 coalesced 512-byte rows
constexpr int K1_CTA_BYTES = K1_WARPS * K1_WARP_BYTES;

struct K1Params {
  const uint8_t* src;
  unsigned long long* hist;  // [frame][3][256]
  long long n_pixels;
  long long frame_stride;    // bytes
  long long units_per_frame; // work unit = K1_CTA_BYTES (C == 3) / 16 * K1_THREADS bytes (C == 4)
  long long total_units;
  long long hist_set_stride; // elements between frames' histogram sets: 768, or 0 = one shared set
  int n_frames;
  int keep_in_l2;            // batch small enough to stay L2-resident until Pass 2
};

template <int SHIFT>
__device__ __forceinline__ void k1_count_byte(uint32_t word, uint32_t lane_base) {
  // address = lane_base + byte * (lanes * 4)
  constexpr uint32_t MASK = 0xFFu << K1_BIN_SHIFT;
  const uint32_t off = (SHIFT < K1_BIN_SHIFT) ? ((word << (K1_BIN_SHIFT - SHIFT)) & MASK) : ((word >> (SHIFT - K1_BIN_SHIFT)) & MASK);
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
__global__ void __launch_bounds__(K1_THREADS, K1_CTAS_PER_SM) wb_hist_u8_kernel(const K1Params p) {
  extern __shared__ __align__(16) uint32_t k1_hist[];  // [3][256][K1_LANES]
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const long long frame_bytes = p.n_pixels * C;
  const int hl = lane & (K1_LANES - 1);
  const uint32_t a0 = smem_u32(k1_hist) + 4u * hl, a1 = a0 + 256u * K1_LANES * 4u, a2 = a1 + 256u * K1_LANES * 4u;
  const long long G = gridDim.x;
  long long u = ((long long)blockIdx.x * p.total_units) / G;
  const long long u_end = ((long long)(blockIdx.x + 1) * p.total_units) / G;
  // Pass 2 re-reads these bytes.  When the whole batch fits in L2 ask the cache to keep them
  // (evict-last): the fused pass then reads its input from L2 and DRAM only sees its writes.
  const uint64_t pol = p.keep_in_l2 ? l2_policy_evict_last() : l2_policy_evict_normal();
#define K1_LOAD(ptr) ldg_v4_policy((ptr), pol)

  while (u < u_end) {
    const long long frame = u / p.units_per_frame;
    const long long fu0 = frame * p.units_per_frame;
    const long long span_end = (fu0 + p.units_per_frame < u_end) ? fu0 + p.units_per_frame : u_end;
    const long long b0 = (u - fu0) * K1Unit<C>::BYTES;
    long long b1 = (span_end - fu0) * K1Unit<C>::BYTES;
    if (b1 > frame_bytes) b1 = frame_bytes;
    const uint8_t* fsrc = p.src + frame * p.frame_stride;
    u = span_end;

    for (int i = tid; i < 3 * 256 * K1_LANES; i += K1_THREADS) k1_hist[i] = 0u;
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
        const uint4 x0 = K1_LOAD(q0), x1 = K1_LOAD(q0 + 512), x2 = K1_LOAD(q0 + 1024);
        const uint4 y0 = K1_LOAD(q1), y1 = K1_LOAD(q1 + 512), y2 = K1_LOAD(q1 + 1024);
        k1_count_vec3<0>(x0, c0, c1, c2);
        k1_count_vec3<2>(x1, c0, c1, c2);
        k1_count_vec3<1>(x2, c0, c1, c2);
        k1_count_vec3<0>(y0, c0, c1, c2);
        k1_count_vec3<2>(y1, c0, c1, c2);
        k1_count_vec3<1>(y2, c0, c1, c2);
      }
      for (; off + K1_WARP_BYTES <= vec_end; off += K1_CTA_BYTES) {
        const uint8_t* q0 = fsrc + off + 16 * lane;
        const uint4 x0 = K1_LOAD(q0), x1 = K1_LOAD(q0 + 512), x2 = K1_LOAD(q0 + 1024);
        k1_count_vec3<0>(x0, c0, c1, c2);
        k1_count_vec3<2>(x1, c0, c1, c2);
        k1_count_vec3<1>(x2, c0, c1, c2);
      }
      // scalar tail (< 1536 bytes, only at the end of a frame)
      for (long long o = vec_end + tid; o < b1; o += K1_THREADS)
        atomicAdd(&k1_hist[((int)(o % 3) * 256 + (int)fsrc[o]) * K1_LANES + hl], 1u);
    } else {
      const long long vec_end = b0 + ((b1 - b0) / 16) * 16;  // frame_bytes is a multiple of 4
      for (long long off = b0 + 16ll * tid; off + 16 <= vec_end; off += 16ll * K1_THREADS)
        k1_count_vec4(K1_LOAD(fsrc + off), a0, a1, a2);
      for (long long o = vec_end + tid; o < b1; o += K1_THREADS) {
        const int ch = (int)(o & 3);
        if (ch < 3) atomicAdd(&k1_hist[(ch * 256 + (int)fsrc[o]) * K1_LANES + hl], 1u);
      }
    }
    __syncthreads();

    // fold the 32 lane copies; rotate the start so that concurrent threads hit distinct banks
    for (int b = tid; b < 3 * 256; b += K1_THREADS) {
      uint32_t sum = 0;
#pragma unroll 8
      for (int l = 0; l < K1_LANES; ++l) sum += k1_hist[b * K1_LANES + ((l + tid) & (K1_LANES - 1))];
      if (sum) atomicAdd(&p.hist[frame * p.hist_set_stride + b], (unsigned long long)sum);
    }
    __syncthreads();
  }
}

#undef K1_LOAD

// ------------------------------------------------------------------------------------------
// K1b: percentiles + stretch LUT (process-images.py:437-441), one CTA per (set, channel)
// ------------------------------------------------------------------------------------------
struct K1bParams {
  const unsigned long long* hist;  // [set][3][256]
  uint8_t* lut;                    // [set][3][256]
  double* pct;                     // [set][3][2] or nullptr
  double q_lo, q_hi;
  int chain;                       // LARS_WB_CHAIN_*: which reference expression the table restates
};

// Body of K1b for one (set, channel): `x` is this thread's count of value v = threadIdx.x; sc = set * 3 + channel.
__device__ __forceinline__ void wb_lut_build_body(unsigned long long x, const K1bParams& p, long long sc) {
  __shared__ unsigned long long cum[256];
  __shared__ unsigned long long warp_tot[8];
  __shared__ int order_stat[4];
  __shared__ double pcts[2];
  const int v = threadIdx.x;
  const int lane = v & 31, warp = v >> 5;
  const long long base = sc * 256;

  // inclusive scan of the 256 bins
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
    if (p.pct) p.pct[sc * 2 + v] = r;
  }
  __syncthreads();
  p.lut[base + v] = p.chain == 1 ? lars_wb_lut_entry_rgn((double)v, pcts[0], pcts[1])
                                 : lars_wb_lut_entry((double)v, pcts[0], pcts[1]);
}

__global__ void __launch_bounds__(256) wb_lut_build_u8_kernel(const K1bParams p) {
  wb_lut_build_body(p.hist[(long long)blockIdx.x * 256 + threadIdx.x], p, blockIdx.x);   // blockIdx.x = set * 3 + channel
}

// ------------------------------------------------------------------------------------------
// K1b fused with the exchange of a tile-sharded image (BASELINE config 4): the image-wide white-balance histogram is
// the SUM of the ranks' local histograms (process-images.py:435-438: the percentiles are global to the image).  The
// payload is 3 x 256 counters, so the exchange is pure latency; instead of an NCCL all-reduce between K1 and K1b
// (launch + ring / tree protocol, ~20-65 us observed) this kernel does it itself over NVLink peer memory:
//   every rank owns a symmetric buffer  slots[2][world][3][256] u64 + flags[world][3] u32  that all peers have mapped;
//   CTA ch (one per channel): thread v STOREs its local count into slots[epoch & 1][rank][ch][v] of EVERY peer,
//   fences at system scope, then threads 0..world-1 raise flags[rank][ch] = epoch on every peer (release) and wait for
//   flags[w][ch] >= epoch in their own buffer (acquire); then every rank sums the world slots in rank order
//   (deterministic, identical on all ranks) and goes on to the percentiles and the table.
// Two slot sets alternate by epoch parity: a rank can be at most one exchange ahead of a peer (its next exchange needs
// that peer's next flag), so it never overwrites counters a peer still reads.  The wait is bounded (~10 s): on timeout
// the kernel sets *status and uses what it has instead of hanging the GPU.
// ------------------------------------------------------------------------------------------
struct K1bPeerParams {
  K1bParams k1b;                         // hist: [1][3][256] local counts in, image-wide counts out; lut / pct: one set
  unsigned long long* const* peer_bufs;  // device array [world]: every rank's symmetric buffer as mapped here
  uint32_t* status;                      // [1] set to 1 on a timed-out wait
  int rank, world;
  uint32_t epoch;                        // 1, 2, 3, ... (flags start at 0)
};

__device__ __forceinline__ size_t peer_slot_index(int parity, int world, int src_rank, int ch, int v) {
  return (((size_t)parity * world + src_rank) * 3 + ch) * 256 + v;
}

__global__ void __launch_bounds__(256) wb_lut_build_u8_peers_kernel(const K1bPeerParams p) {
  const int ch = blockIdx.x, v = threadIdx.x;
  const int parity = (int)(p.epoch & 1u);
  const size_t flags_at = (size_t)2 * p.world * 768;                 // in u64 units: flags follow the slots
  unsigned long long* hist = const_cast<unsigned long long*>(p.k1b.hist);
  const unsigned long long mine = hist[ch * 256 + v];
  for (int w = 0; w < p.world; ++w) p.peer_bufs[w][peer_slot_index(parity, p.world, p.rank, ch, v)] = mine;
  __threadfence_system();
  __syncthreads();
  if (v < p.world) {
    uint32_t* remote = reinterpret_cast<uint32_t*>(p.peer_bufs[v] + flags_at) + p.rank * 3 + ch;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(p.epoch) : "memory");
    const uint32_t* local = reinterpret_cast<const uint32_t*>(p.peer_bufs[p.rank] + flags_at) + v * 3 + ch;
    const long long t0 = clock64();
    uint32_t seen;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(local) : "memory");
      if ((int32_t)(seen - p.epoch) >= 0) break;
      if (clock64() - t0 > 20000000000ll) { *p.status = 1u; break; }  // ~10 s at 2 GHz: report, do not hang
      __nanosleep(100);
    }
  }
  __syncthreads();
  unsigned long long total = 0;
  const unsigned long long* own = p.peer_bufs[p.rank];
  for (int w = 0; w < p.world; ++w) total += __ldcg(own + peer_slot_index(parity, p.world, w, ch, v));
  hist[ch * 256 + v] = total;                                         // the caller sees the image-wide histogram
  wb_lut_build_body(total, p.k1b, ch);
}

}  // namespace lars
