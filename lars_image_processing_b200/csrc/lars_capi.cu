// C ABI of liblars_b200.so (declared in include/lars_b200.h).  Host-side validation and
// launches only: no allocation, no host<->device copies of image data, no stream syncs.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "../../include/lars_b200.h"
#include "host_tables.h"
#include "tiff_host.h"
#include "png_host.h"
#include "lars_kernels.cuh"
#include "lars_fused_kernel.cuh"
#include "lars_map_kernels.cuh"
#include "lars_map_f64_kernels.cuh"
#include "lars_u16_kernels.cuh"
#include "lars_resize_kernels.cuh"
#include "lars_lzw_kernels.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define LARS_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(LARS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                    \
  } while (0)

#ifndef LARS_L2_KEEP_BYTES
#define LARS_L2_KEEP_BYTES (40ll << 20)   /* measured: ~one 36 MB frame survives between the passes */
#endif
constexpr int kMaxDevices = 64;
struct DeviceState {
  bool ready = false;
  bool inflate = false;       // the Deflate kernel got its 194 KB of shared memory
  int sm_count = 0;
  int select_ctas_per_sm = 1; // co-resident CTAs of the persistent select kernels (grid barrier)
  uint32_t* cmaps = nullptr;  // [3][256] packed RGB, device
};
DeviceState g_dev[kMaxDevices];
std::mutex g_mu;

int current_state(DeviceState** out) {
  int dev = -1;
  LARS_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return fail(LARS_ERR_INVALID, "device ordinal %d out of range", dev);
  if (!g_dev[dev].ready)
    return fail(LARS_ERR_NOT_INIT, "lars_init(%d) has not been called for the current device", dev);
  *out = &g_dev[dev];
  return LARS_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

namespace {
// One launch for all passes (cooperative: the CTAs meet at a grid barrier after every pass).
template <typename T>
int run_select(const char* who, DeviceState* st, const T* data, int64_t n, uint64_t rank_lo, uint64_t rank_hi, T* out3,
               void* workspace, size_t workspace_bytes, void* stream) {
  if (!data || !out3 || !workspace) return fail(LARS_ERR_INVALID, "%s: NULL pointer", who);
  if (n < 1 || rank_lo >= (uint64_t)n || rank_hi >= (uint64_t)n || rank_lo > rank_hi)
    return fail(LARS_ERR_INVALID, "%s: ranks out of range", who);
  if (!aligned16(data)) return fail(LARS_ERR_INVALID, "%s: data must be 16-byte aligned", who);
  if (workspace_bytes < sizeof(lars::SelectState<T>) || !aligned16(workspace))
    return fail(LARS_ERR_INVALID, "%s: workspace too small or misaligned", who);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  lars::SelectState<T>* state = static_cast<lars::SelectState<T>*>(workspace);
  LARS_CUDA(cudaMemsetAsync(state, 0, sizeof(lars::SelectState<T>), s));      // barrier counter + every pass's histogram
  long long want = (n / lars::SelectTraits<T>::PER_VEC + lars::SEL_THREADS - 1) / lars::SEL_THREADS;
  int grid = st->sm_count * st->select_ctas_per_sm;
  if (want < grid) grid = (int)(want > 0 ? want : 1);
  long long n_ll = n;
  unsigned long long r0 = rank_lo, r1 = rank_hi;
  void* args[] = {(void*)&data, (void*)&n_ll, (void*)&state, (void*)&r0, (void*)&r1};
  LARS_CUDA(cudaLaunchCooperativeKernel((const void*)lars::select_kernel<T>, dim3(grid), dim3(lars::SEL_THREADS), args,
                                        (size_t)lars::SEL_SMEM_BYTES, s));
  LARS_CUDA(cudaMemcpyAsync(out3, reinterpret_cast<const char*>(state) + offsetof(lars::SelectState<T>, value),
                            3 * sizeof(T), cudaMemcpyDeviceToDevice, s));
  return LARS_OK;
}
}  // namespace

extern "C" {

const char* lars_last_error(void) { return g_err; }
int lars_abi_version(void) { return 7; }   // 7: float64 map statistics / select, lars_wb_lut_build_u8_chain, device PNG + LZW ring variant removed; 3: + resize, TIFF ingest; 4: lars_tiff_info grew (tiles, predictor, BigTIFF), lars_tiff_read_region; 5: PNG reader, device-side LZW; 6: experimental device-side Deflate

int lars_init(int device) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (device < 0 || device >= kMaxDevices) return fail(LARS_ERR_INVALID, "bad device ordinal %d", device);
  int count = 0;
  LARS_CUDA(cudaGetDeviceCount(&count));
  if (device >= count) return fail(LARS_ERR_INVALID, "device %d not present (%d visible)", device, count);
  if (g_dev[device].ready) return LARS_OK;
  int prev = -1;
  LARS_CUDA(cudaGetDevice(&prev));
  LARS_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  LARS_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    cudaSetDevice(prev);
    return fail(LARS_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library only contains sm_100a code (B200)",
                device, prop.major, prop.minor);
  }
  DeviceState& st = g_dev[device];
  st.sm_count = prop.multiProcessorCount;

  LARS_CUDA(cudaFuncSetAttribute(lars::wb_hist_u8_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 lars::K1_SMEM_BYTES));
  LARS_CUDA(cudaFuncSetAttribute(lars::wb_hist_u8_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 lars::K1_SMEM_BYTES));
  LARS_CUDA(cudaFuncSetAttribute(lars::fused_index_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 lars::K2Smem<3, 1>::TOTAL));
  LARS_CUDA(cudaFuncSetAttribute(lars::fused_index_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 lars::K2Smem<4, 1>::TOTAL));
  LARS_CUDA(cudaFuncSetAttribute(lars::fused_index_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 lars::K2Smem<3, 2>::TOTAL));
  LARS_CUDA(cudaFuncSetAttribute(lars::fused_index_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 lars::K2Smem<4, 2>::TOTAL));
  LARS_CUDA(cudaFuncSetAttribute(lars::wb_hist_u16_hi_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, lars::U16_HI_SMEM_BYTES));
  LARS_CUDA(cudaFuncSetAttribute(lars::wb_hist_u16_hi_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lars::U16_HI_SMEM_BYTES));
  LARS_CUDA(cudaFuncSetAttribute(lars::wb_hist_u16_sample_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, lars::U16_HI_SMEM_BYTES));
  LARS_CUDA(cudaFuncSetAttribute(lars::wb_hist_u16_sample_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lars::U16_HI_SMEM_BYTES));
  LARS_CUDA(cudaFuncSetAttribute(lars::wb_hist_u16_guided_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, lars::U16_GUIDED_SMEM_BYTES));
  LARS_CUDA(cudaFuncSetAttribute(lars::wb_hist_u16_guided_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lars::U16_GUIDED_SMEM_BYTES));
  LARS_CUDA(cudaFuncSetAttribute(lars::wb_hist_u16_lo_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, lars::U16_LO_SMEM_BYTES));
  LARS_CUDA(cudaFuncSetAttribute(lars::wb_hist_u16_lo_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lars::U16_LO_SMEM_BYTES));
  LARS_CUDA(cudaFuncSetAttribute(lars::select_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 lars::SEL_SMEM_BYTES));
  LARS_CUDA(cudaFuncSetAttribute(lars::select_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 lars::SEL_SMEM_BYTES));
  {  // the select kernels meet at a grid barrier: how many of their CTAs are resident per SM
    int a = 0, b = 0;
    LARS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, lars::select_kernel<float>, lars::SEL_THREADS, lars::SEL_SMEM_BYTES));
    LARS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, lars::select_kernel<double>, lars::SEL_THREADS, lars::SEL_SMEM_BYTES));
    st.select_ctas_per_sm = a < b ? a : b;
    if (st.select_ctas_per_sm < 1) return fail(LARS_ERR_CUDA, "select kernels do not fit on an SM");
    if (st.select_ctas_per_sm > 2) st.select_ctas_per_sm = 2;
  }
  LARS_CUDA(cudaFuncSetAttribute(lars::lzw_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 lars::LZW_SMEM_BYTES));
  st.inflate = cudaFuncSetAttribute(lars::inflate_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    lars::INF_SMEM_BYTES) == cudaSuccess;
  if (!st.inflate) cudaGetLastError();

  // colormap tables: 3 x 256 packed R | G << 8 | B << 16
  static uint32_t packed[3 * 256];
  for (int c = 0; c < 3; ++c) {
    uint8_t t[256][3];
    lars_host::build_colormap(c, t);
    for (int i = 0; i < 256; ++i)
      packed[c * 256 + i] = (uint32_t)t[i][0] | ((uint32_t)t[i][1] << 8) | ((uint32_t)t[i][2] << 16);
  }
  LARS_CUDA(cudaMalloc(&st.cmaps, sizeof(packed)));  // 3 KB of constants, lives until lars_shutdown
  LARS_CUDA(cudaMemcpy(st.cmaps, packed, sizeof(packed), cudaMemcpyHostToDevice));
  st.ready = true;
  if (prev >= 0 && prev != device) LARS_CUDA(cudaSetDevice(prev));
  return LARS_OK;
}

int lars_shutdown(void) {
  std::lock_guard<std::mutex> lock(g_mu);
  for (int d = 0; d < kMaxDevices; ++d) {
    if (g_dev[d].ready) {
      int prev = -1;
      cudaGetDevice(&prev);
      cudaSetDevice(d);
      cudaFree(g_dev[d].cmaps);
      if (prev >= 0) cudaSetDevice(prev);
      g_dev[d] = DeviceState();
    }
  }
  return LARS_OK;
}

int lars_sm_count(void) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  return st->sm_count;
}

int lars_colormap_table(int cmap_id, uint8_t* rgb_out) {
  if (!rgb_out) return fail(LARS_ERR_INVALID, "rgb_out is NULL");
  uint8_t t[256][3];
  if (!lars_host::build_colormap(cmap_id, t)) return fail(LARS_ERR_INVALID, "unknown colormap id %d", cmap_id);
  memcpy(rgb_out, t, sizeof(t));
  return LARS_OK;
}

int lars_histogram_edges_f32(int bins, float* edges_out) {
  if (!edges_out || bins < 1) return fail(LARS_ERR_INVALID, "bad histogram edge request (bins=%d)", bins);
  lars_host::histogram_edges_f32(bins, edges_out);
  return LARS_OK;
}

// ------------------------------------------------------------------------------------------
// Pass 1
// ------------------------------------------------------------------------------------------
int lars_wb_hist_u8(const uint8_t* src, int32_t n_frames, int64_t n_pixels, int32_t channels,
                    int64_t src_frame_stride, uint64_t* hist, int32_t shared_hist, void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!src || !hist) return fail(LARS_ERR_INVALID, "lars_wb_hist_u8: NULL pointer");
  if (n_frames < 1 || n_pixels < 1) return fail(LARS_ERR_INVALID, "lars_wb_hist_u8: empty input (frames=%d, pixels=%lld)", n_frames, (long long)n_pixels);
  if (channels != 3 && channels != 4) return fail(LARS_ERR_UNSUPPORTED, "lars_wb_hist_u8: channels must be 3 or 4, got %d", channels);
  if (!aligned16(src) || (src_frame_stride & 15) || src_frame_stride < n_pixels * channels)
    return fail(LARS_ERR_INVALID, "lars_wb_hist_u8: src / frame stride must be 16-byte aligned and cover a frame");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  LARS_CUDA(cudaMemsetAsync(hist, 0, sizeof(uint64_t) * 768 * (size_t)(shared_hist ? 1 : n_frames), s));

  const long long frame_bytes = (long long)n_pixels * channels;
  const long long unit = (channels == 3) ? lars::K1Unit<3>::BYTES : lars::K1Unit<4>::BYTES;
  lars::K1Params p;
  p.src = src;
  p.hist = reinterpret_cast<unsigned long long*>(hist);
  p.n_pixels = n_pixels;
  p.frame_stride = src_frame_stride;
  p.units_per_frame = (frame_bytes + unit - 1) / unit;
  p.total_units = p.units_per_frame * n_frames;
  p.n_frames = n_frames;
  p.hist_set_stride = shared_hist ? 0 : 768;
  p.keep_in_l2 = (frame_bytes * n_frames <= (long long)LARS_L2_KEEP_BYTES) ? 1 : 0;
  const long long target_ctas = (long long)lars::K1_CTAS_PER_SM * st->sm_count;  // resident CTAs per SM
  const int grid = (int)(p.total_units < target_ctas ? p.total_units : target_ctas);
  if (channels == 3)
    lars::wb_hist_u8_kernel<3><<<grid, lars::K1_THREADS, lars::K1_SMEM_BYTES, s>>>(p);
  else
    lars::wb_hist_u8_kernel<4><<<grid, lars::K1_THREADS, lars::K1_SMEM_BYTES, s>>>(p);
  LARS_CUDA(cudaGetLastError());
  return LARS_OK;
}

int lars_wb_lut_build_u8(const uint64_t* hist, int32_t n_sets, double q_lo, double q_hi, uint8_t* lut,
                         double* pct, void* stream) {
  return lars_wb_lut_build_u8_chain(hist, n_sets, q_lo, q_hi, LARS_WB_CHAIN_IMAGES, lut, pct, stream);
}

int lars_wb_lut_build_u8_chain(const uint64_t* hist, int32_t n_sets, double q_lo, double q_hi, int32_t chain,
                               uint8_t* lut, double* pct, void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!hist || !lut) return fail(LARS_ERR_INVALID, "lars_wb_lut_build_u8: NULL pointer");
  if (n_sets < 1) return fail(LARS_ERR_INVALID, "lars_wb_lut_build_u8: n_sets=%d", n_sets);
  if (!(q_lo >= 0.0 && q_lo <= 1.0 && q_hi >= 0.0 && q_hi <= 1.0))
    return fail(LARS_ERR_INVALID, "lars_wb_lut_build_u8: quantiles must be fractions in [0,1]");
  if (chain != LARS_WB_CHAIN_IMAGES && chain != LARS_WB_CHAIN_RGN)
    return fail(LARS_ERR_INVALID, "lars_wb_lut_build_u8_chain: chain=%d", chain);
  lars::K1bParams p;
  p.chain = chain;
  p.hist = reinterpret_cast<const unsigned long long*>(hist);
  p.lut = lut;
  p.pct = pct;
  p.q_lo = q_lo;
  p.q_hi = q_hi;
  lars::wb_lut_build_u8_kernel<<<n_sets * 3, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  LARS_CUDA(cudaGetLastError());
  return LARS_OK;
}

size_t lars_wb_peer_buffer_bytes(int32_t world) {
  if (world < 1) return 0;
  return ((size_t)2 * world * 768 * 8 + (size_t)world * 3 * 4 + 255) & ~(size_t)255;
}

int lars_wb_lut_build_u8_peers(uint64_t* hist, uint64_t* const* peer_bufs, int32_t rank, int32_t world,
                               uint32_t epoch, double q_lo, double q_hi, int32_t chain, uint8_t* lut, double* pct,
                               uint32_t* status, void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!hist || !peer_bufs || !lut || !status) return fail(LARS_ERR_INVALID, "lars_wb_lut_build_u8_peers: NULL pointer");
  if (world < 1 || world > 256 || rank < 0 || rank >= world)
    return fail(LARS_ERR_INVALID, "lars_wb_lut_build_u8_peers: rank %d of %d", rank, world);
  if (epoch == 0) return fail(LARS_ERR_INVALID, "lars_wb_lut_build_u8_peers: epochs start at 1 (flags start at 0)");
  if (!(q_lo >= 0.0 && q_lo <= 1.0 && q_hi >= 0.0 && q_hi <= 1.0))
    return fail(LARS_ERR_INVALID, "lars_wb_lut_build_u8_peers: quantiles must be fractions in [0,1]");
  if (chain != LARS_WB_CHAIN_IMAGES && chain != LARS_WB_CHAIN_RGN)
    return fail(LARS_ERR_INVALID, "lars_wb_lut_build_u8_peers: chain=%d", chain);
  lars::K1bPeerParams p;
  p.k1b.hist = reinterpret_cast<const unsigned long long*>(hist);
  p.k1b.lut = lut; p.k1b.pct = pct; p.k1b.q_lo = q_lo; p.k1b.q_hi = q_hi; p.k1b.chain = chain;
  p.peer_bufs = reinterpret_cast<unsigned long long* const*>(peer_bufs);
  p.status = status; p.rank = rank; p.world = world; p.epoch = epoch;
  lars::wb_lut_build_u8_peers_kernel<<<3, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  LARS_CUDA(cudaGetLastError());
  return LARS_OK;
}

// ------------------------------------------------------------------------------------------
// Pass 2
// ------------------------------------------------------------------------------------------
static int fused_grid(int sm_count) {
  int g = sm_count * lars::K2_CTAS_PER_SM;
  return g > lars::K2_MAX_GRID ? lars::K2_MAX_GRID : g;
}

size_t lars_fused_workspace_bytes(int32_t n_frames) {
  if (n_frames < 1) return 0;
  const size_t slots = (size_t)lars::K2_MAX_GRID / (size_t)n_frames + 2;
  return slots * (size_t)n_frames * sizeof(lars::K2Partial);
}

static int fused_index_impl(const lars_fused_args* a, void* stream, int BPS);
int lars_fused_index_u8(const lars_fused_args* a, void* stream) { return fused_index_impl(a, stream, 1); }
int lars_fused_index_u16(const lars_fused_args* a, void* stream) { return fused_index_impl(a, stream, 2); }

static int fused_index_impl(const lars_fused_args* a, void* stream, int BPS) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!a) return fail(LARS_ERR_INVALID, "lars_fused_index_u8: args is NULL");
  if (a->struct_bytes != sizeof(lars_fused_args))
    return fail(LARS_ERR_INVALID, "lars_fused_index_u8: struct_bytes=%u, library expects %zu", a->struct_bytes,
                sizeof(lars_fused_args));
  if (!a->src) return fail(LARS_ERR_INVALID, "lars_fused_index_u8: src is NULL");
  if (a->n_frames < 1 || a->n_pixels < 1)
    return fail(LARS_ERR_INVALID, "lars_fused_index_u8: empty input (frames=%d, pixels=%lld)", a->n_frames, (long long)a->n_pixels);
  if (a->channels != 3 && a->channels != 4)
    return fail(LARS_ERR_UNSUPPORTED, "lars_fused_index_u8: channels must be 3 or 4, got %d", a->channels);
  if (a->bins < 1 || a->bins > LARS_MAX_BINS)
    return fail(LARS_ERR_INVALID, "lars_fused_index_u8: bins must be in 1..%d, got %d", LARS_MAX_BINS, a->bins);
  const int C = a->channels;
  const int64_t padded_px = (a->n_pixels + LARS_PIXEL_GROUP - 1) / LARS_PIXEL_GROUP * LARS_PIXEL_GROUP;
  if (!aligned16(a->src) || (a->src_frame_stride & 15) || a->src_frame_stride < padded_px * C * BPS)
    return fail(LARS_ERR_INVALID, "lars_fused_index_u8: src must be 16-byte aligned with a 16-pixel padded frame stride");
  if (BPS == 2 && !a->wb_lut)
    return fail(LARS_ERR_INVALID, "lars_fused_index_u16: wb_lut (lars_stretch_u16[frame][3]) is required");
  if (BPS == 2 && (!aligned16(a->wb_lut) || (a->lut_frame_stride != 0 && a->lut_frame_stride < (int64_t)(3 * sizeof(lars_stretch_u16)))))
    return fail(LARS_ERR_INVALID, "lars_fused_index_u16: stretch tables must be 16-byte aligned, stride 0 or >= %zu", 3 * sizeof(lars_stretch_u16));
  if (a->wb_out && (!aligned16(a->wb_out) || (a->wb_frame_stride & 15) || a->wb_frame_stride < padded_px * C))
    return fail(LARS_ERR_INVALID, "lars_fused_index_u8: wb_out alignment / stride");
  for (int i = 0; i < 3; ++i) {
    if (a->maps[i] && (!aligned16(a->maps[i]) || (a->map_frame_stride & 3) || a->map_frame_stride < padded_px))
      return fail(LARS_ERR_INVALID, "lars_fused_index_u8: maps[%d] alignment / stride", i);
    if (a->rgb[i] && (!aligned16(a->rgb[i]) || (a->rgb_frame_stride & 15) || a->rgb_frame_stride < padded_px * 3))
      return fail(LARS_ERR_INVALID, "lars_fused_index_u8: rgb[%d] alignment / stride", i);
    if (a->cmap[i] < 0 || a->cmap[i] > 2) return fail(LARS_ERR_INVALID, "lars_fused_index_u8: cmap[%d]=%d", i, a->cmap[i]);
  }
  if (a->wb_lut && a->lut_frame_stride != 0 && a->lut_frame_stride < 768)
    return fail(LARS_ERR_INVALID, "lars_fused_index_u8: lut_frame_stride must be 0 or >= 768");
  if (a->stats) {
    if (!a->workspace || !aligned16(a->workspace) || a->workspace_bytes < lars_fused_workspace_bytes(a->n_frames))
      return fail(LARS_ERR_INVALID, "lars_fused_index_u8: workspace too small (%zu < %zu) or misaligned",
                  a->workspace_bytes, lars_fused_workspace_bytes(a->n_frames));
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  lars::K2Params p;
  memset(&p, 0, sizeof(p));
  p.src = a->src;
  p.wb_lut = a->wb_lut;
  p.wb_out = a->wb_out;
  for (int i = 0; i < 3; ++i) {
    p.maps[i] = a->maps[i];
    p.rgb[i] = a->rgb[i];
    p.cmap_id[i] = a->cmap[i];
    p.thresholds[i] = a->thresholds[i];
  }
  p.cmaps = st->cmaps;
  p.n_pixels = a->n_pixels;
  p.src_frame_stride = a->src_frame_stride;
  p.lut_frame_stride = a->lut_frame_stride;
  p.wb_frame_stride = a->wb_frame_stride;
  p.map_frame_stride = a->map_frame_stride;
  p.rgb_frame_stride = a->rgb_frame_stride;
  const int tile_px = lars::k2_tile_px(BPS);
  p.tiles_per_frame = (a->n_pixels + tile_px - 1) / tile_px;
  p.total_tiles = p.tiles_per_frame * a->n_frames;
  p.n_frames = a->n_frames;
  p.bins = a->bins;
  long long grid = fused_grid(st->sm_count);
  if (grid > p.total_tiles) grid = p.total_tiles;
  p.slots_per_frame = (int)(grid / a->n_frames + 2);
  p.partials = nullptr;
  if (a->stats) {
    p.partials = static_cast<lars::K2Partial*>(a->workspace);
    LARS_CUDA(cudaMemsetAsync(a->workspace, 0,
                              (size_t)p.slots_per_frame * a->n_frames * sizeof(lars::K2Partial), s));
  }
  if (BPS == 1 && C == 3)
    lars::fused_index_kernel<3, 1><<<(int)grid, lars::K2_THREADS, lars::K2Smem<3, 1>::TOTAL, s>>>(p);
  else if (BPS == 1)
    lars::fused_index_kernel<4, 1><<<(int)grid, lars::K2_THREADS, lars::K2Smem<4, 1>::TOTAL, s>>>(p);
  else if (C == 3)
    lars::fused_index_kernel<3, 2><<<(int)grid, lars::K2_THREADS, lars::K2Smem<3, 2>::TOTAL, s>>>(p);
  else
    lars::fused_index_kernel<4, 2><<<(int)grid, lars::K2_THREADS, lars::K2Smem<4, 2>::TOTAL, s>>>(p);
  LARS_CUDA(cudaGetLastError());
  if (a->stats) {
    lars::K2fParams f;
    f.partials = p.partials;
    f.stats = a->stats;
    f.slots_per_frame = p.slots_per_frame;
    f.bins = a->bins;
    for (int i = 0; i < 3; ++i) f.thresholds[i] = a->thresholds[i];
    lars::fused_finalize_kernel<<<a->n_frames, 3 * lars::K2_BINS_PAD * lars::K2F_SPLIT, 0, s>>>(f);
    LARS_CUDA(cudaGetLastError());
  }
  return LARS_OK;
}

// ------------------------------------------------------------------------------------------
// float-map operations
// ------------------------------------------------------------------------------------------
static int map_parts(int sm_count) { return sm_count * 4; }

size_t lars_map_stats_workspace_bytes(int32_t n_maps) {
  if (n_maps < 1) return 0;
  return (size_t)n_maps * (size_t)(148 * 4) * sizeof(lars::MapPartial) + lars::MAP_SUBBIN_BYTES;
}

int lars_map_stats_f32(const float* data, int32_t n_maps, int64_t n, int64_t stride, int32_t bins,
                       float threshold, lars_index_stats* stats, void* workspace, size_t workspace_bytes,
                       void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!data || !stats || !workspace) return fail(LARS_ERR_INVALID, "lars_map_stats_f32: NULL pointer");
  if (n_maps < 1 || n < 1) return fail(LARS_ERR_INVALID, "lars_map_stats_f32: empty input");
  if (bins < 1 || bins > LARS_MAX_BINS) return fail(LARS_ERR_INVALID, "lars_map_stats_f32: bins must be in 1..%d", LARS_MAX_BINS);
  if (!aligned16(data) || (n_maps > 1 && (stride & 3))) return fail(LARS_ERR_INVALID, "lars_map_stats_f32: rows must be 16-byte aligned");
  // about 4 CTAs per SM over the whole batch: a CTA's prologue (counters, table copy) and epilogue (32-copy
  // histogram fold) cost as much as ~20,000 elements, and round 1 launched 592 CTAs per map
  int parts = map_parts(st->sm_count) / n_maps;
  if (parts < 8) parts = 8;
  if (parts > 148 * 4) parts = 148 * 4;
  const long long nvec = n / 4;
  if (parts > nvec) parts = (int)(nvec > 0 ? nvec : 1);
  const size_t partial_bytes = (size_t)n_maps * (size_t)(148 * 4) * sizeof(lars::MapPartial);
  if (workspace_bytes < partial_bytes + lars::MAP_SUBBIN_BYTES || !aligned16(workspace))
    return fail(LARS_ERR_INVALID, "lars_map_stats_f32: workspace too small or misaligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* subbin = static_cast<uint8_t*>(workspace) + partial_bytes;
  lars::map_subbin_table_kernel<<<lars::MAP_SUBBIN_BYTES / 256, 256, 0, s>>>(subbin, bins);
  LARS_CUDA(cudaGetLastError());
  lars::MapStatsParams p;
  p.data = data; p.n = n; p.stride = stride;
  p.partials = static_cast<lars::MapPartial*>(workspace);
  p.subbin = subbin;
  p.threshold = threshold; p.bins = bins;
  const size_t smem = lars::MAP_SUBBIN_BYTES + (size_t)(bins + 1) * 128 + (size_t)((bins + 1 + 3) / 4) * 16 + 8 * 8 * 8;
  lars::map_stats_f32_kernel<<<dim3(parts, n_maps), lars::MAP_THREADS, smem, s>>>(p);
  LARS_CUDA(cudaGetLastError());
  lars::MapFinalizeParams f;
  f.partials = p.partials; f.data = data; f.stride = stride; f.stats = stats;
  f.n_parts = parts; f.bins = bins; f.threshold = threshold;
  lars::map_stats_finalize_kernel<<<n_maps, lars::MAP_HIST_ROWS * lars::MAP_FIN_SPLIT, 0, s>>>(f);
  LARS_CUDA(cudaGetLastError());
  return LARS_OK;
}


size_t lars_select_workspace_bytes(void) { return sizeof(lars::SelectState<float>); }

int lars_select_f32(const float* data, int64_t n, uint64_t rank_lo, uint64_t rank_hi, float* out3,
                    void* workspace, size_t workspace_bytes, void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  return run_select<float>("lars_select_f32", st, data, n, rank_lo, rank_hi, out3, workspace, workspace_bytes, stream);
}

size_t lars_map_stats_f64_workspace_bytes(void) { return (size_t)(148 * 4) * sizeof(lars::MapPartialF64); }

int lars_map_stats_f64(const double* data, int64_t n, int32_t bins, double threshold, lars_map_record_f64* stats,
                       void* workspace, size_t workspace_bytes, void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!data || !stats || !workspace) return fail(LARS_ERR_INVALID, "lars_map_stats_f64: NULL pointer");
  if (n < 1) return fail(LARS_ERR_INVALID, "lars_map_stats_f64: empty input");
  if (bins < 1 || bins > LARS_MAX_BINS) return fail(LARS_ERR_INVALID, "lars_map_stats_f64: bins must be in 1..%d", LARS_MAX_BINS);
  if (!aligned16(data)) return fail(LARS_ERR_INVALID, "lars_map_stats_f64: data must be 16-byte aligned");
  int parts = map_parts(st->sm_count);
  if (parts > 148 * 4) parts = 148 * 4;
  const long long nvec = n / 2;
  if (parts > nvec) parts = (int)(nvec > 0 ? nvec : 1);
  if (workspace_bytes < (size_t)parts * sizeof(lars::MapPartialF64) || !aligned16(workspace))
    return fail(LARS_ERR_INVALID, "lars_map_stats_f64: workspace too small or misaligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  lars::MapStatsF64Params p;
  p.data = data; p.n = n; p.partials = static_cast<lars::MapPartialF64*>(workspace);
  p.threshold = threshold; p.bins = bins;
  const size_t smem = (size_t)bins * 128 + (size_t)((bins + 2) & ~1) * 8 + 8 * 8 * 8;
  lars::map_stats_f64_kernel<<<parts, lars::MAP_THREADS, smem, s>>>(p);
  LARS_CUDA(cudaGetLastError());
  lars::map_stats_f64_finalize_kernel<<<1, lars::MAP_HIST_ROWS, 0, s>>>(p.partials, parts, bins, threshold, stats);
  LARS_CUDA(cudaGetLastError());
  return LARS_OK;
}

size_t lars_select_f64_workspace_bytes(void) { return sizeof(lars::SelectState<double>); }

int lars_select_f64(const double* data, int64_t n, uint64_t rank_lo, uint64_t rank_hi, double* out3,
                    void* workspace, size_t workspace_bytes, void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  return run_select<double>("lars_select_f64", st, data, n, rank_lo, rank_hi, out3, workspace, workspace_bytes, stream);
}

int lars_colormap_f32(const float* data, int64_t n, int32_t cmap_id, float vmin, float vmax, uint8_t* rgb,
                      void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!data || !rgb) return fail(LARS_ERR_INVALID, "lars_colormap_f32: NULL pointer");
  if (n < 1) return fail(LARS_ERR_INVALID, "lars_colormap_f32: empty input");
  if (cmap_id < 0 || cmap_id > 2) return fail(LARS_ERR_INVALID, "lars_colormap_f32: unknown colormap %d", cmap_id);
  if (!(vmax > vmin)) return fail(LARS_ERR_INVALID, "lars_colormap_f32: vmax must exceed vmin");
  if (!aligned16(data) || (reinterpret_cast<uintptr_t>(rgb) & 3))
    return fail(LARS_ERR_INVALID, "lars_colormap_f32: data must be 16-byte and rgb 4-byte aligned");
  lars::ColormapParams p;
  p.data = data; p.rgb = rgb; p.cmap = st->cmaps + cmap_id * 256; p.n = n;
  p.vmin = vmin; p.vmax = vmax; p.unit_range = (vmin == -1.0f && vmax == 1.0f) ? 1 : 0;
  long long groups = (n + 127) / 128;
  long long want = (groups + 7) / 8;
  int grid = st->sm_count * 8;
  if (want < grid) grid = (int)want;
  lars::colormap_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  LARS_CUDA(cudaGetLastError());
  return LARS_OK;
}

int lars_ndvi_f64_u8(const uint8_t* src, int64_t n_pixels, int32_t channels, double* out, void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!src || !out) return fail(LARS_ERR_INVALID, "lars_ndvi_f64_u8: NULL pointer");
  if (n_pixels < 1 || channels < 3) return fail(LARS_ERR_INVALID, "lars_ndvi_f64_u8: need >= 1 pixel and >= 3 channels");
  long long want = (n_pixels + 255) / 256;
  int grid = st->sm_count * 8;
  if (want < grid) grid = (int)want;
  lars::ndvi_f64_u8_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, out, n_pixels, channels);
  LARS_CUDA(cudaGetLastError());
  return LARS_OK;
}

int lars_index_planes_f32(const float* hi, const float* lo, int64_t n, float* out, void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!hi || !lo || !out) return fail(LARS_ERR_INVALID, "lars_index_planes_f32: NULL pointer");
  if (n < 1) return fail(LARS_ERR_INVALID, "lars_index_planes_f32: empty input");
  if (!aligned16(hi) || !aligned16(lo) || !aligned16(out))
    return fail(LARS_ERR_INVALID, "lars_index_planes_f32: planes must be 16-byte aligned");
  long long want = (n / 4 + 255) / 256;
  int grid = st->sm_count * 8;
  if (want < grid) grid = (int)(want > 0 ? want : 1);
  lars::index_planes_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(hi, lo, out, n);
  LARS_CUDA(cudaGetLastError());
  return LARS_OK;
}

int lars_stats_merge(const lars_index_stats* in, int32_t n_sets, lars_index_stats* out, void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!in || !out) return fail(LARS_ERR_INVALID, "lars_stats_merge: NULL pointer");
  if (n_sets < 1) return fail(LARS_ERR_INVALID, "lars_stats_merge: n_sets=%d", n_sets);
  lars::stats_merge_kernel<<<3, LARS_MAX_BINS, 0, static_cast<cudaStream_t>(stream)>>>(in, n_sets, out);
  LARS_CUDA(cudaGetLastError());
  return LARS_OK;
}

// ------------------------------------------------------------------------------------------
// uint16 Pass 1
// ------------------------------------------------------------------------------------------
namespace {
struct U16Workspace {
  unsigned long long* hist_hi;
  unsigned long long* hist_lo;
  lars::U16Select* select;
};
// per set: high-byte histogram, level-B low-byte histograms, selection records (the public layout of the staged call),
// then the guided pass's private part: sampled high-byte histogram, candidate classes, candidate low-byte histograms
constexpr size_t kU16PublicBytes = 3 * 256 * 8 + 3 * lars::U16_MAX_BUCKETS * 256 * 8 + 3 * sizeof(lars::U16Select);
constexpr size_t kU16SetBytes = kU16PublicBytes + 3 * 256 * 8 + 3 * sizeof(lars::U16Candidates) +
                                3 * lars::U16_CAND_SLOTS * 256 * 8;
}  // namespace

size_t lars_wb_u16_workspace_bytes(int32_t n_sets) { return n_sets < 1 ? 0 : (size_t)n_sets * kU16SetBytes; }

int lars_wb_stretch_build_u16(const uint16_t* src, int32_t n_frames, int64_t n_pixels, int32_t channels,
                              int64_t src_frame_stride, double q_lo, double q_hi, lars_stretch_u16* stretch,
                              double* pct, void* workspace, size_t workspace_bytes, int32_t shared_hist,
                              void* stream) {
  return lars_wb_stretch_build_u16_staged(src, n_frames, n_pixels, channels, src_frame_stride, q_lo, q_hi, stretch, pct,
                                          workspace, workspace_bytes, shared_hist, LARS_U16_STAGE_ALL, stream);
}

int lars_wb_stretch_build_u16_staged(const uint16_t* src, int32_t n_frames, int64_t n_pixels, int32_t channels,
                                     int64_t src_frame_stride, double q_lo, double q_hi, lars_stretch_u16* stretch,
                                     double* pct, void* workspace, size_t workspace_bytes, int32_t shared_hist,
                                     int32_t stage, void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!src || !stretch || !workspace) return fail(LARS_ERR_INVALID, "lars_wb_stretch_build_u16: NULL pointer");
  if (n_frames < 1 || n_pixels < 1) return fail(LARS_ERR_INVALID, "lars_wb_stretch_build_u16: empty input");
  if (channels != 3 && channels != 4) return fail(LARS_ERR_UNSUPPORTED, "lars_wb_stretch_build_u16: channels must be 3 or 4, got %d", channels);
  if (!aligned16(src) || (src_frame_stride & 15) || src_frame_stride < n_pixels * channels * 2)
    return fail(LARS_ERR_INVALID, "lars_wb_stretch_build_u16: src / frame stride must be 16-byte aligned and cover a frame");
  if (!(q_lo >= 0.0 && q_lo <= 1.0 && q_hi >= 0.0 && q_hi <= 1.0))
    return fail(LARS_ERR_INVALID, "lars_wb_stretch_build_u16: quantiles must be fractions in [0,1]");
  const int n_sets = shared_hist ? 1 : n_frames;
  if (!aligned16(workspace) || !aligned16(stretch) || workspace_bytes < lars_wb_u16_workspace_bytes(n_sets))
    return fail(LARS_ERR_INVALID, "lars_wb_stretch_build_u16: workspace too small or misaligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  unsigned long long* hist_hi = reinterpret_cast<unsigned long long*>(ws);
  unsigned long long* hist_lo = reinterpret_cast<unsigned long long*>(ws + (size_t)n_sets * 3 * 256 * 8);
  lars::U16Select* select = reinterpret_cast<lars::U16Select*>(ws + (size_t)n_sets * (3 * 256 * 8 + 3 * lars::U16_MAX_BUCKETS * 256 * 8));
  char* priv = ws + (size_t)n_sets * kU16PublicBytes;
  unsigned long long* hist_sample = reinterpret_cast<unsigned long long*>(priv);
  lars::U16Candidates* cand = reinterpret_cast<lars::U16Candidates*>(priv + (size_t)n_sets * 3 * 256 * 8);
  unsigned long long* cand_lo = reinterpret_cast<unsigned long long*>(priv + (size_t)n_sets * (3 * 256 * 8 + 3 * sizeof(lars::U16Candidates)));
  if (stage < LARS_U16_STAGE_ALL || stage > LARS_U16_STAGE_ALL_TWO_LEVEL)
    return fail(LARS_ERR_INVALID, "lars_wb_stretch_build_u16: unknown stage %d", stage);
  const bool all = stage == LARS_U16_STAGE_ALL || stage == LARS_U16_STAGE_ALL_TWO_LEVEL;
  // per-frame statistics in one call: the guided single pass; tiles of one image and the staged form keep two levels
  const bool guided = stage == LARS_U16_STAGE_ALL && !shared_hist;
  const bool do_hi = all || stage == LARS_U16_STAGE_HIST_HI;
  const bool do_lo = all || stage == LARS_U16_STAGE_HIST_LO;
  const bool do_build = all || stage == LARS_U16_STAGE_BUILD;
  if (do_hi) LARS_CUDA(cudaMemsetAsync(workspace, 0, (size_t)n_sets * kU16SetBytes, s));

  const long long frame_bytes = (long long)n_pixels * channels * 2;
  const long long unit = (channels == 3) ? lars::U16Unit<3>::BYTES : lars::U16Unit<4>::BYTES;
  lars::U16HistParams p;
  p.src = reinterpret_cast<const uint8_t*>(src);
  p.hist_hi = hist_hi; p.hist_lo = hist_lo; p.select = select;
  p.n_pixels = n_pixels; p.frame_stride = src_frame_stride;
  p.units_per_frame = (frame_bytes + unit - 1) / unit;
  p.total_units = p.units_per_frame * n_frames;
  p.set_stride = shared_hist ? 0 : 1;
  p.n_frames = n_frames; p.lo_pass = 0;
  p.hist_sample = hist_sample; p.cand = cand; p.cand_lo = cand_lo;
  const long long target = 2ll * st->sm_count;
  const int grid = (int)(p.total_units < target ? p.total_units : target);
  if (guided) {
    const int sstep = lars::u16_sample_step(p.units_per_frame);
    const long long sampled = ((p.units_per_frame + sstep - 1) / sstep) * n_frames;
    const int sgrid = (int)(sampled < target ? sampled : target);
    if (channels == 3) lars::wb_hist_u16_sample_kernel<3><<<sgrid, lars::K1_THREADS, lars::U16_HI_SMEM_BYTES, s>>>(p);
    else lars::wb_hist_u16_sample_kernel<4><<<sgrid, lars::K1_THREADS, lars::U16_HI_SMEM_BYTES, s>>>(p);
    LARS_CUDA(cudaGetLastError());
    lars::U16CandParams cp; cp.hist_sample = hist_sample; cp.cand = cand; cp.q_lo = q_lo; cp.q_hi = q_hi;
    lars::wb_u16_candidates_kernel<<<n_sets * 3, 256, 0, s>>>(cp);
    LARS_CUDA(cudaGetLastError());
    const int ggrid = (int)(p.total_units < st->sm_count ? p.total_units : st->sm_count);   // one 1,024-thread CTA per SM
    if (channels == 3) lars::wb_hist_u16_guided_kernel<3><<<ggrid, lars::U16_GUIDED_THREADS, lars::U16_GUIDED_SMEM_BYTES, s>>>(p);
    else lars::wb_hist_u16_guided_kernel<4><<<ggrid, lars::U16_GUIDED_THREADS, lars::U16_GUIDED_SMEM_BYTES, s>>>(p);
    LARS_CUDA(cudaGetLastError());
  } else if (do_hi) {
    if (channels == 3) lars::wb_hist_u16_hi_kernel<3><<<grid, lars::K1_THREADS, lars::U16_HI_SMEM_BYTES, s>>>(p);
    else lars::wb_hist_u16_hi_kernel<4><<<grid, lars::K1_THREADS, lars::U16_HI_SMEM_BYTES, s>>>(p);
    LARS_CUDA(cudaGetLastError());
  }
  if (do_lo) {
    lars::U16SelectParams sp; sp.hist_hi = hist_hi; sp.select = select; sp.q_lo = q_lo; sp.q_hi = q_hi;
    sp.cand = guided ? cand : nullptr; sp.cand_lo = guided ? cand_lo : nullptr; sp.hist_lo = guided ? hist_lo : nullptr;
    lars::wb_u16_select_kernel<<<n_sets * 3, 256, 0, s>>>(sp);
    LARS_CUDA(cudaGetLastError());
    for (int pass = 0; pass < 2; ++pass) {
      p.lo_pass = pass;
      if (channels == 3) lars::wb_hist_u16_lo_kernel<3><<<grid, lars::K1_THREADS, lars::U16_LO_SMEM_BYTES, s>>>(p);
      else lars::wb_hist_u16_lo_kernel<4><<<grid, lars::K1_THREADS, lars::U16_LO_SMEM_BYTES, s>>>(p);
      LARS_CUDA(cudaGetLastError());
    }
  }
  if (do_build) {
    lars::U16BuildParams bp;
    bp.hist_hi = hist_hi; bp.hist_lo = hist_lo; bp.select = select; bp.stretch = stretch; bp.pct = pct;
    bp.q_lo = q_lo; bp.q_hi = q_hi;
    lars::wb_stretch_build_u16_kernel<<<n_sets * 3, 256, 0, s>>>(bp);
    LARS_CUDA(cudaGetLastError());
  }
  return LARS_OK;
}

static bool index_channels(int index, int* hi_c, int* lo_c) {
  switch (index) {
    case LARS_NDVI: *hi_c = 2; *lo_c = 0; return true;
    case LARS_GNDVI: *hi_c = 2; *lo_c = 1; return true;
    case LARS_NDWI: *hi_c = 1; *lo_c = 2; return true;
    default: return false;
  }
}

int lars_index_hwc(const void* src, int32_t dtype, int64_t n_pixels, int32_t channels, int32_t index,
                   float* out, void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!src || !out) return fail(LARS_ERR_INVALID, "lars_index_hwc: NULL pointer");
  if (n_pixels < 1 || channels < 3) return fail(LARS_ERR_INVALID, "lars_index_hwc: need >= 1 pixel and >= 3 channels");
  int hi_c, lo_c;
  if (!index_channels(index, &hi_c, &lo_c)) return fail(LARS_ERR_INVALID, "lars_index_hwc: unknown index %d", index);
  long long want = (n_pixels + 255) / 256;
  int grid = st->sm_count * 8;
  if (want < grid) grid = (int)want;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto launch = [&](auto kern, const auto* typed) { kern<<<grid, 256, 0, s>>>(typed, out, n_pixels, channels); };
#define LARS_HWC_DISPATCH(T)                                                                   \
  do {                                                                                         \
    const T* typed = static_cast<const T*>(src);                                               \
    if (index == LARS_NDVI) launch(lars::index_hwc_kernel<T, 2, 0>, typed);                    \
    else if (index == LARS_GNDVI) launch(lars::index_hwc_kernel<T, 2, 1>, typed);              \
    else launch(lars::index_hwc_kernel<T, 1, 2>, typed);                                       \
  } while (0)
  switch (dtype) {
    case LARS_DTYPE_U16: LARS_HWC_DISPATCH(uint16_t); break;
    case LARS_DTYPE_F32: LARS_HWC_DISPATCH(float); break;
    case LARS_DTYPE_F64: LARS_HWC_DISPATCH(double); break;
    default:
      return fail(LARS_ERR_UNSUPPORTED, "lars_index_hwc: dtype %d (uint8 frames go through lars_fused_index_u8)", dtype);
  }
#undef LARS_HWC_DISPATCH
  LARS_CUDA(cudaGetLastError());
  return LARS_OK;
}

int lars_index_change_u8(const uint8_t* early, const uint8_t* late, int64_t n_pixels, int32_t channels,
                         int32_t index, float vmin, float vmax, float* early_map, float* late_map,
                         float* diff, uint8_t* rgb, void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!early || !late || !diff) return fail(LARS_ERR_INVALID, "lars_index_change_u8: NULL pointer");
  if (n_pixels < 1 || channels < 3) return fail(LARS_ERR_INVALID, "lars_index_change_u8: need >= 1 pixel and >= 3 channels");
  if (!(vmax > vmin)) return fail(LARS_ERR_INVALID, "lars_index_change_u8: vmax must exceed vmin");
  lars::ChangeParams p;
  if (!index_channels(index, &p.hi_c, &p.lo_c)) return fail(LARS_ERR_INVALID, "lars_index_change_u8: unknown index %d", index);
  p.early = early; p.late = late; p.early_map = early_map; p.late_map = late_map; p.diff = diff; p.rgb = rgb;
  p.cmap = st->cmaps + LARS_CMAP_BWR * 256; p.n = n_pixels; p.channels = channels; p.vmin = vmin; p.vmax = vmax;
  long long want = (n_pixels + 255) / 256;
  int grid = st->sm_count * 8;
  if (want < grid) grid = (int)want;
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  if (index == LARS_NDVI) lars::index_change_u8_kernel<2, 0><<<grid, 256, 0, cs>>>(p);
  else if (index == LARS_GNDVI) lars::index_change_u8_kernel<2, 1><<<grid, 256, 0, cs>>>(p);
  else lars::index_change_u8_kernel<1, 2><<<grid, 256, 0, cs>>>(p);
  LARS_CUDA(cudaGetLastError());
  return LARS_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// Lanczos resize (Pillow Resample.c restated; preprocess_large_image, process-images.py:398-422)
// ------------------------------------------------------------------------------------------
namespace {

constexpr double kLanczosSupport = 3.0;
constexpr int kResizeSmemBudget = 96 * 1024;   // two CTAs per SM
constexpr int kResizeSmemMax = 200 * 1024;

inline double rs_sinc(double x) {
  if (x == 0.0) return 1.0;
  x = x * M_PI;
  return sin(x) / x;
}
inline double rs_lanczos(double x) {
  if (-3.0 <= x && x < 3.0) return rs_sinc(x) * rs_sinc(x / 3);
  return 0.0;
}
inline int rs_ksize(int in_size, int out_size) {
  double filterscale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  return (int)ceil(kLanczosSupport * filterscale) * 2 + 1;
}
// window of output sample xx: first source sample and tap count (Resample.c precompute_coeffs)
inline void rs_window(int in_size, int out_size, int xx, int* first, int* count) {
  const double scale = (double)in_size / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = kLanczosSupport * filterscale;
  const double center = (xx + 0.5) * scale;
  int xmin = (int)(center - support + 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)(center + support + 0.5);
  if (xmax > in_size) xmax = in_size;
  *first = xmin;
  *count = xmax - xmin;
}
// bounds [out][2] and fixed-point coefficients [out][ksize]; `scratch` holds ksize doubles
void rs_coeffs(int in_size, int out_size, int ksize, int* bounds, int* kk, double* scratch) {
  const double scale = (double)in_size / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    int xmin, cnt;
    rs_window(in_size, out_size, xx, &xmin, &cnt);
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    for (int x = 0; x < cnt; ++x) {
      const double w = rs_lanczos((x + xmin - center + 0.5) * ss);
      scratch[x] = w;
      ww += w;
    }
    int* k = kk + (size_t)xx * ksize;
    for (int x = 0; x < cnt; ++x) {
      double v = scratch[x];
      if (ww != 0.0) v /= ww;
      k[x] = (v < 0) ? (int)(-0.5 + v * (1 << lars::RS_PRECISION_BITS)) : (int)(0.5 + v * (1 << lars::RS_PRECISION_BITS));
    }
    for (int x = cnt; x < ksize; ++x) k[x] = 0;
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = cnt;
  }
}
// shared bytes of the horizontal pass for `tile` output columns per CTA
int rs_h_smem(int in_w, int out_w, int channels, int tile, int* plane_words, int* out_pitch) {
  int span_max = 0;                              // pixels
  for (int xo0 = 0; xo0 < out_w; xo0 += tile) {
    const int last = (xo0 + tile < out_w ? xo0 + tile : out_w) - 1;
    int f0, c0, f1, c1;
    rs_window(in_w, out_w, xo0, &f0, &c0);
    rs_window(in_w, out_w, last, &f1, &c1);
    const int span = f1 + c1 - f0;
    if (span > span_max) span_max = span;
  }
  *plane_words = (span_max + 3) / 4 + 4;         // + slack: tile origin rounded down to 4 pixels, last tap group
  int pw = (tile * channels + 3) / 4;
  if ((pw & 1) == 0) ++pw;                       // odd word pitch: lanes (rows) fall on distinct banks
  *out_pitch = pw * 4;
  return channels * *plane_words * lars::RS_IN_PITCH * 4 + lars::RS_ROWS * *out_pitch;
}
// Tensor-core horizontal pass: blocks of 8 output columns; block nb starts at input pixel
// kstart = first(8 nb) rounded down to 4 and spans `ksteps` steps of 32 pixels.
constexpr int kMmaOutPitch = lars::RS_MMA_OUT_PITCH * 4;   // bytes per row of the output tile: one packed pixel per word
int rs_mma_kstart(int in_w, int out_w, int nb) {
  int f, c;
  rs_window(in_w, out_w, nb * 8, &f, &c);
  return f & ~3;
}
// returns ksteps (0 = not usable) and the shared words per row per plane
int rs_mma_geometry(int in_w, int out_w, int* plane_words) {
  const int nblocks = (out_w + 7) / 8;
  int ksteps = 0;
  for (int nb = 0; nb < nblocks; ++nb) {
    const int last = (nb * 8 + 8 < out_w ? nb * 8 + 8 : out_w) - 1;
    int f, c;
    rs_window(in_w, out_w, last, &f, &c);
    const int need = (f + c - rs_mma_kstart(in_w, out_w, nb) + 31) / 32;
    if (need > ksteps) ksteps = need;
  }
  int span = 0;                                  // pixels a 64-column tile can touch
  for (int nb0 = 0; nb0 < nblocks; nb0 += 8) {
    const int nbl = (nb0 + 8 < nblocks ? nb0 + 8 : nblocks) - 1;
    const int s = rs_mma_kstart(in_w, out_w, nbl) + ksteps * 32 - rs_mma_kstart(in_w, out_w, nb0);
    if (s > span) span = s;
  }
  *plane_words = span / 4;
  const long long smem = 3ll * *plane_words * 32 * 4 + 32ll * kMmaOutPitch;
  return smem <= 100 * 1024 ? ksteps : 0;
}

// [n][ksize] int coefficients -> [n][groups][3] byte planes of 4 taps (bits 0-7, 8-15 unsigned; 16-23 signed)
bool rs_pack_planes(const int* kk, int n, int ksize, int groups, uint32_t* out) {
  for (int i = 0; i < n; ++i)
    for (int g = 0; g < groups; ++g) {
      uint32_t w0 = 0, w1 = 0, w2 = 0;
      for (int t = 0; t < 4; ++t) {
        const int idx = 4 * g + t;
        const int k = idx < ksize ? kk[(size_t)i * ksize + idx] : 0;
        const int hi = k >> 16;                  // arithmetic shift: floor
        if (hi < -128 || hi > 127) return false;
        w0 |= (uint32_t)(k & 255) << (8 * t);
        w1 |= (uint32_t)((k >> 8) & 255) << (8 * t);
        w2 |= (uint32_t)(hi & 255) << (8 * t);
      }
      uint32_t* o = out + ((size_t)i * groups + g) * 3;
      o[0] = w0; o[1] = w1; o[2] = w2;
    }
  return true;
}

}  // namespace

extern "C" {

int lars_resize_plan_lanczos(int32_t in_h, int32_t in_w, int32_t out_h, int32_t out_w, int32_t channels,
                             lars_resize_plan* plan) {
  if (!plan) return fail(LARS_ERR_INVALID, "lars_resize_plan_lanczos: NULL plan");
  if (in_h < 1 || in_w < 1 || out_h < 1 || out_w < 1)
    return fail(LARS_ERR_INVALID, "lars_resize_plan_lanczos: sizes must be positive");
  if (channels != 1 && channels != 3 && channels != 4)
    return fail(LARS_ERR_UNSUPPORTED, "lars_resize_plan_lanczos: %d channels (1, 3 or 4)", channels);
  if ((long long)in_w * channels > 0x7fffffffll / 2 || (long long)out_w * channels > 0x7fffffffll / 2)
    return fail(LARS_ERR_UNSUPPORTED, "lars_resize_plan_lanczos: rows longer than 2^30 bytes");
  memset(plan, 0, sizeof(*plan));
  plan->in_h = in_h; plan->in_w = in_w; plan->out_h = out_h; plan->out_w = out_w; plan->channels = channels;
  plan->need_h = (out_w != in_w);
  plan->need_v = (out_h != in_h);
  plan->ksize_h = plan->need_h ? rs_ksize(in_w, out_w) : 0;
  plan->ksize_v = plan->need_v ? rs_ksize(in_h, out_h) : 0;
  plan->row_first = 0;
  plan->row_count = in_h;
  if (plan->need_h && plan->need_v) {
    int f0, c0, f1, c1;
    rs_window(in_h, out_h, 0, &f0, &c0);
    rs_window(in_h, out_h, out_h - 1, &f1, &c1);
    plan->row_first = f0;                       // Resample.c ybox_first
    plan->row_count = f1 + c1 - f0;             // ybox_last - ybox_first
  }
  if (plan->need_h) {
    int tile = 64, smem = 0;
    for (;; tile >>= 1) {
      smem = rs_h_smem(in_w, out_w, channels, tile, &plan->plane_words, &plan->out_pitch);
      if (smem <= kResizeSmemBudget || tile == 1) break;
    }
    if (smem > kResizeSmemMax)
      return fail(LARS_ERR_UNSUPPORTED, "lars_resize_plan_lanczos: a %d -> %d filter window needs %d bytes of shared memory",
                  in_w, out_w, smem);
    plan->xo_tile = tile;
  }
  plan->groups_h = (plan->ksize_h + 3) / 4;
  plan->groups_v = (plan->ksize_v + 3) / 4;
  plan->table_bytes = 4ull * ((uint64_t)(plan->need_h ? out_w : 0) * (2 + 3 * plan->groups_h) +
                              (uint64_t)(plan->need_v ? out_h : 0) * (2 + 3 * plan->groups_v));
  if (plan->table_bytes == 0) plan->table_bytes = 4;
  if (plan->need_h && channels == 3) {
    plan->mma_ksteps = rs_mma_geometry(in_w, out_w, &plan->mma_plane_words);
    if (plan->mma_ksteps > 0) {
      const uint64_t nblocks = (uint64_t)(out_w + 7) / 8;
      plan->mma_table_offset = (plan->table_bytes + 7ull) & ~7ull;
      plan->table_bytes = plan->mma_table_offset + ((nblocks * 4 + 7ull) & ~7ull) +
                          nblocks * (uint64_t)plan->mma_ksteps * 3 * 32 * 8;
    }
  }
  plan->temp_frame_bytes = (plan->need_h && plan->need_v)
                               ? (((uint64_t)plan->row_count * out_w * channels + 15ull) & ~15ull) : 0;
  return LARS_OK;
}

int lars_resize_tables_lanczos(const lars_resize_plan* plan, void* tables_host) {
  if (!plan || !tables_host) return fail(LARS_ERR_INVALID, "lars_resize_tables_lanczos: NULL pointer");
  int* t = static_cast<int*>(tables_host);
  const int kmax = plan->ksize_h > plan->ksize_v ? plan->ksize_h : plan->ksize_v;
  double* scratch = static_cast<double*>(malloc(sizeof(double) * (size_t)(kmax + 1)));
  if (!scratch) return fail(LARS_ERR_INVALID, "lars_resize_tables_lanczos: out of host memory");
  if (plan->need_h) {
    int* bounds = t;
    uint32_t* planes = reinterpret_cast<uint32_t*>(t + 2 * (size_t)plan->out_w);
    int* kk = static_cast<int*>(malloc(sizeof(int) * (size_t)plan->out_w * plan->ksize_h));
    if (!kk) { free(scratch); return fail(LARS_ERR_INVALID, "lars_resize_tables_lanczos: out of host memory"); }
    rs_coeffs(plan->in_w, plan->out_w, plan->ksize_h, bounds, kk, scratch);
    const bool ok = rs_pack_planes(kk, plan->out_w, plan->ksize_h, plan->groups_h, planes);
    if (ok && plan->mma_ksteps > 0) {
      // B fragments of mma.m16n8k32 (col-major 32 x 8): lane = 4 g + t holds k = 4 t .. 4 t + 3 (b0) and
      // 16 + 4 t .. (b1) of output column g; one uint2 per (block, k-step, byte plane, lane)
      char* base = static_cast<char*>(tables_host) + plan->mma_table_offset;
      const int nblocks = (plan->out_w + 7) / 8;
      int* kstart = reinterpret_cast<int*>(base);
      uint32_t* frag = reinterpret_cast<uint32_t*>(base + (((size_t)nblocks * 4 + 7) & ~(size_t)7));
      for (int nb = 0; nb < nblocks; ++nb) {
        kstart[nb] = rs_mma_kstart(plan->in_w, plan->out_w, nb);
        for (int ks = 0; ks < plan->mma_ksteps; ++ks)
          for (int pl = 0; pl < 3; ++pl)
            for (int lane = 0; lane < 32; ++lane) {
              const int g = lane >> 2, t4 = (lane & 3) * 4;
              const int xo = nb * 8 + g;
              uint32_t w[2] = {0u, 0u};
              for (int half = 0; half < 2; ++half)
                for (int j = 0; j < 4; ++j) {
                  const int px = kstart[nb] + ks * 32 + half * 16 + t4 + j;
                  int coef = 0;
                  if (xo < plan->out_w) {
                    const int tap = px - bounds[2 * xo];
                    if (tap >= 0 && tap < bounds[2 * xo + 1]) coef = kk[(size_t)xo * plan->ksize_h + tap];
                  }
                  const uint32_t byte = pl == 0 ? (uint32_t)(coef & 255) : pl == 1 ? (uint32_t)((coef >> 8) & 255)
                                                                                  : (uint32_t)((coef >> 16) & 255);
                  w[half] |= byte << (8 * j);
                }
              uint32_t* o = frag + ((((size_t)nb * plan->mma_ksteps + ks) * 3 + pl) * 32 + lane) * 2;
              o[0] = w[0];
              o[1] = w[1];
            }
      }
    }
    free(kk);
    if (!ok) { free(scratch); return fail(LARS_ERR_UNSUPPORTED, "lars_resize_tables_lanczos: coefficient outside 24 bits"); }
    t = reinterpret_cast<int*>(planes + (size_t)plan->out_w * plan->groups_h * 3);
  }
  if (plan->need_v) {
    int* bounds = t;
    uint32_t* planes = reinterpret_cast<uint32_t*>(t + 2 * (size_t)plan->out_h);
    int* kk = static_cast<int*>(malloc(sizeof(int) * (size_t)plan->out_h * plan->ksize_v));
    if (!kk) { free(scratch); return fail(LARS_ERR_INVALID, "lars_resize_tables_lanczos: out of host memory"); }
    rs_coeffs(plan->in_h, plan->out_h, plan->ksize_v, bounds, kk, scratch);
    const bool ok = rs_pack_planes(kk, plan->out_h, plan->ksize_v, plan->groups_v, planes);
    free(kk);
    if (!ok) { free(scratch); return fail(LARS_ERR_UNSUPPORTED, "lars_resize_tables_lanczos: coefficient outside 24 bits"); }
    for (int i = 0; i < plan->out_h; ++i) bounds[2 * i] -= plan->row_first;   // rows of the intermediate image
  }
  if (!plan->need_h && !plan->need_v) t[0] = 0;
  free(scratch);
  return LARS_OK;
}

int lars_rgba_alpha_u8(uint8_t* data, int32_t n_frames, int64_t n_pixels, int64_t frame_stride, int32_t premultiply,
                       void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!data) return fail(LARS_ERR_INVALID, "lars_rgba_alpha_u8: NULL pointer");
  if (n_frames < 1 || n_frames > 65535 || n_pixels < 1) return fail(LARS_ERR_INVALID, "lars_rgba_alpha_u8: empty input");
  if ((reinterpret_cast<uintptr_t>(data) & 3u) || (frame_stride & 3) || frame_stride < n_pixels * 4)
    return fail(LARS_ERR_INVALID, "lars_rgba_alpha_u8: frames must be 4-byte aligned RGBA");
  long long want = (n_pixels + 255) / 256;
  const long long cap = (long long)st->sm_count * 8;
  const dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)n_frames);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (premultiply) lars::rgba_alpha_kernel<true><<<grid, 256, 0, s>>>(data, frame_stride, n_pixels);
  else lars::rgba_alpha_kernel<false><<<grid, 256, 0, s>>>(data, frame_stride, n_pixels);
  LARS_CUDA(cudaGetLastError());
  return LARS_OK;
}

int lars_resize_lanczos_u8(const lars_resize_plan* plan, const void* tables_dev, const uint8_t* src,
                           int64_t src_frame_stride, int32_t n_frames, uint8_t* dst, int64_t dst_frame_stride,
                           void* temp, size_t temp_bytes, void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!plan || !tables_dev || !src || !dst) return fail(LARS_ERR_INVALID, "lars_resize_lanczos_u8: NULL pointer");
  if (n_frames < 1 || n_frames > 65535) return fail(LARS_ERR_INVALID, "lars_resize_lanczos_u8: n_frames must be 1..65535");
  if (reinterpret_cast<uintptr_t>(tables_dev) & 3u) return fail(LARS_ERR_INVALID, "lars_resize_lanczos_u8: tables must be 4-byte aligned");
  const int C = plan->channels;
  const long long in_frame = (long long)plan->in_h * plan->in_w * C, out_frame = (long long)plan->out_h * plan->out_w * C;
  if (src_frame_stride < in_frame || dst_frame_stride < out_frame)
    return fail(LARS_ERR_INVALID, "lars_resize_lanczos_u8: frame strides smaller than a frame");
  if (plan->temp_frame_bytes && (!temp || temp_bytes < plan->temp_frame_bytes * (uint64_t)n_frames))
    return fail(LARS_ERR_INVALID, "lars_resize_lanczos_u8: temp must hold %llu bytes",
                (unsigned long long)(plan->temp_frame_bytes * (uint64_t)n_frames));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int* t = static_cast<const int*>(tables_dev);
  if (!plan->need_h && !plan->need_v) {          // Image.resize returns a copy
    LARS_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_frame_stride, src, (size_t)src_frame_stride, (size_t)in_frame,
                                (size_t)n_frames, cudaMemcpyDeviceToDevice, s));
    return LARS_OK;
  }
  const uint8_t* v_src = src;
  long long v_src_stride = src_frame_stride;
  if (plan->need_h) {
    lars::ResizeHParams p;
    p.src = src; p.src_frame_stride = src_frame_stride;
    p.dst = plan->need_v ? static_cast<uint8_t*>(temp) : dst;
    p.dst_frame_stride = plan->need_v ? (long long)plan->temp_frame_bytes : dst_frame_stride;
    p.bounds = t; p.kk = reinterpret_cast<const uint32_t*>(t + 2 * (size_t)plan->out_w);
    p.in_w = plan->in_w; p.out_w = plan->out_w; p.groups = plan->groups_h;
    p.row_first = plan->row_first; p.row_count = plan->row_count;
    p.xo_tile = plan->xo_tile; p.plane_words = plan->plane_words; p.out_pitch = plan->out_pitch;
    const int smem = C * plan->plane_words * lars::RS_IN_PITCH * 4 + lars::RS_ROWS * plan->out_pitch;
    if (smem > kResizeSmemMax) return fail(LARS_ERR_INVALID, "lars_resize_lanczos_u8: inconsistent plan");
    const long long gy = (plan->row_count + lars::RS_ROWS - 1) / lars::RS_ROWS;
    if (gy > 65535) return fail(LARS_ERR_UNSUPPORTED, "lars_resize_lanczos_u8: more than 2,097,120 rows");
    dim3 grid((plan->out_w + plan->xo_tile - 1) / plan->xo_tile, (unsigned)gy, n_frames);
    auto launch = [&](auto kern) -> int {
      if (smem > 48 * 1024) LARS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      kern<<<grid, lars::RS_THREADS, smem, s>>>(p);
      return LARS_OK;
    };
    const bool fast = (C == 3) && ((long long)plan->in_w * 3 % 4 == 0) && (src_frame_stride % 4 == 0) &&
                      !(reinterpret_cast<uintptr_t>(src) & 3u);
    if (fast && plan->mma_ksteps > 0 && !(plan->mma_table_offset & 7u) && !(reinterpret_cast<uintptr_t>(tables_dev) & 7u)) {
      const char* mbase = static_cast<const char*>(tables_dev) + plan->mma_table_offset;
      const size_t nblocks = (size_t)(plan->out_w + 7) / 8;
      lars::ResizeHMmaParams m;
      m.src = p.src; m.dst = p.dst; m.bounds = p.bounds;
      m.kstart = reinterpret_cast<const int*>(mbase);
      m.bfrag = reinterpret_cast<const uint2*>(mbase + ((nblocks * 4 + 7) & ~(size_t)7));
      m.src_frame_stride = p.src_frame_stride; m.dst_frame_stride = p.dst_frame_stride;
      m.in_w = plan->in_w; m.out_w = plan->out_w; m.ksteps = plan->mma_ksteps;
      m.row_first = plan->row_first; m.row_count = plan->row_count;
      m.plane_words = plan->mma_plane_words; m.out_pitch = kMmaOutPitch;
      const int msmem = 3 * plan->mma_plane_words * 32 * 4 + 32 * kMmaOutPitch;
      dim3 mgrid((plan->out_w + lars::RS_MMA_COLS - 1) / lars::RS_MMA_COLS, (unsigned)gy, n_frames);
      auto mlaunch = [&](auto kern) -> int {
        if (msmem > 48 * 1024) LARS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, msmem));
        kern<<<mgrid, lars::RS_THREADS, msmem, s>>>(m);
        return LARS_OK;
      };
      switch (plan->mma_ksteps) {               // k-step counts of ordinary scale factors keep B in registers
        case 1: rc = mlaunch(lars::resize_h_mma_kernel<1>); break;
        case 2: rc = mlaunch(lars::resize_h_mma_kernel<2>); break;
        case 3: rc = mlaunch(lars::resize_h_mma_kernel<3>); break;
        case 4: rc = mlaunch(lars::resize_h_mma_kernel<4>); break;
        default: rc = mlaunch(lars::resize_h_mma_kernel<0>); break;
      }
    } else
    rc = (C == 1) ? launch(lars::resize_h_kernel<1, false>)
         : (C == 4) ? launch(lars::resize_h_kernel<4, false>)
         : fast ? launch(lars::resize_h_kernel<3, true>) : launch(lars::resize_h_kernel<3, false>);
    if (rc != LARS_OK) return rc;
    LARS_CUDA(cudaGetLastError());
    t = reinterpret_cast<const int*>(p.kk + (size_t)plan->out_w * plan->groups_h * 3);
    v_src = static_cast<const uint8_t*>(temp);
    v_src_stride = (long long)plan->temp_frame_bytes;
  }
  if (plan->need_v) {
    lars::ResizeVParams p;
    p.src = v_src; p.src_frame_stride = v_src_stride;
    p.dst = dst; p.dst_frame_stride = dst_frame_stride;
    p.bounds = t; p.kk = reinterpret_cast<const uint32_t*>(t + 2 * (size_t)plan->out_h);
    p.row_bytes = plan->out_w * C; p.out_h = plan->out_h; p.groups = plan->groups_v;
    p.src_rows = plan->row_count;
    if (plan->out_h > 65535) return fail(LARS_ERR_UNSUPPORTED, "lars_resize_lanczos_u8: more than 65535 output rows");
    const bool vec4 = (p.row_bytes % 4 == 0) && (v_src_stride % 4 == 0) && (dst_frame_stride % 4 == 0) &&
                      !(reinterpret_cast<uintptr_t>(v_src) & 3u) && !(reinterpret_cast<uintptr_t>(dst) & 3u);
    const int per = lars::RS_THREADS * (vec4 ? 4 : 1);
    dim3 grid((p.row_bytes + per - 1) / per, plan->out_h, n_frames);
    if (vec4) lars::resize_v_kernel<4><<<grid, lars::RS_THREADS, 0, s>>>(p);
    else lars::resize_v_kernel<1><<<grid, lars::RS_THREADS, 0, s>>>(p);
    LARS_CUDA(cudaGetLastError());
  }
  return LARS_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// ingest: TIFF reader (host only)
// ------------------------------------------------------------------------------------------
extern "C" {

int lars_tiff_probe(const void* file, size_t file_bytes, lars_tiff_info* info) {
  if (!file || !info) return fail(LARS_ERR_INVALID, "lars_tiff_probe: NULL pointer");
  bool unsupported = false;
  const char* why = lars_host::tiff_probe(file, file_bytes, info, &unsupported);
  if (why) return fail(unsupported ? LARS_ERR_UNSUPPORTED : LARS_ERR_INVALID, "lars_tiff_probe: %s", why);
  return LARS_OK;
}

int lars_tiff_read_region(const void* file, size_t file_bytes, const lars_tiff_info* info, int32_t row0,
                          int32_t row1, int32_t col0, int32_t col1, void* dst, size_t dst_bytes,
                          int32_t n_threads) {
  if (!file || !info || !dst) return fail(LARS_ERR_INVALID, "lars_tiff_read: NULL pointer");
  lars_tiff_info check;
  bool unsupported = false;
  const char* why = lars_host::tiff_probe(file, file_bytes, &check, &unsupported);   // never trust a stale info block
  if (why) return fail(unsupported ? LARS_ERR_UNSUPPORTED : LARS_ERR_INVALID, "lars_tiff_read: %s", why);
  if (memcmp(&check, info, sizeof(check)) != 0) return fail(LARS_ERR_INVALID, "lars_tiff_read: info does not describe this file");
  why = lars_host::tiff_read_region(file, file_bytes, &check, row0, row1, col0, col1, dst, dst_bytes, n_threads);
  if (why) return fail(LARS_ERR_INVALID, "lars_tiff_read: %s", why);
  return LARS_OK;
}

int lars_tiff_read(const void* file, size_t file_bytes, const lars_tiff_info* info, void* dst, size_t dst_bytes) {
  if (!info) return fail(LARS_ERR_INVALID, "lars_tiff_read: NULL pointer");
  return lars_tiff_read_region(file, file_bytes, info, 0, info->height, 0, info->width, dst, dst_bytes, 1);
}

int lars_png_probe(const void* file, size_t file_bytes, lars_png_info* info) {
  if (!file || !info) return fail(LARS_ERR_INVALID, "lars_png_probe: NULL pointer");
  bool unsupported = false;
  const char* why = lars_host::png_probe(file, file_bytes, info, &unsupported);
  if (why) return fail(unsupported ? LARS_ERR_UNSUPPORTED : LARS_ERR_INVALID, "lars_png_probe: %s", why);
  return LARS_OK;
}

int lars_png_read(const void* file, size_t file_bytes, const lars_png_info* info, void* dst, size_t dst_bytes) {
  if (!file || !info || !dst) return fail(LARS_ERR_INVALID, "lars_png_read: NULL pointer");
  lars_png_info check;
  bool unsupported = false;
  const char* why = lars_host::png_probe(file, file_bytes, &check, &unsupported);   // never trust a stale info block
  if (why) return fail(unsupported ? LARS_ERR_UNSUPPORTED : LARS_ERR_INVALID, "lars_png_read: %s", why);
  if (memcmp(&check, info, sizeof(check)) != 0) return fail(LARS_ERR_INVALID, "lars_png_read: info does not describe this file");
  why = lars_host::png_read(file, file_bytes, &check, dst, dst_bytes);
  if (why) return fail(LARS_ERR_INVALID, "lars_png_read: %s", why);
  return LARS_OK;
}

static int tiff_device_chunks(const void* file, size_t file_bytes, const lars_tiff_info* info, lars_lzw_chunk* chunks,
                              int32_t max_chunks, int compression) {
  if (!file || !info || !chunks) return fail(LARS_ERR_INVALID, "lars_tiff_lzw_chunks: NULL pointer");
  lars_tiff_info check;
  bool unsupported = false;
  const char* why = lars_host::tiff_probe(file, file_bytes, &check, &unsupported);
  if (why) return fail(unsupported ? LARS_ERR_UNSUPPORTED : LARS_ERR_INVALID, "lars_tiff_lzw_chunks: %s", why);
  if (memcmp(&check, info, sizeof(check)) != 0) return fail(LARS_ERR_INVALID, "lars_tiff_lzw_chunks: info does not describe this file");
  if (check.compression != compression || check.tile_width > 0 || check.planar_config != 1 || check.bits_per_sample > 16)
    return fail(LARS_ERR_UNSUPPORTED, "lars_tiff_lzw_chunks: the device decoder takes %s-compressed chunky strips",
                compression == 5 ? "LZW" : "Deflate");
  if (check.n_strips > max_chunks) return fail(LARS_ERR_INVALID, "lars_tiff_lzw_chunks: %d strips, room for %d", check.n_strips, max_chunks);
  lars_host::TiffCursor c{static_cast<const uint8_t*>(file), file_bytes, check.big_endian != 0};
  const uint64_t row_bytes = (uint64_t)check.width * check.samples_per_pixel * (check.bits_per_sample / 8);
  for (int32_t s = 0; s < check.n_strips; ++s) {
    const int64_t left = (int64_t)check.height - (int64_t)s * check.rows_per_strip;
    const uint64_t rows = (uint64_t)(left < check.rows_per_strip ? left : check.rows_per_strip);
    const uint64_t need = rows * row_bytes;
    const uint64_t off = lars_host::tiff_value(c, check.strip_offsets_pos, check.strip_offsets_type, (uint64_t)s);
    const uint64_t cnt = lars_host::tiff_value(c, check.strip_counts_pos, check.strip_counts_type, (uint64_t)s);
    if (need > LARS_LZW_MAX_CHUNK || cnt > 0xffffffffull)
      return fail(LARS_ERR_UNSUPPORTED, "lars_tiff_lzw_chunks: a strip decodes to more than 1 MB");
    chunks[s].src_offset = off;
    chunks[s].dst_offset = (uint64_t)s * check.rows_per_strip * row_bytes;
    chunks[s].src_bytes = (uint32_t)cnt;
    chunks[s].dst_bytes = (uint32_t)need;
  }
  return check.n_strips;
}

int lars_tiff_lzw_chunks(const void* file, size_t file_bytes, const lars_tiff_info* info, lars_lzw_chunk* chunks,
                         int32_t max_chunks) {
  return tiff_device_chunks(file, file_bytes, info, chunks, max_chunks, 5);
}

int lars_tiff_deflate_chunks(const void* file, size_t file_bytes, const lars_tiff_info* info, lars_lzw_chunk* chunks,
                             int32_t max_chunks) {
  return tiff_device_chunks(file, file_bytes, info, chunks, max_chunks, 8);
}

int lars_inflate_decode_device(const uint8_t* src, const lars_lzw_chunk* chunks, int32_t n_chunks, uint8_t* dst,
                               uint32_t* counters, void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!src || !chunks || !dst || !counters) return fail(LARS_ERR_INVALID, "lars_inflate_decode_device: NULL pointer");
  if (n_chunks < 1) return fail(LARS_ERR_INVALID, "lars_inflate_decode_device: n_chunks=%d", n_chunks);
  if (reinterpret_cast<uintptr_t>(chunks) & 7u) return fail(LARS_ERR_INVALID, "lars_inflate_decode_device: chunks must be 8-byte aligned");
  if (!st->inflate) return fail(LARS_ERR_UNSUPPORTED, "lars_inflate_decode_device: the kernel's shared memory was refused on this device");
  lars::LzwParams p;
  p.src = src; p.chunks = chunks; p.dst = dst; p.status = counters; p.next = counters + 1; p.n_chunks = n_chunks;
  const int want = (n_chunks + lars::INF_WARPS - 1) / lars::INF_WARPS;
  lars::inflate_decode_kernel<<<want < st->sm_count ? want : st->sm_count, lars::INF_WARPS * 32, lars::INF_SMEM_BYTES,
                                static_cast<cudaStream_t>(stream)>>>(p);
  LARS_CUDA(cudaGetLastError());
  return LARS_OK;
}

int lars_untile_device(const uint8_t* scratch, int64_t slot_bytes, int64_t slot_row_bytes, const int64_t* table,
                       int32_t n_chunks, uint8_t* dst, int64_t dst_row_bytes, void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!scratch || !table || !dst) return fail(LARS_ERR_INVALID, "lars_untile_device: NULL pointer");
  if (n_chunks < 1 || slot_bytes < 1 || slot_row_bytes < 1 || dst_row_bytes < 1)
    return fail(LARS_ERR_INVALID, "lars_untile_device: bad geometry");
  if (reinterpret_cast<uintptr_t>(table) & 7u) return fail(LARS_ERR_INVALID, "lars_untile_device: table must be 8-byte aligned");
  lars::UntileParams p;
  p.scratch = scratch; p.table = reinterpret_cast<const long long*>(table); p.dst = dst;
  p.slot_bytes = slot_bytes; p.slot_row_bytes = slot_row_bytes; p.dst_row_bytes = dst_row_bytes; p.n_chunks = n_chunks;
  const dim3 grid((unsigned)n_chunks, 8);
  lars::untile_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  LARS_CUDA(cudaGetLastError());
  return LARS_OK;
}

int lars_lzw_decode_device(const uint8_t* src, const lars_lzw_chunk* chunks, int32_t n_chunks, uint8_t* dst,
                           uint32_t* counters, void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!src || !chunks || !dst || !counters) return fail(LARS_ERR_INVALID, "lars_lzw_decode_device: NULL pointer");
  if (n_chunks < 1) return fail(LARS_ERR_INVALID, "lars_lzw_decode_device: n_chunks=%d", n_chunks);
  if (reinterpret_cast<uintptr_t>(chunks) & 7u) return fail(LARS_ERR_INVALID, "lars_lzw_decode_device: chunks must be 8-byte aligned");
  lars::LzwParams p;
  p.src = src; p.chunks = chunks; p.dst = dst; p.status = counters; p.next = counters + 1; p.n_chunks = n_chunks;
  const int want = (n_chunks + lars::LZW_WARPS - 1) / lars::LZW_WARPS;
  const int full = st->sm_count * 3;
  lars::lzw_decode_kernel<<<want < full ? want : full, lars::LZW_WARPS * 32, lars::LZW_SMEM_BYTES,
                            static_cast<cudaStream_t>(stream)>>>(p);
  LARS_CUDA(cudaGetLastError());
  return LARS_OK;
}

int lars_tiff_post_device(uint8_t* dst, int32_t n_frames, int64_t frame_stride, int32_t rows, int32_t width,
                          int32_t samples_per_pixel, int32_t sample_bytes, int32_t predictor, int32_t swap16,
                          void* stream) {
  DeviceState* st = nullptr;
  int rc = current_state(&st);
  if (rc != LARS_OK) return rc;
  if (!dst) return fail(LARS_ERR_INVALID, "lars_tiff_post_device: NULL pointer");
  if (n_frames < 1 || rows < 1 || width < 1 || samples_per_pixel < 1 || samples_per_pixel > 4 ||
      (sample_bytes != 1 && sample_bytes != 2) || (predictor != 1 && predictor != 2))
    return fail(LARS_ERR_INVALID, "lars_tiff_post_device: bad geometry");
  if (sample_bytes == 2 && ((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)frame_stride) & 1u))
    return fail(LARS_ERR_INVALID, "lars_tiff_post_device: 16-bit frames must be 2-byte aligned");
  if (predictor == 1 && !(swap16 && sample_bytes == 2)) return LARS_OK;      // nothing to undo
  lars::TiffPostParams p;
  p.dst = dst; p.frame_stride = frame_stride;
  p.row_bytes = (long long)width * samples_per_pixel * sample_bytes;
  p.n_frames = n_frames; p.rows = rows; p.width = width; p.spp = samples_per_pixel; p.sample_bytes = sample_bytes;
  p.predictor = predictor; p.swap16 = swap16 ? 1 : 0;
  const long long threads = (long long)n_frames * rows * samples_per_pixel;
  lars::tiff_post_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  LARS_CUDA(cudaGetLastError());
  return LARS_OK;
}

}  // extern "C"
