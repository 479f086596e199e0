// TEST INFRASTRUCTURE ONLY.  Compiles the device arithmetic header (pixel_math.h) for the
// host with -ffp-contract=off so that the -m "not gpu" tests can check, exhaustively and
// bit for bit, the exact formulas the kernels evaluate against the NumPy oracle.  The product
// never loads this library.
#include <stdint.h>
#include <string.h>

#include <vector>
#include "../../lars_image_processing_b200/csrc/pixel_math.h"
#include "../../lars_image_processing_b200/csrc/lzw_warp.h"
#include "../../lars_image_processing_b200/csrc/inflate_warp.h"
#include "../../lars_image_processing_b200/csrc/tiff_host.h"

extern "C" {

// All 65,536 (hi, lo) uint8 pairs: value bits, pair-domain bin, literal-edge bin, cmap slot,
// coverage flag, and the NDWI value derived by negation.
void hc_pair_tables(int bins, const float* edges, float threshold, float* value, int32_t* bin_pair,
                    int32_t* bin_edges, int32_t* cmap, uint8_t* above, float* negated) {
  const float half = 0.5f * (float)bins;
  const float bias = half + LARS_HIST_BIAS;
  for (int hi = 0; hi < 256; ++hi)
    for (int lo = 0; lo < 256; ++lo) {
      const int k = hi * 256 + lo;
      const float x = lars_ratio_f32((float)hi, (float)lo);
      value[k] = x;
      bin_pair[k] = lars_hist_bin_pair(x, half, bias, bins - 1);
      bin_edges[k] = lars_hist_bin_edges(x, edges, bins);
      cmap[k] = lars_cmap_index(x);
      above[k] = x > threshold ? 1 : 0;
      negated[k] = lars_negate_index(x);
    }
}

void hc_hist_bin_edges(const float* x, int64_t n, int bins, const float* edges, int32_t* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = lars_hist_bin_edges(x[i], edges, bins);
}

// K4's fast path: the sub-bin table, with the literal chain behind its ambiguous entries
void hc_hist_bin_subbin(const float* x, int64_t n, int bins, const float* edges, int32_t* out, int32_t* n_ambiguous, int rn) {
  std::vector<uint8_t> table(LARS_SUBBIN_COUNT);
  int amb = 0;
  for (int k = 0; k < LARS_SUBBIN_COUNT; ++k) {
    table[k] = lars_hist_subbin_entry(k, edges, bins);
    amb += table[k] == LARS_SUBBIN_AMBIGUOUS;
  }
  *n_ambiguous = amb;
  for (int64_t i = 0; i < n; ++i) {
    int b = table[rn ? (int)lars_hist_subbin_index_rn(x[i]) : lars_hist_subbin_index(x[i])];
    if (b == LARS_SUBBIN_AMBIGUOUS) b = lars_hist_bin_edges(x[i], edges, bins);
    out[i] = b;
  }
}

// float64 flavour; the edges are built here with the device kernel's own expression (arange * step + start)
void hc_hist_bin_edges_f64(const double* x, int64_t n, int bins, int32_t* out, double* edges_out) {
  std::vector<double> edges(bins + 1);
  const double step = LARS_DDIV(2.0, (double)bins);
  for (int i = 0; i <= bins; ++i) edges[i] = (i == bins) ? 1.0 : LARS_DADD(LARS_DMUL((double)i, step), -1.0);
  for (int i = 0; i <= bins; ++i) edges_out[i] = edges[i];
  for (int64_t i = 0; i < n; ++i) out[i] = lars_hist_bin_edges_f64(x[i], edges.data(), bins);
}

void hc_cmap_index_range(const float* x, int64_t n, float vmin, float vmax, int32_t* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = lars_cmap_index_range(x[i], vmin, vmax);
}

void hc_wb_lut(double lo, double hi, int domain, uint8_t* out) {
  for (int v = 0; v < domain; ++v) out[v] = lars_wb_lut_entry((double)v, lo, hi);
}

void hc_wb_lut_rgn(double lo, double hi, int domain, uint8_t* out) {
  for (int v = 0; v < domain; ++v) out[v] = lars_wb_lut_entry_rgn((double)v, lo, hi);
}

double hc_percentile_lerp(double a, double b, double gamma) { return lars_percentile_lerp(a, b, gamma); }

void hc_ratio_clip_f64(const double* hi, const double* lo, int64_t n, double* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = lars_ratio_clip_f64(hi[i], lo[i]);
}

void hc_ratio_clip_f32(const float* hi, const float* lo, int64_t n, float* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = lars_ratio_clip_f32(hi[i], lo[i]);
}
// The conversion-free forms the fused kernel uses, over all pairs.
void hc_pair_tables_fast(int bins, float* value, int32_t* row, int32_t* slot) {
  const float half = 0.5f * (float)bins;
  const float bias = half + LARS_HIST_BIAS - 0.5f;
  for (int hi = 0; hi < 256; ++hi)
    for (int lo = 0; lo < 256; ++lo) {
      const int k = hi * 256 + lo;
      const float x = lars_ratio_pair_u8(hi, lo);
      value[k] = x;
      row[k] = (int32_t)(lars_hist_row_bits(x, half, bias) - LARS_MAGIC_U);
      slot[k] = (int32_t)(lars_cmap_slot_bits(x) - LARS_MAGIC_U);
      const float xn = lars_negate_index(x);
      row[65536 + k] = (int32_t)(lars_hist_row_bits(xn, half, bias) - LARS_MAGIC_U);
      slot[65536 + k] = (int32_t)(lars_cmap_slot_bits(xn) - LARS_MAGIC_U);
    }
}

// The warp LZW decoder of the device-side TIFF path (lzw_warp.h) with its 32 lanes run one after the other.
uint32_t hc_lzw_chunk_host(const uint8_t* in, uint32_t n_in, uint8_t* out, uint32_t cap) {
  return (uint32_t)lars_host::lzw_chunk(in, n_in, out, cap);       // the product's host decoder, for comparison
}
// The warp inflate of inflate_warp.h, lanes run one after the other; the stream sits at byte offset `skew` of an
// aligned, padded buffer as on the device.
uint32_t hc_inflate_warp(const uint8_t* in, uint32_t n_in, uint8_t* out, uint32_t cap, uint32_t skew) {
  static thread_local LarsInflateSmem sm;
  std::vector<uint32_t> padded((n_in + 16) / 4 + 2, 0xA5A5A5A5u);
  uint8_t* base = reinterpret_cast<uint8_t*>(padded.data()) + 4 + (skew & 3u);
  memcpy(base, in, n_in);
  return lars_inflate_warp(base, n_in, out, cap, &sm);
}
uint32_t hc_lzw_decode_warp(const uint8_t* in, uint32_t n_in, uint8_t* out, uint32_t cap) {
  static thread_local uint32_t table[4096];
  return lars_lzw_decode_warp(in, n_in, out, cap, table);
}
}
