// Host-side constant tables: matplotlib colormaps and NumPy histogram edges.
//
// matplotlib is a third-party dependency of the reference (requirements.txt:4, unpinned) and
// is not vendored in it, so its published algorithm is restated here:
// LinearSegmentedColormap.from_list(name, anchors, N=256) -> _create_lookup_table ->
// bytes = trunc(lut * 255).  Call sites in the reference: process-images.py:689-695, :755-760,
// :935-956; backend-process.py:42-43; process-ndvi.py:38.  All arithmetic is IEEE double with
// the same operation order as NumPy (compile with -ffp-contract=off).
#pragma once
#include <stdint.h>

namespace lars_host {

static const uint8_t kRdYlGn[11][3] = {{165, 0, 38},    {215, 48, 39},   {244, 109, 67}, {253, 174, 97},
                                       {254, 224, 139}, {255, 255, 191}, {217, 239, 139}, {166, 217, 106},
                                       {102, 189, 99},  {26, 152, 80},   {0, 104, 55}};
static const uint8_t kRdYlBu[11][3] = {{165, 0, 38},    {215, 48, 39},   {244, 109, 67},  {253, 174, 97},
                                       {254, 224, 144}, {255, 255, 191}, {224, 243, 248}, {171, 217, 233},
                                       {116, 173, 209}, {69, 117, 180},  {49, 54, 149}};
static const uint8_t kBwr[3][3] = {{0, 0, 255}, {255, 255, 255}, {255, 0, 0}};

// np.linspace(0.0, 1.0, num)[i]
static inline double unit_linspace(int i, int num) {
  if (num > 1 && i == num - 1) return 1.0;
  const double step = 1.0 / (double)(num - 1);
  return (double)i * step + 0.0;
}

static inline bool build_colormap(int cmap_id, uint8_t out[256][3]) {
  const uint8_t(*anchors)[3];
  int m;
  switch (cmap_id) {
    case 0: anchors = kRdYlGn; m = 11; break;
    case 1: anchors = kRdYlBu; m = 11; break;
    case 2: anchors = kBwr; m = 3; break;
    default: return false;
  }
  const int n = 256;
  double pos[11];
  for (int j = 0; j < m; ++j) pos[j] = unit_linspace(j, m) * (double)(n - 1);
  for (int ch = 0; ch < 3; ++ch) {
    double y[11];
    for (int j = 0; j < m; ++j) y[j] = (double)anchors[j][ch] / 255.0;
    for (int i = 0; i < n; ++i) {
      double v;
      if (i == 0) {
        v = y[0];
      } else if (i == n - 1) {
        v = y[m - 1];
      } else {
        const double x = (double)(n - 1) * unit_linspace(i, n);
        int ind = 0;  // np.searchsorted(pos, x, side='left')
        while (ind < m && pos[ind] < x) ++ind;
        const double frac = (x - pos[ind - 1]) / (pos[ind] - pos[ind - 1]);
        v = frac * (y[ind] - y[ind - 1]) + y[ind - 1];
      }
      v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
      out[i][ch] = (uint8_t)(v * 255.0);
    }
  }
  return true;
}

// np.linspace(-1, 1, bins + 1, dtype=float32): float64 arithmetic, cast at the end, last = stop.
static inline void histogram_edges_f32(int bins, float* edges) {
  const double step = 2.0 / (double)bins;
  for (int i = 0; i <= bins; ++i) edges[i] = (float)((double)i * step + -1.0);
  edges[bins] = 1.0f;
}

}  // namespace lars_host
