#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <random>
#include "../../lars_image_processing_b200/csrc/tiff_host.h"
int main(int argc, char** argv) {
  std::mt19937_64 rng(12345);
  long ok = 0, rej = 0, readok = 0, readrej = 0;
  for (int a = 1; a < argc; ++a) {
    FILE* f = fopen(argv[a], "rb"); if (!f) continue;
    std::vector<uint8_t> seed; uint8_t buf[65536]; size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) seed.insert(seed.end(), buf, buf + n);
    fclose(f);
    for (int it = 0; it < 4000; ++it) {
      std::vector<uint8_t> raw(seed);
      int flips = 1 + rng() % 4;
      for (int k = 0; k < flips; ++k) {
        size_t span = (it % 3 == 0) ? raw.size() : std::min<size_t>(raw.size(), 400);
        raw[rng() % span] = (uint8_t)rng();
      }
      if (it % 7 == 0) raw.resize(1 + rng() % raw.size());
      // exact-size heap copy so ASAN sees overreads
      uint8_t* p = (uint8_t*)malloc(raw.size()); memcpy(p, raw.data(), raw.size());
      lars_tiff_info info; bool unsup;
      const char* why = lars_host::tiff_probe(p, raw.size(), &info, &unsup);
      if (!why) {
        ++ok;
        if (info.frame_bytes <= (1u << 26)) {
          uint8_t* dst = (uint8_t*)malloc(info.frame_bytes);
          const char* w2 = lars_host::tiff_read_region(p, raw.size(), &info, 0, info.height, 0, info.width, dst, info.frame_bytes, 1 + it % 3);
          if (w2) ++readrej; else ++readok;
          // a sub-region too
          int r0 = rng() % info.height, c0 = rng() % info.width;
          int r1 = r0 + 1 + rng() % (info.height - r0), c1 = c0 + 1 + rng() % (info.width - c0);
          size_t need = (size_t)(r1 - r0) * (c1 - c0) * info.samples_per_pixel * (info.bits_per_sample / 8);
          uint8_t* d2 = (uint8_t*)malloc(need);
          lars_host::tiff_read_region(p, raw.size(), &info, r0, r1, c0, c1, d2, need, 2);
          free(d2);
          free(dst);
        }
      } else ++rej;
      free(p);
    }
  }
  printf("probe ok %ld rejected %ld; read ok %ld rejected %ld\n", ok, rej, readok, readrej);
  return 0;
}
