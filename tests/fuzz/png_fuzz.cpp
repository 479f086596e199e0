#include <stdio.h>
#include <random>
#include <vector>
#include "../../lars_image_processing_b200/csrc/png_host.h"
int main(int argc, char** argv) {
  std::mt19937_64 rng(99);
  long ok = 0, rej = 0, rok = 0, rrej = 0;
  for (int a = 1; a < argc; ++a) {
    FILE* f = fopen(argv[a], "rb"); if (!f) continue;
    std::vector<uint8_t> seed; uint8_t buf[65536]; size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) seed.insert(seed.end(), buf, buf + n);
    fclose(f);
    for (int it = 0; it < 6000; ++it) {
      std::vector<uint8_t> raw(seed);
      int flips = 1 + rng() % 4;
      for (int k = 0; k < flips; ++k) {
        size_t span = (it % 3 == 0) ? raw.size() : std::min<size_t>(raw.size(), 64);
        raw[rng() % span] = (uint8_t)rng();
      }
      if (it % 7 == 0) raw.resize(1 + rng() % raw.size());
      uint8_t* p = (uint8_t*)malloc(raw.size()); memcpy(p, raw.data(), raw.size());
      lars_png_info info; bool unsup;
      const char* why = lars_host::png_probe(p, raw.size(), &info, &unsup);
      if (!why) {
        ++ok;
        if (info.frame_bytes <= (1u << 26)) {
          uint8_t* dst = (uint8_t*)malloc(info.frame_bytes);
          if (lars_host::png_read(p, raw.size(), &info, dst, info.frame_bytes)) ++rrej; else ++rok;
          free(dst);
        }
      } else ++rej;
      free(p);
    }
  }
  printf("probe ok %ld rejected %ld; read ok %ld rejected %ld\n", ok, rej, rok, rrej);
}
