// Practical HBM roofline for the access mix of the fused pass (read 3 B/px, write 24 B/px over 7
// streams) next to plain copy / fill, to tell "kernel inefficiency" from "what HBM gives for a
// write-dominated mix".  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench membench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void copy_k(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void fill_k(uint4* __restrict__ b, size_t n, uint32_t v) {
  const uint4 x = make_uint4(v, v, v, v);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = x;
}
__global__ void fill_cs_k(uint4* __restrict__ b, size_t n, uint32_t v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(b + i), "r"(v) : "memory");
}
__global__ void read_k(const uint4* __restrict__ a, size_t n, uint32_t* out) {
  uint32_t acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { uint4 x = a[i]; acc += x.x ^ x.y ^ x.z ^ x.w; }
  if (acc == 0x12345678u) *out = acc;
}
// mix: per group of 16 px: read 48 B (3 x uint4), write 3 x 64 B floats + 4 x 48 B bytes.  Contiguous per CTA tile.
__global__ void mix_k(const uint4* __restrict__ src, uint4* __restrict__ m0, uint4* __restrict__ m1, uint4* __restrict__ m2,
                      uint4* __restrict__ b0, uint4* __restrict__ b1, uint4* __restrict__ b2, uint4* __restrict__ b3, size_t ngroups4) {
  // one thread handles 4 px: reads 12 B (as part of a uint4 every 4/3 threads -> approximate with 3 threads share), writes 16 B x 3 + 12 B x 4
  // implemented per 16-px group per 4 threads for simplicity of alignment: thread q of the group reads uint4 q (q<3) and writes.
  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < ngroups4; g += (size_t)gridDim.x * blockDim.x) {
    const size_t grp = g >> 2; const int q = (int)(g & 3);
    uint4 x = make_uint4(1, 2, 3, 4);
    if (q < 3) x = src[grp * 3 + q];
    m0[g] = x; m1[g] = x; m2[g] = x;
    if (q < 3) { b0[grp * 3 + q] = x; b1[grp * 3 + q] = x; b2[grp * 3 + q] = x; b3[grp * 3 + q] = x; }
  }
}

// 7 output streams, no read: element-interleaved (thread g writes its piece of every stream)
__global__ void fill7_k(uint4* __restrict__ m0, uint4* __restrict__ m1, uint4* __restrict__ m2,
                        uint4* __restrict__ b0, uint4* __restrict__ b1, uint4* __restrict__ b2, uint4* __restrict__ b3, size_t ngroups4) {
  const uint4 x = make_uint4(1, 2, 3, 4);
  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < ngroups4; g += (size_t)gridDim.x * blockDim.x) {
    const size_t grp = g >> 2; const int q = (int)(g & 3);
    m0[g] = x; m1[g] = x; m2[g] = x;
    if (q < 3) { b0[grp * 3 + q] = x; b1[grp * 3 + q] = x; b2[grp * 3 + q] = x; b3[grp * 3 + q] = x; }
  }
}
// 7 output streams, CTA-chunked: a CTA owns a contiguous pixel range and writes CHUNK pixels of stream 0, then of stream 1, ...
__global__ void fill7_chunk_k(uint4* __restrict__ m0, uint4* __restrict__ m1, uint4* __restrict__ m2,
                              uint4* __restrict__ b0, uint4* __restrict__ b1, uint4* __restrict__ b2, uint4* __restrict__ b3,
                              size_t npx, int chunk_px) {
  const uint4 x = make_uint4(1, 2, 3, 4);
  const size_t per = (npx / gridDim.x) / chunk_px * chunk_px;
  const size_t p0 = blockIdx.x * per, p1 = (blockIdx.x == gridDim.x - 1) ? npx / chunk_px * chunk_px : p0 + per;
  uint4* ms[3] = {m0, m1, m2}; uint4* bs[4] = {b0, b1, b2, b3};
  for (size_t c = p0; c < p1; c += chunk_px) {
    for (int s = 0; s < 3; ++s) for (int i = threadIdx.x; i < chunk_px / 4; i += blockDim.x) ms[s][c / 4 + i] = x;
    for (int s = 0; s < 4; ++s) for (int i = threadIdx.x; i < chunk_px * 3 / 16; i += blockDim.x) bs[s][c * 3 / 16 + i] = x;
  }
}
__global__ void fill1_k(uint4* __restrict__ b, size_t n) {
  const uint4 x = make_uint4(1, 2, 3, 4);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = x;
}

// 3r + 24w with stores independent of the loads (no dependency stalls): the pure DRAM mix
__global__ void mix_indep_k(const uint4* __restrict__ src, uint4* __restrict__ m0, uint4* __restrict__ m1, uint4* __restrict__ m2,
                            uint4* __restrict__ b0, uint4* __restrict__ b1, uint4* __restrict__ b2, uint4* __restrict__ b3,
                            size_t ngroups4, uint32_t* sink) {
  const uint4 x = make_uint4(1, 2, 3, 4);
  uint32_t acc = 0;
  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < ngroups4; g += (size_t)gridDim.x * blockDim.x) {
    const size_t grp = g >> 2; const int q = (int)(g & 3);
    if (q < 3) { const uint4 v = src[grp * 3 + q]; acc += v.x ^ v.y ^ v.z ^ v.w; }
    m0[g] = x; m1[g] = x; m2[g] = x;
    if (q < 3) { b0[grp * 3 + q] = x; b1[grp * 3 + q] = x; b2[grp * 3 + q] = x; b3[grp * 3 + q] = x; }
  }
  if (acc == 0x12345678u) *sink = acc;
}
template <class F> float timeit(F f, int it = 10) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f(); CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
  float best = 1e30f;
  for (int i = 0; i < it; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}
int main() {
  const size_t npx = 16ull * 12000000ull;   // the bench batch
  uint8_t *src, *b0, *b1, *b2, *b3; float *m0, *m1, *m2; uint32_t* flag;
  CK(cudaMalloc(&src, npx * 3)); CK(cudaMalloc(&b0, npx * 3)); CK(cudaMalloc(&b1, npx * 3)); CK(cudaMalloc(&b2, npx * 3)); CK(cudaMalloc(&b3, npx * 3));
  CK(cudaMalloc(&m0, npx * 4)); CK(cudaMalloc(&m1, npx * 4)); CK(cudaMalloc(&m2, npx * 4)); CK(cudaMalloc(&flag, 4));
  CK(cudaMemset(src, 1, npx * 3));
  const int grid = 148 * 8, blk = 512;
  const size_t nv = npx * 4 / 16;  // uint4 in a float map (768 MB)
  float t;
  t = timeit([&] { copy_k<<<grid, blk>>>((uint4*)m0, (uint4*)m1, nv); });
  printf("copy   768MB->768MB : %.3f ms  %.0f GB/s (read+write)\n", t, 2.0 * npx * 4 / t / 1e6);
  t = timeit([&] { fill_k<<<grid, blk>>>((uint4*)m0, nv, 7); fill_k<<<grid, blk>>>((uint4*)m1, nv, 7); fill_k<<<grid, blk>>>((uint4*)m2, nv, 7); });
  printf("fill   3x768MB      : %.3f ms  %.0f GB/s (write)\n", t, 3.0 * npx * 4 / t / 1e6);
  t = timeit([&] { fill_cs_k<<<grid, blk>>>((uint4*)m0, nv, 7); fill_cs_k<<<grid, blk>>>((uint4*)m1, nv, 7); fill_cs_k<<<grid, blk>>>((uint4*)m2, nv, 7); });
  printf("fill.cs 3x768MB     : %.3f ms  %.0f GB/s (write)\n", t, 3.0 * npx * 4 / t / 1e6);
  t = timeit([&] { cudaMemsetAsync(m0, 0, npx * 4); cudaMemsetAsync(m1, 0, npx * 4); cudaMemsetAsync(m2, 0, npx * 4); });
  printf("memset 3x768MB      : %.3f ms  %.0f GB/s (write)\n", t, 3.0 * npx * 4 / t / 1e6);
  t = timeit([&] { read_k<<<grid, blk>>>((uint4*)m0, nv, flag); read_k<<<grid, blk>>>((uint4*)m1, nv, flag); });
  printf("read   2x768MB      : %.3f ms  %.0f GB/s (read)\n", t, 2.0 * npx * 4 / t / 1e6);
  for (int g : {148 * 2, 148 * 4, 148 * 8, 148 * 16}) {
    t = timeit([&] { mix_k<<<g, 512>>>((uint4*)src, (uint4*)m0, (uint4*)m1, (uint4*)m2, (uint4*)b0, (uint4*)b1, (uint4*)b2, (uint4*)b3, npx / 4); });
    printf("mix 3r+24w B/px grid %5d: %.3f ms  %.0f GB/s  (%.1f Gpix/s)\n", g, t, 27.0 * npx / t / 1e6, npx / t / 1e6);
  }

  for (int g : {148 * 2, 148 * 8}) {
    t = timeit([&] { fill7_k<<<g, 512>>>((uint4*)m0, (uint4*)m1, (uint4*)m2, (uint4*)b0, (uint4*)b1, (uint4*)b2, (uint4*)b3, npx / 4); });
    printf("fill7 interleaved 24w B/px grid %5d: %.3f ms  %.0f GB/s\n", g, t, 24.0 * npx / t / 1e6);
  }
  for (int chunk : {1024, 4096})
    for (int g : {148 * 2, 148 * 4}) {
      t = timeit([&] { fill7_chunk_k<<<g, 512>>>((uint4*)m0, (uint4*)m1, (uint4*)m2, (uint4*)b0, (uint4*)b1, (uint4*)b2, (uint4*)b3, npx, chunk); });
      printf("fill7 chunk %6d px grid %5d: %.3f ms  %.0f GB/s\n", chunk, g, t, 24.0 * npx / t / 1e6);
    }
  {
    uint8_t* big; CK(cudaMalloc(&big, npx * 24));
    t = timeit([&] { fill1_k<<<148 * 8, 512>>>((uint4*)big, npx * 24 / 16); });
    printf("fill1 single 4.6 GB stream: %.3f ms  %.0f GB/s\n", t, 24.0 * npx / t / 1e6);
    t = timeit([&] { cudaMemsetAsync(big, 0, npx * 24); });
    printf("memset single 4.6 GB      : %.3f ms  %.0f GB/s\n", t, 24.0 * npx / t / 1e6);
    t = timeit([&] { copy_k<<<148 * 8, 512>>>((uint4*)big, (uint4*)(big + npx * 12), npx * 12 / 16); });
    printf("copy 2.3 GB -> 2.3 GB     : %.3f ms  %.0f GB/s (read+write)\n", t, 24.0 * npx / t / 1e6);
    cudaFree(big);
  }

  for (int g : {148 * 4, 148 * 8, 148 * 16}) {
    t = timeit([&] { mix_indep_k<<<g, 512>>>((uint4*)src, (uint4*)m0, (uint4*)m1, (uint4*)m2, (uint4*)b0, (uint4*)b1, (uint4*)b2, (uint4*)b3, npx / 4, flag); });
    printf("mix independent 3r+24w grid %5d: %.3f ms  %.0f GB/s\n", g, t, 27.0 * npx / t / 1e6);
  }
  return 0;
}
