// Does DRAM like the 3r+24w mix better when the reads arrive as GPU-wide synchronous bursts?
// Persistent CTAs; per iteration a CTA loads B bytes of input into shared memory and then writes
// 8 B bytes of output (7 streams of the fused pass collapsed to one interleaved stream per CTA slice).
// sync = 1: a global barrier (one atomic per CTA + spin) precedes every load burst, so every SM
// reads at the same time and DRAM sees read bursts separated by long pure-write periods.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench3 membench3.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ void stg_cs(uint4* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ldg_cs(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

template <int SYNC>
__global__ void __launch_bounds__(512, 2) burst_k(const uint4* __restrict__ src, uint4* __restrict__ out, long long in_v, int bvec,
                                                  unsigned int* counter, uint32_t* sink) {
  extern __shared__ uint4 buf[];
  const long long per = in_v / gridDim.x;                 // uint4 of input per CTA
  const uint4* s = src + per * blockIdx.x;
  uint4* o = out + per * 8 * blockIdx.x;
  unsigned int epoch = 0;
  uint32_t acc = 0;
  for (long long base = 0; base + bvec <= per; base += bvec) {
    if (SYNC) {
      if (threadIdx.x == 0) {
        ++epoch;
        atomicAdd(counter, 1u);
        const unsigned int target = epoch * gridDim.x;
        while (*(volatile unsigned int*)counter < target) { }
      }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < bvec; i += blockDim.x) buf[i] = ldg_cs(s + base + i);
    __syncthreads();
    for (int i = threadIdx.x; i < bvec * 8; i += blockDim.x) {
      uint4 v = buf[i >> 3]; v.x += i;
      stg_cs(o + base * 8 + i, v);
    }
    __syncthreads();
  }
  if (acc == 0x12345678u) *sink = acc;
}

// Same structure, but the 24 B/px go to SEVEN arrays as in the fused pass: 3 float maps (4 B/px) and
// 4 byte images (3 B/px); per iteration each stream receives one contiguous run.
template <int SYNC>
__global__ void __launch_bounds__(512, 2) burst7_k(const uint4* __restrict__ src, uint4* __restrict__ m0, uint4* __restrict__ m1,
                                                   uint4* __restrict__ m2, uint4* __restrict__ b0, uint4* __restrict__ b1,
                                                   uint4* __restrict__ b2, uint4* __restrict__ b3, long long in_v, int bvec,
                                                   unsigned int* counter, uint32_t* sink) {
  extern __shared__ uint4 buf[];
  const long long per = in_v / gridDim.x / 3 * 3;         // uint4 of input per CTA (multiple of 3 -> whole 16-px groups)
  const uint4* s = src + per * blockIdx.x;
  uint4* ms[3] = {m0 + per / 3 * 4 * blockIdx.x, m1 + per / 3 * 4 * blockIdx.x, m2 + per / 3 * 4 * blockIdx.x};
  uint4* bs[4] = {b0 + per * blockIdx.x, b1 + per * blockIdx.x, b2 + per * blockIdx.x, b3 + per * blockIdx.x};
  unsigned int epoch = 0;
  for (long long base = 0; base + bvec <= per; base += bvec) {
    if (SYNC) {
      if (threadIdx.x == 0) {
        ++epoch;
        atomicAdd(counter, 1u);
        const unsigned int target = epoch * gridDim.x;
        while (*(volatile unsigned int*)counter < target) { }
      }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < bvec; i += blockDim.x) buf[i] = ldg_cs(s + base + i);
    __syncthreads();
    const int mvec = bvec / 3 * 4;                         // uint4 per map for this chunk
#pragma unroll
    for (int k = 0; k < 3; ++k)
      for (int i = threadIdx.x; i < mvec; i += blockDim.x) { uint4 v = buf[(i * 3) >> 2]; v.x += i; stg_cs(ms[k] + base / 3 * 4 + i, v); }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      for (int i = threadIdx.x; i < bvec; i += blockDim.x) { uint4 v = buf[i]; v.y += k; stg_cs(bs[k] + base + i, v); }
    __syncthreads();
  }
  if (epoch == 0x12345678u) *sink = epoch;
}

// Same as burst7_k, but the four byte streams leave through cp.async.bulk (TMA) stores issued by one
// thread from the staged tile, as the fused pass does; the float maps stay st.global.
__global__ void __launch_bounds__(512, 2) burst7_tma_k(const uint4* __restrict__ src, uint4* __restrict__ m0, uint4* __restrict__ m1,
                                                       uint4* __restrict__ m2, uint4* __restrict__ b0, uint4* __restrict__ b1,
                                                       uint4* __restrict__ b2, uint4* __restrict__ b3, long long in_v, int bvec) {
  extern __shared__ __align__(128) uint4 buf[];
  const long long per = in_v / gridDim.x / 3 * 3;
  const uint4* s = src + per * blockIdx.x;
  uint4* ms[3] = {m0 + per / 3 * 4 * blockIdx.x, m1 + per / 3 * 4 * blockIdx.x, m2 + per / 3 * 4 * blockIdx.x};
  uint4* bs[4] = {b0 + per * blockIdx.x, b1 + per * blockIdx.x, b2 + per * blockIdx.x, b3 + per * blockIdx.x};
  const uint32_t sbuf = (uint32_t)__cvta_generic_to_shared(buf);
  for (long long base = 0; base + bvec <= per; base += bvec) {
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous tile drained
    __syncthreads();
    for (int i = threadIdx.x; i < bvec; i += blockDim.x) buf[i] = ldg_cs(s + base + i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(bs[k] + base), "r"(sbuf), "r"(bvec * 16) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    const int mvec = bvec / 3 * 4;
#pragma unroll
    for (int k = 0; k < 3; ++k)
      for (int i = threadIdx.x; i < mvec; i += blockDim.x) { uint4 v = buf[(i * 3) >> 2]; v.x += i; stg_cs(ms[k] + base / 3 * 4 + i, v); }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const long long npx = 16ll * 12000000;
  const long long in_bytes = npx * 3, out_bytes = npx * 24;
  uint8_t *src, *out; uint32_t* sink; unsigned int* counter;
  CK(cudaMalloc(&src, in_bytes)); CK(cudaMalloc(&out, out_bytes)); CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&counter, 4));
  CK(cudaMemset(src, 1, in_bytes));
  CK(cudaFuncSetAttribute(burst_k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CK(cudaFuncSetAttribute(burst_k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int grid = 296;
  for (int kb : {6, 12, 24, 48, 96})
    for (int sync = 0; sync < 2; ++sync) {
      const int bvec = kb * 1024 / 16;
      float best = 1e30f;
      for (int it = 0; it < 5; ++it) {
        CK(cudaMemset(counter, 0, 4));
        cudaEventRecord(a);
        if (sync) burst_k<1><<<grid, 512, kb * 1024>>>((const uint4*)src, (uint4*)out, in_bytes / 16, bvec, counter, sink);
        else burst_k<0><<<grid, 512, kb * 1024>>>((const uint4*)src, (uint4*)out, in_bytes / 16, bvec, counter, sink);
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); if (it > 0 && ms < best) best = ms;
      }
      CK(cudaGetLastError());
      printf("burst %3d KB/CTA  sync=%d : %.3f ms  %.0f GB/s on 27 B/px  (%.1f Gpix/s)\n", kb, sync, best, 27.0 * npx / best / 1e6, npx / best / 1e6);
    }
  {
    uint8_t *m[3], *bb[4];
    for (int k = 0; k < 3; ++k) CK(cudaMalloc(&m[k], npx * 4));
    for (int k = 0; k < 4; ++k) CK(cudaMalloc(&bb[k], npx * 3));
    CK(cudaFuncSetAttribute(burst7_k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(burst7_k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (int kb : {6, 12, 24, 48, 96})
      for (int sync = 0; sync < 2; ++sync) {
        const int bvec = kb * 1024 / 16;
        float best = 1e30f;
        for (int it = 0; it < 5; ++it) {
          CK(cudaMemset(counter, 0, 4));
          cudaEventRecord(a);
          if (sync) burst7_k<1><<<grid, 512, kb * 1024>>>((const uint4*)src, (uint4*)m[0], (uint4*)m[1], (uint4*)m[2], (uint4*)bb[0], (uint4*)bb[1], (uint4*)bb[2], (uint4*)bb[3], in_bytes / 16, bvec, counter, sink);
          else burst7_k<0><<<grid, 512, kb * 1024>>>((const uint4*)src, (uint4*)m[0], (uint4*)m[1], (uint4*)m[2], (uint4*)bb[0], (uint4*)bb[1], (uint4*)bb[2], (uint4*)bb[3], in_bytes / 16, bvec, counter, sink);
          cudaEventRecord(b); CK(cudaEventSynchronize(b));
          float ms; cudaEventElapsedTime(&ms, a, b); if (it > 0 && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        if (sync == 0) {
          CK(cudaFuncSetAttribute(burst7_tma_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
          float bt = 1e30f;
          for (int it = 0; it < 5; ++it) {
            cudaEventRecord(a);
            burst7_tma_k<<<grid, 512, kb * 1024>>>((const uint4*)src, (uint4*)m[0], (uint4*)m[1], (uint4*)m[2], (uint4*)bb[0], (uint4*)bb[1], (uint4*)bb[2], (uint4*)bb[3], in_bytes / 16, bvec);
            cudaEventRecord(b); CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b); if (it > 0 && ms < bt) bt = ms;
          }
          CK(cudaGetLastError());
          printf("7 streams, burst %3d KB/CTA  bytes via TMA store: %.3f ms  %.0f GB/s on 27 B/px  (%.1f Gpix/s)\n", kb, bt, 27.0 * npx / bt / 1e6, npx / bt / 1e6);
        }
        printf("7 streams, burst %3d KB/CTA  sync=%d : %.3f ms  %.0f GB/s on 27 B/px  (%.1f Gpix/s)\n", kb, sync, best, 27.0 * npx / best / 1e6, npx / best / 1e6);
      }
  }
  return 0;
}
