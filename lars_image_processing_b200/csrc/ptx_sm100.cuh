// Thin inline-PTX wrappers for the sm_100a features the path uses: mbarrier, 1-D TMA bulk
// copies (cp.async.bulk, SASS UBLKCP) in both directions, proxy fences, named barriers and
// cache-hinted vector loads/stores.  No tensor maps are needed: frames are flat 1-D arrays.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lars {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t tx_bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx_bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// Wait of a producer / drain warp that is NOT on the consumers' critical path: let the hardware park
// the thread (suspend-time hint) and back off between polls, so the poll loop does not take issue
// slots from the warps doing the arithmetic (ncu: the two tight poll loops of the fused pass were
// 26 % of all executed warp instructions).
#ifndef LARS_RELAXED_WAIT_NS
#define LARS_RELAXED_WAIT_NS 256
#endif
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
#if LARS_RELAXED_WAIT_NS < 0
  mbar_wait(bar, parity);   // tuning variant: the original tight poll
  return;
#endif
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(4096u)
        : "memory");
    if (done) return;
#if LARS_RELAXED_WAIT_NS > 0
    __nanosleep(LARS_RELAXED_WAIT_NS);
#endif
  }
}

// ---- TMA 1-D bulk copies ----------------------------------------------------------------
// global -> shared, completion signalled on an mbarrier (complete_tx).  16-byte aligned
// addresses, size a multiple of 16.
__device__ __forceinline__ void tma_load_1d(uint32_t dst_smem, const void* src_gmem, uint32_t bytes,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          dst_smem),
      "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(bar)
      : "memory");
}
// Same with an L2 eviction-priority hint (createpolicy result).
__device__ __forceinline__ void tma_load_1d_hint(uint32_t dst_smem, const void* src_gmem,
                                                 uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
      "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group.
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                   __cvta_generic_to_global(dst_gmem)),
               "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// Wait until all committed bulk groups of this thread have finished READING shared memory.
__device__ __forceinline__ void tma_store_wait_read_all() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// Wait until they are complete (global writes performed).
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// Make generic-proxy shared-memory writes visible to the async proxy (before a bulk store).
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ---- named barrier among a subset of the CTA's warps ------------------------------------
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// ---- streaming vector access -------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_v4(void* p, float a, float b, float c, float d) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// 128-bit load / store with an explicit L2 eviction policy (createpolicy result)
__device__ __forceinline__ uint4 ldg_v4_policy(const void* p, uint64_t policy) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(policy));
  return r;
}
__device__ __forceinline__ void stg_v4_policy(void* p, float a, float b, float c, float d, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d),
               "l"(policy)
               : "memory");
}
__device__ __forceinline__ void tma_store_1d_hint(void* dst_gmem, uint32_t src_smem, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(
                   __cvta_generic_to_global(dst_gmem)),
               "r"(src_smem), "r"(bytes), "l"(policy)
               : "memory");
}

// Bulk prefetch of a contiguous global range into L2 (no shared-memory destination).
__device__ __forceinline__ void l2_prefetch_bulk(const void* src_gmem, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(__cvta_generic_to_global(src_gmem)),
               "r"(bytes), "l"(policy)
               : "memory");
}

// PRMT byte permute: picks 4 bytes out of {a (bytes 0-3), b (bytes 4-7)}.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  return __byte_perm(a, b, sel);
}

}  // namespace lars
