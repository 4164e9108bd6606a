// fg_async.cuh - PTX wrappers shared by the kernels that use the asynchronous copy engine: mbarrier, bulk copy
// (cp.async.bulk = TMA engine without a tensor map, SASS UBLKCP), proxy fences, acquire / release flags.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fg {

// PTX wrappers: mbarrier, bulk copy (TMA engine, SASS UBLKCP), acquire / release flags
__device__ __forceinline__ uint32_t rq_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rq_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rq_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void rq_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rq_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool rq_mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      " selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(ok)
      : "r"(rq_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void rq_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   rq_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(rq_smem_u32(bar))
               : "memory");
}
// shared -> global bulk copy tracked by the issuing thread's bulk groups (16-byte aligned, size a multiple of 16)
__device__ __forceinline__ void rq_bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(rq_smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void rq_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's bulk groups may still be READING their shared-memory source
template <int N>
__device__ __forceinline__ void rq_bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void rq_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// shared-memory only: orders this thread's generic-proxy accesses to the tile before a bulk copy into it, without
// waiting for the thread's outstanding global stores (the R blocks)
__device__ __forceinline__ void rq_fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// read-once input streams: L2 evict-first, so that they do not displace the ring and the factors
__device__ __forceinline__ unsigned long long rq_policy_evict_first() {
  unsigned long long p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long rq_policy_evict_last() {
  unsigned long long p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// 16-byte store that asks L2 to keep the line (rows of W that are still being completed by later patches)
__device__ __forceinline__ void rq_st_keep(double2* p, double2 v) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(rq_policy_evict_last())
               : "memory");
}
__device__ __forceinline__ double rq_ldg_stream(const double* p) {
  double v;
  asm("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(rq_policy_evict_first()));
  return v;
}
__device__ __forceinline__ unsigned rq_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void rq_st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}


}  // namespace fg
