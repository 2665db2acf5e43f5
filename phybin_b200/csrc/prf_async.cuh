// prf_async.cuh -- shared-memory address, mbarrier and bulk-copy helpers shared by the matrix kernels.
#pragma once

#include "prf_common.cuh"

namespace prf {

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_addr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Waiting warps back off with nanosleep so that they leave the issue slots to the working warps.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t sleep_ns = 0) {
  while (!mbar_try(bar, parity)) {
    if (sleep_ns) __nanosleep(sleep_ns);
  }
}

}  // namespace prf
