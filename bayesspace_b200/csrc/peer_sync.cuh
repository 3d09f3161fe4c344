// peer_sync.cuh -- system-scope flag primitives of the peer transport (device side), shared by the stand-alone
// exchange kernels (comm.cpp) and by the sweep kernel's fused halo exchange (sweep_fast.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "comm.hpp"

namespace bsb {

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait for a peer's epoch.  Once any error is latched nobody waits any more (flags are still sent), so a
// failed rank drains the captured graph instead of stalling its peers for one timeout per wait.
__device__ __forceinline__ void wait_flag(const uint32_t *flag, uint32_t epoch, int *err, unsigned long long timeout_ns) {
  if ((int32_t) (ld_acquire_sys(flag) - epoch) >= 0) return;
  if (*(volatile int *) err) return;
  const unsigned long long t0 = global_ns();
  while ((int32_t) (ld_acquire_sys(flag) - epoch) < 0) {
    __nanosleep(32);
    if (global_ns() - t0 > timeout_ns) {
      atomicCAS(err, 0, BSB_DEVERR_PEER_TIMEOUT);
      return;
    }
  }
}

}   // namespace bsb
