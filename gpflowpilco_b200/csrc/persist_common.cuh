// Synchronisation primitives of the persistent rollout kernels (rollout_persist.cu).
//
// One cooperative launch runs a whole sweep of H steps.  The CTAs are warp specialised — warps 0..3 ("scalar group") run the
// per-rollout stages of a step, the remaining warps contract Psi2 tiles — and the two roles of ALL CTAs hand work to each other
// through monotone counters in global memory:
//   writer:  data stores ... ; group/role barrier ; one thread: __threadfence(); red.release / st.release on the counter
//   reader:  one thread spins with ld.acquire.gpu on the counter ; role barrier ; everybody reads the data past L1 (__ldcg)
// which is the message-passing pattern of the PTX memory model (release/acquire at gpu scope, barrier for CTA-level cumulativity) —
// the same one cooperative_groups::grid_group::sync() is built from, without making every CTA wait for every other.
// A wait that exceeds kSpinLimit clock cycles is a bug (or a lost CTA): the kernel traps instead of hanging the GPU.
#pragma once
#include "common.cuh"

namespace gpp {

constexpr long long kSpinLimit = 60LL * 2000000000LL;   // ~60 s at 2 GHz

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// spin until *flag >= target (one thread; callers follow with their role's barrier)
__device__ __forceinline__ void spin_wait_ge(const unsigned* flag, unsigned target) {
  if (ld_acquire_u32(flag) >= target) return;
  const long long t0 = clock64();
  unsigned ns = 20;
  while (ld_acquire_u32(flag) < target) {
    __nanosleep(ns);
    if (ns < 200) ns += 20;
    if (clock64() - t0 > kSpinLimit) {
      printf("gpp persistent rollout: CTA %d thread %d waited too long for flag %p >= %u (now %u)\n", (int)blockIdx.x, (int)threadIdx.x,
             (const void*)flag, target, ld_acquire_u32(flag));
      asm volatile("trap;");
    }
  }
}

// warp-group register reallocation (sm_90a+): the contraction warps run at <= 96 registers, which lets the scalar warps — serial
// D x D factorisations written for one thread — keep their ~200 live values in registers instead of spilling
#ifndef GPP_PERSIST_SETMAXNREG
#define GPP_PERSIST_SETMAXNREG 1
#endif
template <int REGS>
__device__ __forceinline__ void warpgroup_reg_inc() {
#if GPP_PERSIST_SETMAXNREG
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(REGS));
#endif
}
template <int REGS>
__device__ __forceinline__ void warpgroup_reg_dec() {
#if GPP_PERSIST_SETMAXNREG
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REGS));
#endif
}

template <int ID, int COUNT>
__device__ __forceinline__ void role_bar_sync() {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}

}  // namespace gpp
