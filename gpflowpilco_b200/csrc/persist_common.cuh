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

constexpr long long kSpinLimit = 4LL * 2000000000LL;    // ~4 s at 2 GHz

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

// Flag-in-data publication of one double (the "LL" protocol of collective libraries): the value travels as two 8-byte words
// {32 data bits, 32-bit tag}; an aligned 8-byte store is single-copy atomic, so a reader that sees the expected tag in both words
// has the data — no fence on either side.  Tags are the (1-based) step number; buffers start zeroed.
__device__ __forceinline__ void ll_store(unsigned long long* slot, double v, unsigned tag) {
  const unsigned long long w0 = ((unsigned long long)tag << 32) | (unsigned)__double2loint(v);
  const unsigned long long w1 = ((unsigned long long)tag << 32) | (unsigned)__double2hiint(v);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(slot), "l"(w0) : "memory");
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(slot + 1), "l"(w1) : "memory");
}
__device__ __forceinline__ double ll_load(const unsigned long long* slot, unsigned tag) {
  long long t0 = 0;
  for (;;) {
    unsigned long long w0, w1;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w0) : "l"(slot) : "memory");
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w1) : "l"(slot + 1) : "memory");
    if ((unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag) return __hiloint2double((int)(unsigned)w1, (int)(unsigned)w0);
    if (t0 == 0) t0 = clock64();
    __nanosleep(200);
    if (clock64() - t0 > kSpinLimit) {
      printf("gpp persistent rollout: CTA %d thread %d waited too long for value %p (tag %u)\n", (int)blockIdx.x, (int)threadIdx.x,
             (const void*)slot, tag);
      asm volatile("trap;");
    }
  }
}

// mbarrier wait with the same time-out policy as the global-memory spins (no printf here: these sit in the hot contraction roles)
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, unsigned parity) {
  if (mbar_test(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_test(bar, parity))
    if (clock64() - t0 > kSpinLimit) asm volatile("trap;");
}

// raw halves of a tagged value, for readers that issue the loads early and test the tags later
__device__ __forceinline__ void ll_load_raw(const unsigned long long* slot, unsigned long long& w0, unsigned long long& w1) {
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w0) : "l"(slot) : "memory");
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w1) : "l"(slot + 1) : "memory");
}
__device__ __forceinline__ void ll_load_raw_cg(const unsigned long long* slot, unsigned long long& w0, unsigned long long& w1) {
  asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");   // one 16-byte L2 load
}
__device__ __forceinline__ bool ll_ready(unsigned long long w0, unsigned long long w1, unsigned tag) {
  return (unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag;
}
__device__ __forceinline__ double ll_value(unsigned long long w0, unsigned long long w1) {
  return __hiloint2double((int)(unsigned)w1, (int)(unsigned)w0);
}

// barrier among COUNT threads that also AND-reduces a predicate (everybody learns whether everybody's words had arrived)
template <int ID, int COUNT>
__device__ __forceinline__ bool role_bar_and(bool pred) {
  unsigned out;
  __syncwarp();
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.and.pred p, %2, %3, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(out)
      : "r"((unsigned)pred), "n"(ID), "n"(COUNT)
      : "memory");
  return out != 0;
}

// HINT counters.  Thousands of threads spinning on tagged words would load the L2 with polling traffic (measured: the finalize
// stage's 128 threads x 64 groups polling 16-byte entries every ~0.1 us cost the contraction ~15 %).  Writers therefore also bump
// a per-rollout counter with a relaxed, fire-and-forget reduction (no fence: it orders nothing), ONE thread of the reader polls
// that counter with back-off, and only then does the group read the tagged words — which still carry the correctness: a word
// that has not landed when the counter says "all published" is simply polled a little longer.
__device__ __forceinline__ void hint_add(unsigned* counter) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
__device__ __forceinline__ void hint_wait_ge(const unsigned* counter, unsigned target) {
  unsigned ns = 64;
  const long long t0 = clock64();
  for (;;) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    if (v >= target) return;
    __nanosleep(ns);
    if (ns < 512) ns += 64;
    if (clock64() - t0 > kSpinLimit) return;        // only a hint: the tagged-word reads that follow carry their own time-out
  }
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
  __syncwarp();
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}

}  // namespace gpp
