// Internal helpers shared by the translation units of libgpp_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/gpp_b200.h"
#include "gpp_math.h"

namespace gpp {

// ---- host-side error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(unsigned long long n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define GPP_CUDA_OK(expr)                                                                        \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess) {                                                                    \
      gpp::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return GPP_ERR_CUDA;                                                                       \
    }                                                                                            \
  } while (0)

#define GPP_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      gpp::set_error(__VA_ARGS__);        \
      return (code);                      \
    }                                     \
  } while (0)

// NVTX range around every compute entry point of the C ABI (visible in Nsight Systems / ncu --nvtx as gpp_* ranges); header-only
// nvtx3: a no-op costing one indirect call when no tool is attached
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define GPP_NVTX_RANGE() gpp::NvtxRange gpp_nvtx_range__(__func__)

// profiling hooks (api_common.cu)
void profile_begin(cudaStream_t stream);
void profile_end(cudaStream_t stream);

// Per-DEVICE caches: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count belong to the device that is current when
// they are set / read, and one process may drive several devices (a handle lives on whichever device its tensors are on).
constexpr int kMaxDevices = 64;
inline int current_device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
inline int num_sms() {
  static int n[kMaxDevices] = {};
  const int d = current_device_slot();
  if (!n[d]) {
    cudaDeviceGetAttribute(&n[d], cudaDevAttrMultiProcessorCount, d);
    if (n[d] <= 0) n[d] = 148;
  }
  return n[d];
}
// largest dynamic shared-memory size a kernel has been opted into, per device: `raise(bytes)` is true when the opt-in must be (re)done
struct PerDeviceSmemOptIn {
  size_t bytes[kMaxDevices] = {};
  bool raise(size_t want) {
    const int d = current_device_slot();
    if (want <= bytes[d]) return false;
    bytes[d] = want;
    return true;
  }
};

// ---- device-side helpers ------------------------------------------------------------------------------
#if defined(__CUDACC__)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// The per-rollout ("scalar") stages of the rollouts are written for a GROUP of 128 threads = warps 0..3 of the CTA: threadIdx.x is
// the index inside the group and the group synchronises on named barrier 6.  In the stand-alone kernels the CTA is exactly one
// group (bar.sync 6, 128 == __syncthreads()); in the persistent rollout kernels (rollout_persist.cu) the same code runs on the
// scalar warps of a warp-specialised CTA while the other warps contract, so it must never use a CTA-wide barrier.
constexpr int kGroupThreads = 128;
// (bar.sync is warp-ALIGNED: every thread of a warp must execute it together.  The compiler reconverges a warp before
// __syncthreads(), but an inline-asm barrier is opaque to it — after a divergent section such as `if (tid == 0) spin();` nothing
// guarantees reconvergence, and a partially arrived warp hangs the barrier.  Every asm barrier in this library is therefore
// preceded by __syncwarp().)
__device__ __forceinline__ void group_sync() {
  __syncwarp();
  asm volatile("bar.sync 6, 128;" ::: "memory");
}

__device__ __forceinline__ void flag_not_pd(int* info, int index) {
  if (info) atomicCAS(info, 0, index + 1);   // first writer wins; 0 means "all fine"
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
#endif

// ---- per-(input, kernel pair) coefficients of  log Q_ij = r_i + s_j + z1'_i^T R z2'_j  ------------------
//  (re-derived in coordinates centred at mu_n so that both quadratic forms are negative
//   semi-definite and nothing cancels:  with V = V1 V2/(V1+V2), A1 = V/V1, A2 = V/V2 (A1 + A2 = I), G = (Sigma+V)^-1
//     log Q_ij = c0 - 1/2 (z1-z2)^T (V1+V2)^-1 (z1-z2) - 1/2 (A1 z1' + A2 z2')^T G (A1 z1' + A2 z2'),  z' = z - mu
//   which equals kernel_expectation.py:125-187 term by term.)
template <int D>
struct PairPack {
  static constexpr int TRI = D * (D + 1) / 2;
  static constexpr int R = 0;            // [D][D]   cross matrix
  static constexpr int P1 = D * D;       // packed upper triangle, off-diagonals doubled, -1/2 folded in
  static constexpr int P2 = P1 + TRI;
  static constexpr int C0 = P2 + TRI;    // log(var1 var2) + 1/2 sum log V - sum log diag chol(Sigma+V)
  static constexpr int MU = C0 + 1;      // [D]
  static constexpr int SIZE = ((MU + D + 1) / 2) * 2;   // doubles, even => 16-byte multiple
};

template <int D>
GPP_HD bool make_pair_pack(const double* mu, const double* Sigma, const double* V1, const double* V2,
                           double log_amp, double* out, double* G_out = nullptr /* optional (Sigma + V)^-1, [D][D] */) {
  using PP = PairPack<D>;
  Mat<D> S, Li, G;
  double A1[D], A2[D], dg[D];
  double v_prod = 1.0;                       // log of a product instead of D logs (D <= 8 factors of moderate size)
#pragma unroll
  for (int d = 0; d < D; ++d) {
    double s12 = V1[d] + V2[d];
    dg[d] = 1.0 / s12;
    A1[d] = V2[d] * dg[d];
    A2[d] = V1[d] * dg[d];
    double V = V1[d] * A1[d];
    v_prod *= V;
#pragma unroll
    for (int e = 0; e < D; ++e) S(d, e) = Sigma[d * D + e] + (d == e ? V : 0.0);
  }
  bool ok = cholesky<D>(S);
  double diag_prod = 1.0;
#pragma unroll
  for (int d = 0; d < D; ++d) diag_prod *= S(d, d);
  tri_inverse<D>(S, Li);
  gram_inverse<D>(Li, G);
  if (G_out) {
#pragma unroll
    for (int d = 0; d < D * D; ++d) G_out[d] = G.a[d];
  }
#pragma unroll
  for (int d = 0; d < D; ++d)
#pragma unroll
    for (int e = 0; e < D; ++e)
      out[PP::R + d * D + e] = (d == e ? dg[d] : 0.0) - A1[d] * G(d, e) * A2[e];
  int t = 0;
#pragma unroll
  for (int d = 0; d < D; ++d)
#pragma unroll
    for (int e = d; e < D; ++e, ++t) {
      double g1 = A1[d] * G(d, e) * A1[e], g2 = A2[d] * G(d, e) * A2[e];
      out[PP::P1 + t] = (d == e) ? -0.5 * (dg[d] + g1) : -g1;
      out[PP::P2 + t] = (d == e) ? -0.5 * (dg[d] + g2) : -g2;
    }
  out[PP::C0] = log_amp + log(sqrt(v_prod) / diag_prod);
#pragma unroll
  for (int d = 0; d < D; ++d) out[PP::MU + d] = mu[d];
  if (PP::MU + D < PP::SIZE) out[PP::MU + D] = 0.0;
  return ok;
}

// quadratic form with packed upper-triangular coefficients (off-diagonals pre-doubled)
template <int D>
GPP_HD double packed_quad(const double* P, const double* z) {
  double acc = 0.0;
  int t = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    double row = 0.0;
#pragma unroll
    for (int e = d; e < D; ++e, ++t) row = fma_(P[t], z[e], row);
    acc = fma_(row, z[d], acc);
  }
  return acc;
}

// ---- shared-memory mbarriers (phase-tracked hand-over between warp roles; TMA completion in pathwise.cu) -----------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// try_wait without blocking: true when the phase with the given parity has completed
__device__ __forceinline__ bool mbar_test(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

}  // namespace gpp
