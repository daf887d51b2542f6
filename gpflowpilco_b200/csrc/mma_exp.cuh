// Device helpers shared by the two Psi2 kernels (contract_kernel.cuh: fused contraction, psi.cu: materialised tensor):
// the FP64 tensor-core step that evaluates the exponent of an 8 x 8 block of entries, and the table-driven exp that follows it.
#pragma once
#include "common.cuh"

namespace gpp {

// D (8x8) += A (8x4, row) * B (4x8, col), FP64: lane l holds A[l>>2][l&3], B[l&3][l>>2], D[l>>2][2(l&3) + {0,1}]
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// exp for the Psi2 loops: the 256-entry table algorithm of gpp_math.h (exp_tab256_ref: 8 FP64-pipe ops), with the integer
// tail written for the issue port (every non-FP64 instruction issued beside the FP64 pipe costs ~0.8 cycles, fp64_pipe.cu):
//   * the argument is clamped to >= -707 by an unsigned min on its high word (1 op) instead of selecting 0 afterwards
//     (4 ops): entries below exp(-707) ~ 1e-307 are numerically zero in the contraction either way;
//   * table address = lane base + ((n & 255) << 7) and exponent insert = hi + ((n >> 8) << 20), 2 ops each.
// `tab_addr` is the shared-space byte address of this lane's table replica (etab + (lane & (REP-1))); entry j of replica c lives at
// etab[j * REP + c]: with REP = 16 the 64-bit loads of a half-warp hit 16 distinct bank pairs whatever the j's are (REP = 8,
// used when D >= 7 needs the shared memory for a third k-step, allows 2-way conflicts).
constexpr int kContractTab = 256;

template <int K, int REP>
__device__ __forceinline__ void exp_tab_contract(double (&x)[K], unsigned tab_addr) {
  constexpr int SHIFT = REP == 16 ? 7 : 6;      // entry stride in bytes: 8 * REP
  static_assert(REP == 16 || REP == 8, "table replication");
  const double MAGIC = 6755399441055744.0;
  double t[K], r[K], q[K], tj[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    unsigned hi = (unsigned)__double2hiint(x[k]);
    hi = min(hi, 0xC0861800u);                                   // x >= -707 (positive x: hi < 2^31, untouched)
    x[k] = __hiloint2double((int)hi, __double2loint(x[k]));
  }
#pragma unroll
  for (int k = 0; k < K; ++k) t[k] = fma(x[k], kExpT256[0], MAGIC);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    unsigned addr;
    asm("{\n\t.reg .b32 m;\n\tand.b32 m, %1, 255;\n\tshl.b32 m, m, %3;\n\tadd.u32 %0, m, %2;\n\t}" : "=r"(addr) : "r"(__double2loint(t[k])), "r"(tab_addr), "n"(SHIFT));
    asm("ld.shared.f64 %0, [%1];" : "=d"(tj[k]) : "r"(addr));
  }
#pragma unroll
  for (int k = 0; k < K; ++k) r[k] = t[k] - MAGIC;
#pragma unroll
  for (int k = 0; k < K; ++k) r[k] = fma(r[k], kExpT256[1], x[k]);
#pragma unroll
  for (int k = 0; k < K; ++k) q[k] = fma(r[k], kExpT256[4], kExpT256[3]);
#pragma unroll
  for (int k = 0; k < K; ++k) q[k] = fma(q[k], r[k], kExpT256[2]);
#pragma unroll
  for (int k = 0; k < K; ++k) q[k] = fma(q[k], r[k], 1.0);
#pragma unroll
  for (int k = 0; k < K; ++k) q[k] = q[k] * r[k];                 // exp(r) - 1
#pragma unroll
  for (int k = 0; k < K; ++k) q[k] = fma(tj[k], q[k], tj[k]);      // T_j exp(r)
#pragma unroll
  for (int k = 0; k < K; ++k) {
    int hi2;
    asm("{\n\t.reg .s32 e;\n\tshr.s32 e, %1, 8;\n\tmad.lo.s32 %0, e, 1048576, %2;\n\t}" : "=r"(hi2) : "r"(__double2loint(t[k])), "r"(__double2hiint(q[k])));
    x[k] = __hiloint2double(hi2, __double2loint(q[k]));
  }
}

// extended row / column vectors of the one-inner-product form  log Q_ij = A_i . B_j :
//   A_i = [R^T z1'_i (D), c0 + z1'^T P1 z1', 1, 0..],  B_j = [z2'_j (D), 1, z2'^T P2 z2', 0..],  length 4 KS
template <int D>
struct ExtLayout {
  static constexpr int KS = (D + 2 + 3) / 4;        // k-steps of 4
};

}  // namespace gpp
