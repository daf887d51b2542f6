// Microbenchmark of the SCALAR (pre-DMMA) k_contract consumer inner loop in isolation (developer tool; history of DESIGN.md §4.1).
// Same data layout as the real kernel; no producers, no barriers: measures the FP64 loop's own ceiling.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../gpflowpilco_b200/csrc/gpp_math.h"

#ifndef VARIANT
#define VARIANT 0
#endif

constexpr int D = 6, T = 128, RPT = 4, CS = 8;

__constant__ double kC[16];

template <int K>
__device__ __forceinline__ void exp_v1(double (&x)[K]) {   // constants from the constant bank, single low clamp
  const double MAGIC = 6755399441055744.0;
  double t[K], r[K], p[K], s2[K], q[K];
#pragma unroll
  for (int k = 0; k < K; ++k) t[k] = fma(x[k], kC[0], MAGIC);
#pragma unroll
  for (int k = 0; k < K; ++k) r[k] = t[k] - MAGIC;
#pragma unroll
  for (int k = 0; k < K; ++k) p[k] = fma(r[k], kC[1], x[k]);
#pragma unroll
  for (int k = 0; k < K; ++k) r[k] = fma(r[k], kC[2], p[k]);
#pragma unroll
  for (int k = 0; k < K; ++k) s2[k] = r[k] * r[k];
#pragma unroll
  for (int k = 0; k < K; ++k) { p[k] = fma(kC[13], s2[k], kC[11]); q[k] = fma(kC[14], s2[k], kC[12]); }
#pragma unroll
  for (int k = 0; k < K; ++k) { p[k] = fma(p[k], s2[k], kC[9]); q[k] = fma(q[k], s2[k], kC[10]); }
#pragma unroll
  for (int k = 0; k < K; ++k) { p[k] = fma(p[k], s2[k], kC[7]); q[k] = fma(q[k], s2[k], kC[8]); }
#pragma unroll
  for (int k = 0; k < K; ++k) { p[k] = fma(p[k], s2[k], kC[5]); q[k] = fma(q[k], s2[k], kC[6]); }
#pragma unroll
  for (int k = 0; k < K; ++k) { p[k] = fma(p[k], s2[k], 1.0); q[k] = fma(q[k], s2[k], 1.0); }
#pragma unroll
  for (int k = 0; k < K; ++k) p[k] = fma(q[k], r[k], p[k]);
#pragma unroll
  for (int k = 0; k < K; ++k) {
#if VARIANT == 2 || VARIANT == 4
    x[k] = p[k] + t[k];   // timing-only: no integer exponent insert
#else
    int e = max(gpp::lo_int(t[k]), -1021);
    x[k] = gpp::make_double(gpp::hi_int(p[k]) + (e << 20), gpp::lo_int(p[k]));
#endif
  }
}

__global__ void __launch_bounds__(512) k(int reps, int ncons_warps, int first_cons, const double* seed, double* sink, int diag) {
  extern __shared__ __align__(16) double smem[];
  double* Ct = smem;                 // [T][T]
  double* colbuf = Ct + T * T;       // [T][CS]
  double* rowbuf = colbuf + T * CS;  // [D+2][T]
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) Ct[i] = seed[i & 1023] * 1e-3;
  for (int i = threadIdx.x; i < T * CS; i += blockDim.x) colbuf[i] = -0.01 * seed[(i * 7) & 1023];
  for (int i = threadIdx.x; i < (D + 2) * T; i += blockDim.x) rowbuf[i] = 0.02 * seed[(i * 3) & 1023];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cwarp = warp - first_cons;
  if (cwarp < 0 || cwarp >= ncons_warps) return;
  const int cw = T / ncons_warps + (cwarp < T % ncons_warps ? 1 : 0);
  const int c0 = cwarp * (T / ncons_warps) + min(cwarp, T % ncons_warps);
  const double* ct = Ct + c0 * T + lane;
  double total = 0.0;
  for (int rep = 0; rep < reps; ++rep) {
    const double* rb = rowbuf + lane;
    double g[RPT][D], r[RPT], acc[RPT];
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
#pragma unroll
      for (int d = 0; d < D; ++d) g[q][d] = rb[d * T + 32 * q] + rep * 1e-9;
      r[q] = rb[D * T + 32 * q];
      acc[q] = 0.0;
    }
    const double* cb = colbuf + c0 * CS;
#pragma unroll 2
    for (int jj = 0; jj < cw; ++jj) {
      const double* c = cb + jj * CS;
      double zc[D];
#pragma unroll
      for (int d = 0; d < D; ++d) zc[d] = c[d];
      const double sj = c[D];
      double t[RPT];
#pragma unroll
      for (int q = 0; q < RPT; ++q) t[q] = r[q] + sj;
#pragma unroll
      for (int d = 0; d < D; ++d)
#pragma unroll
#if VARIANT == 3 || VARIANT == 4
        for (int q = 0; q < RPT; ++q) t[q] = fma(g[q][d], kC[d], t[q]);   // timing-only: 2 register operands
#else
        for (int q = 0; q < RPT; ++q) t[q] = fma(g[q][d], zc[d], t[q]);
#endif
#if VARIANT == 0
      gpp::fast_exp_n<RPT>(t);
#else
      exp_v1<RPT>(t);
#endif
      if (diag) {
#pragma unroll
        for (int q = 0; q < RPT; ++q) acc[q] = fma(t[q], ct[jj * T + 32 * q], acc[q]);
      } else {
        const double wj = c[D + 1];
#pragma unroll
        for (int q = 0; q < RPT; ++q) acc[q] = fma(t[q], wj, acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < RPT; ++q) total += acc[q];
  }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = total;
}

int main() {
  double h[1024];
  for (int i = 0; i < 1024; ++i) h[i] = 0.5 + (i * 37 % 101) / 101.0;
  double *seed, *sink;
  cudaMalloc(&seed, sizeof(h)); cudaMalloc(&sink, 8 * 148 * 512);
  cudaMemcpy(seed, h, sizeof(h), cudaMemcpyHostToDevice);
  double hc[16] = {1.4426950408889634074, -6.93147180369123816490e-01, -1.90821492927058770002e-10, 0, 0,
                   0x1.0000000000001p-1, 0x1.5555555555556p-3, 0x1.5555555553d63p-5, 0x1.11111111109b3p-7,
                   0x1.6c16c1788bd90p-10, 0x1.a01a01a7c41d5p-13, 0x1.a019b90d2ae7ap-16, 0x1.71de0dae63bb3p-19,
                   0x1.289185613a3d6p-22, 0x1.af38a9b0ec855p-26, 0};
  cudaMemcpyToSymbol(kC, hc, sizeof(hc));
  size_t smem = sizeof(double) * (T * T + T * CS + (D + 2) * T);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int reps = 400;
  int cfgs[4][2] = {{16, 0}, {12, 4}, {12, 0}, {8, 0}};
  for (int diag = 0; diag < 2; ++diag)
    for (auto& c : cfgs) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      k<<<148, 512, smem>>>(10, c[0], c[1], seed, sink, diag);
      float best = 1e9;
      for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<<<148, 512, smem>>>(reps, c[0], c[1], seed, sink, diag); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      double entries = 148.0 * reps * T * T;
      printf("variant %d diag=%d consumers=%2d: %7.3f ms  %.1f entries/us/SM  algorithmic %.2f TFLOP/s  (cycles/n-iter @1.96GHz: %.0f)\n",
             VARIANT, diag, c[0], best, entries / 148 / (best * 1e3), entries * 38 / (best * 1e-3) / 1e12, best * 1e-3 * 1.96e9 / reps);
    }
  return 0;
}
