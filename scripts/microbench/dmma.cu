// Does DMMA (mma.sync m8n8k4 f64) run beside the DFMA pipe on B200?  (developer tool)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int V>
__global__ void k(int iters, double* sink, const double* seed) {
  double a[8], acc[8][2];
  for (int i = 0; i < 8; ++i) { a[i] = seed[i] + threadIdx.x * 1e-9; acc[i][0] = 0; acc[i][1] = 0; }
  const double m = 0.999999, cc = 1e-7;
  double A = seed[3] * 1e-3, B = seed[5] * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (V == 0 || V == 2) a[i] = fma(a[i], m, cc);
      if (V == 1 || V == 2) if (i < 4) dmma(acc[i][0], acc[i][1], A, B);     // 4 DMMA (512 flop each) per 8 DFMA (64 flop each)
      if (V == 3) dmma(acc[i][0], acc[i][1], A, B);
    }
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + acc[i][0] + acc[i][1];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V>
void run(const char* name, double dfma_per_it, double dmma_per_it, double* sink, double* seed) {
  int iters = 20000, blocks = 148 * 4, threads = 256;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V><<<blocks, threads>>>(100, sink, seed);
  float best = 1e9;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); k<V><<<blocks, threads>>>(iters, sink, seed); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double thr = (double)blocks * threads * iters, warps = thr / 32;
  printf("%-28s %8.3f ms   DFMA %6.2f TFLOP/s   DMMA %6.2f TFLOP/s\n", name, best, 2 * thr * dfma_per_it / (best * 1e-3) / 1e12,
         512 * warps * dmma_per_it / (best * 1e-3) / 1e12);
}
int main() {
  double *sink, *seed;
  cudaMalloc(&sink, 8 * 148 * 4 * 256); cudaMalloc(&seed, 8 * 64);
  double h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0 + 0.001 * i; cudaMemcpy(seed, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("DFMA only (8/iter)", 8, 0, sink, seed);
  run<1>("DMMA only (4/iter)", 0, 4, sink, seed);
  run<3>("DMMA only (8/iter)", 0, 8, sink, seed);
  run<2>("DFMA 8 + DMMA 4 per iter", 8, 4, sink, seed);
  return 0;
}
