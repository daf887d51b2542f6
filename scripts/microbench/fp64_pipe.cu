// FP64 pipe microbenchmarks for B200 (developer tool): what does a DFMA really cost next to other instructions?
#include <cstdio>
#include <cuda_runtime.h>
#define CHAINS 8
template <int VARIANT>
__global__ void k(int iters, double* sink, const double* seed, int* isink) {
  __shared__ double sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = seed[threadIdx.x];
  __syncthreads();
  double a[CHAINS], b[CHAINS], c[CHAINS];
  int x[CHAINS];
  for (int i = 0; i < CHAINS; ++i) { a[i] = seed[i] + threadIdx.x * 1e-9; b[i] = 0.999999 + 1e-9 * seed[i + 8] * threadIdx.x; c[i] = 1e-7 * seed[i + 16] + threadIdx.x * 1e-12; x[i] = threadIdx.x + i; }
  const double m = 0.999999, cc = 1e-7;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (VARIANT == 0) a[i] = fma(a[i], m, cc);                 // 1 reg + uniform consts
      if (VARIANT == 1) a[i] = fma(a[i], b[i], c[i]);            // 3 distinct register operands
      if (VARIANT == 2) { a[i] = fma(a[i], m, cc); x[i] = x[i] * 3 + it; }          // + IMAD
      if (VARIANT == 3) { a[i] = fma(a[i], b[i], c[i]); x[i] = x[i] * 3 + it; }     // 3 regs + IMAD
      if (VARIANT == 4) { a[i] = fma(a[i], sm[(it + i) & 63], cc); }                // + LDS broadcast operand
      if (VARIANT == 5) { a[i] = fma(a[i], m, cc); x[i] = max(min(x[i] + it, 1023), -1021); }   // + 2 VIMNMX-ish
      if (VARIANT == 6) { a[i] = fma(a[i], b[(i + 1) % CHAINS], a[(i + 3) % CHAINS]); }   // 3 regs, shifting operands
    }
  }
  double s = 0; int xs = 0;
  for (int i = 0; i < CHAINS; ++i) { s += a[i]; xs += x[i]; }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  isink[blockIdx.x * blockDim.x + threadIdx.x] = xs;
}
template <int V>
void run(const char* name, int blocks, int threads, double* sink, double* seed, int* isink) {
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V><<<blocks, threads>>>(100, sink, seed, isink);
  float best = 1e9;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); k<V><<<blocks, threads>>>(iters, sink, seed, isink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double dfma = (double)blocks * threads * iters * CHAINS;
  printf("%-38s blocks=%4d threads=%4d : %7.3f ms  %6.2f TFLOP/s (DFMA only)\n", name, blocks, threads, best, 2 * dfma / (best * 1e-3) / 1e12);
}
int main() {
  double *sink, *seed; int* isink;
  cudaMalloc(&sink, 8 * 148 * 16 * 1024); cudaMalloc(&isink, 4 * 148 * 16 * 1024); cudaMalloc(&seed, 8 * 64);
  double h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0 + 0.001 * i; cudaMemcpy(seed, h, sizeof(h), cudaMemcpyHostToDevice);
  int cfgs[4][2] = {{148 * 8, 256}, {148, 512}, {148, 256}, {148, 128}};
  for (auto& c : cfgs) {
    run<0>("DFMA reg,uniform,uniform", c[0], c[1], sink, seed, isink);
    run<1>("DFMA 3 distinct regs", c[0], c[1], sink, seed, isink);
    run<6>("DFMA 3 regs shifting", c[0], c[1], sink, seed, isink);
    run<2>("DFMA(uniform) + IMAD", c[0], c[1], sink, seed, isink);
    run<3>("DFMA(3 regs) + IMAD", c[0], c[1], sink, seed, isink);
    run<4>("DFMA + LDS operand", c[0], c[1], sink, seed, isink);
    run<5>("DFMA + IADD + 2 VIMNMX", c[0], c[1], sink, seed, isink);
    printf("\n");
  }
  return 0;
}
