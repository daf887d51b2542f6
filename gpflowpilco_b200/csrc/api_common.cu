// Error reporting, version and launch accounting for libgpp_b200.so.
#include "common.cuh"

namespace gpp {

static thread_local char t_error[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

static bool g_profile = false;
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;

void profile_begin(cudaStream_t stream) {
  if (g_profile) cudaEventRecord(g_ev0, stream);
}
void profile_end(cudaStream_t stream) {
  if (g_profile) cudaEventRecord(g_ev1, stream);
}

__global__ void k_microbench_fp64(int iters, double* sink) {
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
  double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
  const double m = 0.999999, c = 1e-7;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace gpp

extern "C" {

int gpp_profile_enable(int on) {
  if (on && !gpp::g_ev0) {
    GPP_CUDA_OK(cudaEventCreate(&gpp::g_ev0));
    GPP_CUDA_OK(cudaEventCreate(&gpp::g_ev1));
  }
  gpp::g_profile = on != 0;
  return GPP_OK;
}

int gpp_profile_last_ms(float* ms) {
  GPP_REQUIRE(ms && gpp::g_ev0, GPP_ERR_NULL, "gpp_profile_last_ms: profiling was never enabled");
  GPP_CUDA_OK(cudaEventSynchronize(gpp::g_ev1));
  GPP_CUDA_OK(cudaEventElapsedTime(ms, gpp::g_ev0, gpp::g_ev1));
  return GPP_OK;
}

int gpp_microbench_fp64(int blocks, int threads, int iters, double* sink, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(sink && blocks > 0 && threads > 0 && threads <= 1024 && iters > 0, GPP_ERR_BAD_SHAPE, "gpp_microbench_fp64: bad arguments");
  gpp::k_microbench_fp64<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink);
  gpp::count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

int gpp_version(void) { return 100; }

const char* gpp_last_error(void) { return gpp::t_error; }

unsigned long long gpp_launch_count(void) { return gpp::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
