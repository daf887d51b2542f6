// Device-side generation of pathwise function draws (the set-up half of gpflow_sampling's decoupled sampler, called by
// upstream at gpflow_pilco/loops/pilco.py:281-284: fresh paths for every closure evaluation), written directly in the
// particle-minor layout the rollout kernel streams.  Randomness: Philox4x32-10 streams keyed by the GLOBAL particle
// index (philox.cuh), so any sharding of particles over launches or GPUs draws identical numbers.
//
//   w[l,i,s]  ~ N(0,1)                                                        (prior weights)
//   u_{s,l}   = q_mu_l + tril(q_sqrt_l) eps_{s,l}   (then Luu_l u if whitened)
//   v[l,:,s]  = (Kuu_l + jitter I)^-1 (u_{s,l} - Phi_l(Z_l) w_{s,l} - sqrt(jitter) xi_{s,l})     (canonical-basis update)
// The M x F by F x S product and the two triangular solves are plain FP64 library GEMM/TRSM calls (cuBLAS).
#include <cublas_v2.h>

#include "model.cuh"
#include "philox.cuh"

namespace gpp {

__global__ void k_philox_raw(uint64_t first, int count, uint32_t stream_id, uint64_t seed, uint32_t* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Philox4 w = philox4x32_10(first + i, stream_id, seed);
  for (int k = 0; k < 4; ++k) out[(size_t)i * 4 + k] = w.v[k];
}

__global__ void k_draw_basis(int L, int F, int D, uint64_t seed, double* omega, double* phase) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < L * F * D) omega[idx] = philox_normal((uint64_t)idx, STREAM_OMEGA, seed);
  if (idx < L * F) phase[idx] = 6.283185307179586476925 * philox_uniform((uint64_t)idx, STREAM_PHASE, seed);
}

__global__ void k_draw_x0(int S, uint64_t first, int Dx, const double* m0, const double* chol0, uint64_t seed, double* x0) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double n[8];
  for (int d = 0; d < Dx; ++d) n[d] = philox_normal((first + s) * (uint64_t)Dx + d, STREAM_X0, seed);
  for (int i = 0; i < Dx; ++i) {
    double v = m0[i];
    for (int k = 0; k <= i; ++k) v = fma(chol0[i * Dx + k], n[k], v);
    x0[(size_t)s * Dx + i] = v;
  }
}

// out[row, s] = normal(((first+s) L + l) K + row) for row < K rows, s < S; zero padding up to ldS
__global__ void k_draw_rows(int rows, int K, int L, int l, int S, int ldS, uint64_t first, uint32_t stream_id, uint64_t seed, double* out) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  int row = blockIdx.y;
  if (s >= ldS) return;
  double v = 0.0;
  if (s < S && row < rows) v = philox_normal(((first + s) * (uint64_t)L + l) * (uint64_t)K + row, stream_id, seed);
  out[(size_t)row * ldS + s] = v;
}

// E[j,s] = q_mu[j,l] + E[j,s]   (E holds tril(q_sqrt) eps)
__global__ void k_add_qmu(int M, int L, int l, int ldS, const double* q_mu, double* E) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y;
  if (s < ldS) E[(size_t)j * ldS + s] += q_mu[(size_t)j * L + l];
}

// E[j,s] -= sqrt(jitter) xi
__global__ void k_sub_noise(int M, int L, int l, int S, int ldS, uint64_t first, uint64_t seed, double sq_jitter, double* E) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y;
  if (s < S) E[(size_t)j * ldS + s] -= sq_jitter * philox_normal(((first + s) * (uint64_t)L + l) * (uint64_t)M + j, STREAM_UPDATE_XI, seed);
}

// PhiZ[j,f] = sqrt(2 var/F) cos(omega_f . (z_j/ell) + b_f)
__global__ void k_phi_z(int M, int F, int D, const double* Z, const double* ell, double var, const double* omega, const double* phase,
                        double* PhiZ) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y;
  if (f >= F) return;
  double a = phase[f];
  for (int d = 0; d < D; ++d) a = fma(omega[(size_t)f * D + d], Z[(size_t)j * D + d] / ell[d], a);
  PhiZ[(size_t)j * F + f] = sqrt(2.0 * var / F) * cos(a);
}

}  // namespace gpp

extern "C" {

int gpp_philox_raw(unsigned long long first_index, int count, unsigned stream_id, unsigned long long seed, unsigned* out, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(out && count >= 0, GPP_ERR_NULL, "gpp_philox_raw: bad arguments");
  if (count == 0) return GPP_OK;
  gpp::k_philox_raw<<<(count + 127) / 128, 128, 0, (cudaStream_t)stream>>>(first_index, count, stream_id, seed, out);
  gpp::count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

int gpp_pathwise_draw_basis(int L, int F, int D, unsigned long long seed, double* omega, double* phase, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(omega && phase, GPP_ERR_NULL, "gpp_pathwise_draw_basis: null argument");
  int n = L * F * D;
  gpp::k_draw_basis<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(L, F, D, seed, omega, phase);
  gpp::count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

int gpp_pathwise_draw_x0(int S, unsigned long long first_particle, int Dx, const double* m0, const double* chol0,
                         unsigned long long seed, double* x0, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(m0 && chol0 && x0, GPP_ERR_NULL, "gpp_pathwise_draw_x0: null argument");
  GPP_REQUIRE(S >= 1 && Dx >= 1 && Dx <= 8, GPP_ERR_BAD_SHAPE, "gpp_pathwise_draw_x0: bad sizes");
  gpp::k_draw_x0<<<(S + 127) / 128, 128, 0, (cudaStream_t)stream>>>(S, first_particle, Dx, m0, chol0, seed, x0);
  gpp::count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

size_t gpp_pathwise_generate_workspace_bytes(const gpp_gp_model* model, int ldS, int F) {
  if (!model) return 0;
  return sizeof(double) * ((size_t)model->M * ldS + (size_t)model->M * F) + 512;
}

int gpp_pathwise_generate(gpp_gp_model* model, int S, int ldS, unsigned long long first_particle, int F, int Mpad,
                          unsigned long long seed, const double* omega, const double* phase, double* w, double* v,
                          void* workspace, size_t workspace_bytes, void* stream_) {
  GPP_NVTX_RANGE();
  using namespace gpp;
  GPP_REQUIRE(model && omega && phase && w && v && workspace, GPP_ERR_NULL, "gpp_pathwise_generate: null argument");
  const int L = model->L, M = model->M, D = model->D;
  GPP_REQUIRE(S >= 1 && ldS >= S && Mpad >= M && F >= 1, GPP_ERR_BAD_SHAPE, "gpp_pathwise_generate: bad sizes");
  GPP_REQUIRE(workspace_bytes >= gpp_pathwise_generate_workspace_bytes(model, ldS, F), GPP_ERR_WORKSPACE, "gpp_pathwise_generate: workspace too small");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!model->blas) {
    cublasHandle_t h;
    GPP_REQUIRE(cublasCreate(&h) == CUBLAS_STATUS_SUCCESS, GPP_ERR_CUDA, "gpp_pathwise_generate: cublasCreate failed");
    model->blas = h;
  }
  cublasHandle_t blas = (cublasHandle_t)model->blas;
  GPP_REQUIRE(cublasSetStream(blas, stream) == CUBLAS_STATUS_SUCCESS, GPP_ERR_CUDA, "gpp_pathwise_generate: cublasSetStream failed");
  double* E = (double*)workspace;               // [M, ldS]
  double* PhiZ = E + (size_t)M * ldS;           // [M, F]
  const double one = 1.0, minus_one = -1.0;
  const size_t mm = (size_t)M * M;
  dim3 blk(128);
#define GPP_BLAS(expr) GPP_REQUIRE((expr) == CUBLAS_STATUS_SUCCESS, GPP_ERR_CUDA, "gpp_pathwise_generate: cuBLAS call failed: " #expr)
  for (int l = 0; l < L; ++l) {
    double* wl = w + (size_t)l * F * ldS;
    double* vl = v + (size_t)l * Mpad * ldS;
    const double* Lu = model->Luu + (size_t)l * mm;     // row-major lower == column-major upper U, Kuu = U^T U
    k_draw_rows<<<dim3((ldS + 127) / 128, F), blk, 0, stream>>>(F, F, L, l, S, ldS, first_particle, STREAM_PRIOR_W, seed, wl);
    k_draw_rows<<<dim3((ldS + 127) / 128, M), blk, 0, stream>>>(M, M, L, l, S, ldS, first_particle, STREAM_U_EPS, seed, vl);
    // E = tril(q_sqrt) eps  (column-major: E_cm = eps_cm * Q_c with Q_c the column-major view of the row-major lower factor)
    GPP_BLAS(cublasDtrmm(blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, ldS, M, &one,
                         model->q_sqrt + (size_t)l * mm, M, vl, ldS, E, ldS));
    k_add_qmu<<<dim3((ldS + 127) / 128, M), blk, 0, stream>>>(M, L, l, ldS, model->q_mu, E);
    if (model->whiten)   // u <- Luu u
      GPP_BLAS(cublasDtrmm(blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, ldS, M, &one, Lu, M, E, ldS, E, ldS));
    k_phi_z<<<dim3((F + 127) / 128, M), blk, 0, stream>>>(M, F, D, model->Z + (size_t)l * M * D, model->ell + (size_t)l * D,
                                                          model->h_var[l], omega + (size_t)l * F * D, phase + (size_t)l * F, PhiZ);
    // E -= PhiZ W_l      (column-major: E_cm[S,M] -= W_cm[S,F] * PhiZ_cm[F,M])
    GPP_BLAS(cublasDgemm(blas, CUBLAS_OP_N, CUBLAS_OP_N, ldS, M, F, &minus_one, wl, ldS, PhiZ, F, &one, E, ldS));
    k_sub_noise<<<dim3((ldS + 127) / 128, M), blk, 0, stream>>>(M, L, l, S, ldS, first_particle, seed, sqrt(model->h_jitter[l]), E);
    // v = Kuu^-1 E = L^-T L^-1 E   (column-major: X U = E, then Y U^T = X)
    GPP_BLAS(cublasDtrsm(blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, ldS, M, &one, Lu, M, E, ldS));
    GPP_BLAS(cublasDtrsm(blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, ldS, M, &one, Lu, M, E, ldS));
    GPP_CUDA_OK(cudaMemcpyAsync(vl, E, sizeof(double) * (size_t)M * ldS, cudaMemcpyDeviceToDevice, stream));
    if (Mpad > M) GPP_CUDA_OK(cudaMemsetAsync(vl + (size_t)M * ldS, 0, sizeof(double) * (size_t)(Mpad - M) * ldS, stream));
    count_launch(9);
  }
#undef GPP_BLAS
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

}  // extern "C"
