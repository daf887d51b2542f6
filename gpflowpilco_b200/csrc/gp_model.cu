// GP model handle: caches the step-invariant solves of the moment-matching rules.
//
// The reference recomputes Kuu, its Cholesky and the Luu^-1 {q_mu, q_sqrt} solves in EVERY moment_matching call
// (gpflow_pilco/moment_matching/models.py:145-157 single output, :216-235 multi output, :66-75 GPR), although they
// do not depend on the input distribution.  Here they are done once per model with cuSOLVER/cuBLAS (plain
// library Cholesky / triangular solves, FP64) and folded into
//     beta_l = Kuu_l^-1 m_l                                  (models.py:235)
//     C_l    = beta_l beta_l^T - [unc] Luu^-T (I - R R^T) Luu^-1,  R = Luu^-1 tril(q_sqrt)  (R = tril(q_sqrt) if whitened)
// so that   f2_ab = beta_a^T Q_ab beta_b   and   Sff_aa - var_a + f1_a^2 = sum_ij Q_aa[i,j] C_a[i,j]   (models.py:245-259).
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <algorithm>

#include "model.cuh"

namespace {

__global__ void kuu_kernel(const double* __restrict__ Z, const double* __restrict__ ell, double var, double jitter,
                           int M, int D, double* __restrict__ K) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j >= M) return;
  double acc = 0.0;
  for (int d = 0; d < D; ++d) {
    double t = (Z[i * D + d] - Z[j * D + d]) / ell[d];
    acc = fma(t, t, acc);
  }
  K[(size_t)i * M + j] = var * exp(-0.5 * acc) + (i == j ? jitter : 0.0);
}

__global__ void tril_kernel(const double* __restrict__ src, int M, double* __restrict__ dst) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j >= M) return;
  dst[(size_t)i * M + j] = (src != nullptr && j <= i) ? src[(size_t)i * M + j] : 0.0;
}

__global__ void eye_kernel(int M, double* __restrict__ dst) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j >= M) return;
  dst[(size_t)i * M + j] = (i == j) ? 1.0 : 0.0;
}

__global__ void gather_col_kernel(const double* __restrict__ q_mu, int M, int L, int l, double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) out[i] = q_mu[(size_t)i * L + l];
}

// C = beta beta^T - unc * B   (B symmetrised on the fly to wash out trsm round-off asymmetry)
__global__ void finish_c_kernel(const double* __restrict__ beta, const double* __restrict__ B, int M, int unc,
                                double* __restrict__ C) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j >= M) return;
  double v = beta[i] * beta[j];
  if (unc) v -= 0.5 * (B[(size_t)i * M + j] + B[(size_t)j * M + i]);
  C[(size_t)i * M + j] = v;
}

__global__ void clear_upper_kernel(int M, double* __restrict__ Lmat) {   // keep row-major lower triangle only
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j < M && j > i) Lmat[(size_t)i * M + j] = 0.0;
}

void build_slots(gpp_gp_model::SlotTable& tab, int L, int M, int tile, bool diag_only) {
  tab.tile = tile;
  int nt = (M + tile - 1) / tile;
  std::vector<int> pair_start, pair_ab;
  int pair = 0;
  for (int a = 0; a < L; ++a)
    for (int b = a; b < L; ++b) {
      if (diag_only && a != b) continue;
      pair_start.push_back((int)tab.h_slots.size());
      pair_ab.push_back(a);
      pair_ab.push_back(b);
      for (int ti = 0; ti < nt; ++ti)
        for (int tj = (a == b ? ti : 0); tj < nt; ++tj) {
          gpp_slot s;
          s.pair = pair; s.a = a; s.b = b; s.ti = ti; s.tj = tj;
          s.weight = (a == b && tj > ti) ? 2.0 : 1.0;
          tab.h_slots.push_back(s);
        }
      ++pair;
    }
  pair_start.push_back((int)tab.h_slots.size());
  tab.npairs = pair;
  tab.nslots = (int)tab.h_slots.size();
  cudaMalloc(&tab.d_slots, sizeof(gpp_slot) * tab.nslots);
  cudaMemcpy(tab.d_slots, tab.h_slots.data(), sizeof(gpp_slot) * tab.nslots, cudaMemcpyHostToDevice);
  cudaMalloc(&tab.d_pair_start, sizeof(int) * pair_start.size());
  cudaMemcpy(tab.d_pair_start, pair_start.data(), sizeof(int) * pair_start.size(), cudaMemcpyHostToDevice);
  cudaMalloc(&tab.d_pair_ab, sizeof(int) * pair_ab.size());
  cudaMemcpy(tab.d_pair_ab, pair_ab.data(), sizeof(int) * pair_ab.size(), cudaMemcpyHostToDevice);
}

#define GPP_LIB_OK(expr, what)                                                     \
  do {                                                                             \
    int s__ = (int)(expr);                                                         \
    if (s__ != 0) {                                                                \
      gpp::set_error("%s:%d %s failed with status %d", __FILE__, __LINE__, what, s__); \
      status = GPP_ERR_CUDA;                                                       \
      goto done;                                                                   \
    }                                                                              \
  } while (0)

}  // namespace

extern "C" {

int gpp_gp_model_create(gpp_gp_model** out, int L, int M, int D, const double* Z, const double* lengthscales,
                        const double* variance, const double* q_mu, const double* q_sqrt, int whiten,
                        const double* mean_const, const double* W, int P, const double* kuu_jitter,
                        int model_uncertainty, void* stream_) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(out && Z && lengthscales && variance && q_mu && kuu_jitter, GPP_ERR_NULL, "gpp_gp_model_create: null argument");
  GPP_REQUIRE(L >= 1 && L <= GPP_MAX_L && M >= 1 && D >= 1, GPP_ERR_BAD_SHAPE, "gpp_gp_model_create: bad sizes L=%d M=%d D=%d", L, M, D);
  GPP_REQUIRE(D <= GPP_MAX_D, GPP_ERR_UNSUPPORTED, "gpp_gp_model_create: D=%d exceeds GPP_MAX_D=%d", D, GPP_MAX_D);
  GPP_REQUIRE(W != nullptr || P == L, GPP_ERR_BAD_SHAPE, "gpp_gp_model_create: P=%d must equal L=%d without W", P, L);
  GPP_REQUIRE(P >= 1 && P <= GPP_MAX_L, GPP_ERR_BAD_SHAPE, "gpp_gp_model_create: bad P=%d", P);
  cudaStream_t stream = (cudaStream_t)stream_;

  gpp_gp_model* m = new gpp_gp_model();
  m->L = L; m->M = M; m->D = D; m->P = P;
  m->whiten = whiten; m->model_uncertainty = model_uncertainty; m->coreg = (W != nullptr);
  size_t mm = (size_t)M * M;
  int status = GPP_OK;
  cusolverDnHandle_t solver = nullptr;
  cublasHandle_t blas = nullptr;
  double *work = nullptr, *R = nullptr, *inner = nullptr, *wvec = nullptr;
  int* dinfo = nullptr;
  int lwork = 0;
  const double one = 1.0, minus_one = -1.0;
  dim3 blk(128), grd((M + 127) / 128, M);

  if (cudaMalloc(&m->Z, sizeof(double) * L * M * D) != cudaSuccess || cudaMalloc(&m->ell, sizeof(double) * L * D) != cudaSuccess ||
      cudaMalloc(&m->var, sizeof(double) * L) != cudaSuccess || cudaMalloc(&m->beta, sizeof(double) * L * M) != cudaSuccess ||
      cudaMalloc(&m->C, sizeof(double) * L * mm) != cudaSuccess || cudaMalloc(&m->Luu, sizeof(double) * L * mm) != cudaSuccess ||
      cudaMalloc(&m->mean, sizeof(double) * P) != cudaSuccess || cudaMalloc(&R, sizeof(double) * mm) != cudaSuccess ||
      cudaMalloc(&inner, sizeof(double) * mm) != cudaSuccess || cudaMalloc(&wvec, sizeof(double) * M) != cudaSuccess ||
      cudaMalloc(&dinfo, sizeof(int)) != cudaSuccess) {
    gpp::set_error("gpp_gp_model_create: device allocation failed (M=%d, L=%d)", M, L);
    status = GPP_ERR_CUDA;
    goto done;
  }
  if (cudaMalloc(&m->q_mu, sizeof(double) * M * L) != cudaSuccess || cudaMalloc(&m->q_sqrt, sizeof(double) * L * mm) != cudaSuccess) {
    gpp::set_error("gpp_gp_model_create: device allocation failed (q_mu/q_sqrt copies)");
    status = GPP_ERR_CUDA;
    goto done;
  }
  cudaMemcpyAsync(m->q_mu, q_mu, sizeof(double) * M * L, cudaMemcpyDeviceToDevice, stream);
  for (int l = 0; l < L; ++l) {
    tril_kernel<<<grd, blk, 0, stream>>>(q_sqrt ? q_sqrt + (size_t)l * mm : nullptr, M, m->q_sqrt + (size_t)l * mm);
    m->h_jitter.push_back(kuu_jitter[l]);
  }
  if (W) {
    if (cudaMalloc(&m->W, sizeof(double) * P * L) != cudaSuccess) { status = GPP_ERR_CUDA; goto done; }
    cudaMemcpyAsync(m->W, W, sizeof(double) * P * L, cudaMemcpyDeviceToDevice, stream);
  }
  cudaMemcpyAsync(m->Z, Z, sizeof(double) * L * M * D, cudaMemcpyDeviceToDevice, stream);
  cudaMemcpyAsync(m->ell, lengthscales, sizeof(double) * L * D, cudaMemcpyDeviceToDevice, stream);
  cudaMemcpyAsync(m->var, variance, sizeof(double) * L, cudaMemcpyDeviceToDevice, stream);
  if (mean_const) cudaMemcpyAsync(m->mean, mean_const, sizeof(double) * P, cudaMemcpyDeviceToDevice, stream);
  else cudaMemsetAsync(m->mean, 0, sizeof(double) * P, stream);
  m->h_ell.resize((size_t)L * D);
  m->h_var.resize(L);
  cudaMemcpyAsync(m->h_ell.data(), lengthscales, sizeof(double) * L * D, cudaMemcpyDeviceToHost, stream);
  cudaMemcpyAsync(m->h_var.data(), variance, sizeof(double) * L, cudaMemcpyDeviceToHost, stream);
  if (cudaStreamSynchronize(stream) != cudaSuccess) { gpp::set_error("gpp_gp_model_create: parameter copy failed"); status = GPP_ERR_CUDA; goto done; }

  GPP_LIB_OK(cusolverDnCreate(&solver), "cusolverDnCreate");
  GPP_LIB_OK(cublasCreate(&blas), "cublasCreate");
  GPP_LIB_OK(cusolverDnSetStream(solver, stream), "cusolverDnSetStream");
  GPP_LIB_OK(cublasSetStream(blas, stream), "cublasSetStream");
  GPP_LIB_OK(cusolverDnDpotrf_bufferSize(solver, CUBLAS_FILL_MODE_UPPER, M, m->Luu, M, &lwork), "potrf_bufferSize");
  if (cudaMalloc(&work, sizeof(double) * std::max(lwork, 1)) != cudaSuccess) { status = GPP_ERR_CUDA; goto done; }

  for (int l = 0; l < L; ++l) {
    double* Lu = m->Luu + (size_t)l * mm;   // row-major lower L  ==  column-major upper U = L^T,  Kuu = U^T U
    kuu_kernel<<<grd, blk, 0, stream>>>(m->Z + (size_t)l * M * D, m->ell + (size_t)l * D, m->h_var[l], kuu_jitter[l], M, D, Lu);
    gpp::count_launch();
    GPP_LIB_OK(cusolverDnDpotrf(solver, CUBLAS_FILL_MODE_UPPER, M, Lu, M, work, lwork, dinfo), "potrf");
    int hinfo = 0;
    cudaMemcpyAsync(&hinfo, dinfo, sizeof(int), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    if (hinfo != 0) {
      gpp::set_error("gpp_gp_model_create: Kuu of latent %d is not positive definite (leading minor %d)", l, hinfo);
      status = GPP_ERR_NOT_PD;
      goto done;
    }
    clear_upper_kernel<<<grd, blk, 0, stream>>>(M, Lu);
    // w = q_mu[:, l]  (whitened)  or  L^-1 q_mu[:, l];   beta = L^-T w
    gather_col_kernel<<<(M + 127) / 128, 128, 0, stream>>>(q_mu, M, L, l, wvec);
    if (!whiten) GPP_LIB_OK(cublasDtrsv(blas, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, M, Lu, M, wvec, 1), "trsv");
    GPP_LIB_OK(cublasDtrsv(blas, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, M, Lu, M, wvec, 1), "trsv");
    cudaMemcpyAsync(m->beta + (size_t)l * M, wvec, sizeof(double) * M, cudaMemcpyDeviceToDevice, stream);
    if (model_uncertainty) {
      // R (row-major lower) ; column-major view of the buffer is R^T
      tril_kernel<<<grd, blk, 0, stream>>>(q_sqrt ? q_sqrt + (size_t)l * mm : nullptr, M, R);
      if (!whiten && q_sqrt)   // R <- L^-1 R   <=>   R^T <- R^T U^-1
        GPP_LIB_OK(cublasDtrsm(blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, M, M, &one, Lu, M, R, M), "trsm");
      eye_kernel<<<grd, blk, 0, stream>>>(M, inner);
      if (q_sqrt)              // inner = I - R R^T = I - (R^T)^T (R^T)
        GPP_LIB_OK(cublasDgemm(blas, CUBLAS_OP_T, CUBLAS_OP_N, M, M, M, &minus_one, R, M, R, M, &one, inner, M), "gemm");
      // B = L^-T inner L^-1 = U^-1 inner U^-T
      GPP_LIB_OK(cublasDtrsm(blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, M, M, &one, Lu, M, inner, M), "trsm");
      GPP_LIB_OK(cublasDtrsm(blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, M, M, &one, Lu, M, inner, M), "trsm");
      gpp::count_launch(3);
    }
    finish_c_kernel<<<grd, blk, 0, stream>>>(m->beta + (size_t)l * M, inner, M, model_uncertainty, m->C + (size_t)l * mm);
    gpp::count_launch(3);
  }
  if (cudaStreamSynchronize(stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
    gpp::set_error("gpp_gp_model_create: CUDA failure while preparing weights: %s", cudaGetErrorString(cudaGetLastError()));
    status = GPP_ERR_CUDA;
    goto done;
  }
  for (int t = 0; t < 2; ++t)
    for (int d = 0; d < 2; ++d) build_slots(m->tables[t][d], L, M, t == 0 ? 64 : 128, d == 1);

done:
  if (solver) cusolverDnDestroy(solver);
  if (blas) cublasDestroy(blas);
  cudaFree(work); cudaFree(R); cudaFree(inner); cudaFree(wvec); cudaFree(dinfo);
  if (status != GPP_OK) {
    gpp_gp_model_destroy(m);
    *out = nullptr;
    return status;
  }
  *out = m;
  return GPP_OK;
}

int gpp_gp_model_destroy(gpp_gp_model* m) {
  if (!m) return GPP_OK;
  cudaFree(m->Z); cudaFree(m->ell); cudaFree(m->var); cudaFree(m->beta); cudaFree(m->C);
  cudaFree(m->mean); cudaFree(m->W); cudaFree(m->Luu); cudaFree(m->q_mu); cudaFree(m->q_sqrt);
  if (m->blas) cublasDestroy((cublasHandle_t)m->blas);
  for (int t = 0; t < 2; ++t)
    for (int d = 0; d < 2; ++d) {
      cudaFree(m->tables[t][d].d_slots);
      cudaFree(m->tables[t][d].d_pair_start);
      cudaFree(m->tables[t][d].d_pair_ab);
    }
  delete m;
  return GPP_OK;
}

int gpp_gp_model_weights(const gpp_gp_model* m, double* beta, double* C, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(m, GPP_ERR_NULL, "gpp_gp_model_weights: null model");
  if (beta) GPP_CUDA_OK(cudaMemcpyAsync(beta, m->beta, sizeof(double) * m->L * m->M, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  if (C) GPP_CUDA_OK(cudaMemcpyAsync(C, m->C, sizeof(double) * m->L * m->M * m->M, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return GPP_OK;
}

}  // extern "C"
