// Materialised kernel expectations: Psi1 = eKxz [N,M] and Psi2 = eKzxKxz [N,M1,M2].
//
// Psi2 replaces `_E`, gpflow_pilco/utils/kernel_expectation.py:72-187.  The reference builds >= 4 [N,M1,M2]
// temporaries (tile of Z, two triangular solves, a K=D matmul, two exp tensors); here each entry is produced
// once from  log Q_ij = r_i + s_j + z1'_i^T R z2'_j  (common.cuh) and written with 128-bit coalesced stores.
// The kernel is HBM-write bound (8 B/entry) with the FP64 pipe a close second (D + 17 FP64 ops/entry).
#include "common.cuh"

namespace gpp {

template <int D>
__global__ void __launch_bounds__(128) k_ekxz(const double* __restrict__ mu, const double* __restrict__ cov, int N, int M,
                                              const double* __restrict__ Z, const double* __restrict__ ell, double variance,
                                              double* __restrict__ out, int* info) {
  const int n = blockIdx.y;
  Mat<D> A, Li;
  double m[D], half_log_v = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    double e = ell[d];
    half_log_v += log(e);
    m[d] = mu[(size_t)n * D + d];
#pragma unroll
    for (int e2 = 0; e2 < D; ++e2) A(d, e2) = cov[(size_t)n * D * D + d * D + e2] + (d == e2 ? e * e : 0.0);
  }
  bool ok = cholesky<D>(A);
  if (!ok && threadIdx.x == 0 && blockIdx.x == 0) flag_not_pd(info, n);
  double log_det = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) log_det += log(A(d, d));
  tri_inverse<D>(A, Li);
  const double c0 = log(variance) + half_log_v - log_det;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
    double dz[D];
#pragma unroll
    for (int d = 0; d < D; ++d) dz[d] = Z[(size_t)j * D + d] - m[d];
    double maha = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double y = 0.0;
#pragma unroll
      for (int k = 0; k <= i; ++k) y = fma(Li(i, k), dz[k], y);
      maha = fma(y, y, maha);
    }
    out[(size_t)n * M + j] = fast_exp(c0 - 0.5 * maha);
  }
}

template <int D>
__global__ void __launch_bounds__(64) k_pack_single(const double* __restrict__ mu, const double* __restrict__ cov, int N,
                                                    const double* __restrict__ ell1, const double* __restrict__ ell2,
                                                    double log_amp, double* __restrict__ packs, int* info) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double V1[D], V2[D], m[D], Sg[D * D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    V1[d] = ell1[d] * ell1[d];
    V2[d] = ell2[d] * ell2[d];
    m[d] = mu[(size_t)n * D + d];
  }
#pragma unroll
  for (int d = 0; d < D * D; ++d) Sg[d] = cov[(size_t)n * D * D + d];
  double out[PairPack<D>::SIZE];
  if (!make_pair_pack<D>(m, Sg, V1, V2, log_amp, out)) flag_not_pd(info, n);
#pragma unroll
  for (int t = 0; t < PairPack<D>::SIZE; ++t) packs[(size_t)n * PairPack<D>::SIZE + t] = out[t];
}

// grid (column strips of CT columns, n); 128 threads, each owns 2 adjacent columns of the strip and walks all rows.
template <int D>
__global__ void __launch_bounds__(128, 4) k_ekzxkxz(const double* __restrict__ packs, const double* __restrict__ Z1, int M1,
                                                 const double* __restrict__ Z2, int M2, double* __restrict__ out) {
  using PP = PairPack<D>;
  constexpr int RT = 64;                 // rows staged per pass
  constexpr int RS = (D + 1 + 1) & ~1;   // z1'[D], r
  __shared__ __align__(16) double pk[PP::SIZE];
  __shared__ __align__(16) double rowbuf[RT * RS];
  __shared__ double etab[64 * GPP_EXP_TAB_REP];   // replicated 2^(j/64) table of the 10-op exp (gpp_math.h)
  const int n = blockIdx.y, tid = threadIdx.x;
  for (int t = tid; t < PP::SIZE; t += blockDim.x) pk[t] = packs[(size_t)n * PP::SIZE + t];
  for (int t = tid; t < 64 * GPP_EXP_TAB_REP; t += blockDim.x) etab[t] = kExp2Tab[t / GPP_EXP_TAB_REP];
  __syncthreads();
  const double* etab_lane = etab + (tid & (GPP_EXP_TAB_REP - 1));
  const int j0 = (blockIdx.x * 128 + tid) * 2;
  const bool pair_ok = ((M2 & 1) == 0);          // 16-byte aligned pair stores need an even row length
  double g[2][D], s[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    int j = j0 + c;
    double zc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) zc[d] = (j < M2 ? Z2[(size_t)j * D + d] : 0.0) - pk[PP::MU + d];
    s[c] = packed_quad<D>(pk + PP::P2, zc);
#pragma unroll
    for (int d = 0; d < D; ++d) {       // g = R z2'
      double t = 0.0;
#pragma unroll
      for (int e = 0; e < D; ++e) t = fma(pk[PP::R + d * D + e], zc[e], t);
      g[c][d] = t;
    }
  }
  double* outn = out + (size_t)n * M1 * M2;
  for (int i0 = 0; i0 < M1; i0 += RT) {
    __syncthreads();
    if (tid < RT) {
      int i = i0 + tid;
      double zr[D];
#pragma unroll
      for (int d = 0; d < D; ++d) zr[d] = (i < M1 ? Z1[(size_t)i * D + d] : 0.0) - pk[PP::MU + d];
      double r = pk[PP::C0] + packed_quad<D>(pk + PP::P1, zr);
#pragma unroll
      for (int d = 0; d < D; ++d) rowbuf[tid * RS + d] = zr[d];
      rowbuf[tid * RS + D] = r;
    }
    __syncthreads();
    const int rows = min(RT, M1 - i0);
    if (j0 >= M2) continue;
    // 2 rows x 2 columns in lock-step (4 independent FP64 chains per thread); rowbuf is zero-padded past M1
#pragma unroll 1
    for (int ii = 0; ii < rows; ii += 2) {
      const double* rb0 = rowbuf + ii * RS;
      const double* rb1 = rb0 + RS;
      double t[4] = {rb0[D] + s[0], rb0[D] + s[1], rb1[D] + s[0], rb1[D] + s[1]};
#pragma unroll
      for (int d = 0; d < D; ++d) {
        t[0] = fma(rb0[d], g[0][d], t[0]);
        t[1] = fma(rb0[d], g[1][d], t[1]);
        t[2] = fma(rb1[d], g[0][d], t[2]);
        t[3] = fma(rb1[d], g[1][d], t[3]);
      }
      fast_exp_tab_n<4>(t, etab_lane);
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        if (ii + rr >= rows) break;
        double* dst = outn + (size_t)(i0 + ii + rr) * M2 + j0;
        if (pair_ok) {
          __stcs(reinterpret_cast<double2*>(dst), make_double2(t[2 * rr], t[2 * rr + 1]));   // streaming: never re-read
        } else {
          __stcs(dst, t[2 * rr]);
          if (j0 + 1 < M2) __stcs(dst + 1, t[2 * rr + 1]);
        }
      }
    }
  }
}

template <int D>
static int run_ekxz(const double* mu, const double* cov, int N, const double* Z, int M, const double* ell, double variance,
                    double* out, int* info, cudaStream_t stream) {
  dim3 grid(std::min((M + 127) / 128, 64), N);
  k_ekxz<D><<<grid, 128, 0, stream>>>(mu, cov, N, M, Z, ell, variance, out, info);
  count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

template <int D>
static int run_ekzxkxz(const double* mu, const double* cov, int N, const double* Z1, int M1, const double* ell1, double var1,
                       const double* Z2, int M2, const double* ell2, double var2, double* out, int* info,
                       cudaStream_t stream) {
  double* packs = nullptr;
  GPP_CUDA_OK(cudaMallocAsync(&packs, sizeof(double) * PairPack<D>::SIZE * (size_t)N, stream));
  k_pack_single<D><<<(N + 63) / 64, 64, 0, stream>>>(mu, cov, N, ell1, ell2, log(var1 * var2), packs, info);
  dim3 grid((M2 + 255) / 256, N);
  profile_begin(stream);
  k_ekzxkxz<D><<<grid, 128, 0, stream>>>(packs, Z1, M1, Z2, M2, out);
  profile_end(stream);
  count_launch(2);
  GPP_CUDA_OK(cudaGetLastError());
  GPP_CUDA_OK(cudaFreeAsync(packs, stream));
  return GPP_OK;
}

}  // namespace gpp

extern "C" {

int gpp_ekxz(const double* mu, const double* cov, int N, int D, const double* Z, int M, const double* lengthscales,
             double variance, double* out, int* info, void* stream) {
  GPP_REQUIRE(mu && cov && Z && lengthscales && out, GPP_ERR_NULL, "gpp_ekxz: null argument");
  GPP_REQUIRE(N >= 0 && M >= 0 && D >= 1, GPP_ERR_BAD_SHAPE, "gpp_ekxz: bad sizes N=%d M=%d D=%d", N, M, D);
  GPP_REQUIRE(N <= 65535, GPP_ERR_UNSUPPORTED, "gpp_ekxz: N=%d > 65535 inputs per call", N);
  if (N == 0 || M == 0) return GPP_OK;
  switch (D) {
#define GPP_CASE(d) case d: return gpp::run_ekxz<d>(mu, cov, N, Z, M, lengthscales, variance, out, info, (cudaStream_t)stream);
    GPP_CASE(1) GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7) GPP_CASE(8)
#undef GPP_CASE
    default: gpp::set_error("gpp_ekxz: unsupported D=%d (max %d)", D, GPP_MAX_D); return GPP_ERR_UNSUPPORTED;
  }
}

int gpp_ekzxkxz(const double* mu, const double* cov, int N, int D, const double* Z1, int M1, const double* lengthscales1,
                double variance1, const double* Z2, int M2, const double* lengthscales2, double variance2, double* out,
                int* info, void* stream) {
  GPP_REQUIRE(mu && cov && Z1 && lengthscales1 && out, GPP_ERR_NULL, "gpp_ekzxkxz: null argument");
  if (!Z2) { Z2 = Z1; M2 = M1; }
  if (!lengthscales2) { lengthscales2 = lengthscales1; variance2 = variance1; }
  GPP_REQUIRE(N >= 0 && M1 >= 0 && M2 >= 0 && D >= 1, GPP_ERR_BAD_SHAPE, "gpp_ekzxkxz: bad sizes N=%d M1=%d M2=%d D=%d", N, M1, M2, D);
  GPP_REQUIRE(N <= 65535, GPP_ERR_UNSUPPORTED, "gpp_ekzxkxz: N=%d > 65535 inputs per call", N);
  if (N == 0 || M1 == 0 || M2 == 0) return GPP_OK;
  switch (D) {
#define GPP_CASE(d) \
  case d: return gpp::run_ekzxkxz<d>(mu, cov, N, Z1, M1, lengthscales1, variance1, Z2, M2, lengthscales2, variance2, out, info, (cudaStream_t)stream);
    GPP_CASE(1) GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7) GPP_CASE(8)
#undef GPP_CASE
    default: gpp::set_error("gpp_ekzxkxz: unsupported D=%d (max %d)", D, GPP_MAX_D); return GPP_ERR_UNSUPPORTED;
  }
}

}  // extern "C"
