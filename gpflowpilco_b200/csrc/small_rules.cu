// Batched entry points for the small moment-matching rules (one thread per Gaussian state), used by the Python
// façade when the rules are applied one at a time (moment_matching(x, encoder), moment_matching(x, bijector), objective(x))
// instead of through the fused rollout.  Same device code as the rollout kernels (mm_small.cuh).
#include "bvn.cuh"
#include "mm_small.cuh"

namespace gpp {

__global__ void k_mm_encoder(EncoderSpec es, int N, const double* m, const double* S, double* me, double* See, double* Cxe) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int Dx = es.Dx, De = es.De();
  double lm[GPP_SMALL_MAX], lS[GPP_SMALL_MAX * GPP_SMALL_MAX], ome[GPP_SMALL_MAX], oS[GPP_SMALL_MAX * GPP_SMALL_MAX],
      oC[GPP_SMALL_MAX * GPP_SMALL_MAX];
  for (int i = 0; i < Dx; ++i) lm[i] = m[(size_t)n * Dx + i];
  for (int i = 0; i < Dx * Dx; ++i) lS[i] = S[(size_t)n * Dx * Dx + i];
  mm_encoder<double>(es, lm, lS, ome, oS, oC);
  for (int i = 0; i < De; ++i) me[(size_t)n * De + i] = ome[i];
  for (int i = 0; i < De * De; ++i) See[(size_t)n * De * De + i] = oS[i];
  for (int i = 0; i < Dx * De; ++i) Cxe[(size_t)n * Dx * De + i] = oC[i];
}

__global__ void k_mm_squash(int N, const double* mf, const double* vf, double scale, double shift, double* mu, double* vu, double* gain) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double a, b, c;
  mm_squash_1d<double>(mf[n], vf[n], scale, shift, a, b, c);
  mu[n] = a; vu[n] = b; gain[n] = c;
}

__global__ void k_cost_gaussian(int N, int De, const double* me, const double* See, const double* target, const double* W, double* out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double lm[GPP_SMALL_MAX], lS[GPP_SMALL_MAX * GPP_SMALL_MAX];
  for (int i = 0; i < De; ++i) lm[i] = me[(size_t)n * De + i];
  for (int i = 0; i < De * De; ++i) lS[i] = See[(size_t)n * De * De + i];
  out[n] = expected_cost<double>(De, lm, lS, target, W);
}

__global__ void k_cost_samples(int N, int De, const double* e, const double* target, const double* W, double* out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double le[GPP_SMALL_MAX];
  for (int i = 0; i < De; ++i) le[i] = e[(size_t)n * De + i];
  out[n] = sample_cost(De, le, target, W);
}

// u = scale (Phi(f) + shift) for an A-dimensional Gaussian f ~ N(mf, Sf): one thread per (state, i, j) entry of the covariance.
// E[Phi(f_i) Phi(f_j)] = BVN(-9 < w_i < h_i, -9 < w_j < h_j; rho_ij) with h = mf / sqrt(1 + diag Sf), rho_ij = Sf_ij /
// sqrt((1 + v_i)(1 + v_j)) — including i = j — and the lower limit -9 of upstream moment_matching/bijectors.py:59-63.
__global__ void k_mm_squash_nd(int N, int A, const double* __restrict__ mf, const double* __restrict__ Sf, double scale, double shift,
                               double* __restrict__ mu, double* __restrict__ Su, double* __restrict__ gain) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * A * A) return;
  const int j = idx % A, i = (idx / A) % A, n = idx / (A * A);
  const double* S = Sf + (size_t)n * A * A;
  const double vi = S[i * A + i], vj = S[j * A + j];
  const double qi = rsqrt(vi + 1.0), qj = rsqrt(vj + 1.0);
  const double hi = mf[(size_t)n * A + i] * qi, hj = mf[(size_t)n * A + j] * qj;
  const double y1i = bvn_ndtr(hi), y1j = bvn_ndtr(hj);
  const double y2 = bvn_box(-9.0, hi, -9.0, hj, S[i * A + j] * qi * qj);
  Su[idx] = (y2 - y1i * y1j) * scale * scale;
  if (i == j) {
    mu[(size_t)n * A + i] = (y1i + shift) * scale;
    gain[(size_t)n * A + i] = qi * 0.39894228040143267794 * exp(-0.5 * hi * hi) * scale;
  }
}

// Reverse mode of k_mm_squash_nd, one thread per state.  No bivariate probability is needed: with Phi2 the BVN distribution function
//   d Phi2(h, k; r) / dh = phi(h) Phi((k - r h) / sqrt(1 - r^2)),     d Phi2 / dr = phi2(h, k; r)  (the bivariate density),
// applied to the four corners of the box [-9, h_i] x [-9, h_j] the forward integrates over.
__device__ __forceinline__ double bvn_pdf(double h, double k, double r) {
  const double om = (1.0 - r) * (1.0 + r);
  return 0.15915494309189533577 * rsqrt(om) * exp(-0.5 * (h * h - 2.0 * r * h * k + k * k) / om);
}
__device__ __forceinline__ double box_dh(double h, double kl, double ku, double r) {   // d/dh P(. < h, kl < . < ku)
  const double is = rsqrt((1.0 - r) * (1.0 + r));
  return 0.39894228040143267794 * exp(-0.5 * h * h) * (bvn_ndtr((ku - r * h) * is) - bvn_ndtr((kl - r * h) * is));
}
__global__ void k_mm_squash_nd_bwd(int N, int A, const double* __restrict__ mf, const double* __restrict__ Sf, double scale, double shift,
                                   const double* __restrict__ mu_bar, const double* __restrict__ Su_bar,
                                   const double* __restrict__ gain_bar, double* __restrict__ mf_bar, double* __restrict__ Sf_bar) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const double* S = Sf + (size_t)n * A * A;
  double q[GPP_SMALL_MAX], h[GPP_SMALL_MAX], y1[GPP_SMALL_MAX], ph[GPP_SMALL_MAX], hb[GPP_SMALL_MAX], qb[GPP_SMALL_MAX];
  for (int i = 0; i < A; ++i) {
    q[i] = rsqrt(S[i * A + i] + 1.0);
    h[i] = mf[(size_t)n * A + i] * q[i];
    y1[i] = bvn_ndtr(h[i]);
    ph[i] = 0.39894228040143267794 * exp(-0.5 * h[i] * h[i]);
    const double mb = mu_bar ? mu_bar[(size_t)n * A + i] : 0.0, gb = gain_bar ? gain_bar[(size_t)n * A + i] : 0.0;
    hb[i] = scale * ph[i] * (mb - gb * q[i] * h[i]);       // mu = scale (Phi(h) + shift), gain = scale q phi(h)
    qb[i] = gb * scale * ph[i];
  }
  double* Sb = Sf_bar + (size_t)n * A * A;
  const double s2 = scale * scale;
  for (int i = 0; i < A; ++i)
    for (int j = 0; j < A; ++j) {
      const double g = Su_bar ? s2 * Su_bar[((size_t)n * A + i) * A + j] : 0.0;      // Su_ij = scale^2 (y2 - y1_i y1_j)
      const double r = S[i * A + j] * q[i] * q[j];
      hb[i] += g * (box_dh(h[i], -9.0, h[j], r) - y1[j] * ph[i]);
      hb[j] += g * (box_dh(h[j], -9.0, h[i], r) - y1[i] * ph[j]);
      const double rb = g * (bvn_pdf(h[i], h[j], r) - bvn_pdf(h[i], -9.0, r) - bvn_pdf(-9.0, h[j], r) + bvn_pdf(-9.0, -9.0, r));
      Sb[i * A + j] = rb * q[i] * q[j];                    // rho = S_ij q_i q_j
      qb[i] += rb * S[i * A + j] * q[j];
      qb[j] += rb * S[i * A + j] * q[i];
    }
  for (int i = 0; i < A; ++i) {
    mf_bar[(size_t)n * A + i] = hb[i] * q[i];             // h = m q
    const double qbi = qb[i] + hb[i] * mf[(size_t)n * A + i];
    Sb[i * A + i] += -0.5 * qbi * q[i] * q[i] * q[i];     // q = (1 + S_ii)^-1/2
  }
}

__global__ void k_owens_t(int N, const double* h, const double* a, double* out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < N) out[n] = owens_t<double>(h[n], a[n]);
}

}  // namespace gpp

extern "C" {

int gpp_mm_encoder(int N, int Dx, int num_active, const int* active_dims, const double* m, const double* S, double* me, double* See,
                   double* Cxe, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(m && S && me && See && Cxe, GPP_ERR_NULL, "gpp_mm_encoder: null argument");
  GPP_REQUIRE(N >= 0 && Dx >= 1 && num_active >= 0 && num_active <= 4 && num_active <= Dx && Dx + num_active <= GPP_SMALL_MAX,
              GPP_ERR_BAD_SHAPE, "gpp_mm_encoder: bad sizes Dx=%d active=%d", Dx, num_active);
  if (N == 0) return GPP_OK;
  gpp::EncoderSpec es{};
  es.Dx = Dx; es.na = num_active;
  for (int k = 0; k < num_active; ++k) {
    GPP_REQUIRE(active_dims[k] >= 0 && active_dims[k] < Dx, GPP_ERR_BAD_SHAPE, "gpp_mm_encoder: active dim out of range");
    es.active[k] = active_dims[k];
  }
  es.finish();
  gpp::k_mm_encoder<<<(N + 63) / 64, 64, 0, (cudaStream_t)stream>>>(es, N, m, S, me, See, Cxe);
  gpp::count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

int gpp_mm_squash(int N, const double* mf, const double* vf, double scale, double shift, double* mu, double* vu, double* gain, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(mf && vf && mu && vu && gain, GPP_ERR_NULL, "gpp_mm_squash: null argument");
  if (N <= 0) return GPP_OK;
  gpp::k_mm_squash<<<(N + 63) / 64, 64, 0, (cudaStream_t)stream>>>(N, mf, vf, scale, shift, mu, vu, gain);
  gpp::count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

int gpp_mm_squash_nd(int N, int A, const double* mf, const double* Sf, double scale, double shift, double* mu, double* Su, double* gain,
                     void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(mf && Sf && mu && Su && gain, GPP_ERR_NULL, "gpp_mm_squash_nd: null argument");
  GPP_REQUIRE(A >= 1 && A <= GPP_SMALL_MAX, GPP_ERR_BAD_SHAPE, "gpp_mm_squash_nd: A=%d", A);
  if (N <= 0) return GPP_OK;
  const long long total = (long long)N * A * A;
  gpp::k_mm_squash_nd<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(N, A, mf, Sf, scale, shift, mu, Su, gain);
  gpp::count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

int gpp_mm_squash_nd_bwd(int N, int A, const double* mf, const double* Sf, double scale, double shift, const double* mu_bar,
                         const double* Su_bar, const double* gain_bar, double* mf_bar, double* Sf_bar, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(mf && Sf && mf_bar && Sf_bar, GPP_ERR_NULL, "gpp_mm_squash_nd_bwd: null argument");
  GPP_REQUIRE(A >= 1 && A <= GPP_SMALL_MAX, GPP_ERR_BAD_SHAPE, "gpp_mm_squash_nd_bwd: A=%d", A);
  if (N <= 0) return GPP_OK;
  gpp::k_mm_squash_nd_bwd<<<(N + 63) / 64, 64, 0, (cudaStream_t)stream>>>(N, A, mf, Sf, scale, shift, mu_bar, Su_bar, gain_bar, mf_bar,
                                                                           Sf_bar);
  gpp::count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

int gpp_cost_gaussian(int N, int De, const double* me, const double* See, const double* target, const double* W, double* out, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(me && See && target && W && out, GPP_ERR_NULL, "gpp_cost_gaussian: null argument");
  GPP_REQUIRE(De >= 1 && De <= GPP_SMALL_MAX, GPP_ERR_BAD_SHAPE, "gpp_cost_gaussian: De=%d", De);
  if (N <= 0) return GPP_OK;
  gpp::k_cost_gaussian<<<(N + 63) / 64, 64, 0, (cudaStream_t)stream>>>(N, De, me, See, target, W, out);
  gpp::count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

int gpp_cost_samples(int N, int De, const double* e, const double* target, const double* W, double* out, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(e && target && W && out, GPP_ERR_NULL, "gpp_cost_samples: null argument");
  GPP_REQUIRE(De >= 1 && De <= GPP_SMALL_MAX, GPP_ERR_BAD_SHAPE, "gpp_cost_samples: De=%d", De);
  if (N <= 0) return GPP_OK;
  gpp::k_cost_samples<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(N, De, e, target, W, out);
  gpp::count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

int gpp_owens_t(int N, const double* h, const double* a, double* out, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(h && a && out, GPP_ERR_NULL, "gpp_owens_t: null argument");
  if (N <= 0) return GPP_OK;
  gpp::k_owens_t<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(N, h, a, out);
  gpp::count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

}  // extern "C"
