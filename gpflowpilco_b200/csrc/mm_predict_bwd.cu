// Backward of the fused moment-matched GP predict: given the adjoints of (f1, Sff, cross) return the adjoints of the input
// moments (m, S).  Upstream obtains these by TensorFlow autodiff through gpflow_pilco/moment_matching/models.py:200-299 and
// gpflow_pilco/utils/kernel_expectation.py:96-187; here they are closed form:
//
//   Psi2 part   d f2_ab / d mu = G (sum A e),  d f2_ab / d Sigma = 1/2 (G (sum A e e^T) G - (sum A) G),
//               A = C_ab o Q_ab,  e_ij = A1 z1'_i + A2 z2'_j  (z' = z - mu),  G = (Sigma + V_ab)^-1,
//               sum A e     = A1 r1_ab + A2 r1_ba,
//               sum A e e^T = A1 R2_ab A1 + A2 R2_ba A2 + A1 X_ab A2 + (A1 X_ab A2)^T,
//               with the ROW statistics of the ordered pair (a,b):  a_i = sum_j A_ij, u_i = sum_j A_ij z2'_j,
//               S0 = sum_i a_i, r1 = sum_i a_i z1'_i, R2 = sum_i a_i z1'_i z1'_i^T, X = sum_i z1'_i u_i^T.
//               Column statistics of (a,b) are the row statistics of (b,a) (Q_ba = Q_ab^T); of those only r1_ba = sum_i u_i and
//               R2_ba = sum_j (sum_i A_ij) z2'_j z2'_j^T are needed beyond the row pass, so each unordered pair is contracted once.
//   Psi1 part   s = f1bar_l f1_l + crossbar_l . cross_l = sum_m w_m (f1bar + y . dz_m),  w = beta psi1, y = G1 crossbar,
//               ds/dmu = G1 (sum w e dz) - (sum w) y,
//               ds/dSigma = 1/2 (G1 (sum w e dz dz^T) G1 - (sum w e) G1) - 1/2 (y c^T + c y^T),  c = G1 sum w dz.
//
//   k_pack_psi1 (coefficient packs | per input: forward latent means, un-mix W and the Sff = f2 - f1 f1^T chain rule, Psi1 adjoints)
//   -> k_contract_grad (DMMA exponents, warp per 8-row strip, CTA per (input, unordered pair, 128-row block))
//   -> k_bwd_finalize (CTA per input: D x D algebra per unordered pair, fixed-order sums).
// Gradients w.r.t. the model parameters are not produced: the dynamics model is constant during policy optimisation
// (upstream differentiates w.r.t. policy.trainable_variables only, gpflow_pilco/utils/optimizers.py:52-56).
#include <algorithm>
#include <type_traits>

#include "predict_bwd_kernels.cuh"

namespace gpp {

static size_t bwd_align(size_t x) { return (x + 255) / 256 * 256; }

struct BwdLayout {
  size_t packs, Gs, stats, f1lat, crosslat, f1lat_bar, crosslat_bar, omega, gm, gS, total;
  int nrb;
};

static BwdLayout bwd_layout(const gpp_gp_model* m, int N) {
  BwdLayout lo{};
  const int D = m->D, L = m->L;
  size_t pack_doubles = 0, stat_doubles = 0;
  switch (D) {
#define GPP_CASE(d) case d: pack_doubles = PairPack<d>::SIZE; stat_doubles = GradStats<d>::SIZE; break;
    GPP_CASE(1) GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7) GPP_CASE(8)
#undef GPP_CASE
  }
  lo.nrb = (m->M + kGradRows - 1) / kGradRows;
  size_t off = 0;
  auto take = [&](size_t doubles) { size_t o = off; off = bwd_align(off + doubles * sizeof(double)); return o; };
  lo.packs = take(pack_doubles * L * L * N);
  lo.Gs = take((size_t)D * D * L * L * N);
  lo.stats = take(stat_doubles * L * L * lo.nrb * N);
  lo.f1lat = take((size_t)L * N);
  lo.crosslat = take((size_t)L * D * N);
  lo.f1lat_bar = take((size_t)L * N);
  lo.crosslat_bar = take((size_t)L * D * N);
  lo.omega = take((size_t)L * L * N);
  lo.gm = take((size_t)L * D * N);
  lo.gS = take((size_t)L * D * D * N);
  lo.total = off;
  return lo;
}

template <int D>
static int predict_bwd(const gpp_gp_model* m, const double* mu, const double* S, int N, const double* f1_bar, const double* Sff_bar,
                       const double* cross_bar, int full_output_cov, double* m_bar, double* S_bar, char* ws, const BwdLayout& lo,
                       int* info, cudaStream_t stream) {
  const int L = m->L;
  double* packs = (double*)(ws + lo.packs);
  double* Gs = (double*)(ws + lo.Gs);
  double* stats = (double*)(ws + lo.stats);
  double* f1lat = (double*)(ws + lo.f1lat);
  double* crosslat = (double*)(ws + lo.crosslat);
  double* f1lat_bar = (double*)(ws + lo.f1lat_bar);
  double* crosslat_bar = (double*)(ws + lo.crosslat_bar);
  double* omega = (double*)(ws + lo.omega);
  double* gm = (double*)(ws + lo.gm);
  double* gS = (double*)(ws + lo.gS);
  PackPsi1Params pp;
  pp.m = mu; pp.S = S; pp.Z = m->Z; pp.ell = m->ell; pp.var = m->var; pp.beta = m->beta; pp.pair_ab = nullptr;
  pp.packs = packs; pp.f1lat = f1lat; pp.crosslat = crosslat; pp.Gs = Gs; pp.counter = nullptr; pp.info = info;
  pp.N = N; pp.L = L; pp.M = m->M; pp.npairs = L * L;
  BwdEpilogue<D> epi;
  epi.m = mu; epi.S = S; epi.Z = m->Z; epi.ell = m->ell; epi.var = m->var; epi.beta = m->beta; epi.gm = gm; epi.gS = gS; epi.M = m->M;
  BwdPrepareParams& bp = epi.bp;
  bp.f1_bar = f1_bar; bp.Sff_bar = Sff_bar; bp.cross_bar = cross_bar; bp.f1lat = f1lat; bp.W = m->W;
  bp.f1lat_bar = f1lat_bar; bp.crosslat_bar = crosslat_bar; bp.omega = omega;
  bp.N = N; bp.L = L; bp.P = m->P; bp.D = D; bp.full_cov = full_output_cov;
  launch_pack_psi1<D, BwdEpilogue<D>>(pp, stream, epi);
  profile_begin(stream);
  {
    const size_t smem = sizeof(double) * (GradCfg<D>::FIXED + (size_t)lo.nrb * kGradCols);   // nrb == number of column blocks
    static PerDeviceSmemOptIn configured;
    GPP_REQUIRE(smem <= 227 * 1024, GPP_ERR_UNSUPPORTED, "gpp_mm_gp_predict_bwd: M=%d needs %zu bytes of shared memory", m->M, smem);
    if (configured.raise(smem))
      GPP_CUDA_OK(cudaFuncSetAttribute(k_contract_grad<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_contract_grad<D><<<N * (L * (L + 1) / 2) * lo.nrb, kGradThreads, smem, stream>>>(m->Z, m->beta, m->C, packs, omega, stats, m->M,
                                                                                     L, lo.nrb);
  }
  profile_end(stream);
  BwdFinalizeParams fp;
  fp.m = mu; fp.S = S; fp.ell = m->ell; fp.stats = stats; fp.ll_tag = 0; fp.omega = omega; fp.Gs = Gs; fp.gm = gm; fp.gS = gS;
  fp.m_bar = m_bar; fp.S_bar = S_bar; fp.N = N; fp.L = L; fp.nrb = lo.nrb;
  {
    const int npairs = L * (L + 1) / 2;
    const size_t smem = sizeof(double) * (size_t)npairs * FinalizeSmem<D>::PER_PAIR;
    static PerDeviceSmemOptIn configured;
    if (smem > 48 * 1024 && configured.raise(smem))
      GPP_CUDA_OK(cudaFuncSetAttribute(k_bwd_finalize<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_bwd_finalize<D><<<N, 128, smem, stream>>>(fp);
  }
  count_launch(3);
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

// typed-stream alias used by the rollout backward
int mm_predict_bwd_enqueue(const gpp_gp_model* model, const double* m, const double* S, int N, const double* f1_bar,
                           const double* Sff_bar, const double* cross_bar, int full_output_cov, double* m_bar, double* S_bar,
                           void* workspace, size_t workspace_bytes, int* info, cudaStream_t stream) {
  return gpp_mm_gp_predict_bwd(model, m, S, N, f1_bar, Sff_bar, cross_bar, full_output_cov, m_bar, S_bar, workspace,
                               workspace_bytes, info, (void*)stream);
}

}  // namespace gpp

extern "C" {

size_t gpp_mm_gp_predict_bwd_workspace_bytes(const gpp_gp_model* model, int N) {
  if (!model || N <= 0) return 0;
  return gpp::bwd_layout(model, N).total;
}

int gpp_mm_gp_predict_bwd(const gpp_gp_model* model, const double* m, const double* S, int N, const double* f1_bar,
                          const double* Sff_bar, const double* cross_bar, int full_output_cov, double* m_bar, double* S_bar,
                          void* workspace, size_t workspace_bytes, int* info, void* stream_) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(model && m && S && m_bar && S_bar && workspace, GPP_ERR_NULL, "gpp_mm_gp_predict_bwd: null argument");
  GPP_REQUIRE(N >= 1, GPP_ERR_BAD_SHAPE, "gpp_mm_gp_predict_bwd: N=%d", N);
  gpp::BwdLayout lo = gpp::bwd_layout(model, N);
  GPP_REQUIRE(workspace_bytes >= lo.total, GPP_ERR_WORKSPACE, "gpp_mm_gp_predict_bwd: workspace %zu < required %zu", workspace_bytes, lo.total);
  cudaStream_t stream = (cudaStream_t)stream_;
  char* ws = (char*)workspace;
  switch (model->D) {
#define GPP_CASE(d) \
  case d: return gpp::predict_bwd<d>(model, m, S, N, f1_bar, Sff_bar, cross_bar, full_output_cov, m_bar, S_bar, ws, lo, info, stream);
    GPP_CASE(1) GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7) GPP_CASE(8)
#undef GPP_CASE
    default:
      gpp::set_error("gpp_mm_gp_predict_bwd: unsupported D=%d", model->D);
      return GPP_ERR_UNSUPPORTED;
  }
}

}  // extern "C"
