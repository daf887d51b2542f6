// Backward of the moment-matched rollout: reverse sweep over the H steps of gpp_rollout_mm_fwd.
//
// Upstream differentiates the closure of MomentMatchingPILCO (gpflow_pilco/loops/pilco.py:192-220) with tape.gradient
// w.r.t. policy.trainable_variables (gpflow_pilco/utils/optimizers.py:52-56).  Here, for t = H-1 .. 0, from the stored
// trajectory (m_t, S_t) and the per-step (md, Sd, Sxd, cross, pre-stage block) kept by gpp_rollout_mm_fwd_save — or, without them:
//   k_step_pre            recompute the pre stage of step t (md, Sd, Sxd)                       [rollout_mm_common.cuh]
//   mm_predict_enqueue    recompute (f1, Sff, cross) of step t with the fused forward kernels   [mm_predict.cu]
// then
//   k_cost_grad_ring      d cost / d(m, S) of 32 consecutive trajectory states at once (dual numbers through the encoder and
//                         expected-cost rules, one (state, rollout, direction) per thread)
//   k_bwd_post            adjoint of (m_{t+1}, S_{t+1}) += loss_bar * that gradient, then the adjoint of the Euler moment update
//                         (dynamics/solvers.py:128-129) and of Sxf = Sxd cross  (forward_sde.py:126)
//   mm_predict_bwd        closed-form adjoint of the GP dynamics prediction                     [mm_predict_bwd.cu]
//   k_bwd_pre             reload (or recompute) the pre stage's shared block; adjoint of the joint assembly (forward_sde.py:105-124,
//                         gaussian.py:53-63), of the squashing link (3x2 Jacobian by dual numbers), closed-form adjoint of the policy's Psi1/Psi2 sums w.r.t. the
//                         encoded moments AND the policy parameters (centres Z, weights beta = Kuu^-1 m, lengthscales),
//                         then the encoder adjoint (dual numbers, one direction per thread).
// Matrix adjoints follow one convention throughout: S_bar is symmetric and dLoss = sum_ij S_bar_ij dS_ij for symmetric dS.
#include "rollout_mm_bwd_common.cuh"
#include "rollout_persist.h"

namespace gpp {

int mm_predict_bwd_enqueue(const gpp_gp_model* model, const double* m, const double* S, int N, const double* f1_bar,
                           const double* Sff_bar, const double* cross_bar, int full_output_cov, double* m_bar, double* S_bar,
                           void* workspace, size_t workspace_bytes, int* info, cudaStream_t stream);   // mm_predict_bwd.cu

// out[r, k] = sum over the rollouts that used parameter set r (fixed order): R == N copies, R == 1 sums over n
__global__ void k_reduce_param_grads(const double* __restrict__ g, int N, int R, int K, double* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= R * K) return;
  if (R == N) {
    out[k] = g[k];
  } else {
    double s = 0.0;
    for (int n = 0; n < N; ++n) s += g[(size_t)n * K + k];
    out[k] = s;
  }
}

struct RolloutBwdLayout {
  size_t md, Sd, Sxd, f1, Sff, cross, mb, Sb, f1_bar, Sff_bar, cross_bar, Sxd_bar, md_bar, Sd_bar, gZ, gEll, gBeta, cg, predict, predict_bwd, persist, total;
  size_t predict_bytes, predict_bwd_bytes;
};

static RolloutBwdLayout rollout_bwd_layout(const gpp_gp_model* dyn, int N, int Dx, int Mp, int H) {
  RolloutBwdLayout lo{};
  const int D = dyn->D, L = dyn->P, De = D - 1;
  size_t off = 0;
  auto take = [&](size_t doubles) { size_t o = off; off = rollout_align_up(off + doubles * sizeof(double), 256); return o; };
  lo.md = take((size_t)N * D);
  lo.Sd = take((size_t)N * D * D);
  lo.Sxd = take((size_t)N * Dx * D);
  lo.f1 = take((size_t)N * L);
  lo.Sff = take((size_t)N * L * L);
  lo.cross = take((size_t)N * D * L);
  lo.mb = take((size_t)N * Dx);
  lo.Sb = take((size_t)N * Dx * Dx);
  lo.f1_bar = take((size_t)N * L);
  lo.Sff_bar = take((size_t)N * L * L);
  lo.cross_bar = take((size_t)N * D * L);
  lo.Sxd_bar = take((size_t)N * Dx * D);
  lo.md_bar = take((size_t)N * D);
  lo.Sd_bar = take((size_t)N * D * D);
  lo.gZ = take((size_t)N * Mp * De);
  lo.gEll = take((size_t)N * De);
  lo.gBeta = take((size_t)N * Mp);
  lo.cg = take((size_t)kGradRing * N * (Dx + Dx * (Dx + 1) / 2));
  lo.predict_bytes = gpp_mm_gp_predict_workspace_bytes(dyn, N);
  lo.predict_bwd_bytes = gpp_mm_gp_predict_bwd_workspace_bytes(dyn, N);
  lo.predict = off;
  off += rollout_align_up(lo.predict_bytes, 256);
  lo.predict_bwd = off;
  off += rollout_align_up(lo.predict_bwd_bytes, 256);
  lo.persist = off;                         // region of the persistent reverse sweep (rollout_persist.cu)
  if (persist_bwd_supported(dyn, N, Dx)) off += rollout_align_up(persist_bwd_layout(dyn, N, Dx, H).total, 256);
  lo.total = off;
  return lo;
}

}  // namespace gpp

extern "C" {

size_t gpp_rollout_mm_bwd_workspace_bytes(const gpp_gp_model* dynamics, int N, int Dx, int Mp, int H) {
  if (!dynamics || N <= 0 || Mp <= 0 || H < 0) return 0;
  return gpp::rollout_bwd_layout(dynamics, N, Dx, Mp, H).total;
}

int gpp_rollout_mm_bwd(const gpp_gp_model* dynamics, int N, int Dx, int num_active, const int* active_dims /*host*/,
                       int R, int Mp, const double* policy_Z, const double* policy_lengthscales, const double* policy_variance,
                       const double* policy_beta, double squash_scale, double squash_shift,
                       const double* cost_target, const double* cost_W, int H, const double* traj_m, const double* traj_S,
                       const double* saved, const double* loss_bar, double* Z_bar, double* lengthscales_bar, double* beta_bar,
                       double* m0_bar, double* S0_bar, void* workspace, size_t workspace_bytes, int* info, void* stream_) {
  GPP_NVTX_RANGE();
  using namespace gpp;
  GPP_REQUIRE(dynamics && policy_Z && policy_lengthscales && policy_variance && policy_beta && cost_target && cost_W && traj_m && traj_S &&
                  Z_bar && lengthscales_bar && beta_bar && workspace, GPP_ERR_NULL, "gpp_rollout_mm_bwd: null argument");
  GPP_REQUIRE(N >= 1 && H >= 0 && Dx >= 1 && Dx <= GPP_SMALL_MAX && Mp >= 1, GPP_ERR_BAD_SHAPE, "gpp_rollout_mm_bwd: bad sizes N=%d H=%d Dx=%d Mp=%d", N, H, Dx, Mp);
  GPP_REQUIRE(num_active >= 0 && num_active <= 4 && num_active <= Dx, GPP_ERR_BAD_SHAPE, "gpp_rollout_mm_bwd: bad number of encoded dims %d", num_active);
  GPP_REQUIRE(R == 1 || R == N, GPP_ERR_BAD_SHAPE, "gpp_rollout_mm_bwd: R=%d must be 1 (shared policy) or N=%d", R, N);
  RolloutMMParams p{};
  p.enc.Dx = Dx; p.enc.na = num_active;
  for (int k = 0; k < num_active; ++k) {
    GPP_REQUIRE(active_dims[k] >= 0 && active_dims[k] < Dx, GPP_ERR_BAD_SHAPE, "gpp_rollout_mm_bwd: active dim %d out of range", active_dims[k]);
    p.enc.active[k] = active_dims[k];
  }
  p.enc.finish();
  p.N = N; p.Dx = Dx; p.De = Dx + num_active; p.D = p.De + 1; p.L = dynamics->P;
  GPP_REQUIRE(p.De <= GPP_SMALL_MAX - 1, GPP_ERR_UNSUPPORTED, "gpp_rollout_mm_bwd: encoded dimension %d too large", p.De);
  GPP_REQUIRE(dynamics->D == p.D && dynamics->P == Dx, GPP_ERR_BAD_SHAPE, "gpp_rollout_mm_bwd: dynamics dims (%d in, %d out) do not match the state (%d, %d)",
              dynamics->D, dynamics->P, p.D, Dx);
  GPP_REQUIRE(Dx + Dx * (Dx + 1) / 2 <= 128, GPP_ERR_UNSUPPORTED, "gpp_rollout_mm_bwd: state dimension too large");
  RolloutBwdLayout lo = rollout_bwd_layout(dynamics, N, Dx, Mp, H);
  GPP_REQUIRE(workspace_bytes >= lo.total, GPP_ERR_WORKSPACE, "gpp_rollout_mm_bwd: workspace %zu < required %zu", workspace_bytes, lo.total);
  cudaStream_t stream = (cudaStream_t)stream_;
  char* ws = (char*)workspace;
  auto D_ = [&](size_t off) { return (double*)(ws + off); };
  p.R = R; p.Mp = Mp; p.pZ = policy_Z; p.pEll = policy_lengthscales; p.pVar = policy_variance; p.pBeta = policy_beta;
  p.scale = squash_scale; p.shift = squash_shift; p.target = cost_target; p.W = cost_W;
  p.md = D_(lo.md); p.Sd = D_(lo.Sd); p.Sxd = D_(lo.Sxd); p.f1 = D_(lo.f1); p.Sff = D_(lo.Sff); p.cross = D_(lo.cross);
  p.loss = nullptr; p.traj_m = nullptr; p.traj_S = nullptr; p.info = info;
  RolloutBwdBuffers bw;
  bw.mb = D_(lo.mb); bw.Sb = D_(lo.Sb); bw.f1_bar = D_(lo.f1_bar); bw.Sff_bar = D_(lo.Sff_bar); bw.cross_bar = D_(lo.cross_bar);
  bw.Sxd_bar = D_(lo.Sxd_bar); bw.md_bar = D_(lo.md_bar); bw.Sd_bar = D_(lo.Sd_bar); bw.gZ = D_(lo.gZ); bw.gEll = D_(lo.gEll); bw.gBeta = D_(lo.gBeta);
  // zero the running adjoint and the per-rollout parameter gradients (contiguous ranges of the workspace)
  GPP_CUDA_OK(cudaMemsetAsync(ws + lo.mb, 0, lo.f1_bar - lo.mb, stream));
  GPP_CUDA_OK(cudaMemsetAsync(ws + lo.gZ, 0, lo.cg - lo.gZ, stream));
  const size_t sm = (size_t)N * Dx, sS = (size_t)N * Dx * Dx;
  GPP_REQUIRE(p.De >= 1 && p.De <= 7, GPP_ERR_UNSUPPORTED, "gpp_rollout_mm_bwd: unsupported encoded dimension %d", p.De);
  const RolloutSaved sv(N, Dx, p.D, p.L);
  const int ndir = Dx + Dx * (Dx + 1) / 2;
  double* cg = D_(lo.cg);
  // the reverse sweep on the device: one persistent cooperative launch (rollout_persist.cu); needs the forward's per-step block
  const int mode = rollout_mode();
  const bool can_persist = saved != nullptr && persist_bwd_supported(dynamics, N, Dx);
  GPP_REQUIRE(mode != GPP_ROLLOUT_PERSIST || can_persist, GPP_ERR_UNSUPPORTED,
              "gpp_rollout_mm_bwd: the persistent kernel needs `saved` and a supported model (D=%d, M=%d)", dynamics->D, dynamics->M);
  const bool persist = mode != GPP_ROLLOUT_LEGACY && can_persist;
  if (persist) {
    const int rc = rollout_mm_bwd_persist(dynamics, p, bw, H, traj_m, traj_S, saved, loss_bar, ws + lo.persist, stream);
    if (rc != GPP_OK) return rc;
  }
  for (int t = H - 1; t >= 0 && !persist; --t) {
    if (t == H - 1 || (t + 1) % kGradRing == 0) {   // cost gradients of the states t+1 of this ring of steps, one launch
      const int t0 = t / kGradRing * kGradRing, count = t - t0 + 1;
      const int total = count * N * ndir;
      k_cost_grad_ring<<<(total + 63) / 64, 64, 0, stream>>>(p, traj_m + (size_t)(t0 + 1) * sm, traj_S + (size_t)(t0 + 1) * sS, count, ndir, cg);
      count_launch();
    }
    p.m = const_cast<double*>(traj_m) + (size_t)t * sm;     // read-only in the kernels launched below
    p.S = const_cast<double*>(traj_S) + (size_t)t * sS;
    int rc = GPP_OK;
    if (saved) {   // step t's joint moments, Cov(x, d) and cross term as the forward stored them (read-only here)
      double* base = const_cast<double*>(saved) + (size_t)t * sv.per_step;
      p.md = base + sv.md; p.Sd = base + sv.Sd; p.Sxd = base + sv.Sxd; p.cross = base + sv.cross; p.pre = base + sv.pre;
    } else {       // recompute them: pre stage + fused forward predict
      switch (p.De) {
#define GPP_CASE(d) case d: k_step_pre<d><<<N, 128, 0, stream>>>(p); break;
        GPP_CASE(1) GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7)
#undef GPP_CASE
        default: break;
      }
      count_launch();
      rc = mm_predict_enqueue(dynamics, p.md, p.Sd, N, p.f1, p.Sff, p.cross, 1, 0.0, ws + lo.predict, lo.predict_bytes, info, stream);
      if (rc != GPP_OK) return rc;
    }
    k_bwd_post<<<N, 32, 0, stream>>>(p, cg + (size_t)(t % kGradRing) * N * ndir, loss_bar, bw);
    rc = mm_predict_bwd_enqueue(dynamics, p.md, p.Sd, N, bw.f1_bar, bw.Sff_bar, bw.cross_bar, 1, bw.md_bar, bw.Sd_bar,
                                ws + lo.predict_bwd, lo.predict_bwd_bytes, info, stream);
    if (rc != GPP_OK) return rc;
    switch (p.De) {
#define GPP_CASE(d) case d: k_bwd_pre<d><<<N, 128, 0, stream>>>(p, bw); break;
      GPP_CASE(1) GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7)
#undef GPP_CASE
      default: break;
    }
    count_launch(2);
  }
  if (m0_bar) GPP_CUDA_OK(cudaMemcpyAsync(m0_bar, bw.mb, sizeof(double) * sm, cudaMemcpyDeviceToDevice, stream));
  if (S0_bar) GPP_CUDA_OK(cudaMemcpyAsync(S0_bar, bw.Sb, sizeof(double) * sS, cudaMemcpyDeviceToDevice, stream));
  const int De = p.De;
  k_reduce_param_grads<<<(R * Mp * De + 127) / 128, 128, 0, stream>>>(bw.gZ, N, R, Mp * De, Z_bar);
  k_reduce_param_grads<<<(R * De + 127) / 128, 128, 0, stream>>>(bw.gEll, N, R, De, lengthscales_bar);
  k_reduce_param_grads<<<(R * Mp + 127) / 128, 128, 0, stream>>>(bw.gBeta, N, R, Mp, beta_bar);
  count_launch(3);
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

}  // extern "C"
