// Moment-matched rollout: H sequential steps of  encoder -> RBF policy -> squash -> joint (e,u) -> GP dynamics ->
// cross-covariance re-assembly -> Euler moment update -> expected cost, for N independent Gaussian initial states
// (and R = 1 or N policy parameter sets), entirely on the device.
//
// Replaces, per step, upstream gpflow_pilco/dynamics/forward_sde.py:95-137 (+ the rules it dispatches to),
// gpflow_pilco/dynamics/solvers.py:110-135 and the loss callback gpflow_pilco/loops/pilco.py:199-205; the H-loop
// is upstream's tf.foldl (solvers.py:84-105).  Launch structure per step (all on the caller's stream, capturable in
// a CUDA graph — no allocation, no synchronisation):
//   k_step_pre   one CTA per rollout: encoder rule, policy moment matching (Psi1/Psi2 of the small policy GP done by the
//                CTA's threads), squashing link, joint moments of d = (e,u), and Sxd = Cov(x, d) rows (forward_sde.py:105-124)
//   GP predict   the fused kernels of mm_predict.cu on the N joint states (3 launches); its finalize kernel carries the step's
//                epilogue (model.cuh EulerPost): Sxf = Sxd cross, Euler update (solvers.py:128-129), trajectory slice, cost ring slot
//   k_cost_ring + k_cost_accumulate   every 32 steps: encoder rule + expected cost (components.py:30-37) of the ring's states in
//                one parallel launch, added to the loss in step order (the cost does not feed back into the state).
// gpp_rollout_mm_fwd_save additionally keeps each step's (md, Sd, Sxd, cross, pre-stage block) for gpp_rollout_mm_bwd (rollout_mm_bwd.cu);
// gpp_policy_prepare / gpp_policy_prepare_bwd map the policy's q_mu to beta = Kuu^-1 m and back.
#include "rollout_mm_common.cuh"
#include "rollout_persist.h"

namespace gpp {


// ---------------------------------------------------------------------------------------------------------
// policy weights: beta_r = Kuu_r^-1 m_r  for R small kernel regressors (one CTA each; Cholesky in shared memory)
//   whitened: beta = Luu^-T q_mu     else: beta = Kuu^-1 q_mu        (upstream moment_matching/models.py:228-235)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_policy_prepare(int Mp, int Dp, const double* __restrict__ Z, const double* __restrict__ ell,
                                                        const double* __restrict__ var, const double* __restrict__ q_mu,
                                                        int whiten, double jitter, double* __restrict__ beta, int* info) {
  extern __shared__ double sm[];
  double* K = sm;            // [Mp][Mp+1]
  double* v = K + Mp * (Mp + 1);
  const int r = blockIdx.x, tid = threadIdx.x, ld = Mp + 1;
  const double* Zr = Z + (size_t)r * Mp * Dp;
  const double* er = ell + (size_t)r * Dp;
  for (int idx = tid; idx < Mp * Mp; idx += blockDim.x) {
    int i = idx / Mp, j = idx % Mp;
    double acc = 0.0;
    for (int d = 0; d < Dp; ++d) {
      double t = (Zr[i * Dp + d] - Zr[j * Dp + d]) / er[d];
      acc = fma(t, t, acc);
    }
    K[i * ld + j] = var[r] * exp(-0.5 * acc) + (i == j ? jitter : 0.0);
  }
  for (int i = tid; i < Mp; i += blockDim.x) v[i] = q_mu[(size_t)r * Mp + i];
  __syncthreads();
  // right-looking Cholesky, column by column
  for (int j = 0; j < Mp; ++j) {
    if (tid == 0) {
      double d = K[j * ld + j];
      if (!(d > 0.0)) flag_not_pd(info, r);
      K[j * ld + j] = sqrt(d);
    }
    __syncthreads();
    double djj = K[j * ld + j];
    for (int i = j + 1 + tid; i < Mp; i += blockDim.x) K[i * ld + j] /= djj;
    __syncthreads();
    for (int idx = tid; idx < (Mp - j - 1) * (Mp - j - 1); idx += blockDim.x) {
      int a = j + 1 + idx / (Mp - j - 1), b = j + 1 + idx % (Mp - j - 1);
      if (b <= a) K[a * ld + b] -= K[a * ld + j] * K[b * ld + j];
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (!whiten)   // w = L^-1 q
      for (int i = 0; i < Mp; ++i) {
        double t = v[i];
        for (int k = 0; k < i; ++k) t -= K[i * ld + k] * v[k];
        v[i] = t / K[i * ld + i];
      }
    for (int i = Mp - 1; i >= 0; --i) {   // beta = L^-T w
      double t = v[i];
      for (int k = i + 1; k < Mp; ++k) t -= K[k * ld + i] * v[k];
      v[i] = t / K[i * ld + i];
    }
  }
  __syncthreads();
  for (int i = tid; i < Mp; i += blockDim.x) beta[(size_t)r * Mp + i] = v[i];
}

// ---------------------------------------------------------------------------------------------------------
// adjoint of k_policy_prepare: beta_bar -> (q_mu_bar, Z_bar +=, lengthscales_bar +=) through beta = Kuu^-1 m.
//   not whitened: w = Kuu^-1 beta_bar,  q_bar = w,  K_bar = -(w beta^T + beta w^T)/2
//   whitened (beta = L^-T q): w = L^-1 beta_bar, q_bar = w, L_bar = -tril(beta w^T), P = Phi(L^T L_bar) (lower triangle, halved
//                             diagonal), K_bar = L^-T (P + P^T)/2 L^-1   (reverse-mode Cholesky)
//   K_ij = var exp(-|z_i - z_j|^2_ell / 2):  Z_bar_id -= sum_j 2 K_bar_ij K_ij (z_id - z_jd) / ell_d^2,
//                                            ell_bar_d += sum_ij K_bar_ij K_ij (z_id - z_jd)^2 / ell_d^3
// One CTA per parameter set, three Mp x Mp matrices in shared memory (upstream: TF autodiff through gpflow.covariances.Kuu,
// tf.linalg.cholesky and the triangular solves of moment_matching/models.py:216-235).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_policy_prepare_bwd(int Mp, int Dp, const double* __restrict__ Z, const double* __restrict__ ell,
                                                            const double* __restrict__ var, const double* __restrict__ beta,
                                                            const double* __restrict__ beta_bar, int whiten, double jitter,
                                                            double* __restrict__ Z_bar, double* __restrict__ ell_bar, double* __restrict__ q_bar) {
  extern __shared__ double sm[];
  const int ld = Mp + 1;
  double* Lm = sm;                 // Cholesky factor of Kuu (lower)
  double* A = Lm + Mp * ld;        // K_bar
  double* B = A + Mp * ld;         // scratch
  double* w = B + Mp * ld;
  double* b = w + Mp;
  const int r = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const double* Zr = Z + (size_t)r * Mp * Dp;
  const double* er = ell + (size_t)r * Dp;
  for (int idx = tid; idx < Mp * Mp; idx += nt) {
    const int i = idx / Mp, j = idx % Mp;
    double acc = 0.0;
    for (int d = 0; d < Dp; ++d) {
      const double t = (Zr[i * Dp + d] - Zr[j * Dp + d]) / er[d];
      acc = fma(t, t, acc);
    }
    const double k = var[r] * exp(-0.5 * acc);
    B[i * ld + j] = k;                                   // kernel matrix without jitter (needed again at the end)
    Lm[i * ld + j] = k + (i == j ? jitter : 0.0);
  }
  for (int i = tid; i < Mp; i += nt) { w[i] = beta_bar[(size_t)r * Mp + i]; b[i] = beta[(size_t)r * Mp + i]; }
  __syncthreads();
  for (int j = 0; j < Mp; ++j) {                         // right-looking Cholesky, as in k_policy_prepare
    if (tid == 0) Lm[j * ld + j] = sqrt(Lm[j * ld + j]);
    __syncthreads();
    const double djj = Lm[j * ld + j];
    for (int i = j + 1 + tid; i < Mp; i += nt) Lm[i * ld + j] /= djj;
    __syncthreads();
    for (int idx = tid; idx < (Mp - j - 1) * (Mp - j - 1); idx += nt) {
      const int a = j + 1 + idx / (Mp - j - 1), c = j + 1 + idx % (Mp - j - 1);
      if (c <= a) Lm[a * ld + c] -= Lm[a * ld + j] * Lm[c * ld + j];
    }
    __syncthreads();
  }
  if (tid == 0) {
    for (int i = 0; i < Mp; ++i) {                       // w = L^-1 beta_bar
      double t = w[i];
      for (int k = 0; k < i; ++k) t -= Lm[i * ld + k] * w[k];
      w[i] = t / Lm[i * ld + i];
    }
    if (!whiten)
      for (int i = Mp - 1; i >= 0; --i) {                // w = L^-T w = Kuu^-1 beta_bar
        double t = w[i];
        for (int k = i + 1; k < Mp; ++k) t -= Lm[k * ld + i] * w[k];
        w[i] = t / Lm[i * ld + i];
      }
  }
  __syncthreads();
  for (int i = tid; i < Mp; i += nt) q_bar[(size_t)r * Mp + i] = w[i];
  if (!whiten) {
    for (int idx = tid; idx < Mp * Mp; idx += nt) {
      const int i = idx / Mp, j = idx % Mp;
      A[i * ld + j] = -0.5 * (w[i] * b[j] + b[i] * w[j]);
    }
    __syncthreads();
  } else {
    // P = Phi(L^T L_bar), L_bar = -tril(beta w^T):  P_ij = -sum_{k >= i} L_ki b_k w_j  (i >= j)
    double* P = A;
    for (int idx = tid; idx < Mp * Mp; idx += nt) {
      const int i = idx / Mp, j = idx % Mp;
      double v = 0.0;
      if (i >= j) {
        for (int k = i; k < Mp; ++k) v = fma(Lm[k * ld + i], b[k], v);
        v *= -w[j];
        if (i == j) v *= 0.5;
      }
      P[i * ld + j] = v;
    }
    __syncthreads();
    // S = (P + P^T)/2 (symmetric) -> X = L^-T S, one column per thread; then K_bar = L^-T X^T
    double* X = B;                    // scratch: the kernel matrix kept in B is recomputed after the two solves
    for (int c = tid; c < Mp; c += nt) {
      for (int i = Mp - 1; i >= 0; --i) {
        double t = 0.5 * (P[i * ld + c] + P[c * ld + i]);
        for (int k = i + 1; k < Mp; ++k) t -= Lm[k * ld + i] * X[k * ld + c];
        X[i * ld + c] = t / Lm[i * ld + i];
      }
    }
    __syncthreads();
    for (int c = tid; c < Mp; c += nt) {                 // column c of K_bar = L^-T (row c of X)^T
      for (int i = Mp - 1; i >= 0; --i) {
        double t = X[c * ld + i];
        for (int k = i + 1; k < Mp; ++k) t -= Lm[k * ld + i] * A[k * ld + c];
        A[i * ld + c] = t / Lm[i * ld + i];
      }
    }
    __syncthreads();
    for (int idx = tid; idx < Mp * Mp; idx += nt) {      // B <- kernel matrix again
      const int i = idx / Mp, j = idx % Mp;
      double acc = 0.0;
      for (int d = 0; d < Dp; ++d) {
        const double t = (Zr[i * Dp + d] - Zr[j * Dp + d]) / er[d];
        acc = fma(t, t, acc);
      }
      B[i * ld + j] = var[r] * exp(-0.5 * acc);
    }
    __syncthreads();
  }
  // G = K_bar o K, then the centre and lengthscale gradients (fixed summation order)
  for (int idx = tid; idx < Mp * Dp; idx += nt) {
    const int i = idx / Dp, d = idx % Dp;
    double acc = 0.0;
    for (int j = 0; j < Mp; ++j)
      acc = fma((A[i * ld + j] + A[j * ld + i]) * B[i * ld + j], Zr[i * Dp + d] - Zr[j * Dp + d], acc);
    Z_bar[(size_t)r * Mp * Dp + idx] -= acc / (er[d] * er[d]);
  }
  for (int d = tid; d < Dp; d += nt) {
    double acc = 0.0;
    for (int i = 0; i < Mp; ++i)
      for (int j = 0; j < Mp; ++j) {
        const double df = Zr[i * Dp + d] - Zr[j * Dp + d];
        acc = fma(A[i * ld + j] * B[i * ld + j], df * df, acc);
      }
    ell_bar[(size_t)r * Dp + d] += acc / (er[d] * er[d] * er[d]);
  }
}

constexpr int kCostRing = 32;   // steps whose costs are evaluated by one k_cost_ring launch

// expected cost of `count` ring slots x N rollouts, one thread each  (upstream loops/pilco.py:199-205 with components.py:30-37)
__global__ void k_cost_ring(RolloutMMParams p, const double* __restrict__ ring_m, const double* __restrict__ ring_S, int count,
                            double* __restrict__ costbuf) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count * p.N) return;
  const int Dx = p.Dx;
  double m[GPP_SMALL_MAX], S[GPP_SMALL_MAX * GPP_SMALL_MAX];
  for (int i = 0; i < Dx; ++i) m[i] = ring_m[(size_t)idx * Dx + i];
  for (int i = 0; i < Dx * Dx; ++i) S[i] = ring_S[(size_t)idx * Dx * Dx + i];
  double me[GPP_SMALL_MAX], See[GPP_SMALL_MAX * GPP_SMALL_MAX], Cxe[GPP_SMALL_MAX * GPP_SMALL_MAX];
  mm_encoder<double>(p.enc, m, S, me, See, Cxe);
  costbuf[idx] = expected_cost<double>(p.De, me, See, p.target, p.W);
}

// loss[n] += costs of the ring slots in step order (fixed summation order: same result as a per-step accumulation)
__global__ void k_cost_accumulate(int N, int count, const double* __restrict__ costbuf, double* __restrict__ loss) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double acc = loss[n];
  for (int t = 0; t < count; ++t) acc += costbuf[(size_t)t * N + n];
  loss[n] = acc;
}

__global__ void k_rollout_init(RolloutMMParams p, const double* m0, const double* S0) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= p.N) return;
  const int Dx = p.Dx;
  for (int i = 0; i < Dx; ++i) {
    double v = m0[(size_t)n * Dx + i];
    p.m[(size_t)n * Dx + i] = v;
    if (p.traj_m) p.traj_m[(size_t)n * Dx + i] = v;
  }
  for (int i = 0; i < Dx * Dx; ++i) {
    double v = S0[(size_t)n * Dx * Dx + i];
    p.S[(size_t)n * Dx * Dx + i] = v;
    if (p.traj_S) p.traj_S[(size_t)n * Dx * Dx + i] = v;
  }
  p.loss[n] = 0.0;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct RolloutLayout {
  size_t m, S, md, Sd, Sxd, f1, Sff, cross, ring_m, ring_S, costbuf, predict, persist, total, predict_bytes;
};

static RolloutLayout rollout_layout(const gpp_gp_model* dyn, int N, int Dx) {
  RolloutLayout lo{};
  const int D = dyn->D, L = dyn->P;
  size_t off = 0;
  auto take = [&](size_t doubles) { size_t o = off; off = align_up(off + doubles * sizeof(double), 256); return o; };
  lo.m = take((size_t)N * Dx);
  lo.S = take((size_t)N * Dx * Dx);
  lo.md = take((size_t)N * D);
  lo.Sd = take((size_t)N * D * D);
  lo.Sxd = take((size_t)N * Dx * D);
  lo.f1 = take((size_t)N * L);
  lo.Sff = take((size_t)N * L * L);
  lo.cross = take((size_t)N * D * L);
  lo.ring_m = take((size_t)kCostRing * N * Dx);
  lo.ring_S = take((size_t)kCostRing * N * Dx * Dx);
  lo.costbuf = take((size_t)kCostRing * N);
  lo.predict = off;
  lo.predict_bytes = gpp_mm_gp_predict_workspace_bytes(dyn, N);
  off += align_up(lo.predict_bytes, 256);
  lo.persist = off;                         // region of the persistent kernel (packs, tile partials, counters)
  if (persist_fwd_supported(dyn, N, Dx)) off += align_up(persist_fwd_layout(dyn, N).total, 256);
  lo.total = off;
  return lo;
}

}  // namespace gpp

static int rollout_mm_fwd_impl(const gpp_gp_model* dynamics, int N, int Dx, int num_active, const int* active_dims /*host*/,
                               int R, int Mp, const double* policy_Z, const double* policy_lengthscales, const double* policy_variance,
                               const double* policy_beta, double squash_scale, double squash_shift,
                               const double* cost_target, const double* cost_W, int H, const double* m0, const double* S0,
                               double* loss, double* traj_m, double* traj_S, double* m_final, double* S_final, double* saved,
                               void* workspace, size_t workspace_bytes, int* info, void* stream_) {
  using namespace gpp;
  GPP_REQUIRE(dynamics && policy_Z && policy_lengthscales && policy_variance && policy_beta && cost_target && cost_W && m0 && S0 &&
                  loss && workspace, GPP_ERR_NULL, "gpp_rollout_mm_fwd: null argument");
  GPP_REQUIRE(N >= 1 && H >= 0 && Dx >= 1 && Dx <= GPP_SMALL_MAX, GPP_ERR_BAD_SHAPE, "gpp_rollout_mm_fwd: bad sizes N=%d H=%d Dx=%d", N, H, Dx);
  GPP_REQUIRE(num_active >= 0 && num_active <= 4 && num_active <= Dx, GPP_ERR_BAD_SHAPE, "gpp_rollout_mm_fwd: bad number of encoded dims %d", num_active);
  GPP_REQUIRE(R == 1 || R == N, GPP_ERR_BAD_SHAPE, "gpp_rollout_mm_fwd: R=%d must be 1 (shared policy) or N=%d", R, N);
  RolloutMMParams p{};
  p.enc.Dx = Dx; p.enc.na = num_active;
  for (int k = 0; k < num_active; ++k) {
    GPP_REQUIRE(active_dims[k] >= 0 && active_dims[k] < Dx, GPP_ERR_BAD_SHAPE, "gpp_rollout_mm_fwd: active dim %d out of range", active_dims[k]);
    p.enc.active[k] = active_dims[k];
  }
  p.enc.finish();
  p.N = N; p.Dx = Dx; p.De = Dx + num_active; p.D = p.De + 1; p.L = dynamics->P;
  GPP_REQUIRE(p.De <= GPP_SMALL_MAX - 1, GPP_ERR_UNSUPPORTED, "gpp_rollout_mm_fwd: encoded dimension %d too large", p.De);
  GPP_REQUIRE(dynamics->D == p.D, GPP_ERR_BAD_SHAPE, "gpp_rollout_mm_fwd: dynamics input dim %d != encoded state + action = %d", dynamics->D, p.D);
  GPP_REQUIRE(dynamics->P == Dx, GPP_ERR_BAD_SHAPE, "gpp_rollout_mm_fwd: dynamics output dim %d != state dim %d", dynamics->P, Dx);
  RolloutLayout lo = rollout_layout(dynamics, N, Dx);
  GPP_REQUIRE(workspace_bytes >= lo.total, GPP_ERR_WORKSPACE, "gpp_rollout_mm_fwd: workspace %zu < required %zu", workspace_bytes, lo.total);
  cudaStream_t stream = (cudaStream_t)stream_;
  char* ws = (char*)workspace;
  p.R = R; p.Mp = Mp; p.pZ = policy_Z; p.pEll = policy_lengthscales; p.pVar = policy_variance; p.pBeta = policy_beta;
  p.scale = squash_scale; p.shift = squash_shift; p.target = cost_target; p.W = cost_W;
  p.m = (double*)(ws + lo.m); p.S = (double*)(ws + lo.S); p.loss = loss;
  p.md = (double*)(ws + lo.md); p.Sd = (double*)(ws + lo.Sd); p.Sxd = (double*)(ws + lo.Sxd);
  p.f1 = (double*)(ws + lo.f1); p.Sff = (double*)(ws + lo.Sff); p.cross = (double*)(ws + lo.cross);
  p.traj_m = traj_m; p.traj_S = traj_S; p.info = info;
  GPP_REQUIRE((traj_m == nullptr) == (traj_S == nullptr), GPP_ERR_NULL, "gpp_rollout_mm_fwd: traj_m and traj_S go together");

  // the H-loop on the device: one persistent cooperative launch for the whole sweep (rollout_persist.cu)
  const int mode = rollout_mode();
  const bool can_persist = persist_fwd_supported(dynamics, N, Dx);
  GPP_REQUIRE(mode != GPP_ROLLOUT_PERSIST || can_persist, GPP_ERR_UNSUPPORTED,
              "gpp_rollout_mm_fwd: the persistent kernel does not support this model (D=%d, %d tiles of 64 x 64)", dynamics->D,
              dynamics->tables[0][0].nslots);
  if (mode != GPP_ROLLOUT_LEGACY && can_persist)
    return rollout_mm_fwd_persist(dynamics, p, H, m0, S0, m_final, S_final, saved, ws + lo.persist, stream);

  const int tb = 64, gb = (N + tb - 1) / tb;
  k_rollout_init<<<gb, tb, 0, stream>>>(p, m0, S0);
  count_launch();
  const RolloutSaved sv(N, Dx, p.D, p.L);
  double* ring_m = (double*)(ws + lo.ring_m);
  double* ring_S = (double*)(ws + lo.ring_S);
  double* costbuf = (double*)(ws + lo.costbuf);
  for (int t = 0; t < H; ++t) {
    if (saved) {   // the joint moments, Cov(x, d) and the pre-inverted cross term of every step are written where the backward reads them
      double* base = saved + (size_t)t * sv.per_step;
      p.md = base + sv.md; p.Sd = base + sv.Sd; p.Sxd = base + sv.Sxd; p.cross = base + sv.cross; p.pre = base + sv.pre;
    }
    switch (p.De) {
#define GPP_CASE(d) case d: k_step_pre<d><<<N, 128, 0, stream>>>(p); break;
      GPP_CASE(1) GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7)
#undef GPP_CASE
      default: set_error("gpp_rollout_mm_fwd: unsupported encoded dimension %d", p.De); return GPP_ERR_UNSUPPORTED;
    }
    count_launch();
    // Euler update, trajectory slice and cost-ring slot of step t ride on the predict's finalize kernel.  The expected cost of the
    // new state is evaluated later, for a whole ring of steps at once (k_cost_ring): it does not feed back into the state, so the
    // serial LU / determinant of every step stays off the critical path of the rollout
    EulerPost post;
    post.m = p.m; post.S = p.S; post.Sxd = p.Sxd; post.Dx = Dx;
    post.traj_m = traj_m ? traj_m + (size_t)(t + 1) * N * Dx : nullptr;
    post.traj_S = traj_S ? traj_S + (size_t)(t + 1) * N * Dx * Dx : nullptr;
    post.ring_m = ring_m + (size_t)(t % kCostRing) * N * Dx;
    post.ring_S = ring_S + (size_t)(t % kCostRing) * N * Dx * Dx;
    int rc = mm_predict_enqueue(dynamics, p.md, p.Sd, N, p.f1, p.Sff, p.cross, 1, 0.0, ws + lo.predict, lo.predict_bytes, info, stream, &post);
    if (rc != GPP_OK) return rc;
    if ((t + 1) % kCostRing == 0 || t == H - 1) {
      const int count = t % kCostRing + 1;
      k_cost_ring<<<(count * N + 63) / 64, 64, 0, stream>>>(p, ring_m, ring_S, count, costbuf);
      k_cost_accumulate<<<gb, tb, 0, stream>>>(N, count, costbuf, loss);
      count_launch(2);
    }
  }
  if (m_final) GPP_CUDA_OK(cudaMemcpyAsync(m_final, p.m, sizeof(double) * N * Dx, cudaMemcpyDeviceToDevice, stream));
  if (S_final) GPP_CUDA_OK(cudaMemcpyAsync(S_final, p.S, sizeof(double) * N * Dx * Dx, cudaMemcpyDeviceToDevice, stream));
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}


extern "C" {

int gpp_policy_prepare(int R, int Mp, int Dp, const double* Z, const double* lengthscales, const double* variance,
                       const double* q_mu, int whiten, double jitter, double* beta, int* info, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(Z && lengthscales && variance && q_mu && beta, GPP_ERR_NULL, "gpp_policy_prepare: null argument");
  GPP_REQUIRE(R >= 1 && Mp >= 1 && Dp >= 1, GPP_ERR_BAD_SHAPE, "gpp_policy_prepare: bad sizes R=%d Mp=%d Dp=%d", R, Mp, Dp);
  size_t smem = sizeof(double) * ((size_t)Mp * (Mp + 1) + Mp);
  GPP_REQUIRE(smem <= 200 * 1024, GPP_ERR_UNSUPPORTED, "gpp_policy_prepare: Mp=%d too large for the in-CTA Cholesky (use gpp_gp_model_create)", Mp);
  if (smem > 48 * 1024) GPP_CUDA_OK(cudaFuncSetAttribute(gpp::k_policy_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gpp::k_policy_prepare<<<R, 128, smem, (cudaStream_t)stream>>>(Mp, Dp, Z, lengthscales, variance, q_mu, whiten, jitter, beta, info);
  gpp::count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

int gpp_policy_prepare_bwd(int R, int Mp, int Dp, const double* Z, const double* lengthscales, const double* variance, const double* beta,
                           const double* beta_bar, int whiten, double jitter, double* Z_bar, double* lengthscales_bar, double* q_mu_bar,
                           void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(Z && lengthscales && variance && beta && beta_bar && Z_bar && lengthscales_bar && q_mu_bar, GPP_ERR_NULL,
              "gpp_policy_prepare_bwd: null argument");
  GPP_REQUIRE(R >= 1 && Mp >= 1 && Dp >= 1, GPP_ERR_BAD_SHAPE, "gpp_policy_prepare_bwd: bad sizes R=%d Mp=%d Dp=%d", R, Mp, Dp);
  size_t smem = sizeof(double) * (3 * (size_t)Mp * (Mp + 1) + 2 * Mp);
  GPP_REQUIRE(smem <= 200 * 1024, GPP_ERR_UNSUPPORTED, "gpp_policy_prepare_bwd: Mp=%d too large for the in-CTA Cholesky adjoint", Mp);
  if (smem > 48 * 1024) GPP_CUDA_OK(cudaFuncSetAttribute(gpp::k_policy_prepare_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gpp::k_policy_prepare_bwd<<<R, 128, smem, (cudaStream_t)stream>>>(Mp, Dp, Z, lengthscales, variance, beta, beta_bar, whiten, jitter, Z_bar,
                                                                    lengthscales_bar, q_mu_bar);
  gpp::count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

size_t gpp_rollout_mm_workspace_bytes(const gpp_gp_model* dynamics, int N, int Dx) {
  if (!dynamics || N <= 0) return 0;
  return gpp::rollout_layout(dynamics, N, Dx).total;
}

size_t gpp_rollout_mm_saved_doubles(const gpp_gp_model* dynamics, int N, int Dx, int H) {
  if (!dynamics || N <= 0 || H < 0) return 0;
  return gpp::RolloutSaved(N, Dx, dynamics->D, dynamics->P).per_step * (size_t)H;
}

int gpp_rollout_mm_fwd(const gpp_gp_model* dynamics, int N, int Dx, int num_active, const int* active_dims /*host*/,
                       int R, int Mp, const double* policy_Z, const double* policy_lengthscales, const double* policy_variance,
                       const double* policy_beta, double squash_scale, double squash_shift,
                       const double* cost_target, const double* cost_W, int H, const double* m0, const double* S0,
                       double* loss, double* traj_m, double* traj_S, double* m_final, double* S_final,
                       void* workspace, size_t workspace_bytes, int* info, void* stream) {
  GPP_NVTX_RANGE();
  return rollout_mm_fwd_impl(dynamics, N, Dx, num_active, active_dims, R, Mp, policy_Z, policy_lengthscales, policy_variance, policy_beta,
                             squash_scale, squash_shift, cost_target, cost_W, H, m0, S0, loss, traj_m, traj_S, m_final, S_final, nullptr,
                             workspace, workspace_bytes, info, stream);
}

int gpp_rollout_mm_fwd_save(const gpp_gp_model* dynamics, int N, int Dx, int num_active, const int* active_dims /*host*/,
                            int R, int Mp, const double* policy_Z, const double* policy_lengthscales, const double* policy_variance,
                            const double* policy_beta, double squash_scale, double squash_shift,
                            const double* cost_target, const double* cost_W, int H, const double* m0, const double* S0,
                            double* loss, double* traj_m, double* traj_S, double* m_final, double* S_final, double* saved,
                            void* workspace, size_t workspace_bytes, int* info, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(traj_m && traj_S && saved, GPP_ERR_NULL, "gpp_rollout_mm_fwd_save: traj_m, traj_S and saved are required (the backward reads them)");
  return rollout_mm_fwd_impl(dynamics, N, Dx, num_active, active_dims, R, Mp, policy_Z, policy_lengthscales, policy_variance, policy_beta,
                             squash_scale, squash_shift, cost_target, cost_W, H, m0, S0, loss, traj_m, traj_S, m_final, S_final, saved,
                             workspace, workspace_bytes, info, stream);
}

}  // extern "C"
