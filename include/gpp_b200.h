/* gpp_b200.h — C ABI of libgpp_b200.so: the B200-native (sm_100a) moment-matching / pathwise rollout path.
 *
 * Drop-in boundary for the hot path of j-wilson/GPflowPILCO.  Every entry point names the upstream
 * interface (file:line, relative to the upstream repo root) whose arithmetic it replaces; the Python
 * host side (gpflowpilco_b200/) re-registers the same dispatch keys and marshals tensors to these calls.
 *
 * Conventions
 *   - all array arguments are DEVICE pointers to float64, row-major, batch-leading ([N,...]) unless a
 *     parameter is documented as host; sizes are explicit; nothing is retained past the call except by
 *     gpp_gp_model_* handles; caller owns every buffer (inputs, outputs, workspace).
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no implicit synchronisation.
 *   - return value: GPP_OK or a negative gpp_status; gpp_last_error() gives a thread-local message.
 *   - `info` (device int32[1], may be NULL): set to 1+index of the first batch element whose D x D
 *     Cholesky failed (0 = all positive definite).  It is written asynchronously; read it after syncing.
 */
#ifndef GPP_B200_H
#define GPP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum gpp_status {
  GPP_OK = 0,
  GPP_ERR_BAD_SHAPE = -1,      /* inconsistent / non-positive sizes */
  GPP_ERR_UNSUPPORTED = -2,    /* e.g. D > GPP_MAX_D */
  GPP_ERR_NOT_PD = -3,         /* Kuu / Kyy Cholesky failed while building a model handle */
  GPP_ERR_CUDA = -4,
  GPP_ERR_WORKSPACE = -5,      /* workspace too small */
  GPP_ERR_NULL = -6
} gpp_status;

#define GPP_MAX_D 8            /* GP input dimension supported by the compiled kernels (1..8) */
#define GPP_MAX_L 8            /* latent GPs per model */

int gpp_version(void);
const char* gpp_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
unsigned long long gpp_launch_count(void);

/* ---- Psi1: eKxz[n,m] = E_{x~N(mu_n,cov_n)} k(x, z_m) ------------------------------------------------
 * replaces GPflow expectation(p, (kernel, Z)) as called at
 * gpflow_pilco/moment_matching/models.py:62,141,212 and gpflow_pilco/utils/kernel_expectation.py:87-88. */
int gpp_ekxz(const double* mu, const double* cov, int N, int D,
             const double* Z, int M, const double* lengthscales /*[D]*/, double variance,
             double* out /*[N,M]*/, int* info, void* stream);

/* ---- Psi2: eKzxKxz[n,i,j] = E[k1(z1_i,x) k2(x,z2_j)], materialised -------------------------------------
 * replaces `_E`, gpflow_pilco/utils/kernel_expectation.py:72-187.  Pass Z2 = NULL (and lengthscales2 = NULL)
 * for the same-kernel/same-feature case (:96-97). */
int gpp_ekzxkxz(const double* mu, const double* cov, int N, int D,
                const double* Z1, int M1, const double* lengthscales1, double variance1,
                const double* Z2, int M2, const double* lengthscales2, double variance2,
                double* out /*[N,M1,M2]*/, int* info, void* stream);

/* ---- sparse / exact GP model handle --------------------------------------------------------------------
 * Caches the step-invariant solves the reference redoes in every call
 * (gpflow_pilco/moment_matching/models.py:216-235: Kuu, its Cholesky, Luu^-1 q_mu, Luu^-1 q_sqrt):
 *   beta_l = Kuu_l^-1 m_l,   C_l = beta_l beta_l^T - [model_uncertainty](Kuu_l^-1 - Kuu_l^-1 S_l Kuu_l^-1).
 * Covers gpflow.models.SVGP with SeparateIndependent / LinearCoregionalization kernels (models.py:200-299),
 * single-output SVGP (L = 1, :129-197) and exact GPR (:44-111: pass Z = X, q_mu = Y - c, q_sqrt = NULL,
 * whiten = 0, kuu_jitter = noise variance).
 *   Z [L,M,D], lengthscales [L,D], variance [L], q_mu [M,L], q_sqrt [L,M,M] lower (NULL = zero),
 *   mean_const [P] (NULL = Zero mean), W [P,L] (NULL = no coregionalisation, then P = L). */
typedef struct gpp_gp_model gpp_gp_model;

int gpp_gp_model_create(gpp_gp_model** out, int L, int M, int D,
                        const double* Z, const double* lengthscales, const double* variance,
                        const double* q_mu, const double* q_sqrt, int whiten,
                        const double* mean_const, const double* W, int P,
                        const double* kuu_jitter /*host [L]*/, int model_uncertainty, void* stream);
int gpp_gp_model_destroy(gpp_gp_model* model);
/* Copies the cached weights into caller buffers (either may be NULL): beta [L,M], C [L,M,M]. */
int gpp_gp_model_weights(const gpp_gp_model* model, double* beta, double* C, void* stream);

/* ---- fused exact moment matching through the GP ---------------------------------------------------------
 * replaces _mm_gauss_svgp_mo / _so / _gpr, gpflow_pilco/moment_matching/models.py:44-299:
 *   f1 [N,P] mean, Sff [N,P,P] covariance (diagonal only filled when full_output_cov = 0),
 *   cross [N,D,P] = Cov(x,x)^-1 Cov(x,f)  (the reference's pre-inverted cross term, :264-277),
 *   `jitter` added to diag(Sff) (:293-296).  Psi2 is never materialised. */
size_t gpp_mm_gp_predict_workspace_bytes(const gpp_gp_model* model, int N);
int gpp_mm_gp_predict_fwd(const gpp_gp_model* model, const double* m /*[N,D]*/, const double* S /*[N,D,D]*/, int N,
                          double* f1, double* Sff, double* cross,
                          int full_output_cov, double jitter,
                          void* workspace, size_t workspace_bytes, int* info, void* stream);

/* Backward of gpp_mm_gp_predict_fwd (upstream: TensorFlow autodiff through moment_matching/models.py:200-299 and
 * utils/kernel_expectation.py:96-187, driven by tape.gradient in utils/optimizers.py:52-56): given the adjoints of the
 * outputs, f1_bar [N,P], Sff_bar [N,P,P] (only its diagonal is read when full_output_cov = 0), cross_bar [N,D,P] (any of
 * the three may be NULL = zero), returns m_bar [N,D] and the SYMMETRIC S_bar [N,D,D].  Closed form: Psi2 is re-contracted
 * with first/second-moment accumulators, never materialised.  Model parameters are treated as constants. */
size_t gpp_mm_gp_predict_bwd_workspace_bytes(const gpp_gp_model* model, int N);
int gpp_mm_gp_predict_bwd(const gpp_gp_model* model, const double* m, const double* S, int N,
                          const double* f1_bar, const double* Sff_bar, const double* cross_bar, int full_output_cov,
                          double* m_bar, double* S_bar, void* workspace, size_t workspace_bytes, int* info, void* stream);

/* ---- small moment-matching rules, batched (one thread per Gaussian state) -----------------------------------
 *   gpp_mm_encoder    e = [sin x_a, cos x_a, x_b]: mean me [N,De], covariance See [N,De,De], Cxe = Cov(x,e) [N,Dx,De]
 *                     (upstream moment_matching/components.py:19-57 with maths.py:143-176; De = Dx + num_active)
 *   gpp_mm_squash     u = scale (Phi(f) + shift), f ~ N(mf, vf): mean, variance and gain = Cov(f,f)^-1 Cov(f,u)
 *                     (upstream moment_matching/bijectors.py:21-69 chained with maths.py:47-78; Owen's T on device)
 *   gpp_mm_squash_nd  the same link on an A-dimensional Gaussian f ~ N(mf [N,A], Sf [N,A,A]): mean mu [N,A], covariance Su [N,A,A],
 *                     diagonal gain [N,A]; E[Phi(f_i) Phi(f_j)] by Genz's bivariate normal probabilities with upstream's lower
 *                     limit -9 (upstream moment_matching/bijectors.py:59-63 with utils/bvn.py:67-232)
 *   gpp_cost_gaussian E[-exp(-1/2 (e-t)^T W (e-t))] for e ~ N(me, See)   (upstream components.py:30-37)
 *   gpp_cost_samples  -exp(-1/2 (e-t)^T W (e-t))                           (upstream components.py:39-41)
 *   gpp_owens_t       Owen's T(h, a), 0 < a <= 1 (stands in for tfp.math.owens_t, upstream bijectors.py:15,58) */
int gpp_mm_encoder(int N, int Dx, int num_active, const int* active_dims /*host*/, const double* m, const double* S,
                   double* me, double* See, double* Cxe, void* stream);
int gpp_mm_squash(int N, const double* mf, const double* vf, double scale, double shift, double* mu, double* vu, double* gain,
                  void* stream);
int gpp_mm_squash_nd(int N, int A, const double* mf, const double* Sf, double scale, double shift, double* mu, double* Su, double* gain,
                     void* stream);
/* reverse mode of gpp_mm_squash_nd (upstream differentiates bijectors.py:59-63 through utils/bvn.py with the tape): adjoints of
 * (mu, Su, gain) -> adjoints of (mf, Sf), every entry of Sf treated as an independent variable like the forward reads it; any of the
 * three incoming adjoints may be NULL (= zero).  Closed form: Phi2's partial derivatives are phi * Phi and the bivariate density. */
int gpp_mm_squash_nd_bwd(int N, int A, const double* mf, const double* Sf, double scale, double shift, const double* mu_bar,
                         const double* Su_bar, const double* gain_bar, double* mf_bar, double* Sf_bar, void* stream);
int gpp_cost_gaussian(int N, int De, const double* me, const double* See, const double* target, const double* W, double* out,
                      void* stream);
int gpp_cost_samples(int N, int De, const double* e, const double* target, const double* W, double* out, void* stream);
int gpp_owens_t(int N, const double* h, const double* a, double* out, void* stream);

/* ---- policy weights ---------------------------------------------------------------------------------------
 * beta_r = Kuu_r^-1 m_r for R small SE-ARD kernel regressors (KernelRegressor(SVGP), upstream models/core.py:61-63,
 * moment_matching/models.py:228-235): whitened beta = Luu^-T q_mu, else Kuu^-1 q_mu; Kuu includes `jitter`.
 *   Z [R,Mp,Dp], lengthscales [R,Dp], variance [R], q_mu [R,Mp] -> beta [R,Mp]. */
int gpp_policy_prepare(int R, int Mp, int Dp, const double* Z, const double* lengthscales, const double* variance,
                       const double* q_mu, int whiten, double jitter, double* beta, int* info, void* stream);

/* Adjoint of gpp_policy_prepare: beta_bar [R,Mp] -> q_mu_bar [R,Mp] (written) and the contributions through Kuu to
 * Z_bar [R,Mp,Dp] and lengthscales_bar [R,Dp] (ADDED to what the buffers hold, e.g. the fixed-beta gradients returned by
 * gpp_rollout_mm_bwd / gpp_rollout_pathwise_bwd).  Upstream: TF autodiff through Kuu, its Cholesky and the triangular solves. */
int gpp_policy_prepare_bwd(int R, int Mp, int Dp, const double* Z, const double* lengthscales, const double* variance,
                           const double* beta, const double* beta_bar, int whiten, double jitter,
                           double* Z_bar, double* lengthscales_bar, double* q_mu_bar, void* stream);

/* ---- moment-matched rollout ------------------------------------------------------------------------------
 * replaces the closure body of MomentMatchingPILCO._policy_loss_closure (upstream loops/pilco.py:192-220): H steps of
 * forward_sde (dynamics/forward_sde.py:95-137: TrigonometricEncoder -> InverseLinkWrapper(KernelRegressor) with the
 * Chain[Scale,Shift,NormalCDF] link -> GP dynamics) + MomentMatchingEuler.step (dynamics/solvers.py:110-135, dt = 1)
 * + the GaussianObjective callback (components.py:30-37), for N independent initial states.
 *   dynamics      handle with D = Dx + num_active + 1 inputs and Dx outputs
 *   active_dims   HOST array: state dims encoded as (sin, cos)  (TrigonometricEncoder.active_dims)
 *   policy_*      R = 1 (shared) or R = N parameter sets; action dimension 1; beta from gpp_policy_prepare
 *   cost_target [De], cost_W [De,De] with De = Dx + num_active
 *   m0 [N,Dx], S0 [N,Dx,Dx] -> loss [N]; optional traj_m [H+1,N,Dx], traj_S [H+1,N,Dx,Dx], m_final, S_final.
 * All launches go to `stream` without synchronisation (capturable in a CUDA graph).
 *
 * Execution mode.  By default (GPP_ROLLOUT_AUTO) the H-loop runs ON THE DEVICE: one persistent, warp-specialised cooperative
 * kernel per sweep direction (csrc/rollout_persist.cu) — upstream's loop is a compiled tf.foldl / tf.scan
 * (dynamics/solvers.py:84-105) — whenever the model's Psi2 tiles fit the GPU (<= 2 tiles of 64 x 64 per SM) and, for the backward,
 * the forward kept its per-step block (`saved`).  GPP_ROLLOUT_LEGACY forces one launch per stage and step, GPP_ROLLOUT_PERSIST
 * makes an unsupported shape an error instead of a fallback.  Initial value: environment variable GPP_ROLLOUT_MODE (0/1/2). */
#define GPP_ROLLOUT_AUTO 0
#define GPP_ROLLOUT_LEGACY 1
#define GPP_ROLLOUT_PERSIST 2
int gpp_rollout_mm_set_mode(int mode);
int gpp_rollout_mm_get_mode(void);
size_t gpp_rollout_mm_workspace_bytes(const gpp_gp_model* dynamics, int N, int Dx);
int gpp_rollout_mm_fwd(const gpp_gp_model* dynamics, int N, int Dx, int num_active, const int* active_dims,
                       int R, int Mp, const double* policy_Z, const double* policy_lengthscales,
                       const double* policy_variance, const double* policy_beta, double squash_scale, double squash_shift,
                       const double* cost_target, const double* cost_W, int H, const double* m0, const double* S0,
                       double* loss, double* traj_m, double* traj_S, double* m_final, double* S_final,
                       void* workspace, size_t workspace_bytes, int* info, void* stream);

/* Backward of gpp_rollout_mm_fwd: reverse sweep over the H steps from the trajectory the forward stored (traj_m, traj_S as
 * written by gpp_rollout_mm_fwd).  Upstream: tape.gradient(loss, policy.trainable_variables), utils/optimizers.py:52-56.
 *   loss_bar [N] (NULL = ones)
 *   -> Z_bar [R,Mp,De], lengthscales_bar [R,De]: gradients w.r.t. the policy centres / lengthscales AT FIXED beta,
 *      beta_bar [R,Mp]: gradient w.r.t. beta = Kuu^-1 m (the caller chains beta to (Z, lengthscales, q_mu); the policy
 *      variance is frozen upstream, loops/pilco.py:99-103), m0_bar [N,Dx], S0_bar [N,Dx,Dx] (symmetric; either may be NULL).
 *   R = 1 sums the gradient over the N rollouts in a fixed order; R = N returns one gradient per restart. */
size_t gpp_rollout_mm_bwd_workspace_bytes(const gpp_gp_model* dynamics, int N, int Dx, int Mp, int H);
/* Forward that also keeps, for every step, what the backward would otherwise recompute (joint moments of (e, u), Cov(x, d),
 * pre-inverted cross term): saved [gpp_rollout_mm_saved_doubles(dynamics, N, Dx, H)] doubles; traj_m / traj_S are required.
 * Pass the buffer to gpp_rollout_mm_bwd (`saved`; NULL there = recompute from the trajectory). */
size_t gpp_rollout_mm_saved_doubles(const gpp_gp_model* dynamics, int N, int Dx, int H);
int gpp_rollout_mm_fwd_save(const gpp_gp_model* dynamics, int N, int Dx, int num_active, const int* active_dims,
                            int R, int Mp, const double* policy_Z, const double* policy_lengthscales,
                            const double* policy_variance, const double* policy_beta, double squash_scale, double squash_shift,
                            const double* cost_target, const double* cost_W, int H, const double* m0, const double* S0,
                            double* loss, double* traj_m, double* traj_S, double* m_final, double* S_final, double* saved,
                            void* workspace, size_t workspace_bytes, int* info, void* stream);
int gpp_rollout_mm_bwd(const gpp_gp_model* dynamics, int N, int Dx, int num_active, const int* active_dims,
                       int R, int Mp, const double* policy_Z, const double* policy_lengthscales,
                       const double* policy_variance, const double* policy_beta, double squash_scale, double squash_shift,
                       const double* cost_target, const double* cost_W, int H, const double* traj_m, const double* traj_S,
                       const double* saved, const double* loss_bar, double* Z_bar, double* lengthscales_bar, double* beta_bar,
                       double* m0_bar, double* S0_bar, void* workspace, size_t workspace_bytes, int* info, void* stream);

/* ---- pathwise (sample-path) rollout ----------------------------------------------------------------------
 * replaces the closure body of PathwisePILCO._policy_loss_closure (upstream loops/pilco.py:263-303) for paths drawn by
 * gpflow_sampling's decoupled sampler (random-Fourier prior + canonical-basis update; contract in oracle/pathwise.py):
 *   f_{s,l}(d) = mean_l + amp_l sum_i w[l,i,s] cos(omega_{l,i}.(d/ell_l) + b_{l,i}) + var_l sum_j v[l,j,s] exp(-|d/ell_l - z_{l,j}/ell_l|^2/2)
 * Layouts (particle-minor so that a warp streams 256 contiguous bytes per feature):
 *   basis  [L,F,BS]     BS = (D+2)&~1:  4 omega/(2 pi ell) [D], 4 b/(2 pi)      (gpp_pathwise_pack_basis)
 *   zbasis [L,Mpad,BS]  z/ell [D]; rows >= M are zero and must carry zero weights
 *   w [L,F,ldS], v [L,Mpad,ldS]  with ldS a multiple of gpp_pathwise_particles_per_cta(), F and Mpad multiples of gpp_pathwise_tile()
 *   policy: centres / ell_pi [Mp,De], 1/ell_pi [De], alpha = var_pi Kuu^-1 m [Mp] (deterministic mean, models/core.py:61-71)
 *   x0 [S,Dx] -> loss [S], optional x_final [S,Dx], traj [H+1,S,Dx]. */
int gpp_pathwise_tile(void);
int gpp_pathwise_particles_per_cta(void);
int gpp_pathwise_pack_basis(int L, int F, int M, int Mpad, int D, const double* omega /*[L,F,D]*/, const double* phase /*[L,F]*/,
                            const double* Z /*[L,M,D]*/, const double* lengthscales /*[L,D]*/, double* basis, double* zbasis,
                            void* stream);
int gpp_rollout_pathwise_fwd(int S, int ldS, int H, int L, int F, int Mpad, int D, int Dx, int num_active, const int* active_dims,
                             const double* basis, const double* zbasis, const double* w, const double* v, const double* amp,
                             const double* variance, const double* inv_lengthscales, const double* mean_const,
                             int Mp, const double* policy_Zs, const double* policy_inv_lengthscales, const double* policy_alpha,
                             double squash_scale, double squash_shift, const double* cost_target, const double* cost_W,
                             const double* x0, double* loss, double* x_final, double* traj, void* stream);

/* Gradient mode of gpp_rollout_pathwise_fwd: same rollout, and while the weights stream past it also reduces
 * jac [H, L*D, ldS] = d f_{s,l} / d d_b of every particle-step (the derivative of each particle's function draw w.r.t. its
 * input).  traj [H+1,S,Dx] and jac are required: they are what gpp_rollout_pathwise_bwd reads. */
/* Arithmetic of an entry point.  Everything computes in GPP_F64 (the reference's default_float, upstream
 * gpflow_pilco/moment_matching/models.py:145,216) unless the entry point takes a gpp_dtype; GPP_MIXED_F32_WEIGHTS is the opt-in
 * variant described below (pathwise forward rollout only).  gpp_dtype_supported() lets a binding probe an operation by name. */
typedef enum gpp_dtype { GPP_F64 = 0, GPP_MIXED_F32_WEIGHTS = 1 } gpp_dtype;
int gpp_dtype_supported(const char* entry_point, gpp_dtype dtype);   /* 1 / 0 */
/* Mixed-precision variant of gpp_rollout_pathwise_fwd (north_star permits an FP32 / mixed path with a stated tolerance): the Fourier
 * weights are streamed as FP32 (`w32`, same [L][F][ldS] layout; half the HBM traffic of the rollout) and the cosine polynomial and the
 * weight products run in FP32; phases (FP64 tensor path), quarter-turn reduction, the canonical-basis part, the policy, the cost and
 * the state update stay FP64, FP32 partial sums are folded into FP64 every 32 feature tiles.  Stated tolerance: per-step drift within
 * 5e-6 of the FP64 kernel relative to the largest drift entry (measured 1.5e-6 at F = 1024; the cosine polynomial is good to 3.3e-7);
 * over a rollout the states inherit the dynamics' own error growth (tests/test_gpu_pathwise.py: 10 cart-pole steps stay within 2e-5,
 * the mean loss within 2e-6).  D = 2..7; no gradient mode.
 *   gpp_pathwise_weights_f32 converts `count` weights (round to nearest). */
int gpp_rollout_pathwise_fwd_mixed(int S, int ldS, int H, int L, int F, int Mpad, int D, int Dx, int num_active, const int* active_dims,
                                   const double* basis, const double* zbasis, const float* w32, const double* v, const double* amp,
                                   const double* variance, const double* inv_lengthscales, const double* mean_const,
                                   int Mp, const double* policy_Zs, const double* policy_inv_lengthscales, const double* policy_alpha,
                                   double squash_scale, double squash_shift, const double* cost_target, const double* cost_W,
                                   const double* x0, double* loss, double* x_final, double* traj, void* stream);
int gpp_pathwise_weights_f32(long long count, const double* w, float* w32, void* stream);
/* gpp_rollout_pathwise_fwd / _mixed behind one signature: `w` points to double (GPP_F64) or float (GPP_MIXED_F32_WEIGHTS) weights */
int gpp_rollout_pathwise_fwd_typed(gpp_dtype dtype, int S, int ldS, int H, int L, int F, int Mpad, int D, int Dx, int num_active,
                                   const int* active_dims, const double* basis, const double* zbasis, const void* w, const double* v,
                                   const double* amp, const double* variance, const double* inv_lengthscales, const double* mean_const,
                                   int Mp, const double* policy_Zs, const double* policy_inv_lengthscales, const double* policy_alpha,
                                   double squash_scale, double squash_shift, const double* cost_target, const double* cost_W,
                                   const double* x0, double* loss, double* x_final, double* traj, void* stream);
int gpp_rollout_pathwise_fwd_grad(int S, int ldS, int H, int L, int F, int Mpad, int D, int Dx, int num_active, const int* active_dims,
                                  const double* basis, const double* zbasis, const double* w, const double* v, const double* amp,
                                  const double* variance, const double* inv_lengthscales, const double* mean_const,
                                  int Mp, const double* policy_Zs, const double* policy_inv_lengthscales, const double* policy_alpha,
                                  double squash_scale, double squash_shift, const double* cost_target, const double* cost_W,
                                  const double* x0, double* loss, double* x_final, double* traj, double* jac, void* stream);
/* Backward of the pathwise closure (upstream: tape.gradient through loops/pilco.py:263-298): reverse sweep over traj / jac.
 *   policy in its ORIGINAL parametrisation: Z [Mp,De], lengthscales [De], variance, beta = Kuu^-1 m [Mp];  loss_bar [S] (NULL = ones)
 *   -> Z_bar [Mp,De], lengthscales_bar [De] (at fixed beta), beta_bar [Mp]: sums over the S particles in a fixed order;
 *      x0_bar [S,Dx] (may be NULL). */
size_t gpp_rollout_pathwise_bwd_workspace_bytes(int S, int Mp, int De);
int gpp_rollout_pathwise_bwd(int S, int ldS, int H, int L, int D, int Dx, int num_active, const int* active_dims,
                             int Mp, const double* policy_Z, const double* policy_lengthscales, double policy_variance,
                             const double* policy_beta, double squash_scale, const double* cost_target, const double* cost_W,
                             const double* traj, const double* jac, const double* loss_bar,
                             double* Z_bar, double* lengthscales_bar, double* beta_bar, double* x0_bar,
                             void* workspace, size_t workspace_bytes, void* stream);

/* ---- pathwise draws on the device (counter-based, sharding-invariant) --------------------------------------
 * Random streams: Philox4x32-10 keyed by (seed, stream, logical element index) with the GLOBAL particle index in the
 * element index (oracle/philox.py is the contract; raw words are bit-identical).  Replaces the set-up half of
 * gpflow_sampling's generate_paths (upstream call site loops/pilco.py:281-284) and p.sample (loops/pilco.py:300-303).
 *   gpp_philox_raw          raw 4x32-bit words of counters first_index.. (test hook for bit-exactness)
 *   gpp_pathwise_draw_basis omega [L,F,D] ~ N(0,1), phase [L,F] ~ U(0, 2 pi)
 *   gpp_pathwise_draw_x0    x0[s] = m0 + chol0 n_s for global particles first_particle .. first_particle+S-1
 *   gpp_pathwise_generate   w [L,F,ldS] and v [L,Mpad,ldS] for those particles, from the model handle's q(u). */
int gpp_philox_raw(unsigned long long first_index, int count, unsigned stream_id, unsigned long long seed,
                   unsigned* out /*[count,4]*/, void* stream);
int gpp_pathwise_draw_basis(int L, int F, int D, unsigned long long seed, double* omega, double* phase, void* stream);
int gpp_pathwise_draw_x0(int S, unsigned long long first_particle, int Dx, const double* m0, const double* chol0 /*[Dx,Dx] lower*/,
                         unsigned long long seed, double* x0 /*[S,Dx]*/, void* stream);
size_t gpp_pathwise_generate_workspace_bytes(const gpp_gp_model* model, int ldS, int F);
int gpp_pathwise_generate(gpp_gp_model* model, int S, int ldS, unsigned long long first_particle, int F, int Mpad,
                          unsigned long long seed, const double* omega, const double* phase, double* w, double* v,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- measurement hooks (bench.py) -------------------------------------------------------------------------
 * gpp_profile_enable(1): every entry point records CUDA events on its stream around its dominant kernel
 * (k_contract for gpp_mm_gp_predict_fwd, k_ekzxkxz for gpp_ekzxkxz, the rollout kernels for the rollouts).
 * gpp_profile_last_ms synchronises on the last recorded pair and returns its duration.
 * gpp_microbench_fp64 launches `blocks` x `threads` threads, each running 8 independent DFMA chains of `iters`
 * steps (2*8*iters flop per thread): the measured FP64-pipe roofline denominator. */
int gpp_profile_enable(int on);
int gpp_profile_last_ms(float* ms);
int gpp_microbench_fp64(int blocks, int threads, int iters, double* sink /*device [blocks*threads]*/, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GPP_B200_H */
