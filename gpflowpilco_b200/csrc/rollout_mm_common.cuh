// Shared pieces of the moment-matched rollout (forward: rollout_mm.cu, backward: rollout_mm_bwd.cu): the per-step
// parameter block and the "pre" stage of a step — encoder rule, policy moment matching, squashing link, joint moments of
// d = (e, u) and Sxd = Cov(x, d)  (upstream gpflow_pilco/dynamics/forward_sde.py:95-124 and the rules it dispatches to).
#pragma once
#include "mm_small.cuh"
#include "model.cuh"

namespace gpp {

struct RolloutMMParams {
  EncoderSpec enc;
  int N, Dx, De, D, L;       // D = De + 1 (scalar action), L = Dx outputs of the dynamics
  int R, Mp;                 // policy sets (1 = shared, or N) and centres per policy
  const double *pZ, *pEll, *pVar, *pBeta;   // [R,Mp,De], [R,De], [R], [R,Mp]
  double scale, shift;
  const double *target, *W;  // [De], [De,De]
  double *m, *S, *loss;      // [N,Dx], [N,Dx,Dx], [N]   current state / accumulated loss
  double *md, *Sd, *Sxd;     // [N,D], [N,D,D], [N,Dx,D]
  double *f1, *Sff, *cross;  // [N,L], [N,L,L], [N,D,L]
  double *traj_m, *traj_S;   // optional [H+1,N,Dx], [H+1,N,Dx,Dx]
  double* pre;               // optional [N, sizeof(PreShared)/8]: the pre stage's shared block, saved by the forward for the backward
  int* info;
};

// everything the pre stage computes for one rollout, kept in shared memory (the backward reuses all of it)
template <int DP>
struct PreShared {
  double me[GPP_SMALL_MAX], See[GPP_SMALL_MAX * GPP_SMALL_MAX], Cxe[GPP_SMALL_MAX * GPP_SMALL_MAX];
  double pack[PairPack<DP>::SIZE];   // (policy kernel, policy kernel) Psi2 coefficients at (me, See)
  double Li1[DP * DP];               // inverse Cholesky factor of See + Lambda  (Psi1)
  double c01;                        // log(var) + sum log ell - log det chol(See + Lambda)
  double red[4][DP + 2];
  double f1, f2, vf, mu_u, vu, gain; // policy mean / second moment / variance, squashed action moments, link gain
  double t0;                         // Owen's T(h, a) of the squashing rule (warp-evaluated; the backward reuses it)
  double vec[DP], cpre[DP], seu[DP]; // sum beta psi1 (z - me), (See+Lambda)^-1 vec, Cov(e, u)
};

// all 128 threads of the group call this: encoder rule on the current state (p.m, p.S) of rollout n -> sh.me, sh.See, sh.Cxe
// (one output entry per thread: the serial form costs ~8 us of single-thread latency per step).  Ends with a group barrier.
template <int DP>
__device__ void step_pre_encode(const RolloutMMParams& p, int n, PreShared<DP>& sh) {
  const int tid = threadIdx.x;
  const int Dx = p.Dx;
  __shared__ double xm[GPP_SMALL_MAX], xS[GPP_SMALL_MAX * GPP_SMALL_MAX];
  __shared__ EncTrig<double> trig;
  for (int t = tid; t < Dx + Dx * Dx; t += kGroupThreads) {
    if (t < Dx) xm[t] = p.m[(size_t)n * Dx + t];
    else xS[t - Dx] = p.S[(size_t)n * Dx * Dx + (t - Dx)];
  }
  group_sync();
  auto mean_at = [&](int i) { return xm[i]; };
  auto cov_at = [&](int i, int j) { return xS[i * Dx + j]; };
  if (tid < p.enc.na) enc_trig_one<double>(p.enc, tid, mean_at, cov_at, trig);
  group_sync();
  {
    const int De = DP;
    for (int t = tid; t < De + De * De + Dx * De; t += kGroupThreads) {
      if (t < De) sh.me[t] = enc_mean_at<double>(p.enc, t, mean_at, trig);
      else if (t < De + De * De) sh.See[t - De] = enc_cov_at<double>(p.enc, (t - De) / De, (t - De) % De, cov_at, trig);
      else sh.Cxe[t - De - De * De] = enc_cross_at<double>(p.enc, (t - De - De * De) / De, (t - De - De * De) % De, cov_at, trig);
    }
  }
  group_sync();
}

// all 128 threads of the group call this; on return (after a group barrier) `sh` is complete.  With `cost_out` the expected cost of
// the current state (upstream components.py:30-37 on the encoded moments, loops/pilco.py:199-205) is evaluated by an otherwise idle
// warp next to the two factorisations.
template <int DP>
__device__ void step_pre_forward(const RolloutMMParams& p, int n, PreShared<DP>& sh, double* cost_out = nullptr) {
  using PP = PairPack<DP>;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = (p.R == 1) ? 0 : n;
  const double* ell = p.pEll + (size_t)r * DP;
  const double var = p.pVar[r];
  step_pre_encode<DP>(p, n, sh);
  // two independent factorisations of the encoded covariance, one warp each (lane 0)
  if (tid == 0) {
    // coefficient pack of the (policy kernel, policy kernel) pair
    double V[DP], mu[DP], Sg[DP * DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      V[d] = ell[d] * ell[d];
      mu[d] = sh.me[d];
#pragma unroll
      for (int e = 0; e < DP; ++e) Sg[d * DP + e] = sh.See[d * DP + e];
    }
    if (!make_pair_pack<DP>(mu, Sg, V, V, 2.0 * log(var), sh.pack)) flag_not_pd(p.info, n);
  } else if (tid == 32) {
    // Psi1 factorisation: inverse Cholesky factor of See + Lambda and the log normaliser
    Mat<DP> A, Li;
    double ell_prod = 1.0;                   // one log of a ratio of products instead of 2 DP logs
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      ell_prod *= ell[d];
#pragma unroll
      for (int e = 0; e < DP; ++e) A(d, e) = sh.See[d * DP + e] + (d == e ? ell[d] * ell[d] : 0.0);
    }
    if (!cholesky<DP>(A)) flag_not_pd(p.info, n);
    double diag_prod = 1.0;
#pragma unroll
    for (int d = 0; d < DP; ++d) diag_prod *= A(d, d);
    tri_inverse<DP>(A, Li);
#pragma unroll
    for (int d = 0; d < DP * DP; ++d) sh.Li1[d] = Li.a[d];
    sh.c01 = log(var * ell_prod / diag_prod);
  } else if (tid == 64 && cost_out) {
    *cost_out = expected_cost<double>(DP, sh.me, sh.See, p.target, p.W);
  }
  group_sync();
  // ---- policy Psi1 terms: f1 = sum_i beta_i psi1_i, vec = sum_i beta_i psi1_i (z_i - me)
  const double* Zp = p.pZ + (size_t)r * p.Mp * DP;
  const double* beta = p.pBeta + (size_t)r * p.Mp;
  double acc = 0.0, vec[DP];
#pragma unroll
  for (int d = 0; d < DP; ++d) vec[d] = 0.0;
  for (int i = tid; i < p.Mp; i += kGroupThreads) {
    double dz[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) dz[d] = Zp[i * DP + d] - sh.me[d];
    double maha = 0.0;
#pragma unroll
    for (int a = 0; a < DP; ++a) {
      double y = 0.0;
#pragma unroll
      for (int k = 0; k <= a; ++k) y = fma(sh.Li1[a * DP + k], dz[k], y);
      maha = fma(y, y, maha);
    }
    double w = beta[i] * fast_exp(sh.c01 - 0.5 * maha);
    acc += w;
#pragma unroll
    for (int d = 0; d < DP; ++d) vec[d] = fma(w, dz[d], vec[d]);
  }
  // ---- policy Psi2 contraction: f2 = sum_ij beta_i beta_j Q_ij  (KernelRegressor: no model uncertainty, models.py:34-41)
  //      thread = (row i, quarter c of the columns): 4 threads share a row so that all 128 threads work at Mp = 30
  double f2 = 0.0;
  for (int idx = tid; idx < 4 * p.Mp; idx += kGroupThreads) {
    const int i = idx >> 2, c = idx & 3;
    double zr[DP], g[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) zr[d] = Zp[i * DP + d] - sh.pack[PP::MU + d];
#pragma unroll
    for (int e = 0; e < DP; ++e) {
      double t = 0.0;
#pragma unroll
      for (int d = 0; d < DP; ++d) t = fma(zr[d], sh.pack[PP::R + d * DP + e], t);
      g[e] = t;
    }
    double ri = sh.pack[PP::C0] + packed_quad<DP>(sh.pack + PP::P1, zr);
    double row = 0.0;
    for (int j = c; j < p.Mp; j += 4) {
      double zc[DP];
#pragma unroll
      for (int d = 0; d < DP; ++d) zc[d] = Zp[j * DP + d] - sh.pack[PP::MU + d];
      double t = ri + packed_quad<DP>(sh.pack + PP::P2, zc);
#pragma unroll
      for (int d = 0; d < DP; ++d) t = fma(g[d], zc[d], t);
      row = fma(beta[j], fast_exp(t), row);
    }
    f2 = fma(beta[i], row, f2);
  }
  acc = warp_sum(acc);
  f2 = warp_sum(f2);
#pragma unroll
  for (int d = 0; d < DP; ++d) vec[d] = warp_sum(vec[d]);
  if (lane == 0) {
    sh.red[warp][0] = acc;
    sh.red[warp][1] = f2;
#pragma unroll
    for (int d = 0; d < DP; ++d) sh.red[warp][2 + d] = vec[d];
  }
  group_sync();
  if (warp == 0) {
    double f1 = 0.0, f2s = 0.0;
    for (int w = 0; w < 4; ++w) {
      f1 += sh.red[w][0];
      f2s += sh.red[w][1];
    }
    const double vf = f2s - f1 * f1;
    double h, a;
    squash_owens_args(f1, vf, h, a);
    const double t0 = owens_t_warp(h, a);      // the 32-node quadrature of E[Phi^2], one node per lane
    if (lane == 0) {
      double v[DP];
#pragma unroll
      for (int d = 0; d < DP; ++d) v[d] = 0.0;
      for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int d = 0; d < DP; ++d) v[d] += sh.red[w][2 + d];
      // pre-inverted cross term of the regressor: (See + Lambda)^-1 vec  with (See+Lambda)^-1 = Li^T Li
      double y[DP];
#pragma unroll
      for (int a2 = 0; a2 < DP; ++a2) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k <= a2; ++k) t = fma(sh.Li1[a2 * DP + k], v[k], t);
        y[a2] = t;
      }
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        double t = 0.0;
#pragma unroll
        for (int a2 = d; a2 < DP; ++a2) t = fma(sh.Li1[a2 * DP + d], y[a2], t);
        sh.cpre[d] = t;
        sh.vec[d] = v[d];
      }
      sh.f1 = f1;
      sh.f2 = f2s;
      sh.vf = vf;
      sh.t0 = t0;
      double mu_u, vu, gain;
      mm_squash_1d<double, true>(f1, vf, p.scale, p.shift, mu_u, vu, gain, t0);
      sh.mu_u = mu_u; sh.vu = vu; sh.gain = gain;
      // joint moments of d = (e, u)  (gaussian.py:53-63): Seu = See cpre gain
#pragma unroll
      for (int a2 = 0; a2 < DP; ++a2) {
        double t = 0.0;
#pragma unroll
        for (int b = 0; b < DP; ++b) t = fma(sh.See[a2 * DP + b], sh.cpre[b], t);
        sh.seu[a2] = t * gain;
      }
    }
  }
  group_sync();
}

// all threads of the CTA: write md, Sd and Sxd of rollout n (one entry per thread)
template <int DP>
__device__ void step_pre_write(const RolloutMMParams& p, int n, const PreShared<DP>& sh) {
  const int Dx = p.Dx, De = p.De, D = p.D;
  double* md = p.md + (size_t)n * D;
  double* Sd = p.Sd + (size_t)n * D * D;
  double* Sxd = p.Sxd + (size_t)n * Dx * D;
  const int na = p.enc.na, nb = p.enc.nb();
  // joint covariance of d = (e, u): [[See, seu], [seu^T, vu]]
  auto sd_entry = [&](int a, int b) {
    if (a < De && b < De) return sh.See[a * De + b];
    if (a < De) return sh.seu[a];
    if (b < De) return sh.seu[b];
    return sh.vu;
  };
  for (int t = threadIdx.x; t < D + D * D + Dx * D; t += kGroupThreads) {
    if (t < D) {
      md[t] = t < De ? sh.me[t] : sh.mu_u;
    } else if (t < D + D * D) {
      const int a = (t - D) / D, b = (t - D) % D;
      Sd[a * D + b] = sd_entry(a, b);
    } else {
      // Sxd = Cov(x, d): active rows through the encoder linearisation, inactive rows copied from S_d (forward_sde.py:112-124)
      const int i = (t - D - D * D) / D, b = (t - D - D * D) % D;
      bool active = false;
      for (int k = 0; k < na; ++k) active = active || p.enc.active[k] == i;
      double v;
      if (active) {
        if (b < De) {
          v = sh.Cxe[i * De + b];
        } else {
          double sau = 0.0;
          for (int c = 0; c < De; ++c) sau = fma(sh.Cxe[i * De + c], sh.cpre[c], sau);
          v = sau * sh.gain;
        }
      } else {
        int j = 0;
        for (int q = 0; q < nb; ++q) j = p.enc.inactive(q) == i ? q : j;
        v = sd_entry(2 * na + j, b);
      }
      Sxd[i * D + b] = v;
    }
  }
}

template <int DP>
struct PreSharedSize { static constexpr int value = (int)(sizeof(PreShared<DP>) / sizeof(double)); };

inline size_t pre_shared_doubles(int De) {
  switch (De) {
#define GPP_CASE(d) case d: return PreSharedSize<d>::value;
    GPP_CASE(1) GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7)
#undef GPP_CASE
    default: return 0;
  }
}

// k_step_pre: one CTA (128 threads) per rollout
template <int DP>
__global__ void __launch_bounds__(128) k_step_pre(RolloutMMParams p) {
  __shared__ PreShared<DP> sh;
  const int n = blockIdx.x;
  step_pre_forward<DP>(p, n, sh);
  step_pre_write<DP>(p, n, sh);
  if (p.pre) {   // forward-with-save: the backward's k_bwd_pre reloads this block instead of recomputing the stage
    constexpr int PS = PreSharedSize<DP>::value;
    const double* src = reinterpret_cast<const double*>(&sh);
    double* dst = p.pre + (size_t)n * PS;
    for (int t = threadIdx.x; t < PS; t += kGroupThreads) dst[t] = src[t];
  }
}

inline size_t rollout_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// per-step block of the optional `saved` buffer of gpp_rollout_mm_fwd_save (offsets in doubles):
//   md [N,D] | Sd [N,D,D] | Sxd [N,Dx,D] | cross [N,D,L] | pre [N, PreShared<D-1>]
struct RolloutSaved {
  size_t md, Sd, Sxd, cross, pre, per_step;
  RolloutSaved(int N, int Dx, int D, int L) {
    md = 0;
    Sd = md + (size_t)N * D;
    Sxd = Sd + (size_t)N * D * D;
    cross = Sxd + (size_t)N * Dx * D;
    pre = cross + (size_t)N * D * L;
    per_step = pre + (size_t)N * pre_shared_doubles(D - 1);
  }
};

}  // namespace gpp
