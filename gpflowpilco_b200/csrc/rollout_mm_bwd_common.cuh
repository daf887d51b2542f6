// Device code of the reverse sweep of the moment-matched rollout (rollout_mm_bwd.cu: one launch per stage and step;
// rollout_persist.cu: the same stages on the scalar warps of the persistent kernel).
#pragma once
#include "dual.cuh"
#include "rollout_mm_common.cuh"

namespace gpp {

struct RolloutBwdBuffers {
  double *mb, *Sb;                                  // [N,Dx], [N,Dx,Dx]  running state adjoint
  double *f1_bar, *Sff_bar, *cross_bar, *Sxd_bar;   // [N,L], [N,L,L], [N,D,L], [N,Dx,D]
  double *md_bar, *Sd_bar;                          // [N,D], [N,D,D]
  double *gZ, *gEll, *gBeta;                        // per-rollout parameter gradients [N,Mp,De], [N,De], [N,Mp]
};

// direction index k -> state-moment entry: k < Dx is m[k]; otherwise the symmetric pair (i <= j) of S
__device__ __forceinline__ void direction_to_entry(int k, int Dx, int& i, int& j) {
  int rem = k - Dx;
  i = 0;
  while (rem >= Dx - i) { rem -= Dx - i; ++i; }
  j = i + rem;
}

constexpr int kGradRing = 32;   // steps whose cost gradients are evaluated by one k_cost_grad_ring launch

// d cost / d(m, S) of `count` consecutive trajectory states x N rollouts x ndir directions, one thread each: forward-mode dual
// numbers through the same templated encoder / expected-cost code as the forward pass.  The states are known from the stored
// trajectory, so all of a ring's gradients are computed in one parallel launch instead of serially inside the sweep.
static __global__ void k_cost_grad_ring(RolloutMMParams p, const double* __restrict__ tm, const double* __restrict__ tS, int count, int ndir,
                                 double* __restrict__ cg /*[count,N,ndir]*/) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count * p.N * ndir) return;
  const int k = idx % ndir;
  const size_t sn = idx / ndir;                 // (slot, rollout) flattened: trajectory states are [t, n] contiguous
  const int Dx = p.Dx;
  Dual m[GPP_SMALL_MAX], S[GPP_SMALL_MAX * GPP_SMALL_MAX];
  for (int i = 0; i < Dx; ++i) m[i] = Dual(tm[sn * Dx + i]);
  for (int i = 0; i < Dx * Dx; ++i) S[i] = Dual(tS[sn * Dx * Dx + i]);
  if (k < Dx) {
    m[k].d = 1.0;
  } else {
    int di, dj;
    direction_to_entry(k, Dx, di, dj);
    S[di * Dx + dj].d = 1.0;
    S[dj * Dx + di].d = 1.0;
  }
  Dual me[GPP_SMALL_MAX], See[GPP_SMALL_MAX * GPP_SMALL_MAX], Cxe[GPP_SMALL_MAX * GPP_SMALL_MAX];
  mm_encoder<Dual>(p.enc, m, S, me, See, Cxe);
  cg[idx] = expected_cost<Dual>(p.De, me, See, p.target, p.W).d;
}

// one warp per rollout (warp 0 of the group in the persistent sweep)
__device__ __forceinline__ void bwd_post_body(const RolloutMMParams& p, const int n, const double* __restrict__ cg /*[N,ndir] of state t+1*/,
                                              const double* __restrict__ loss_bar, const RolloutBwdBuffers& bw) {
  __shared__ double smb[GPP_SMALL_MAX], sSb[GPP_SMALL_MAX * GPP_SMALL_MAX];
  const int lane = threadIdx.x;
  const int Dx = p.Dx, D = p.D, L = p.L;
  for (int i = lane; i < Dx; i += 32) smb[i] = bw.mb[(size_t)n * Dx + i];
  for (int i = lane; i < Dx * Dx; i += 32) sSb[i] = bw.Sb[(size_t)n * Dx * Dx + i];
  __syncwarp();
  const double lb = loss_bar ? loss_bar[n] : 1.0;
  const int ndir = Dx + Dx * (Dx + 1) / 2;
  for (int k = lane; k < ndir; k += 32) {
    const double g = lb * cg[(size_t)n * ndir + k];
    if (k < Dx) {
      smb[k] += g;
    } else {
      int di, dj;
      direction_to_entry(k, Dx, di, dj);
      if (di == dj) {
        sSb[di * Dx + di] += g;
      } else {
        sSb[di * Dx + dj] += 0.5 * g;
        sSb[dj * Dx + di] += 0.5 * g;
      }
    }
  }
  __syncwarp();
  // Euler update m' = m + f1, S' = S + Sxf + Sxf^T + Sff with Sxf = Sxd cross  (dt = 1)
  const double* cross = p.cross + (size_t)n * D * L;
  const double* Sxd = p.Sxd + (size_t)n * Dx * D;
  for (int l = lane; l < L; l += 32) bw.f1_bar[(size_t)n * L + l] = smb[l];
  for (int t = lane; t < L * L; t += 32) bw.Sff_bar[(size_t)n * L * L + t] = sSb[t];
  for (int t = lane; t < Dx * D; t += 32) {
    const int i = t / D, b = t % D;
    double v = 0.0;
    for (int l = 0; l < L; ++l) v = fma(sSb[i * Dx + l] + sSb[l * Dx + i], cross[b * L + l], v);
    bw.Sxd_bar[(size_t)n * Dx * D + t] = v;
  }
  for (int t = lane; t < D * L; t += 32) {
    const int b = t / L, l = t % L;
    double v = 0.0;
    for (int i = 0; i < Dx; ++i) v = fma(Sxd[i * D + b], sSb[i * Dx + l] + sSb[l * Dx + i], v);
    bw.cross_bar[(size_t)n * D * L + t] = v;
  }
  for (int i = lane; i < Dx; i += 32) bw.mb[(size_t)n * Dx + i] = smb[i];
  for (int i = lane; i < Dx * Dx; i += 32) bw.Sb[(size_t)n * Dx * Dx + i] = sSb[i];
}

template <int DP>
struct PreAdjoint {
  double me_bar[DP], See_bar[DP * DP], Cxe_bar[GPP_SMALL_MAX * DP];
  double f1b, f2b, cpre_bar[DP], y[DP];
  double ubar[3], jac[2][3];   // (mu_u_bar, vu_bar, gain_bar); d(mu_u, vu, gain)/d f1 and /d vf of the squashing link
  double G1[DP * DP], G2[DP * DP];
  double red[4][2 * DP + DP * DP];
};

// all 128 threads of the group, for rollout n
template <int DP>
__device__ void bwd_pre_body(const RolloutMMParams& p, const RolloutBwdBuffers& bw, const int n) {
  using PP = PairPack<DP>;
  constexpr int D = DP + 1;
  constexpr int K = 2 * DP + DP * DP;      // block-reduced accumulators: mu [DP], ell [DP], Sigma [DP*DP]
  __shared__ PreShared<DP> sh;
  __shared__ PreAdjoint<DP> ad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Dx = p.Dx;
  const int r = (p.R == 1) ? 0 : n;
  if (p.pre) {   // the forward saved the stage's shared block
    constexpr int PS = PreSharedSize<DP>::value;
    const double* src = p.pre + (size_t)n * PS;
    double* dst = reinterpret_cast<double*>(&sh);
    for (int t = tid; t < PS; t += kGroupThreads) dst[t] = src[t];
    group_sync();
  } else {
    step_pre_forward<DP>(p, n, sh);
  }

  // Four independent serial pieces, one per warp (lane 0): the adjoint of the joint assembly, the two columns of the squashing
  // link's 3 x 2 Jacobian (dual numbers), and the two Gram matrices of the policy adjoint.
  if (tid == 0) {
    // ---- adjoint of the joint assembly: (md, Sd, Sxd) -> (me, See, Cxe, cpre, mu_u, vu, gain)
    const double* md_bar = bw.md_bar + (size_t)n * D;
    const double* Sxd_bar = bw.Sxd_bar + (size_t)n * Dx * D;
    double Sdb[D * D];
    for (int t = 0; t < D * D; ++t) Sdb[t] = bw.Sd_bar[(size_t)n * D * D + t];
    const int na = p.enc.na, nb = p.enc.nb();
    for (int j = 0; j < nb; ++j) {
      const int i = p.enc.inactive(j);
      for (int b = 0; b < D; ++b) Sdb[(2 * na + j) * D + b] += Sxd_bar[i * D + b];
    }
    double seu_bar[DP], gain_bar = 0.0;
    for (int a = 0; a < DP; ++a) {
      ad.me_bar[a] = md_bar[a];
      ad.cpre_bar[a] = 0.0;
      seu_bar[a] = Sdb[a * D + DP] + Sdb[DP * D + a];
      for (int b = 0; b < DP; ++b) ad.See_bar[a * DP + b] = Sdb[a * D + b];
    }
    for (int t = 0; t < Dx * DP; ++t) ad.Cxe_bar[t] = 0.0;
    for (int k = 0; k < na; ++k) {
      const int i = p.enc.active[k];
      double ri = 0.0;
      for (int b = 0; b < DP; ++b) {
        ad.Cxe_bar[i * DP + b] = Sxd_bar[i * D + b];
        ri = fma(sh.Cxe[i * DP + b], sh.cpre[b], ri);
      }
      const double s = Sxd_bar[i * D + DP];
      gain_bar = fma(s, ri, gain_bar);
      for (int b = 0; b < DP; ++b) {
        ad.Cxe_bar[i * DP + b] = fma(s * sh.gain, sh.cpre[b], ad.Cxe_bar[i * DP + b]);
        ad.cpre_bar[b] = fma(s * sh.gain, sh.Cxe[i * DP + b], ad.cpre_bar[b]);
      }
    }
    for (int a = 0; a < DP; ++a) {
      double qa = 0.0;
      for (int b = 0; b < DP; ++b) qa = fma(sh.See[a * DP + b], sh.cpre[b], qa);
      gain_bar = fma(seu_bar[a], qa, gain_bar);
      const double qb = seu_bar[a] * sh.gain;
      for (int b = 0; b < DP; ++b) {
        ad.See_bar[a * DP + b] = fma(qb, sh.cpre[b], ad.See_bar[a * DP + b]);
        ad.cpre_bar[b] = fma(qb, sh.See[a * DP + b], ad.cpre_bar[b]);
      }
    }
    ad.ubar[0] = md_bar[DP];
    ad.ubar[1] = Sdb[DP * D + DP];
    ad.ubar[2] = gain_bar;
  } else if (tid == 32 || tid == 64) {
    // ---- squashing link: one column of the Jacobian each (Owen's T: value from the forward, closed-form partials)
    const int c = tid == 32 ? 0 : 1;
    Dual mu, vu, gn;
    mm_squash_1d<Dual, true>(Dual(sh.f1, c == 0 ? 1.0 : 0.0), Dual(sh.vf, c == 0 ? 0.0 : 1.0), p.scale, p.shift, mu, vu, gn, sh.t0);
    ad.jac[c][0] = mu.d;
    ad.jac[c][1] = vu.d;
    ad.jac[c][2] = gn.d;
  } else if (tid == 96) {
    // ---- G1 = (See + Lambda)^-1, G2 = (See + Lambda/2)^-1
    Mat<DP> Li, G, A2;
    for (int t = 0; t < DP * DP; ++t) Li.a[t] = sh.Li1[t];
    gram_inverse<DP>(Li, G);
    const double* ell = p.pEll + (size_t)r * DP;
    for (int a = 0; a < DP; ++a)
      for (int b = 0; b < DP; ++b) {
        ad.G1[a * DP + b] = G(a, b);
        A2(a, b) = sh.See[a * DP + b] + (a == b ? 0.5 * ell[a] * ell[a] : 0.0);
      }
    cholesky<DP>(A2);
    tri_inverse<DP>(A2, Li);
    gram_inverse<DP>(Li, G);
    for (int t = 0; t < DP * DP; ++t) ad.G2[t] = G.a[t];
  }
  group_sync();
  if (tid < DP) {                            // y = G1 cpre_bar
    double t = 0.0;
    for (int b = 0; b < DP; ++b) t = fma(ad.G1[tid * DP + b], ad.cpre_bar[b], t);
    ad.y[tid] = t;
  } else if (tid == 32) {
    const double f1b = ad.ubar[0] * ad.jac[0][0] + ad.ubar[1] * ad.jac[0][1] + ad.ubar[2] * ad.jac[0][2];
    const double vfb = ad.ubar[0] * ad.jac[1][0] + ad.ubar[1] * ad.jac[1][1] + ad.ubar[2] * ad.jac[1][2];
    ad.f2b = vfb;                          // vf = f2 - f1^2
    ad.f1b = f1b - 2.0 * sh.f1 * vfb;
  }
  group_sync();

  // ---- closed-form adjoint of the policy's Psi1 / Psi2 sums (rows over threads)
  const double* Zp = p.pZ + (size_t)r * p.Mp * DP;
  const double* beta = p.pBeta + (size_t)r * p.Mp;
  const double* ell = p.pEll + (size_t)r * DP;
  double acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.0;
  double* mu_acc = acc;
  double* ell_acc = acc + DP;
  double* sig_acc = acc + 2 * DP;
  const double f1b = ad.f1b, f2b = ad.f2b;
  // thread = (row i, quarter c of the columns); the 4 threads of a row are neighbouring lanes and combine their row sums by shuffle
  const int nwork = (4 * p.Mp + 31) / 32 * 32;     // whole warps iterate together (shuffles below)
  for (int idx = tid; idx < nwork; idx += kGroupThreads) {
    const bool live = idx < 4 * p.Mp;
    const int i = live ? idx >> 2 : 0, c = idx & 3;
    double zr[DP], zbar[DP], h[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) { zr[d] = Zp[i * DP + d] - sh.me[d]; zbar[d] = 0.0; }
    double beta_bar = 0.0;
    if (live && c == 0) {
      // Psi1
      double maha = 0.0, e = f1b;
#pragma unroll
      for (int a = 0; a < DP; ++a) {
        double t = 0.0;
#pragma unroll
        for (int b = 0; b < DP; ++b) t = fma(ad.G1[a * DP + b], zr[b], t);
        h[a] = t;
        maha = fma(t, zr[a], maha);
        e = fma(ad.y[a], zr[a], e);
      }
      const double psi = fast_exp(sh.c01 - 0.5 * maha);
      const double w = beta[i] * psi;
      beta_bar = psi * e;
#pragma unroll
      for (int a = 0; a < DP; ++a) {
        const double t1 = w * (e * h[a] - ad.y[a]);
        mu_acc[a] += t1;
        zbar[a] -= t1;
        double sdd = 0.0;
#pragma unroll
        for (int b = 0; b < DP; ++b) {
          const double s = w * (0.5 * e * (h[a] * h[b] - ad.G1[a * DP + b]) - 0.5 * (ad.y[a] * h[b] + h[a] * ad.y[b]));
          sig_acc[a * DP + b] += s;
          if (a == b) sdd = s;
        }
        ell_acc[a] += 2.0 * ell[a] * sdd + w * e / ell[a];
      }
    }
    // Psi2 (same kernel, same centres: V = Lambda/2, e_ij = (z_i + z_j)/2 - me)
    double g0[DP];
#pragma unroll
    for (int e2 = 0; e2 < DP; ++e2) {
      double t = 0.0;
#pragma unroll
      for (int d = 0; d < DP; ++d) t = fma(zr[d], sh.pack[PP::R + d * DP + e2], t);
      g0[e2] = t;
    }
    const double ri = sh.pack[PP::C0] + packed_quad<DP>(sh.pack + PP::P1, zr);
    double rowQ = 0.0;
    for (int j = c; live && j < p.Mp; j += 4) {
      double zc[DP], ee[DP], g[DP];
#pragma unroll
      for (int d = 0; d < DP; ++d) { zc[d] = Zp[j * DP + d] - sh.me[d]; ee[d] = 0.5 * (zr[d] + zc[d]); }
      double t = ri + packed_quad<DP>(sh.pack + PP::P2, zc);
#pragma unroll
      for (int d = 0; d < DP; ++d) t = fma(g0[d], zc[d], t);
      const double Q = fast_exp(t);
      rowQ = fma(beta[j], Q, rowQ);
      const double a = f2b * beta[i] * beta[j] * Q;
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < DP; ++b) s = fma(ad.G2[d * DP + b], ee[b], s);
        g[d] = s;
      }
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        const double dl = zr[d] - zc[d], il = 1.0 / ell[d];
        mu_acc[d] = fma(a, g[d], mu_acc[d]);
        zbar[d] = fma(a, -dl * il * il - g[d], zbar[d]);
        ell_acc[d] = fma(a, il - 0.5 * ad.G2[d * DP + d] * ell[d] + 0.5 * dl * dl * il * il * il + 0.5 * g[d] * g[d] * ell[d], ell_acc[d]);
#pragma unroll
        for (int b = 0; b < DP; ++b) sig_acc[d * DP + b] = fma(a, 0.5 * (g[d] * g[b] - ad.G2[d * DP + b]), sig_acc[d * DP + b]);
      }
    }
    beta_bar = fma(2.0 * f2b, rowQ, beta_bar);
    // combine the 4 quarters of the row (fixed order) and let quarter 0 write
    beta_bar += __shfl_xor_sync(0xffffffffu, beta_bar, 1);
    beta_bar += __shfl_xor_sync(0xffffffffu, beta_bar, 2);
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      zbar[d] += __shfl_xor_sync(0xffffffffu, zbar[d], 1);
      zbar[d] += __shfl_xor_sync(0xffffffffu, zbar[d], 2);
    }
    if (live && c == 0) {
      bw.gBeta[(size_t)n * p.Mp + i] += beta_bar;
#pragma unroll
      for (int d = 0; d < DP; ++d) bw.gZ[((size_t)n * p.Mp + i) * DP + d] += zbar[d];
    }
  }
  // fixed-order block reduction of the K accumulators
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double v = warp_sum(acc[k]);
    if (lane == 0) ad.red[warp][k] = v;
  }
  group_sync();
  if (tid < K) {
    const double s = (ad.red[0][tid] + ad.red[1][tid]) + (ad.red[2][tid] + ad.red[3][tid]);
    if (tid < DP) ad.me_bar[tid] += s;
    else if (tid < 2 * DP) bw.gEll[(size_t)n * DP + (tid - DP)] += s;
    else ad.See_bar[tid - 2 * DP] += s;
  }
  group_sync();

  // ---- encoder adjoint (closed-form reverse mode shared by the CTA, mm_small.cuh; the state's symmetric covariance receives the
  //      symmetrised adjoint)
  __shared__ double xm[GPP_SMALL_MAX], xS[GPP_SMALL_MAX * GPP_SMALL_MAX], xmb[GPP_SMALL_MAX], xSb[GPP_SMALL_MAX * GPP_SMALL_MAX];
  __shared__ double enc_scratch[32];
  for (int t = tid; t < Dx + Dx * Dx; t += kGroupThreads) {
    if (t < Dx) xm[t] = p.m[(size_t)n * Dx + t];
    else xS[t - Dx] = p.S[(size_t)n * Dx * Dx + (t - Dx)];
  }
  group_sync();
  mm_encoder_bwd_cta(p.enc, xm, xS, ad.me_bar, ad.See_bar, ad.Cxe_bar, xmb, xSb, enc_scratch);
  for (int t = tid; t < Dx + Dx * Dx; t += kGroupThreads) {
    if (t < Dx) {
      bw.mb[(size_t)n * Dx + t] += xmb[t];
    } else {
      const int i = (t - Dx) / Dx, j = (t - Dx) % Dx;
      bw.Sb[(size_t)n * Dx * Dx + i * Dx + j] += 0.5 * (xSb[i * Dx + j] + xSb[j * Dx + i]);
    }
  }
}

static __global__ void __launch_bounds__(32) k_bwd_post(RolloutMMParams p, const double* __restrict__ cg, const double* __restrict__ loss_bar,
                                                 RolloutBwdBuffers bw) {
  bwd_post_body(p, (int)blockIdx.x, cg, loss_bar, bw);
}

template <int DP>
__global__ void __launch_bounds__(128) k_bwd_pre(RolloutMMParams p, RolloutBwdBuffers bw) {
  bwd_pre_body<DP>(p, bw, (int)blockIdx.x);
}

}  // namespace gpp
