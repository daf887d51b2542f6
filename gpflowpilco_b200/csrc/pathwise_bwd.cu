// Backward of the pathwise particle rollout: reverse sweep over the stored particle trajectories.
//
// Upstream differentiates PathwisePILCO's closure (gpflow_pilco/loops/pilco.py:263-298) with tape.gradient w.r.t. the policy
// variables.  The gradient-mode forward (gpp_rollout_pathwise_fwd_grad, pathwise.cu) has already reduced the only expensive
// derivative — d f_l / d d_b of each particle's function draw, a sum over all F + M features — while it streamed the weights,
// and stored it as jac [H][L*D][ldS]; this sweep therefore touches 8 (L D + Dx) bytes per particle-step instead of 139 KB.
//   x_{t+1} = x_t + f(d_t),  d_t = (e_t, u_t),  e_t = enc(x_t),  u_t = scale (Phi(pi(e_t)) + shift),  loss = sum_t c(enc(x_{t+1}))
//   lambda_{t+1} += loss_bar dc/dx_{t+1};   d_bar = J_t^T lambda_{t+1};   f_bar = d_bar[De] scale phi(pi);
//   e_bar = d_bar[:De] + f_bar dpi/de;      lambda_t = lambda_{t+1} + (de/dx)^T e_bar;   theta_bar += f_bar dpi/dtheta
// One thread per particle.  The policy-parameter gradient (Mp (De + 1) + De numbers per particle-step) is reduced over the
// 32 particles of a warp through a [32][33] shared-memory transpose per batch of 32 parameters, accumulated over time in
// registers by the lane that owns the parameter, and summed over warps in a fixed order by k_pathwise_bwd_reduce.
#include "mm_small.cuh"

namespace gpp {

constexpr int kBwdWarps = 4;   // warps (x 32 particles) per CTA

struct PathwiseBwdParams {
  EncoderSpec enc;
  int S, ldS, H, L, D, Dx, De, Mp;
  const double* pZ;       // [Mp][De]
  const double* pEll;     // [De]
  const double* pBeta;    // [Mp]
  double pVar, scale;
  const double *target, *W;   // [De], [De][De]
  const double* traj;     // [H+1][S][Dx]
  const double* jac;      // [H][L*D][ldS]
  const double* loss_bar; // [S] or null
  double* x0_bar;         // [S][Dx] or null
  double* partial;        // [num_warps][nbatch*32]
  int nslots, nbatch;
};

__global__ void __launch_bounds__(32 * kBwdWarps) k_pathwise_bwd(PathwiseBwdParams p) {
  __shared__ double tile[kBwdWarps][32][33];
  __shared__ double sZ[64 * GPP_SMALL_MAX], sAlpha[64], sEll[GPP_SMALL_MAX], sT[GPP_SMALL_MAX], sW[GPP_SMALL_MAX * GPP_SMALL_MAX];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int De = p.De, Dx = p.Dx, D = p.D, L = p.L, Mp = p.Mp;
  for (int i = tid; i < Mp * De; i += blockDim.x) sZ[i] = p.pZ[i];
  for (int i = tid; i < Mp; i += blockDim.x) sAlpha[i] = p.pVar * p.pBeta[i];
  for (int i = tid; i < De; i += blockDim.x) { sEll[i] = p.pEll[i]; sT[i] = p.target[i]; }
  for (int i = tid; i < De * De; i += blockDim.x) sW[i] = p.W[i];
  __syncthreads();
  const int s = blockIdx.x * blockDim.x + tid;
  const bool valid = s < p.S;
  const double lb = valid ? (p.loss_bar ? p.loss_bar[s] : 1.0) : 0.0;
  const int na = p.enc.na, nb = p.enc.nb();
  double lam[GPP_SMALL_MAX];
  for (int i = 0; i < Dx; ++i) lam[i] = 0.0;
  double acc[8];                      // nbatch <= 8 parameter slots owned by this lane
  for (int b = 0; b < 8; ++b) acc[b] = 0.0;
  double xn[GPP_SMALL_MAX];           // x_{t+1}
  for (int i = 0; i < Dx; ++i) xn[i] = valid ? p.traj[((size_t)p.H * p.S + s) * Dx + i] : 0.0;

  for (int t = p.H - 1; t >= 0; --t) {
    // ---- cost at x_{t+1}:  c = -exp(-1/2 r^T W r), dc/de = -c (W + W^T)/2 r
    {
      double e[GPP_SMALL_MAX], de_sin[4], de_cos[4];
      for (int k = 0; k < na; ++k) {
        double sv, cv;
        sincos(xn[p.enc.active[k]], &sv, &cv);
        e[k] = sv; e[na + k] = cv;
        de_sin[k] = cv; de_cos[k] = -sv;
      }
      for (int j = 0; j < nb; ++j) e[2 * na + j] = xn[p.enc.inactive(j)];
      double d2 = 0.0, g[GPP_SMALL_MAX];
      for (int a = 0; a < De; ++a) {
        double row = 0.0;
        for (int b = 0; b < De; ++b) row = fma(0.5 * (sW[a * De + b] + sW[b * De + a]), e[b] - sT[b], row);
        g[a] = row;
        d2 = fma(row, e[a] - sT[a], d2);
      }
      const double c = lb * exp(-0.5 * d2);     // = -loss_bar * cost
      for (int k = 0; k < na; ++k) lam[p.enc.active[k]] += c * (g[k] * de_sin[k] + g[na + k] * de_cos[k]);
      for (int j = 0; j < nb; ++j) lam[p.enc.inactive(j)] += c * g[2 * na + j];
    }
    // ---- step t
    double x[GPP_SMALL_MAX], e[GPP_SMALL_MAX], de_sin[4], de_cos[4];
    for (int i = 0; i < Dx; ++i) x[i] = valid ? p.traj[((size_t)t * p.S + s) * Dx + i] : 0.0;
    for (int k = 0; k < na; ++k) {
      double sv, cv;
      sincos(x[p.enc.active[k]], &sv, &cv);
      e[k] = sv; e[na + k] = cv;
      de_sin[k] = cv; de_cos[k] = -sv;
    }
    for (int j = 0; j < nb; ++j) e[2 * na + j] = x[p.enc.inactive(j)];
    double dbar[GPP_SMALL_MAX];
    for (int b = 0; b < D; ++b) {
      double v = 0.0;
      if (valid)
        for (int l = 0; l < L; ++l) v = fma(lam[l], p.jac[((size_t)(t * L + l) * D + b) * p.ldS + s], v);
      dbar[b] = v;
    }
    // policy mean and its derivative w.r.t. e
    double f = 0.0, dfde[GPP_SMALL_MAX];
    for (int a = 0; a < De; ++a) dfde[a] = 0.0;
    for (int i = 0; i < Mp; ++i) {
      double d2 = 0.0, df[GPP_SMALL_MAX];
      for (int a = 0; a < De; ++a) {
        df[a] = (e[a] - sZ[i * De + a]) / sEll[a];
        d2 = fma(df[a], df[a], d2);
      }
      const double ak = sAlpha[i] * exp(-0.5 * d2);
      f += ak;
      for (int a = 0; a < De; ++a) dfde[a] = fma(-ak, df[a] / sEll[a], dfde[a]);
    }
    const double fbar = dbar[De] * p.scale * 0.39894228040143267794 * exp(-0.5 * f * f);
    double ebar[GPP_SMALL_MAX];
    for (int a = 0; a < De; ++a) ebar[a] = fma(fbar, dfde[a], dbar[a]);
    for (int k = 0; k < na; ++k) lam[p.enc.active[k]] += ebar[k] * de_sin[k] + ebar[na + k] * de_cos[k];
    for (int j = 0; j < nb; ++j) lam[p.enc.inactive(j)] += ebar[2 * na + j];
    // ---- policy-parameter gradient: slot i (De+1) + 0 -> beta_i, + 1 + a -> Z_i[a]; slot Mp (De+1) + a -> lengthscale a
    for (int b = 0; b < p.nbatch; ++b) {
      int cached = -1;
      double kv = 0.0, df[GPP_SMALL_MAX];
      for (int k = 0; k < 32; ++k) {
        const int slot = 32 * b + k;
        double val = 0.0;
        if (slot < Mp * (De + 1)) {
          const int i = slot / (De + 1), c = slot % (De + 1);
          if (i != cached) {
            double d2 = 0.0;
            for (int a = 0; a < De; ++a) {
              df[a] = (e[a] - sZ[i * De + a]) / sEll[a];
              d2 = fma(df[a], df[a], d2);
            }
            kv = exp(-0.5 * d2);
            cached = i;
          }
          val = (c == 0) ? fbar * p.pVar * kv : fbar * sAlpha[i] * kv * df[c - 1] / sEll[c - 1];
        } else if (slot < p.nslots) {
          const int a = slot - Mp * (De + 1);
          double sum = 0.0;
          for (int i = 0; i < Mp; ++i) {
            double d2 = 0.0, dfa = 0.0;
            for (int a2 = 0; a2 < De; ++a2) {
              const double q = (e[a2] - sZ[i * De + a2]) / sEll[a2];
              d2 = fma(q, q, d2);
              if (a2 == a) dfa = q;
            }
            sum = fma(sAlpha[i] * exp(-0.5 * d2), dfa * dfa, sum);
          }
          val = fbar * sum / sEll[a];
        }
        tile[warp][lane][k] = val;
      }
      __syncwarp();
      double col = 0.0;
      for (int q = 0; q < 32; ++q) col += tile[warp][q][lane];
      acc[b] += col;
      __syncwarp();
    }
    for (int i = 0; i < Dx; ++i) xn[i] = x[i];
  }
  if (valid && p.x0_bar)
    for (int i = 0; i < Dx; ++i) p.x0_bar[(size_t)s * Dx + i] = lam[i];
  const int gw = blockIdx.x * kBwdWarps + warp;
  for (int b = 0; b < p.nbatch; ++b) p.partial[((size_t)gw * p.nbatch + b) * 32 + lane] = acc[b];
}

// fixed-order sum over warps, then scatter into the three parameter gradients
__global__ void k_pathwise_bwd_reduce(const double* __restrict__ partial, int nwarps, int nbatch, int Mp, int De,
                                      double* __restrict__ Z_bar, double* __restrict__ ell_bar, double* __restrict__ beta_bar) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  const int nslots = Mp * (De + 1) + De;
  if (slot >= nslots) return;
  double s = 0.0;
  for (int w = 0; w < nwarps; ++w) s += partial[(size_t)w * nbatch * 32 + slot];
  if (slot < Mp * (De + 1)) {
    const int i = slot / (De + 1), c = slot % (De + 1);
    if (c == 0) beta_bar[i] = s;
    else Z_bar[i * De + c - 1] = s;
  } else {
    ell_bar[slot - Mp * (De + 1)] = s;
  }
}

}  // namespace gpp

extern "C" {

size_t gpp_rollout_pathwise_bwd_workspace_bytes(int S, int Mp, int De) {
  if (S <= 0 || Mp <= 0 || De <= 0) return 0;
  const int nslots = Mp * (De + 1) + De, nbatch = (nslots + 31) / 32;
  const size_t nwarps = (size_t)((S + 32 * gpp::kBwdWarps - 1) / (32 * gpp::kBwdWarps)) * gpp::kBwdWarps;
  return nwarps * nbatch * 32 * sizeof(double);
}

int gpp_rollout_pathwise_bwd(int S, int ldS, int H, int L, int D, int Dx, int num_active, const int* active_dims,
                             int Mp, const double* policy_Z, const double* policy_lengthscales, double policy_variance,
                             const double* policy_beta, double squash_scale, const double* cost_target, const double* cost_W,
                             const double* traj, const double* jac, const double* loss_bar,
                             double* Z_bar, double* lengthscales_bar, double* beta_bar, double* x0_bar,
                             void* workspace, size_t workspace_bytes, void* stream_) {
  GPP_NVTX_RANGE();
  using namespace gpp;
  GPP_REQUIRE(policy_Z && policy_lengthscales && policy_beta && cost_target && cost_W && traj && jac && Z_bar && lengthscales_bar && beta_bar &&
                  workspace, GPP_ERR_NULL, "gpp_rollout_pathwise_bwd: null argument");
  GPP_REQUIRE(S >= 1 && H >= 0 && L >= 1 && L <= GPP_SMALL_MAX && Dx >= 1 && Dx <= GPP_SMALL_MAX && L == Dx && ldS >= S, GPP_ERR_BAD_SHAPE,
              "gpp_rollout_pathwise_bwd: bad sizes S=%d ldS=%d H=%d L=%d Dx=%d", S, ldS, H, L, Dx);
  GPP_REQUIRE(num_active >= 0 && num_active <= 4 && D == Dx + num_active + 1 && D <= GPP_SMALL_MAX, GPP_ERR_BAD_SHAPE,
              "gpp_rollout_pathwise_bwd: D=%d must be Dx + num_active + 1", D);
  GPP_REQUIRE(Mp >= 1 && Mp <= 64, GPP_ERR_UNSUPPORTED, "gpp_rollout_pathwise_bwd: Mp=%d policy centres (max 64)", Mp);
  PathwiseBwdParams p{};
  p.enc.Dx = Dx; p.enc.na = num_active;
  for (int k = 0; k < num_active; ++k) p.enc.active[k] = active_dims[k];
  p.enc.finish();
  p.S = S; p.ldS = ldS; p.H = H; p.L = L; p.D = D; p.Dx = Dx; p.De = Dx + num_active; p.Mp = Mp;
  p.pZ = policy_Z; p.pEll = policy_lengthscales; p.pBeta = policy_beta; p.pVar = policy_variance; p.scale = squash_scale;
  p.target = cost_target; p.W = cost_W; p.traj = traj; p.jac = jac; p.loss_bar = loss_bar; p.x0_bar = x0_bar;
  p.nslots = Mp * (p.De + 1) + p.De;
  p.nbatch = (p.nslots + 31) / 32;
  GPP_REQUIRE(p.nbatch <= 8, GPP_ERR_UNSUPPORTED, "gpp_rollout_pathwise_bwd: %d policy parameters exceed the 256 supported", p.nslots);
  const size_t need = gpp_rollout_pathwise_bwd_workspace_bytes(S, Mp, p.De);
  GPP_REQUIRE(workspace_bytes >= need, GPP_ERR_WORKSPACE, "gpp_rollout_pathwise_bwd: workspace %zu < required %zu", workspace_bytes, need);
  p.partial = (double*)workspace;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int grid = (S + 32 * kBwdWarps - 1) / (32 * kBwdWarps);
  profile_begin(stream);
  k_pathwise_bwd<<<grid, 32 * kBwdWarps, 0, stream>>>(p);
  profile_end(stream);
  k_pathwise_bwd_reduce<<<(p.nslots + 127) / 128, 128, 0, stream>>>(p.partial, grid * kBwdWarps, p.nbatch, Mp, p.De, Z_bar, lengthscales_bar, beta_bar);
  count_launch(2);
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

}  // extern "C"
