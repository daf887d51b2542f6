// finalize stage of the fused moment-matched GP predict (one group of 128 threads per input): deterministic fixed-order sum of the
// tile partials of k_contract, Sff = f2 - f1 f1^T + diag(var), optional W mixing (LinearCoregionalization, upstream
// moment_matching/models.py:279-286), mean constant, jitter; with POST, the Euler moment update of the rollouts (model.cuh EulerPost).
// Shared by k_finalize (mm_predict.cu) and the persistent rollout kernel (rollout_persist.cu).
#pragma once
#include "model.cuh"
#include "persist_common.cuh"

namespace gpp {

struct FinalizeParams {
  const double* part;
  const unsigned long long* part_ll;   // persistent rollout: the partials arrive as tagged words (persist_common.cuh), tag = ll_tag
  unsigned ll_tag;                      // 0: plain doubles in `part`
  const gpp_slot* slots;
  const int* pair_start;
  const int* pair_ab;
  const double* f1lat;      // [N,L]
  const double* crosslat;   // [N,D,L]
  const double* var;        // [L]
  const double* mean;       // [P]
  const double* W;          // [P,L] or null
  double* f1;               // [N,P]
  double* Sff;              // [N,P,P]
  double* cross;            // [N,D,P]
  int N, L, P, D, npairs, nslots, full_cov, model_uncertainty;
  double jitter;
  EulerPost post;           // used by k_finalize<true> only
};

// all 128 threads of the group call this for input n.  In the persistent rollout the tile partials are written by other CTAs of the
// same launch and arrive as tagged words: every thread waits for exactly the entries it sums.
template <bool POST>
__device__ void finalize_body(const FinalizeParams& p, int n) {
  __shared__ double f2[GPP_MAX_L * GPP_MAX_L];
  __shared__ double SffL[GPP_MAX_L * GPP_MAX_L];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L, P = p.P;
  for (int t = threadIdx.x; t < L * L; t += kGroupThreads) f2[t] = 0.0;
  group_sync();
  for (int pr = warp; pr < p.npairs; pr += 4) {
    double s = 0.0;
    for (int k = p.pair_start[pr] + lane; k < p.pair_start[pr + 1]; k += 32)
      {
      const double v = p.ll_tag ? ll_load(p.part_ll + 2 * ((size_t)n * p.nslots + k), p.ll_tag) : p.part[(size_t)n * p.nslots + k];
      s = fma(p.slots[k].weight, v, s);
    }
    s = warp_sum(s);
    if (lane == 0) {
      int a = p.pair_ab[2 * pr], b = p.pair_ab[2 * pr + 1];
      f2[a * L + b] = s;
      f2[b * L + a] = s;
    }
  }
  group_sync();
  const double* f1l = p.f1lat + (size_t)n * L;
  for (int t = threadIdx.x; t < L * L; t += kGroupThreads) {
    int a = t / L, b = t % L;
    double v = f2[t] - f1l[a] * f1l[b];
    if (a == b && p.model_uncertainty) v += p.var[a];
    SffL[t] = v;
  }
  group_sync();
  // outputs (optionally mixed by W)
  for (int t = threadIdx.x; t < P; t += kGroupThreads) {
    double v = p.mean[t];
    if (p.W) {
      for (int l = 0; l < L; ++l) v = fma(p.W[t * L + l], f1l[l], v);
    } else {
      v += f1l[t];
    }
    p.f1[(size_t)n * P + t] = v;
  }
  for (int t = threadIdx.x; t < P * P; t += kGroupThreads) {
    int a = t / P, b = t % P;
    double v;
    if (p.W) {
      v = 0.0;
      for (int l = 0; l < L; ++l)
        for (int k = 0; k < L; ++k) v = fma(p.W[a * L + l] * p.W[b * L + k], SffL[l * L + k], v);
    } else {
      v = SffL[t];
    }
    if (a == b) v += p.jitter;
    if (!p.full_cov && a != b) v = 0.0;
    p.Sff[(size_t)n * P * P + t] = v;
  }
  for (int t = threadIdx.x; t < p.D * P; t += kGroupThreads) {
    int d = t / P, o = t % P;
    const double* cl = p.crosslat + ((size_t)n * p.D + d) * L;
    double v;
    if (p.W) {
      v = 0.0;
      for (int l = 0; l < L; ++l) v = fma(p.W[o * L + l], cl[l], v);
    } else {
      v = cl[o];
    }
    p.cross[((size_t)n * p.D + d) * P + o] = v;
  }
  if (POST) {
    group_sync();                         // this block's f1 / Sff / cross are visible to all its threads
    const EulerPost& e = p.post;
    const int Dx = e.Dx, D = p.D;
    const double* Sxd = e.Sxd + (size_t)n * Dx * D;
    const double* cr = p.cross + (size_t)n * D * P;
    for (int t = threadIdx.x; t < Dx * Dx + Dx; t += kGroupThreads) {
      if (t < Dx * Dx) {
        const int i = t / Dx, j = t % Dx;
        double sij = 0.0, sji = 0.0;
        for (int b = 0; b < D; ++b) {
          sij = fma(Sxd[i * D + b], cr[b * P + j], sij);
          sji = fma(Sxd[j * D + b], cr[b * P + i], sji);
        }
        const double v = e.S[(size_t)n * Dx * Dx + t] + sij + sji + p.Sff[(size_t)n * P * P + i * P + j];
        e.S[(size_t)n * Dx * Dx + t] = v;
        if (e.traj_S) e.traj_S[(size_t)n * Dx * Dx + t] = v;
        if (e.ring_S) e.ring_S[(size_t)n * Dx * Dx + t] = v;
      } else {
        const int i = t - Dx * Dx;
        const double v = e.m[(size_t)n * Dx + i] + p.f1[(size_t)n * P + i];
        e.m[(size_t)n * Dx + i] = v;
        if (e.traj_m) e.traj_m[(size_t)n * Dx + i] = v;
        if (e.ring_m) e.ring_m[(size_t)n * Dx + i] = v;
      }
    }
  }
}

template <bool POST>
__global__ void __launch_bounds__(128) k_finalize(FinalizeParams p) {
  finalize_body<POST>(p, (int)blockIdx.x);
}

}  // namespace gpp
