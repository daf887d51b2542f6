// k_pack_psi1: the per-input prologue shared by the forward (mm_predict.cu) and backward (mm_predict_bwd.cu) moment-matched GP
// predict — the Psi2 coefficient packs and the Psi1 latent means are independent, so one launch does both (blocks [0, N) = Psi1 of
// input n, the remaining blocks = 128 (input, pair) packs each) and also resets the contraction's work counter.
#pragma once
#include "model.cuh"
#include "persist_common.cuh"

namespace gpp {

// ---------------------------------------------------------------------------------------------------------
// pack_body: coefficients of one (input, kernel pair)
// ---------------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ void pack_body(int idx, const double* __restrict__ m, const double* __restrict__ S, int N,
                                          const double* __restrict__ ell, const double* __restrict__ var,
                                          const int* __restrict__ pair_ab, int npairs, int L,
                                          double* __restrict__ packs, double* __restrict__ Gs, int* info,
                                          unsigned ll_tag = 0u /* != 0: `packs` holds tagged words, two per coefficient */) {
  if (idx >= N * npairs) return;
  int n = idx / npairs, p = idx % npairs;
  // pair table (a <= b, forward) or, with pair_ab == nullptr, all L x L ordered pairs (backward)
  int a = pair_ab ? pair_ab[2 * p] : p / L, b = pair_ab ? pair_ab[2 * p + 1] : p % L;
  double V1[D], V2[D], mu[D], Sg[D * D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    double e1 = ell[a * D + d], e2 = ell[b * D + d];
    V1[d] = e1 * e1;
    V2[d] = e2 * e2;
    mu[d] = m[(size_t)n * D + d];
  }
#pragma unroll
  for (int d = 0; d < D * D; ++d) Sg[d] = S[(size_t)n * D * D + d];
  double out[PairPack<D>::SIZE];
  bool ok = make_pair_pack<D>(mu, Sg, V1, V2, log(var[a] * var[b]), out, Gs ? Gs + (size_t)idx * D * D : nullptr);
  if (!ok) flag_not_pd(info, n);
  if (ll_tag) {   // persistent rollout: the contraction CTAs poll the tags (persist_common.cuh), no fence and no flag
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(packs) + 2 * (size_t)idx * PairPack<D>::SIZE;
#pragma unroll
    for (int t = 0; t < PairPack<D>::SIZE; ++t) ll_store(dst + 2 * t, out[t], ll_tag);
    return;
  }
  double* dst = packs + (size_t)idx * PairPack<D>::SIZE;
#pragma unroll
  for (int t = 0; t < PairPack<D>::SIZE; ++t) dst[t] = out[t];
}

// ---------------------------------------------------------------------------------------------------------
// publish_packs_tagged: the coefficient packs of one input as TAGGED words (persist_common.cuh) for the persistent rollout kernels.
// All 128 threads of the group call it.  Pairs are processed in rounds of 12 (lanes 0..2 of the 4 warps run make_pair_pack side by
// side, writing straight into shared memory — no 86-double register array), then the whole group publishes the round's words with
// coalesced strong stores (7 per thread instead of 172 on one lane).  pair_of(q, a, b, dst): latent indices and pack index of pair q.
// ---------------------------------------------------------------------------------------------------------
template <int D, class PairFn>
__device__ void publish_packs_tagged(const double* __restrict__ m, const double* __restrict__ S, const int n, const int npairs, PairFn pair_of,
                                     const double* __restrict__ ell, const double* __restrict__ var, unsigned long long* __restrict__ packs_ll,
                                     double* __restrict__ Gs, const unsigned tag, int* info) {
  using PP = PairPack<D>;
  constexpr int ROUND = 12;
  __shared__ double s_pack[ROUND][PP::SIZE];
  __shared__ int s_dst[ROUND];
  const int tid = threadIdx.x, lane = tid & 31, q_local = lane * 4 + (tid >> 5);
  for (int q0 = 0; q0 < npairs; q0 += ROUND) {
    if (lane < 3 && q0 + q_local < npairs) {
      int a, b, dst;
      pair_of(q0 + q_local, a, b, dst);
      double V1[D], V2[D], mu[D], Sg[D * D];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double e1 = ell[a * D + d], e2 = ell[b * D + d];
        V1[d] = e1 * e1;
        V2[d] = e2 * e2;
        mu[d] = m[(size_t)n * D + d];
      }
#pragma unroll
      for (int d = 0; d < D * D; ++d) Sg[d] = S[(size_t)n * D * D + d];
      if (!make_pair_pack<D>(mu, Sg, V1, V2, log(var[a] * var[b]), s_pack[q_local], Gs ? Gs + (size_t)dst * D * D : nullptr)) flag_not_pd(info, n);
      s_dst[q_local] = dst;
    }
    group_sync();
    const int cnt = min(ROUND, npairs - q0);
    for (int e = tid; e < cnt * PP::SIZE; e += kGroupThreads) {
      const int q = e / PP::SIZE, t = e % PP::SIZE;
      ll_store(packs_ll + 2 * ((size_t)s_dst[q] * PP::SIZE + t), s_pack[q][t], tag);
    }
    group_sync();
  }
}

// ---------------------------------------------------------------------------------------------------------
// psi1_body: latent mean  f1[n,l] = sum_m beta_l[m] Psi1[n,m,l]  and  cross[n,:,l] = (S_n+Lambda_l)^-1 sum_m beta Psi1 (z_m - mu)
//         (models.py:236 and :264-277; Psi1 is GPflow's eKxz)
// ---------------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ void psi1_body(int n, const double* __restrict__ m, const double* __restrict__ S, int N, int L, int M,
                                          const double* __restrict__ Z, const double* __restrict__ ell,
                                          const double* __restrict__ var, const double* __restrict__ beta,
                                          double* __restrict__ f1lat /*[N,L]*/, double* __restrict__ crosslat /*[N,D,L]*/,
                                          int* info, double* li_out /* shared [L][D*D + 1]: Li and c0 per latent, for the backward */) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kGroupThreads >> 5;   // a group of 128 threads (common.cuh)
  double mu[D];
#pragma unroll
  for (int d = 0; d < D; ++d) mu[d] = m[(size_t)n * D + d];
  for (int l = warp; l < L; l += nwarps) {
    Mat<D> A, Li;
    double ell_prod = 1.0;                   // one log of the ratio of products instead of 2 D logs (D <= 8 factors of moderate size)
#pragma unroll
    for (int d = 0; d < D; ++d) {
      double e = ell[l * D + d];
      ell_prod *= e;
#pragma unroll
      for (int e2 = 0; e2 < D; ++e2) A(d, e2) = S[(size_t)n * D * D + d * D + e2] + (d == e2 ? e * e : 0.0);
    }
    bool ok = cholesky<D>(A);
    if (!ok && lane == 0) flag_not_pd(info, n);
    double diag_prod = 1.0;
#pragma unroll
    for (int d = 0; d < D; ++d) diag_prod *= A(d, d);
    tri_inverse<D>(A, Li);
    double c0 = log(var[l] * ell_prod / diag_prod);
    if (li_out && lane == 0) {
#pragma unroll
      for (int t = 0; t < D * D; ++t) li_out[l * (D * D + 1) + t] = Li.a[t];
      li_out[l * (D * D + 1) + D * D] = c0;
    }
    double acc = 0.0, vec[D];
#pragma unroll
    for (int d = 0; d < D; ++d) vec[d] = 0.0;
    const double* Zl = Z + (size_t)l * M * D;
    for (int j = lane; j < M; j += 32) {
      double dz[D];
#pragma unroll
      for (int d = 0; d < D; ++d) dz[d] = Zl[(size_t)j * D + d] - mu[d];
      double maha = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double y = 0.0;
#pragma unroll
        for (int k = 0; k <= i; ++k) y = fma(Li(i, k), dz[k], y);
        maha = fma(y, y, maha);
      }
      double w = beta[(size_t)l * M + j] * fast_exp(c0 - 0.5 * maha);
      acc += w;
#pragma unroll
      for (int d = 0; d < D; ++d) vec[d] = fma(w, dz[d], vec[d]);
    }
    acc = warp_sum(acc);
#pragma unroll
    for (int d = 0; d < D; ++d) vec[d] = warp_sum(vec[d]);
    if (lane == 0) {
      f1lat[(size_t)n * L + l] = acc;
      // G vec with G = Li^T Li
      double y[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k <= i; ++k) t = fma(Li(i, k), vec[k], t);
        y[i] = t;
      }
#pragma unroll
      for (int d = 0; d < D; ++d) {
        double t = 0.0;
#pragma unroll
        for (int i = d; i < D; ++i) t = fma(Li(i, d), y[i], t);
        crosslat[((size_t)n * D + d) * L + l] = t;
      }
    }
  }
}

struct PackPsi1Params {
  const double *m, *S;                    // [N,D], [N,D,D]
  const double *Z, *ell, *var, *beta;     // model
  const int* pair_ab;                     // pair table (a <= b) or nullptr = all L x L ordered pairs
  double *packs, *f1lat, *crosslat;       // [N,npairs,PairPack], [N,L], [N,D,L]
  double* Gs;                             // optional [N,npairs,D,D]: (Sigma_n + V_ab)^-1 of every pair (the backward's finalize needs it)
  unsigned* counter;                      // optional: work counter of k_contract, reset here
  int* info;
  int N, L, M, npairs;
};

// `epi(n, li)` runs in the Psi1 block of input n after its latent means are written (all 128 threads call it); `li` holds the
// inverse Cholesky factors and log normalisers of the block's latents when the epilogue asks for them (kNeedsFactors)
struct NoEpilogue {
  static constexpr bool kNeedsFactors = false;
  __device__ __forceinline__ void operator()(int, const double*) const {}
};

template <int D, class Epilogue>
__global__ void __launch_bounds__(128) k_pack_psi1(PackPsi1Params p, Epilogue epi) {
  if ((int)blockIdx.x < p.N) {
    __shared__ double li_sm[GPP_MAX_L * (D * D + 1)];
    psi1_body<D>(blockIdx.x, p.m, p.S, p.N, p.L, p.M, p.Z, p.ell, p.var, p.beta, p.f1lat, p.crosslat, p.info,
                 Epilogue::kNeedsFactors ? li_sm : nullptr);
    epi((int)blockIdx.x, li_sm);
  } else {
    const int idx = ((int)blockIdx.x - p.N) * 128 + threadIdx.x;
    if (idx == 0 && p.counter) *p.counter = 0u;
    pack_body<D>(idx, p.m, p.S, p.N, p.ell, p.var, p.pair_ab, p.npairs, p.L, p.packs, p.Gs, p.info);
  }
}

template <int D, class Epilogue = NoEpilogue>
inline void launch_pack_psi1(const PackPsi1Params& p, cudaStream_t stream, Epilogue epi = Epilogue()) {
  const int grid = p.N + (p.N * p.npairs + 127) / 128;
  k_pack_psi1<D, Epilogue><<<grid, 128, 0, stream>>>(p, epi);
}

}  // namespace gpp
