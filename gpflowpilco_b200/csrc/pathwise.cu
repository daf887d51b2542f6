// Pathwise (function-space sample) rollout: S particles, each pushed H steps through its own posterior function draw
//   f_{s,l}(x) = c_l + sum_i w[l,i,s] phi_{l,i}(x) + sum_j v[l,j,s] k_l(x, z_{l,j})          (decoupled sampling,
//   random-Fourier prior + canonical-basis update; gpflow_sampling, called at upstream loops/pilco.py:282-288 and
//   models/svgp.py:129-130), with the deterministic RBF policy (models/core.py:61-71) and the sample cost
//   (components.py:39-41) accumulated on the way (loops/pilco.py:272-275).  Replaces the closure body of
//   PathwisePILCO._policy_loss_closure (loops/pilco.py:277-295) — the H-loop is upstream's tf.foldl.
//
// One thread per particle, the whole rollout in one launch.  Per particle-step the kernel streams that particle's
// weights (8 B per feature: L (F + M) doubles ~ 139 KB at L=4, F=4096, M=256) — the HBM roofline term — and evaluates
// L F cosines + L M exps in FP64 — the FP64-pipe term; the two are within 20% of each other on B200.
//   * weights are stored particle-minor ([L,F,S]) so that a warp reads 256 contiguous bytes per feature;
//   * a producer warp moves [TF features x P particles] weight tiles and the matching [TF x 8] basis tile into a ring
//     of shared-memory stages with cp.async.bulk (TMA, SASS UBLKCP) + mbarrier complete_tx; consumer warps never wait
//     on a global load;
//   * cos is evaluated in quarter turns: the basis is pre-scaled by 4/(2 pi ell) so the reduction is an exact
//     round-to-nearest (no Cody-Waite), then a degree-6 polynomial in r^2 with sin/cos coefficients selected by quadrant.
#include <cublas_v2.h>

#include <algorithm>
#include <cstring>

#include "mm_small.cuh"
#include "mma_exp.cuh"
#include "model.cuh"
#include "philox.cuh"

namespace gpp {

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// cos(pi/2 * q) for K values in lock-step; q in quarter turns.
//   q = 2 k + r, k = rint(q / 2), r in [-1, 1] (exact);  cos(pi/2 q) = (-1)^k cos(pi/2 r) = (-1)^k (P(r^2)^2 - 1),
//   P(z) = sqrt(2) cos(pi/4 sqrt(z)): one degree-6 polynomial for every lane (the double-angle step replaces the
//   per-quadrant sin/cos coefficient selects, which cost 14 ALU selects per value and made the loop issue-bound:
//   profiles/r1_pathwise_full.txt).  11 FP64 ops, 2 integer ops; absolute error <= 8e-16.
static __constant__ double kCosH[8] = {0x1.6a09e667f3bccp+0, -0x1.bea5b6072b1ecp-2, 0x1.6f5a49b29085dp-6, -0x1.e36ab0b62dd2bp-12,
                                       0x1.54cb876c4eb32p-18, -0x1.2af4d2ca26f33p-25, 0x1.61741369cd96ep-33, 0.0};

template <int K>
__device__ __forceinline__ void cos_quarter_turns(double (&q)[K]) {
  const double MAGIC = 6755399441055744.0;
  double t[K], r[K], z[K], p[K];
#pragma unroll
  for (int k = 0; k < K; ++k) t[k] = fma(q[k], 0.5, MAGIC);
#pragma unroll
  for (int k = 0; k < K; ++k) r[k] = t[k] - MAGIC;
#pragma unroll
  for (int k = 0; k < K; ++k) r[k] = fma(r[k], -2.0, q[k]);        // exact: r in [-1, 1]
#pragma unroll
  for (int k = 0; k < K; ++k) z[k] = r[k] * r[k];
#pragma unroll
  for (int k = 0; k < K; ++k) p[k] = fma(z[k], kCosH[6], kCosH[5]);
#pragma unroll
  for (int j = 4; j >= 0; --j)
#pragma unroll
    for (int k = 0; k < K; ++k) p[k] = fma(p[k], z[k], kCosH[j]);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double val = fma(p[k], p[k], -1.0);
    // (-1)^n through the sign bit: adding n * 2^31 to the high word toggles bit 31 for odd n (one IMAD instead of shift + xor)
    q[k] = make_double((int)((unsigned)hi_int(val) + (unsigned)lo_int(t[k]) * 0x80000000u), lo_int(val));
  }
}

// Mixed-precision variant (gpp_rollout_pathwise_fwd_mixed).  ONE FP64 add reduces the phase: q + 1.5 * 2^22 has an ulp of 2^-30, so the low
// 32 bits of its mantissa are round(q * 2^30) mod 2^32 — the phase modulo 4 quarter turns (one full turn) as a 32-bit fixed-point
// number, the wrap-around of integer arithmetic doing the modulo (|q| < 2^21; resolution 1e-9 quarter turns).  Quadrant folding is
// integer work (r = phase - 2 round(phase / 2) in [-1, 1), sign = parity of the rounding), then an int -> float conversion and an
// FP32 polynomial on the FMA pipe: degree 4 in z = r^2, least-squares Chebyshev fit of sqrt(2) cos(pi/4 sqrt z) on [0, 1] (fit
// error 7e-11), and the same double-angle step as the FP64 kernel; abs. error of the cosine <= 3.3e-7.  (A double -> float
// conversion of an FP64-reduced r instead costs a quarter-rate F2F per value: measured 0.54 of the halved HBM stream.)
template <int K>
__device__ __forceinline__ void cos_quarter_turns_f32(const double (&q)[K], float (&c)[K]) {
  const double MAGIC30 = 6291456.0;            // 1.5 * 2^22
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const unsigned u = (unsigned)lo_int(q[k] + MAGIC30) + 0x40000000u;        // bit 31: parity of round(phase / 2)
    const int rfix = (int)(u & 0x7fffffffu) - 0x40000000;                     // r * 2^30, r in [-1, 1)
    const float r = (float)rfix * 9.313225746154785e-10f;
    const float z = r * r;
    float p = fmaf(z, 4.991897185391281e-06f, -0.0004609467869158834f);
    p = fmaf(p, z, 0.022421400994062424f);
    p = fmaf(p, z, -0.4361790120601654f);
    p = fmaf(p, z, 1.4142135381698608f);
    const float v = fmaf(p, p, -1.0f);
    c[k] = __int_as_float(__float_as_int(v) ^ (int)(u & 0x80000000u));
  }
}

// cos and sin of pi/2 * q in lock-step (gradient mode): sin(pi/2 r) = 2 sin(pi/4 r) cos(pi/4 r) = P(z) * (r * Qs(z)),
// Qs(z) = sqrt(2) sin(pi/4 sqrt z) / sqrt z  (degree 6; abs. error <= 5e-16).  19 FP64 ops for the pair.
static __constant__ double kSinH[8] = {0x1.1c5831add62e4p+0, -0x1.d3ba5c1c5f465p-4, 0x1.cda106381dcb8p-9, -0x1.b1e9f34807cbep-15,
                                       0x1.dbd6f6b8a9057p-22, -0x1.5588a73e88afep-29, 0x1.5634879bdf5c2p-37, 0.0};

template <int K>
__device__ __forceinline__ void sincos_quarter_turns(double (&q)[K], double (&sn)[K]) {
  const double MAGIC = 6755399441055744.0;
  double t[K], r[K], z[K], p[K], s[K];
#pragma unroll
  for (int k = 0; k < K; ++k) t[k] = fma(q[k], 0.5, MAGIC);
#pragma unroll
  for (int k = 0; k < K; ++k) r[k] = t[k] - MAGIC;
#pragma unroll
  for (int k = 0; k < K; ++k) r[k] = fma(r[k], -2.0, q[k]);
#pragma unroll
  for (int k = 0; k < K; ++k) z[k] = r[k] * r[k];
#pragma unroll
  for (int k = 0; k < K; ++k) { p[k] = fma(z[k], kCosH[6], kCosH[5]); s[k] = fma(z[k], kSinH[6], kSinH[5]); }
#pragma unroll
  for (int j = 4; j >= 0; --j)
#pragma unroll
    for (int k = 0; k < K; ++k) { p[k] = fma(p[k], z[k], kCosH[j]); s[k] = fma(s[k], z[k], kSinH[j]); }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const unsigned nlo = (unsigned)lo_int(t[k]);
    double c = fma(p[k], p[k], -1.0);
    double sv = p[k] * (r[k] * s[k]);
    q[k] = make_double((int)((unsigned)hi_int(c) + nlo * 0x80000000u), lo_int(c));
    sn[k] = make_double((int)((unsigned)hi_int(sv) + nlo * 0x80000000u), lo_int(sv));
  }
}

constexpr int kPathP = 512;    // particles per CTA (one consumer thread each) -> 16 consumer warps + 1 producer warp
constexpr int kPathPGrad = 256;   // gradient mode
constexpr int kPathTF = 8;    // feature rows per pipeline stage

struct PathwiseParams {
  EncoderSpec enc;
  int S, ldS, H, L, F, Mpad, Dx, De, Mp;
  const double* basis;    // [L][F][BS]     4 omega/(2 pi ell) [D], 4 b/(2 pi)
  const double* zbasis;   // [L][Mpad][BS]  z/ell [D]
  const double* w;        // [L][F][ldS]
  const float* w32;       // mixed-precision variant: the same weights in FP32 (w is not read then)
  const double* v;        // [L][Mpad][ldS]
  const double* amp;      // [L] sqrt(2 var/F)
  const double* var;      // [L]
  const double* inv_ell;  // [L][D]
  const double* mean;     // [L]
  const double* pZs;      // [Mp][De]  policy centres / ell_pi
  const double* pInvEll;  // [De]
  const double* pAlpha;   // [Mp]      var_pi * Kuu^-1 m
  double scale, shift;
  const double *target, *W;
  const double* x0;       // [S][Dx]
  double* loss;           // [S]
  double* x_final;        // [S][Dx]
  double* traj;           // optional [H+1][S][Dx]
  double* jac;            // gradient mode: [H][L*D][ldS]  d f_l / d d_b of every particle-step (particle-minor)
};

template <int D, int P, int TF, int NS>
struct PathwiseCfg {
  static constexpr int BS = (D + 1 + 1) & ~1;
  // weight rows are padded by 2 doubles: in the tensor-core form of the phase (below) the lanes of a quad read 4 different feature rows
  // at the same particle; a row stride of 2 (mod 8) doubles spreads a half-warp's 64-bit loads over all 32 banks
  static constexpr int PS = P + 2;
  // mixed-precision variant: FP32 weight rows, padded by 4 floats (row stride = 4 mod 16 words: the 32 lanes (r, c) of a warp read
  // rows 2c (+1) at particles r and hit 32 different banks); 16-byte multiples for the bulk copies
  static constexpr int PS32 = P + 4;
  static_assert((PS32 * 4) % 16 == 0 && PS32 * 4 <= PS * 8, "FP32 rows fit in, and are aligned like, the FP64 rows' space");
  static constexpr int STAGE_DOUBLES = TF * BS + TF * PS;
  static constexpr int NWARPS = P / 32 + 1;
  static_assert((NS & (NS - 1)) == 0, "the 32-bit tile counter may wrap: stage = it % NS and parity = (it / NS) & 1 need a power-of-two NS");
  // phases on the FP64 tensor path: [d, 1] (D + 1 <= 8 entries = 2 k-steps of mma.m8n8k4) times the basis tile; needs TF == 8
  static constexpr bool MMA_PHASE = (D + 1 <= 8) && TF == 8;
  static constexpr int DSTAGE = MMA_PHASE ? P * 8 : 0;            // per-particle inputs [P][8] = [d (D), 1, 0..] for the A fragments
  static constexpr size_t SMEM = sizeof(double) * NS * STAGE_DOUBLES + 2 * NS * sizeof(uint64_t) +
                                 sizeof(double) * (64 * (GPP_SMALL_MAX + 1) + 8 * GPP_SMALL_MAX * GPP_SMALL_MAX) + sizeof(double) * DSTAGE;
};

template <int D, int P, int TF, int NS, bool GRAD, bool W32 = false>
__global__ void __launch_bounds__(P + 32, 1) k_pathwise_rollout(PathwiseParams p) {
  static_assert(!W32 || (!GRAD && PathwiseCfg<D, P, TF, NS>::MMA_PHASE), "FP32 weights: forward only, tensor-core phase form");
  using CF = PathwiseCfg<D, P, TF, NS>;
  constexpr int BS = CF::BS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stages = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(stages + NS * CF::STAGE_DOUBLES);
  uint64_t* empty = full + NS;
  double* pol = reinterpret_cast<double*>(empty + NS);          // policy centres [Mp][De] + alpha [Mp]  (Mp <= 64)
  double* cst = pol + 64 * (GPP_SMALL_MAX + 1);                  // target, W, inv_ell, amp, var, mean
  double* dstage = cst + 8 * GPP_SMALL_MAX * GPP_SMALL_MAX;      // [P][8] (tensor-core phase form only)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s0 = blockIdx.x * P;
  const int tiles_f = p.F / TF, tiles_m = p.Mpad / TF;

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], P / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < p.Mp * p.De; i += blockDim.x) pol[i] = p.pZs[i];
  for (int i = tid; i < p.Mp; i += blockDim.x) pol[64 * GPP_SMALL_MAX + i] = p.pAlpha[i];
  {
    const int De = p.De;
    for (int i = tid; i < De; i += blockDim.x) { cst[i] = p.target[i]; cst[8 + 64 + i] = p.pInvEll[i]; }
    for (int i = tid; i < De * De; i += blockDim.x) cst[8 + i] = p.W[i];
    for (int i = tid; i < p.L * D; i += blockDim.x) cst[8 + 64 + 8 + i] = p.inv_ell[i];
    for (int i = tid; i < p.L; i += blockDim.x) {
      cst[8 + 64 + 8 + 64 + i] = p.amp[i];
      cst[8 + 64 + 8 + 64 + 8 + i] = p.var[i];
      cst[8 + 64 + 8 + 64 + 16 + i] = p.mean[i];
    }
  }
  __syncthreads();
  const double* c_target = cst;
  const double* c_W = cst + 8;
  const double* c_pinv = cst + 8 + 64;
  const double* c_inv_ell = cst + 8 + 64 + 8;
  const double* c_amp = cst + 8 + 64 + 8 + 64;
  const double* c_var = c_amp + 8;
  const double* c_mean = c_amp + 16;

  if (warp == P / 32) {
    // ------------------------------------------------------------------ producer warp: lane 0 = basis tile, lanes 1..TF = weight rows
    // (one elected lane issuing all TF + 1 bulk copies of a tile needs ~150 cycles per copy; that bounded the mixed-precision variant,
    //  whose consumers are done with a tile sooner than 9 serial copies can be issued)
    if (lane <= TF) {
      const int pcount = min(P, p.ldS - s0);        // particles this CTA really has columns for (last CTA of a launch)
      constexpr int PS = CF::PS;
      constexpr unsigned kIssuers = (1u << (TF + 1)) - 1u;
      const unsigned bytes64 = (unsigned)(sizeof(double) * (TF * BS + TF * pcount));
      const unsigned bytes32 = (unsigned)(sizeof(double) * TF * BS + sizeof(float) * TF * pcount);
      const int f = lane - 1;                       // this lane's weight row of the tile
      unsigned it = 0;
      for (int t = 0; t < p.H; ++t)
        for (int l = 0; l < p.L; ++l)
          for (int tile = 0; tile < tiles_f + tiles_m; ++tile, ++it) {
            const int st = (int)(it % NS);
            const unsigned ph = (unsigned)((it / NS) & 1);
            mbar_wait(&empty[st], ph ^ 1u);         // every issuing lane sees the stage released itself
            double* sb = stages + st * CF::STAGE_DOUBLES;
            const bool rff = tile < tiles_f;
            const int row0 = (rff ? tile : tile - tiles_f) * TF;
            if (lane == 0) {
              mbar_expect_tx(&full[st], (W32 && rff) ? bytes32 : bytes64);
              const double* bsrc = rff ? p.basis + ((size_t)l * p.F + row0) * BS : p.zbasis + ((size_t)l * p.Mpad + row0) * BS;
              bulk_g2s(sb, bsrc, (unsigned)(sizeof(double) * TF * BS), &full[st]);
            }
            __syncwarp(kIssuers);                   // the expected byte count is registered before any row can complete
            if (lane == 0) continue;
            if (W32 && rff) {
              const float* wsrc32 = p.w32 + ((size_t)l * p.F + row0 + f) * p.ldS + s0;
              bulk_g2s(reinterpret_cast<float*>(sb + TF * BS) + f * CF::PS32, wsrc32, (unsigned)(sizeof(float) * pcount), &full[st]);
            } else {
              const double* wsrc = rff ? p.w + ((size_t)l * p.F + row0 + f) * p.ldS + s0 : p.v + ((size_t)l * p.Mpad + row0 + f) * p.ldS + s0;
              bulk_g2s(sb + TF * BS + f * PS, wsrc, (unsigned)(sizeof(double) * pcount), &full[st]);
            }
          }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers: one particle per thread
  constexpr int PS = CF::PS;
  constexpr bool MMA = CF::MMA_PHASE && !GRAD;
  const int s = s0 + tid;
  const bool valid = s < p.S;
  const int Dx = p.Dx, De = p.De;
  double x[GPP_SMALL_MAX], loss = 0.0;
  for (int i = 0; i < Dx; ++i) x[i] = valid ? p.x0[(size_t)s * Dx + i] : 0.0;
  if (p.traj && valid)
    for (int i = 0; i < Dx; ++i) p.traj[(size_t)s * Dx + i] = x[i];
  unsigned it = 0;   // 32-bit: H * L * tiles stays far below 2^32 and keeps the stage / parity arithmetic to two ALU ops
  for (int t = 0; t < p.H; ++t) {
    // e = [sin, cos, inactive]; u = scale (Phi(policy mean) + shift); d = (e, u)
    double dd[D];
    {
      const int na = p.enc.na;
      for (int k = 0; k < na; ++k) {
        double sv, cv;
        sincos(x[p.enc.active[k]], &sv, &cv);
        dd[k] = sv;
        dd[na + k] = cv;
      }
      for (int j = 0; j < p.enc.nb(); ++j) dd[2 * na + j] = x[p.enc.inactive(j)];
      double es[GPP_SMALL_MAX], f = 0.0;
      for (int a = 0; a < De; ++a) es[a] = dd[a] * c_pinv[a];
      for (int i = 0; i < p.Mp; ++i) {
        double d2 = 0.0;
        for (int a = 0; a < De; ++a) {
          double df = es[a] - pol[i * De + a];
          d2 = fma(df, df, d2);
        }
        f = fma(pol[64 * GPP_SMALL_MAX + i], fast_exp(-0.5 * d2), f);
      }
      dd[De] = p.scale * (0.5 * erfc(-f * 0.70710678118654752440) + p.shift);
    }
    double fx[GPP_SMALL_MAX];
    for (int l = 0; l < p.L; ++l) {
      double accw[4] = {0.0, 0.0, 0.0, 0.0}, accv[4] = {0.0, 0.0, 0.0, 0.0};
      double jw[GRAD ? D : 1], jv[GRAD ? D : 1];     // gradient mode: sum_i w_i sin_i basis_i[d], sum_j v_j k_j (zs_j - ds)[d]
      if (GRAD) {
#pragma unroll
        for (int d = 0; d < D; ++d) { jw[d] = 0.0; jv[d] = 0.0; }
      }
      if (MMA) {
        // Random-Fourier part on the FP64 tensor path.  The phase of (particle s, feature i) is the length-(D+1) inner product
        // [d_s, 1] . [4 omega_i / (2 pi ell), 4 b_i / (2 pi)]: per warp and 8-feature tile that is a [32 x 8] . [8 x 8] product = 4 row
        // groups x 2 k-steps of mma.sync.m8n8k4.f64 (SASS DMMA) instead of 8 x D three-register DFMAs and as many broadcast loads per
        // thread.  DMMA gives lane (r, c) = (lane >> 2, lane & 3) the phases of particle 8 g + r at features 2 c, 2 c + 1: the lane
        // then does the cosines and the weight FMAs for those 8 (particle, feature) pairs; the per-particle sums are put back
        // together across the 4 lanes of a quad and handed to the particle's own thread at the end of the latent.
        const int r = lane >> 2, c = lane & 3;
        double afr[4][2];
        if (l == 0) {                            // this step's inputs -> A fragments (same for every latent)
          double* mine = dstage + (size_t)tid * 8;
#pragma unroll
          for (int d = 0; d < 8; ++d) mine[d] = d < D ? dd[d] : (d == D ? 1.0 : 0.0);
          __syncwarp();
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) afr[g][ks] = dstage[(size_t)(warp * 32 + 8 * g + r) * 8 + c + 4 * ks];
        double acc[4][2];
        float accf[W32 ? 4 : 1][2];              // mixed precision: FP32 partial sums, folded into `acc` every 32 tiles
#pragma unroll
        for (int g = 0; g < 4; ++g) acc[g][0] = acc[g][1] = 0.0;
#pragma unroll
        for (int g = 0; g < (W32 ? 4 : 1); ++g) accf[g][0] = accf[g][1] = 0.0f;
        // phases of the 8 (particle, feature) pairs of this lane in tile `it_` (waits for the tile's stage)
        auto phases = [&](unsigned it_, double (&qq)[8]) {
          const int st = (int)(it_ % NS);
          mbar_wait(&full[st], (unsigned)((it_ / NS) & 1));
          const double* bs = stages + st * CF::STAGE_DOUBLES;
          // B fragment: lane (n, k) = (lane >> 2, lane & 3) holds entry k + 4 ks of feature n's basis row (rows are BS wide, zero padded)
          const double b0 = bs[r * BS + c];
          const double b1 = (c + 4 < BS) ? bs[r * BS + c + 4] : 0.0;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            qq[2 * g] = 0.0; qq[2 * g + 1] = 0.0;
            dmma_m8n8k4(qq[2 * g], qq[2 * g + 1], afr[g][0], b0);
            dmma_m8n8k4(qq[2 * g], qq[2 * g + 1], afr[g][1], b1);
          }
        };
        // (Issuing the next tile's DMMAs ahead of this tile's FP32 work — software pipelining by one tile inside the warp — was
        //  measured SLOWER for the mixed variant, 7.96 vs 5.90 ms at H = 4: the second set of phases costs 16 registers at the
        //  96-register cap of a 544-thread CTA.)
        // unrolled by 4 tiles: at the 96-register cap ptxas re-materialises the lane-derived shared-memory offsets and re-loads the
        // polynomial constants at the top of every iteration (~33 non-FP64 instructions per tile, each ~0.8 issue cycles beside the
        // FP64 pipe); once per four tiles instead: 8.10 -> 7.52 ms at H = 4 (0.80 -> 0.86 of the HBM peak), by 8: no further gain
#pragma unroll 4
        for (int tile = 0; tile < tiles_f; ++tile, ++it) {
          const int st = (int)(it % NS);
          double q[8];
          phases(it, q);
          if constexpr (W32) {
            const float* wsf = reinterpret_cast<const float*>(stages + st * CF::STAGE_DOUBLES + TF * BS) + warp * 32 + r;
            float cf[8];
            cos_quarter_turns_f32<8>(q, cf);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              accf[g][0] = fmaf(wsf[(2 * c) * CF::PS32 + 8 * g], cf[2 * g], accf[g][0]);
              accf[g][1] = fmaf(wsf[(2 * c + 1) * CF::PS32 + 8 * g], cf[2 * g + 1], accf[g][1]);
            }
            if ((tile & 31) == 31 || tile + 1 == tiles_f) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                acc[g][0] += (double)accf[g][0];
                acc[g][1] += (double)accf[g][1];
                accf[g][0] = accf[g][1] = 0.0f;
              }
            }
          } else {
            const double* ws = stages + st * CF::STAGE_DOUBLES + TF * BS + warp * 32 + r;
            cos_quarter_turns<8>(q);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              acc[g][0] = fma(ws[(2 * c) * PS + 8 * g], q[2 * g], acc[g][0]);
              acc[g][1] = fma(ws[(2 * c + 1) * PS + 8 * g], q[2 * g + 1], acc[g][1]);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[st]);
        }
        // particle 8 g + r: sum over the quad's 4 lanes (fixed order), then to the thread that owns the particle (lane 8 g + r)
        double mine_sum = 0.0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          double v = acc[g][0] + acc[g][1];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          const double got = __shfl_sync(0xffffffffu, v, 4 * (lane & 7));
          if ((lane >> 3) == g) mine_sum = got;
        }
        accw[0] = mine_sum;
      }
      for (int tile = 0; !MMA && tile < tiles_f; ++tile, ++it) {
        const int st = (int)(it % NS);
        mbar_wait(&full[st], (unsigned)((it / NS) & 1));
        const double* bs = stages + st * CF::STAGE_DOUBLES;
        const double* ws = bs + TF * BS + tid;
#pragma unroll
        for (int f0 = 0; f0 < TF; f0 += 4) {
          double q[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) q[k] = bs[(f0 + k) * BS + D];
#pragma unroll
          for (int d = 0; d < D; ++d)
#pragma unroll
            for (int k = 0; k < 4; ++k) q[k] = fma(bs[(f0 + k) * BS + d], dd[d], q[k]);
          if (GRAD) {
            double sn[4];
            sincos_quarter_turns<4>(q, sn);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const double w = ws[(f0 + k) * PS];
              accw[k] = fma(w, q[k], accw[k]);
              sn[k] *= w;
            }
#pragma unroll
            for (int d = 0; d < D; ++d)
#pragma unroll
              for (int k = 0; k < 4; ++k) jw[d] = fma(sn[k], bs[(f0 + k) * BS + d], jw[d]);
          } else {
            cos_quarter_turns<4>(q);
#pragma unroll
            for (int k = 0; k < 4; ++k) accw[k] = fma(ws[(f0 + k) * PS], q[k], accw[k]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
      }
      double ds[D];
#pragma unroll
      for (int d = 0; d < D; ++d) ds[d] = dd[d] * c_inv_ell[l * D + d];
      // (two tiles per iteration in the forward-only kernels, like the Fourier part; the gradient-mode kernel is at its register
      //  limit and spills when unrolled)
      constexpr int kUnrollM = GRAD ? 1 : 2;
#pragma unroll kUnrollM
      for (int tile = 0; tile < tiles_m; ++tile, ++it) {
        const int st = (int)(it % NS);
        mbar_wait(&full[st], (unsigned)((it / NS) & 1));
        const double* bs = stages + st * CF::STAGE_DOUBLES;
        const double* ws = bs + TF * BS + tid;
#pragma unroll
        for (int f0 = 0; f0 < TF; f0 += 4) {
          double q[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
          for (int d = 0; d < D; ++d)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              double df = ds[d] - bs[(f0 + k) * BS + d];
              q[k] = fma(df, df, q[k]);
            }
#pragma unroll
          for (int k = 0; k < 4; ++k) q[k] *= -0.5;
          fast_exp_n<4>(q);
          if (GRAD) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              q[k] *= ws[(f0 + k) * PS];
              accv[k] += q[k];
            }
#pragma unroll
            for (int d = 0; d < D; ++d)
#pragma unroll
              for (int k = 0; k < 4; ++k) jv[d] = fma(q[k], bs[(f0 + k) * BS + d] - ds[d], jv[d]);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) accv[k] = fma(ws[(f0 + k) * PS], q[k], accv[k]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
      }
      fx[l] = c_mean[l] + c_amp[l] * ((accw[0] + accw[1]) + (accw[2] + accw[3])) + c_var[l] * ((accv[0] + accv[1]) + (accv[2] + accv[3]));
      if (GRAD && valid) {
        // d f_l / d d_b = -(pi/2) amp_l sum_i w_i sin(pi/2 q_i) basis_i[b] + var_l sum_j v_j k_j (zs_j - ds)[b] / ell_l[b]
#pragma unroll
        for (int d = 0; d < D; ++d)
          p.jac[((size_t)(t * p.L + l) * D + d) * p.ldS + s] =
              -1.5707963267948966192 * c_amp[l] * jw[d] + c_var[l] * jv[d] * c_inv_ell[l * D + d];
      }
    }
    for (int i = 0; i < Dx; ++i) x[i] += fx[i];                       // Euler, dt = 1, no diffusion (solvers.py:49-65)
    {
      double e[GPP_SMALL_MAX];
      const int na = p.enc.na;
      for (int k = 0; k < na; ++k) {
        double sv, cv;
        sincos(x[p.enc.active[k]], &sv, &cv);
        e[k] = sv;
        e[na + k] = cv;
      }
      for (int j = 0; j < p.enc.nb(); ++j) e[2 * na + j] = x[p.enc.inactive(j)];
      loss += sample_cost(De, e, c_target, c_W);
    }
    if (p.traj && valid)
      for (int i = 0; i < Dx; ++i) p.traj[((size_t)(t + 1) * p.S + s) * Dx + i] = x[i];
  }
  if (valid) {
    p.loss[s] = loss;
    if (p.x_final)
      for (int i = 0; i < Dx; ++i) p.x_final[(size_t)s * Dx + i] = x[i];
  }
}

// ---- basis packing ---------------------------------------------------------------------------------------
// basis[l][i][:] = 4 omega[l,i,:] / (2 pi ell_l), 4 phase/(2 pi);   zbasis[l][j][:] = z_{l,j} / ell_l  (zero rows pad to Mpad)
__global__ void k_pack_basis(int L, int F, int M, int Mpad, int D, int BS, const double* __restrict__ omega,
                             const double* __restrict__ phase, const double* __restrict__ Z, const double* __restrict__ ell,
                             double* __restrict__ basis, double* __restrict__ zbasis) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const double inv2pi = 0.15915494309189533577;
  if (idx < L * F) {
    int l = idx / F;
    for (int d = 0; d < BS; ++d) {
      double v = 0.0;
      if (d < D) v = 4.0 * inv2pi * omega[(size_t)idx * D + d] / ell[l * D + d];
      else if (d == D) v = 4.0 * inv2pi * phase[idx];
      basis[(size_t)idx * BS + d] = v;
    }
  }
  if (idx < L * Mpad) {
    int l = idx / Mpad, j = idx % Mpad;
    for (int d = 0; d < BS; ++d) {
      double v = 0.0;
      if (d < D && j < M) v = Z[((size_t)l * M + j) * D + d] / ell[l * D + d];
      zbasis[(size_t)idx * BS + d] = v;
    }
  }
}

// FP64 -> FP32 copy of the Fourier weights for the mixed-precision rollout (round to nearest even, grid-stride)
__global__ void k_weights_f32(long long count, const double* __restrict__ w, float* __restrict__ w32) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) w32[i] = (float)w[i];
}

template <int D, bool GRAD, bool W32 = false>
static int launch_pathwise(const PathwiseParams& p, cudaStream_t stream) {
  // gradient mode carries 2 D more accumulators per thread: 256-particle CTAs (8 consumer warps + the producer, one CTA per SM)
  // keep them in registers without spills (168 registers); measured faster than 2 spilling CTAs per SM or 480-particle CTAs
  constexpr int P = GRAD ? kPathPGrad : kPathP, TF = kPathTF, NS = 4;
  using CF = PathwiseCfg<D, P, TF, NS>;
  static PerDeviceSmemOptIn configured;
  if (configured.raise(CF::SMEM)) {
    GPP_CUDA_OK(cudaFuncSetAttribute(k_pathwise_rollout<D, P, TF, NS, GRAD, W32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF::SMEM));
  }
  int grid = (p.S + P - 1) / P;
  profile_begin(stream);
  k_pathwise_rollout<D, P, TF, NS, GRAD, W32><<<grid, P + 32, CF::SMEM, stream>>>(p);
  profile_end(stream);
  count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

}  // namespace gpp

static int pathwise_fwd_impl(int S, int ldS, int H, int L, int F, int Mpad, int D, int Dx, int num_active, const int* active_dims,
                             const double* basis, const double* zbasis, const double* w, const double* v, const double* amp,
                             const double* variance, const double* inv_lengthscales, const double* mean_const,
                             int Mp, const double* policy_Zs, const double* policy_inv_lengthscales, const double* policy_alpha,
                             double squash_scale, double squash_shift, const double* cost_target, const double* cost_W,
                             const double* x0, double* loss, double* x_final, double* traj, double* jac, void* stream,
                             const float* w32 = nullptr) {
  using namespace gpp;
  GPP_REQUIRE(basis && zbasis && (w || w32) && v && amp && variance && inv_lengthscales && mean_const && policy_Zs && policy_inv_lengthscales &&
                  policy_alpha && cost_target && cost_W && x0 && loss, GPP_ERR_NULL, "gpp_rollout_pathwise_fwd: null argument");
  GPP_REQUIRE(S >= 1 && H >= 0 && L >= 1 && L <= GPP_SMALL_MAX && Dx >= 1 && Dx <= GPP_SMALL_MAX, GPP_ERR_BAD_SHAPE,
              "gpp_rollout_pathwise_fwd: bad sizes S=%d H=%d L=%d Dx=%d", S, H, L, Dx);
  GPP_REQUIRE(L == Dx, GPP_ERR_BAD_SHAPE, "gpp_rollout_pathwise_fwd: the drift must have one output per state dim (L=%d, Dx=%d)", L, Dx);
  GPP_REQUIRE(F % kPathTF == 0 && Mpad % kPathTF == 0, GPP_ERR_BAD_SHAPE, "gpp_rollout_pathwise_fwd: F=%d and Mpad=%d must be multiples of %d", F, Mpad, kPathTF);
  GPP_REQUIRE(ldS % kPathP == 0 && ldS >= S, GPP_ERR_BAD_SHAPE, "gpp_rollout_pathwise_fwd: ldS=%d must be a multiple of %d and >= S=%d", ldS, kPathP, S);
  GPP_REQUIRE(Mp >= 1 && Mp <= 64, GPP_ERR_UNSUPPORTED, "gpp_rollout_pathwise_fwd: Mp=%d policy centres (max 64)", Mp);
  GPP_REQUIRE(num_active >= 0 && num_active <= 4 && D == Dx + num_active + 1, GPP_ERR_BAD_SHAPE,
              "gpp_rollout_pathwise_fwd: D=%d must be Dx + num_active + 1", D);
  PathwiseParams p{};
  p.enc.Dx = Dx; p.enc.na = num_active;
  for (int k = 0; k < num_active; ++k) p.enc.active[k] = active_dims[k];
  p.enc.finish();
  p.S = S; p.ldS = ldS; p.H = H; p.L = L; p.F = F; p.Mpad = Mpad; p.Dx = Dx; p.De = Dx + num_active; p.Mp = Mp;
  p.basis = basis; p.zbasis = zbasis; p.w = w; p.w32 = w32; p.v = v; p.amp = amp; p.var = variance; p.inv_ell = inv_lengthscales; p.mean = mean_const;
  p.pZs = policy_Zs; p.pInvEll = policy_inv_lengthscales; p.pAlpha = policy_alpha; p.scale = squash_scale; p.shift = squash_shift;
  p.target = cost_target; p.W = cost_W; p.x0 = x0; p.loss = loss; p.x_final = x_final; p.traj = traj; p.jac = jac;
  if (w32) {
    GPP_REQUIRE(!jac, GPP_ERR_UNSUPPORTED, "gpp_rollout_pathwise_fwd_mixed: no gradient mode");
    switch (D) {
#define GPP_CASE(d) case d: return launch_pathwise<d, false, true>(p, (cudaStream_t)stream);
      GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7)
#undef GPP_CASE
      default: set_error("gpp_rollout_pathwise_fwd_mixed: unsupported D=%d (2..7)", D); return GPP_ERR_UNSUPPORTED;
    }
  }
  switch (D) {
#define GPP_CASE(d) case d: return jac ? launch_pathwise<d, true>(p, (cudaStream_t)stream) : launch_pathwise<d, false>(p, (cudaStream_t)stream);
    GPP_CASE(2) GPP_CASE(3) GPP_CASE(4) GPP_CASE(5) GPP_CASE(6) GPP_CASE(7) GPP_CASE(8)
#undef GPP_CASE
    default: set_error("gpp_rollout_pathwise_fwd: unsupported D=%d", D); return GPP_ERR_UNSUPPORTED;
  }
}


extern "C" {

int gpp_pathwise_tile(void) { return gpp::kPathTF; }
int gpp_pathwise_particles_per_cta(void) { return gpp::kPathP; }

int gpp_pathwise_pack_basis(int L, int F, int M, int Mpad, int D, const double* omega, const double* phase, const double* Z,
                            const double* lengthscales, double* basis, double* zbasis, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(omega && phase && Z && lengthscales && basis && zbasis, GPP_ERR_NULL, "gpp_pathwise_pack_basis: null argument");
  GPP_REQUIRE(D >= 1 && D <= GPP_MAX_D && Mpad >= M, GPP_ERR_BAD_SHAPE, "gpp_pathwise_pack_basis: bad sizes");
  int BS = (D + 2) & ~1;
  int n = L * (F > Mpad ? F : Mpad);
  gpp::k_pack_basis<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(L, F, M, Mpad, D, BS, omega, phase, Z, lengthscales, basis, zbasis);
  gpp::count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

int gpp_rollout_pathwise_fwd(int S, int ldS, int H, int L, int F, int Mpad, int D, int Dx, int num_active, const int* active_dims,
                             const double* basis, const double* zbasis, const double* w, const double* v, const double* amp,
                             const double* variance, const double* inv_lengthscales, const double* mean_const,
                             int Mp, const double* policy_Zs, const double* policy_inv_lengthscales, const double* policy_alpha,
                             double squash_scale, double squash_shift, const double* cost_target, const double* cost_W,
                             const double* x0, double* loss, double* x_final, double* traj, void* stream) {
  GPP_NVTX_RANGE();
  return pathwise_fwd_impl(S, ldS, H, L, F, Mpad, D, Dx, num_active, active_dims, basis, zbasis, w, v, amp, variance, inv_lengthscales,
                           mean_const, Mp, policy_Zs, policy_inv_lengthscales, policy_alpha, squash_scale, squash_shift, cost_target,
                           cost_W, x0, loss, x_final, traj, nullptr, stream);
}

int gpp_rollout_pathwise_fwd_mixed(int S, int ldS, int H, int L, int F, int Mpad, int D, int Dx, int num_active, const int* active_dims,
                                   const double* basis, const double* zbasis, const float* w32, const double* v, const double* amp,
                                   const double* variance, const double* inv_lengthscales, const double* mean_const,
                                   int Mp, const double* policy_Zs, const double* policy_inv_lengthscales, const double* policy_alpha,
                                   double squash_scale, double squash_shift, const double* cost_target, const double* cost_W,
                                   const double* x0, double* loss, double* x_final, double* traj, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(w32, GPP_ERR_NULL, "gpp_rollout_pathwise_fwd_mixed: null FP32 weights");
  return pathwise_fwd_impl(S, ldS, H, L, F, Mpad, D, Dx, num_active, active_dims, basis, zbasis, nullptr, v, amp, variance,
                           inv_lengthscales, mean_const, Mp, policy_Zs, policy_inv_lengthscales, policy_alpha, squash_scale,
                           squash_shift, cost_target, cost_W, x0, loss, x_final, traj, nullptr, stream, w32);
}

int gpp_dtype_supported(const char* entry_point, gpp_dtype dtype) {
  if (dtype == GPP_F64) return 1;
  if (dtype == GPP_MIXED_F32_WEIGHTS && entry_point)
    return std::strcmp(entry_point, "gpp_rollout_pathwise_fwd") == 0 || std::strcmp(entry_point, "gpp_rollout_pathwise_fwd_typed") == 0;
  return 0;
}

int gpp_rollout_pathwise_fwd_typed(gpp_dtype dtype, int S, int ldS, int H, int L, int F, int Mpad, int D, int Dx, int num_active,
                                   const int* active_dims, const double* basis, const double* zbasis, const void* w, const double* v,
                                   const double* amp, const double* variance, const double* inv_lengthscales, const double* mean_const,
                                   int Mp, const double* policy_Zs, const double* policy_inv_lengthscales, const double* policy_alpha,
                                   double squash_scale, double squash_shift, const double* cost_target, const double* cost_W,
                                   const double* x0, double* loss, double* x_final, double* traj, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(dtype == GPP_F64 || dtype == GPP_MIXED_F32_WEIGHTS, GPP_ERR_UNSUPPORTED, "gpp_rollout_pathwise_fwd_typed: unknown dtype %d",
              (int)dtype);
  const bool mixed = dtype == GPP_MIXED_F32_WEIGHTS;
  return pathwise_fwd_impl(S, ldS, H, L, F, Mpad, D, Dx, num_active, active_dims, basis, zbasis,
                           mixed ? nullptr : static_cast<const double*>(w), v, amp, variance, inv_lengthscales, mean_const, Mp, policy_Zs,
                           policy_inv_lengthscales, policy_alpha, squash_scale, squash_shift, cost_target, cost_W, x0, loss, x_final,
                           traj, nullptr, stream, mixed ? static_cast<const float*>(w) : nullptr);
}

int gpp_pathwise_weights_f32(long long count, const double* w, float* w32, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(count >= 0 && (count == 0 || (w && w32)), GPP_ERR_NULL, "gpp_pathwise_weights_f32: null argument");
  if (count == 0) return GPP_OK;
  const int threads = 256;
  const long long blocks = std::min<long long>((count + threads - 1) / threads, 148LL * 16);
  gpp::k_weights_f32<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(count, w, w32);
  gpp::count_launch();
  GPP_CUDA_OK(cudaGetLastError());
  return GPP_OK;
}

int gpp_rollout_pathwise_fwd_grad(int S, int ldS, int H, int L, int F, int Mpad, int D, int Dx, int num_active, const int* active_dims,
                                  const double* basis, const double* zbasis, const double* w, const double* v, const double* amp,
                                  const double* variance, const double* inv_lengthscales, const double* mean_const,
                                  int Mp, const double* policy_Zs, const double* policy_inv_lengthscales, const double* policy_alpha,
                                  double squash_scale, double squash_shift, const double* cost_target, const double* cost_W,
                                  const double* x0, double* loss, double* x_final, double* traj, double* jac, void* stream) {
  GPP_NVTX_RANGE();
  GPP_REQUIRE(traj && jac, GPP_ERR_NULL, "gpp_rollout_pathwise_fwd_grad: traj and jac are required (they are what the backward reads)");
  return pathwise_fwd_impl(S, ldS, H, L, F, Mpad, D, Dx, num_active, active_dims, basis, zbasis, w, v, amp, variance, inv_lengthscales,
                           mean_const, Mp, policy_Zs, policy_inv_lengthscales, policy_alpha, squash_scale, squash_shift, cost_target,
                           cost_W, x0, loss, x_final, traj, jac, stream);
}

}  // extern "C"
