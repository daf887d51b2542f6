// Small-dimension moment-matching rules evaluated by one thread per Gaussian state (dims <= GPP_SMALL_MAX):
// the glue between the heavy GP predict kernels inside a rollout step.  Generic in a scalar type S so that the
// same code runs on doubles (forward) and on forward-mode dual numbers (backward, dual.cuh).
//
//   sincos / encoder rule   upstream gpflow_pilco/moment_matching/maths.py:143-176, moment_matching/components.py:19-57
//   NormalCDF rule (1-D)    upstream gpflow_pilco/moment_matching/bijectors.py:37-69  (Owen's T by 32-pt Gauss-Legendre)
//   Shift / Scale           upstream gpflow_pilco/moment_matching/maths.py:47-78
//   GaussianObjective       upstream gpflow_pilco/components.py:30-37
//   Euler moment update     upstream gpflow_pilco/dynamics/solvers.py:121-129
#pragma once
#include "common.cuh"

#define GPP_SMALL_MAX 8

namespace gpp {

struct EncoderSpec {
  int Dx;                 // state dimension
  int na;                 // number of sincos-encoded (active) dims
  int active[4];          // their indices
  __host__ __device__ int De() const { return Dx + na; }
  __host__ __device__ int nb() const { return Dx - na; }
  __host__ __device__ bool is_active(int i) const {
    for (int k = 0; k < na; ++k) if (active[k] == i) return true;
    return false;
  }
  int inact[GPP_SMALL_MAX];   // table of the inactive dims, filled by finish()
  int ready;                  // finish() was called
  // j-th inactive dim in increasing order (upstream components.py:58-67 builds tuple(set(...)), i.e. sorted)
  __host__ __device__ int inactive(int j) const {
    if (ready) return inact[j];
    int c = 0;
    for (int i = 0; i < Dx; ++i) if (!is_active(i)) { if (c == j) return i; ++c; }
    return -1;
  }
  // call on the host once Dx, na and active[] are set: the rules look the inactive dims up many times per state
  void finish() {
    ready = 0;
    for (int j = 0; j < Dx - na; ++j) inact[j] = inactive(j);
    ready = 1;
  }
};

// 32-point Gauss-Legendre on [-1,1], positive half (nodes symmetric)
__device__ const double kGL32_X[16] = {0.048307665687738324, 0.14447196158279649, 0.23928736225213706, 0.33186860228212767,
                                       0.42135127613063533, 0.50689990893222936, 0.5877157572407623, 0.66304426693021523,
                                       0.73218211874028971, 0.79448379596794239, 0.84936761373256997, 0.89632115576605209,
                                       0.93490607593773967, 0.96476225558750639, 0.98561151154526838, 0.99726386184948157};
__device__ const double kGL32_W[16] = {0.096540088514727659, 0.095638720079274708, 0.093844399080804511, 0.09117387869576378,
                                       0.087652093004403783, 0.083311924226946707, 0.078193895787070228, 0.072345794108848338,
                                       0.065822222776361683, 0.058684093478535565, 0.050998059262376091, 0.042835898022226836,
                                       0.034273862913021765, 0.025392065309262024, 0.016274394730905743, 0.0070186100094705058};

// scalar helpers overloaded for double (dual.cuh adds the dual-number overloads)
__device__ __forceinline__ double s_exp(double x) { return exp(x); }
__device__ __forceinline__ double s_sin(double x) { return sin(x); }
__device__ __forceinline__ double s_cos(double x) { return cos(x); }
__device__ __forceinline__ double s_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ double s_erfc(double x) { return erfc(x); }
__device__ __forceinline__ double s_value(double x) { return x; }

// Owen's T(h, a), 0 < a <= 1:  (1/2pi) int_0^a exp(-h^2 (1+x^2)/2) / (1+x^2) dx   (abs. error < 1e-16 for |h| <= 12)
template <typename S>
__device__ S owens_t(S h, S a) {
  S acc = S(0.0);
  S half = a * 0.5;
  S mh2 = h * h * (-0.5);
  for (int k = 0; k < 16; ++k) {
#pragma unroll
    for (int sgn = -1; sgn <= 1; sgn += 2) {
      S x = half * (1.0 + sgn * kGL32_X[k]);
      S d = x * x + 1.0;
      acc = acc + s_exp(mh2 * d) / d * kGL32_W[k];
    }
  }
  return acc * half * 0.15915494309189535;   // 1/(2 pi)
}

template <typename S>
__device__ S ndtr(S x) { return s_erfc(x * (-0.70710678118654752440)) * 0.5; }

// ---- encoder: e = [sin(a), cos(a), b]; returns mean me[De], covariance See[De][De], Cxe = Cov(x, e) [Dx][De] -------------
template <typename S>
__device__ void mm_encoder(const EncoderSpec& es, const S* m, const S* Sx /*[Dx][Dx]*/, S* me, S* See, S* Cxe) {
  const int Dx = es.Dx, na = es.na, De = es.De(), nb = es.nb();
  // One exp, one sin and one cos per active dimension (plus one exp per off-diagonal pair); the moments of the pairs follow from
  // the angle-addition formulas and exp(-(v_i + v_j)/2 -+ c_ij) = ev_i ev_j exp(-+c_ij)  (this runs on a single thread of the
  // rollout's pre stage, where every transcendental is ~0.3 us of serial latency).
  S s1[4], c1[4], sn[4], cs[4], ev[4];
  for (int k = 0; k < na; ++k) {
    int i = es.active[k];
    ev[k] = s_exp(Sx[i * Dx + i] * (-0.5));
    sn[k] = s_sin(m[i]);
    cs[k] = s_cos(m[i]);
    s1[k] = ev[k] * sn[k];
    c1[k] = ev[k] * cs[k];
    me[k] = s1[k];
    me[na + k] = c1[k];
  }
  for (int j = 0; j < nb; ++j) me[2 * na + j] = m[es.inactive(j)];
  // covariance of [sin, cos] block: uncentred moments (maths.py:147-170) minus outer product of the means
  for (int k = 0; k < na; ++k)
    for (int l = 0; l < na; ++l) {
      int i = es.active[k], j = es.active[l];
      S A, B;                                   // exp(-(v_i + v_j + 2 c_ij)/2), exp(-(v_i + v_j - 2 c_ij)/2)
      if (k == l) {
        S e2 = ev[k] * ev[k];
        A = e2 * e2;
        B = S(1.0);
      } else {
        S E = ev[k] * ev[l];
        S t = s_exp((Sx[i * Dx + j] + Sx[j * Dx + i]) * (-0.5));
        A = E * t;
        B = E / t;
      }
      S cc0 = cs[k] * cs[l], ss0 = sn[k] * sn[l];
      S ca = A * (cc0 - ss0), cd = B * (cc0 + ss0);          // A cos(m_i + m_j), B cos(m_i - m_j)
      S sx_cx = sn[k] * cs[l], cx_sx = cs[k] * sn[l];
      S ss = (cd - ca) * 0.5, cc = (cd + ca) * 0.5;
      S sc = (sx_cx * (B + A) - cx_sx * (B - A)) * 0.5;      // E[sin x_k cos x_l]
      See[k * De + l] = ss - s1[k] * s1[l];
      See[(na + k) * De + na + l] = cc - c1[k] * c1[l];
      See[k * De + na + l] = sc - s1[k] * c1[l];
      See[(na + l) * De + k] = sc - s1[k] * c1[l];
    }
  // Cov(x, y) = Sxa diag-pre-inverted cross (maths.py:173: [diag(c1), diag(-s1)])
  for (int i = 0; i < Dx; ++i) {
    for (int k = 0; k < na; ++k) {
      S sxa = Sx[i * Dx + es.active[k]];
      Cxe[i * De + k] = sxa * c1[k];
      Cxe[i * De + na + k] = sxa * s1[k] * (-1.0);
    }
    for (int j = 0; j < nb; ++j) Cxe[i * De + 2 * na + j] = Sx[i * Dx + es.inactive(j)];
  }
  // blocks involving the inactive dims (components.py:44-52)
  for (int j = 0; j < nb; ++j) {
    int bj = es.inactive(j);
    for (int k = 0; k < 2 * na; ++k) {
      See[(2 * na + j) * De + k] = Cxe[bj * De + k];
      See[k * De + 2 * na + j] = Cxe[bj * De + k];
    }
    for (int l = 0; l < nb; ++l) See[(2 * na + j) * De + 2 * na + l] = Sx[bj * Dx + es.inactive(l)];
  }
}

// ---- the same rule, one output entry at a time (for the rollout's pre stage, where the CTA's threads share the work) ------------
// per active dimension k: exp(-v_k / 2), sin(m_k), cos(m_k)
template <typename S>
struct EncTrig {
  S ev[4], sn[4], cs[4];
};

template <typename S, class MeanAt, class CovAt>
__device__ __forceinline__ void enc_trig_one(const EncoderSpec& es, int k, MeanAt m, CovAt Sx, EncTrig<S>& t) {
  const int i = es.active[k];
  t.ev[k] = s_exp(Sx(i, i) * (-0.5));
  t.sn[k] = s_sin(m(i));
  t.cs[k] = s_cos(m(i));
}

// position a of the encoded vector e = [sin (na), cos (na), inactive (nb)]: kind 0 / 1 / 2 and the index inside its group
__device__ __forceinline__ void enc_slot(const EncoderSpec& es, int a, int& kind, int& idx) {
  if (a < es.na) { kind = 0; idx = a; }
  else if (a < 2 * es.na) { kind = 1; idx = a - es.na; }
  else { kind = 2; idx = es.inactive(a - 2 * es.na); }      // idx = state dimension
}

template <typename S, class MeanAt>
__device__ __forceinline__ S enc_mean_at(const EncoderSpec& es, int a, MeanAt m, const EncTrig<S>& t) {
  int kind, idx;
  enc_slot(es, a, kind, idx);
  if (kind == 0) return t.ev[idx] * t.sn[idx];
  if (kind == 1) return t.ev[idx] * t.cs[idx];
  return m(idx);
}

// Cxe[i][a] = Cov(x_i, e_a), pre-inverted form (maths.py:173)
template <typename S, class CovAt>
__device__ __forceinline__ S enc_cross_at(const EncoderSpec& es, int i, int a, CovAt Sx, const EncTrig<S>& t) {
  int kind, idx;
  enc_slot(es, a, kind, idx);
  if (kind == 0) return Sx(i, es.active[idx]) * (t.ev[idx] * t.cs[idx]);
  if (kind == 1) return Sx(i, es.active[idx]) * (t.ev[idx] * t.sn[idx]) * (-1.0);
  return Sx(i, idx);
}

// See[a][b]
template <typename S, class CovAt>
__device__ __forceinline__ S enc_cov_at(const EncoderSpec& es, int a, int b, CovAt Sx, const EncTrig<S>& t) {
  int ka, ia, kb, ib;
  enc_slot(es, a, ka, ia);
  enc_slot(es, b, kb, ib);
  if (ka == 2 && kb == 2) return Sx(ia, ib);
  if (ka == 2) return enc_cross_at<S>(es, ia, b, Sx, t);      // blocks involving the inactive dims (components.py:44-52)
  if (kb == 2) return enc_cross_at<S>(es, ib, a, Sx, t);
  // both trigonometric: uncentred moments (maths.py:147-170) minus the outer product of the means.  (sin_k, cos_l) and
  // (cos_l, sin_k) share one value: order the pair as (k = the sine's dimension, l = the cosine's) when the kinds differ
  const int k = (ka == 1 && kb == 0) ? ib : ia, l = (ka == 1 && kb == 0) ? ia : ib;
  const int i = es.active[k], j = es.active[l];
  S A, B;                                     // exp(-(v_i + v_j + 2 c_ij)/2), exp(-(v_i + v_j - 2 c_ij)/2)
  if (k == l) {
    S e2 = t.ev[k] * t.ev[k];
    A = e2 * e2;
    B = S(1.0);
  } else {
    S E = t.ev[k] * t.ev[l];
    S x = s_exp((Sx(i, j) + Sx(j, i)) * (-0.5));
    A = E * x;
    B = E / x;
  }
  const S s1k = t.ev[k] * t.sn[k], c1k = t.ev[k] * t.cs[k], s1l = t.ev[l] * t.sn[l], c1l = t.ev[l] * t.cs[l];
  if (ka != kb) {                             // E[sin x_k cos x_l] - E sin E cos
    S sx_cx = t.sn[k] * t.cs[l], cx_sx = t.cs[k] * t.sn[l];
    return (sx_cx * (B + A) - cx_sx * (B - A)) * 0.5 - s1k * c1l;
  }
  S cc0 = t.cs[k] * t.cs[l], ss0 = t.sn[k] * t.sn[l];
  S ca = A * (cc0 - ss0), cd = B * (cc0 + ss0);            // A cos(m_i + m_j), B cos(m_i - m_j)
  if (ka == 0) return (cd - ca) * 0.5 - s1k * s1l;
  return (cd + ca) * 0.5 - c1k * c1l;
}

// ---- reverse mode of the encoder rule (closed form; replaces one dual-number evaluation of the rule per input direction on the
// rollout's critical path).  Given the adjoints of (me, See, Cxe) — every matrix entry an independent output, as mm_encoder writes
// them — returns the adjoints of m[Dx] and of every entry of Sx[Dx][Dx] (entries treated as independent inputs; the caller
// symmetrises).  All pointers may be shared memory.
// Shared by the threads of a CTA (all of them must call it; >= Dx*Dx + Dx threads): the linear parts are gathered
// one input entry per thread, the small trigonometric block stays on thread 0.  `tr` is shared scratch of 8 * 4 doubles.
// m_bar / Sx_bar are OVERWRITTEN.  Ends with a group barrier (common.cuh group_sync: the 128 threads of warps 0..3).
__device__ inline void mm_encoder_bwd_cta(const EncoderSpec& es, const double* m, const double* Sx, const double* me_bar,
                                          const double* See_bar, const double* Cxe_bar, double* m_bar, double* Sx_bar, double* tr) {
  const int Dx = es.Dx, na = es.na, De = es.De(), nb = es.nb(), tid = threadIdx.x;
  double *ev = tr, *sn = tr + 4, *cs = tr + 8, *s1 = tr + 12, *c1 = tr + 16, *s1b = tr + 20, *c1b = tr + 24;
  auto pos_inactive = [&](int i) {              // position of state dim i among the inactive dims, or -1
    int ji = -1;
    for (int j = 0; j < nb; ++j) ji = es.inactive(j) == i ? j : ji;
    return ji;
  };
  // adjoints of Cxe[i][k] (cosine-weighted) and Cxe[i][na + k] (sine-weighted) including the See blocks that copy them
  auto bar_c = [&](int i, int ji, int k) {
    double v = Cxe_bar[i * De + k];
    if (ji >= 0) v += See_bar[(2 * na + ji) * De + k] + See_bar[k * De + 2 * na + ji];
    return v;
  };
  auto bar_s = [&](int i, int ji, int k) {
    double v = Cxe_bar[i * De + na + k];
    if (ji >= 0) v += See_bar[(2 * na + ji) * De + na + k] + See_bar[(na + k) * De + 2 * na + ji];
    return v;
  };
  if (tid < na) {
    const int i = es.active[tid];
    double s, c;
    const double e = exp(-0.5 * Sx[i * Dx + i]);
    sincos(m[i], &s, &c);
    ev[tid] = e; sn[tid] = s; cs[tid] = c; s1[tid] = e * s; c1[tid] = e * c;
  }
  group_sync();
  if (tid < Dx * Dx) {                          // linear part of d/dSx[i][j]
    const int i = tid / Dx, j = tid % Dx, ji = pos_inactive(i);
    double v = 0.0;
    int kj = -1;
    for (int k = 0; k < na; ++k) kj = es.active[k] == j ? k : kj;
    if (kj >= 0) {
      v = bar_c(i, ji, kj) * c1[kj] - bar_s(i, ji, kj) * s1[kj];        // Cxe[i][k] = Sx[i][a_k] c1_k,  Cxe[i][na+k] = -Sx[i][a_k] s1_k
    } else {
      const int jj = pos_inactive(j);
      v = Cxe_bar[i * De + 2 * na + jj];                                   // Cxe[i][2na+jj] = Sx[i][b_jj]
      if (ji >= 0) v += See_bar[(2 * na + ji) * De + 2 * na + jj];         // See[2na+ji][2na+jj] = Sx[b_ji][b_jj]
    }
    Sx_bar[tid] = v;
  } else if (tid < Dx * Dx + Dx) {              // d/dm[i] of the inactive dims; adjoints of s1_k, c1_k from me and Cxe
    const int i = tid - Dx * Dx, ji = pos_inactive(i);
    m_bar[i] = ji >= 0 ? me_bar[2 * na + ji] : 0.0;
    int k = -1;
    for (int q = 0; q < na; ++q) k = es.active[q] == i ? q : k;
    if (k >= 0) {
      double sb = me_bar[k], cb = me_bar[na + k];
      for (int r = 0; r < Dx; ++r) {
        const int jr = pos_inactive(r);
        const double sxa = Sx[r * Dx + i];
        cb += bar_c(r, jr, k) * sxa;
        sb -= bar_s(r, jr, k) * sxa;
      }
      s1b[k] = sb;
      c1b[k] = cb;
    }
  }
  group_sync();
  if (tid == 0) {
    double evb[4] = {0.0, 0.0, 0.0, 0.0}, snb[4] = {0.0, 0.0, 0.0, 0.0}, csb[4] = {0.0, 0.0, 0.0, 0.0};
    // trigonometric block of See, ordered pairs (k, l)
    for (int k = 0; k < na; ++k)
      for (int l = 0; l < na; ++l) {
        const int i = es.active[k], j = es.active[l];
        double A, B, x = 1.0;
        if (k == l) {
          const double e2 = ev[k] * ev[k];
          A = e2 * e2;
          B = 1.0;
        } else {
          x = exp(-0.5 * (Sx[i * Dx + j] + Sx[j * Dx + i]));
          A = ev[k] * ev[l] * x;
          B = ev[k] * ev[l] / x;
        }
        const double P = cs[k] * cs[l], Q = sn[k] * sn[l], U = sn[k] * cs[l], V = cs[k] * sn[l];
        const double bs = See_bar[k * De + l], bc = See_bar[(na + k) * De + na + l];
        const double bx = See_bar[k * De + na + l] + See_bar[(na + l) * De + k];
        // ss = (B (P + Q) - A (P - Q)) / 2 - s1_k s1_l;  cc = (B (P + Q) + A (P - Q)) / 2 - c1_k c1_l;
        // sc = (U (B + A) - V (B - A)) / 2 - s1_k c1_l
        const double Bb = 0.5 * ((bs + bc) * (P + Q) + bx * (U - V));
        const double Ab = 0.5 * ((bc - bs) * (P - Q) + bx * (U + V));
        const double Pb = 0.5 * (bs * (B - A) + bc * (B + A));
        const double Qb = 0.5 * (bs * (B + A) + bc * (B - A));
        const double Ub = 0.5 * bx * (B + A), Vb = -0.5 * bx * (B - A);
        s1b[k] -= bs * s1[l] + bx * c1[l];
        s1b[l] -= bs * s1[k];
        c1b[k] -= bc * c1[l];
        c1b[l] -= bc * c1[k] + bx * s1[k];
        csb[k] += Pb * cs[l] + Vb * sn[l];
        csb[l] += Pb * cs[k] + Ub * sn[k];
        snb[k] += Qb * sn[l] + Ub * cs[l];
        snb[l] += Qb * sn[k] + Vb * cs[k];
        if (k == l) {
          evb[k] += Ab * 4.0 * ev[k] * ev[k] * ev[k];
        } else {
          evb[k] += (Ab * x + Bb / x) * ev[l];
          evb[l] += (Ab * x + Bb / x) * ev[k];
          const double xb = (Ab - Bb / (x * x)) * ev[k] * ev[l];
          Sx_bar[i * Dx + j] -= 0.5 * x * xb;
          Sx_bar[j * Dx + i] -= 0.5 * x * xb;
        }
      }
    for (int k = 0; k < na; ++k) {
      const int i = es.active[k];
      evb[k] += s1b[k] * sn[k] + c1b[k] * cs[k];
      snb[k] += s1b[k] * ev[k];
      csb[k] += c1b[k] * ev[k];
      m_bar[i] += snb[k] * cs[k] - csb[k] * sn[k];
      Sx_bar[i * Dx + i] -= 0.5 * ev[k] * evb[k];
    }
  }
  group_sync();
}

// Owen's T evaluated by the 32 lanes of a warp (one Gauss-Legendre node each); every lane returns the sum
__device__ __forceinline__ double owens_t_warp(double h, double a) {
  const int lane = threadIdx.x & 31, k = lane >> 1;
  const double half = 0.5 * a;
  const double x = half * (1.0 + ((lane & 1) ? kGL32_X[k] : -kGL32_X[k]));
  const double d = x * x + 1.0;
  double v = exp(-0.5 * h * h * d) / d * kGL32_W[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v * half * 0.15915494309189535;
}

// T(h, a) as a scalar of type S when its VALUE t0 is already known: double -> t0; dual number -> value t0 with the tangent from
// the closed-form partials dT/dh = -phi(h) erf(a h / sqrt 2) / 2, dT/da = exp(-h^2 (1 + a^2) / 2) / (2 pi (1 + a^2))  
__device__ __forceinline__ double owens_t_given(double, double, double t0) { return t0; }

// ---- squashing link  u = scale * (Phi(f) + shift)  applied to a 1-D Gaussian f ~ N(mf, vf) -----------------------------------
// returns mean, variance of u and d = Cov(f,f)^-1 Cov(f,u)  (chain of bijectors.py:37-69, maths.py:47-78)
// USE_T0: Owen's T value supplied by the caller (owens_t_warp) instead of the serial 32-node quadrature
template <typename S, bool USE_T0 = false>
__device__ void mm_squash_1d(S mf, S vf, double scale, double shift, S& mu, S& vu, S& gain, double t0 = 0.0) {
  S vw = vf + 1.0;
  S isq = S(1.0) / s_sqrt(vw);
  S h = isq * mf;
  S y1 = ndtr(h);
  S a = S(1.0) / s_sqrt(vf * 2.0 + 1.0);
  S y2 = y1 - (USE_T0 ? owens_t_given(h, a, t0) : owens_t(h, a)) * 2.0;       // E[Phi^2]   (bijectors.py:57-58)
  S phi = s_exp(h * h * (-0.5)) * 0.39894228040143267794;
  mu = (y1 + shift) * scale;
  vu = (y2 - y1 * y1) * (scale * scale);
  gain = isq * phi * scale;
}

// (h, a) of the squashing rule, for callers that evaluate Owen's T cooperatively first
__device__ __forceinline__ void squash_owens_args(double mf, double vf, double& h, double& a) {
  h = mf / sqrt(vf + 1.0);
  a = 1.0 / sqrt(2.0 * vf + 1.0);
}

// ---- general small linear algebra (runtime n <= GPP_SMALL_MAX) ----------------------------------------------------------------
// LU with partial pivoting (pivot chosen on values); A overwritten; returns det.  inv may be null.
template <typename S>
__device__ S lu_inverse(int n, S* A, S* inv) {
  int perm[GPP_SMALL_MAX];
  S det = S(1.0);
  for (int i = 0; i < n; ++i) perm[i] = i;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = fabs(s_value(A[c * n + c]));
    for (int r = c + 1; r < n; ++r) {
      double v = fabs(s_value(A[r * n + c]));
      if (v > best) { best = v; piv = r; }
    }
    if (piv != c) {
      for (int k = 0; k < n; ++k) { S t = A[c * n + k]; A[c * n + k] = A[piv * n + k]; A[piv * n + k] = t; }
      int t = perm[c]; perm[c] = perm[piv]; perm[piv] = t;
      det = det * (-1.0);
    }
    det = det * A[c * n + c];
    S ip = S(1.0) / A[c * n + c];
    for (int r = c + 1; r < n; ++r) {
      A[r * n + c] = A[r * n + c] * ip;
      for (int k = c + 1; k < n; ++k) A[r * n + k] = A[r * n + k] - A[r * n + c] * A[c * n + k];
    }
  }
  if (inv) {
    for (int col = 0; col < n; ++col) {
      S y[GPP_SMALL_MAX];
      for (int i = 0; i < n; ++i) {
        S v = S(perm[i] == col ? 1.0 : 0.0);
        for (int k = 0; k < i; ++k) v = v - A[i * n + k] * y[k];
        y[i] = v;
      }
      for (int i = n - 1; i >= 0; --i) {
        S v = y[i];
        for (int k = i + 1; k < n; ++k) v = v - A[i * n + k] * inv[k * n + col];
        inv[i * n + col] = v / A[i * n + i];
      }
    }
  }
  return det;
}

// ---- expected saturating cost  E[-exp(-1/2 (e-t)^T W (e-t))]  for e ~ N(me, See)   (components.py:30-37) ------------------------
template <typename S>
__device__ S expected_cost(int n, const S* me, const S* See, const double* target, const double* W) {
  S IpSW[GPP_SMALL_MAX * GPP_SMALL_MAX], inv[GPP_SMALL_MAX * GPP_SMALL_MAX], err[GPP_SMALL_MAX];
  for (int i = 0; i < n; ++i) {
    err[i] = me[i] - target[i];
    for (int j = 0; j < n; ++j) {
      S v = S(i == j ? 1.0 : 0.0);
      for (int k = 0; k < n; ++k) v = v + See[i * n + k] * W[k * n + j];
      IpSW[i * n + j] = v;
    }
  }
  S det = lu_inverse<S>(n, IpSW, inv);
  // dist2 = err^T W inv(I + S W) err
  S d2 = S(0.0);
  for (int i = 0; i < n; ++i) {
    S row = S(0.0);
    for (int j = 0; j < n; ++j) {
      S wij = S(0.0);
      for (int k = 0; k < n; ++k) wij = wij + inv[k * n + j] * W[i * n + k];
      row = row + wij * err[j];
    }
    d2 = d2 + row * err[i];
  }
  return s_exp(d2 * (-0.5)) / s_sqrt(det) * (-1.0);
}

// sample-path cost  -exp(-1/2 (e-t)^T W (e-t))   (components.py:39-41)
__device__ __forceinline__ double sample_cost(int n, const double* e, const double* target, const double* W) {
  double d2 = 0.0;
  for (int i = 0; i < n; ++i) {
    double row = 0.0;
    for (int j = 0; j < n; ++j) row = fma(W[i * n + j], e[j] - target[j], row);
    d2 = fma(row, e[i] - target[i], d2);
  }
  return -exp(-0.5 * d2);
}

}  // namespace gpp
