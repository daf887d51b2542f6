// Scalar FP64 building blocks shared by every kernel (host+device so they can be unit-tested with gcc).
//
//  * fast_exp   : branch-free double exp, <= 1 ulp-class error on [-745, 709]; 15 FP64-pipe ops
//                 (1 FMA magic-round, 1 ADD, 2 FMA Cody-Waite, 11 FMA Horner) + integer exponent insert.
//                 MUFU is FP32-only, so the FP64 exp is software on every path; this one drops libdevice's
//                 slow-path branch and special cases (inputs here are log-kernel-expectations: finite, <= log var^2).
//  * small dense linear algebra on D x D matrices held in registers (D is a template parameter).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define GPP_HD __host__ __device__ __forceinline__
#else
#define GPP_HD inline
#endif

namespace gpp {

GPP_HD int32_t lo_int(double x) {
#if defined(__CUDA_ARCH__)
  return __double2loint(x);
#else
  int64_t b; std::memcpy(&b, &x, 8); return (int32_t)(b & 0xffffffff);
#endif
}
GPP_HD int32_t hi_int(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  int64_t b; std::memcpy(&b, &x, 8); return (int32_t)(b >> 32);
#endif
}
GPP_HD double make_double(int32_t hi, int32_t lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  int64_t b = ((int64_t)hi << 32) | (uint32_t)lo; double x; std::memcpy(&x, &b, 8); return x;
#endif
}
GPP_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}

// exp(x) for x <= ~709.  x < -744 (incl. -inf) returns 0.
GPP_HD double fast_exp(double x) {
  const double LOG2E = 1.4426950408889634074;
  const double MAGIC = 6755399441055744.0;             // 1.5 * 2^52: low word of (t) holds rint(x*log2e)
  const double LN2_HI = 6.93147180369123816490e-01;    // fdlibm split of ln 2
  const double LN2_LO = 1.90821492927058770002e-10;
  double t = fma_(x, LOG2E, MAGIC);
  int32_t k = lo_int(t);
  double kd = t - MAGIC;
  double r = fma_(kd, -LN2_HI, x);
  r = fma_(kd, -LN2_LO, r);
  // degree-11 polynomial: 1 + r + r^2 Q(r), Q interpolated at Chebyshev nodes on [-ln2/2, ln2/2] in 60-digit
  // arithmetic (max relative error 1.6e-17 before rounding).
  double p = 0x1.af38a9b0ec855p-26;
  p = fma_(p, r, 0x1.289185613a3d6p-22);
  p = fma_(p, r, 0x1.71de0dae63bb3p-19);
  p = fma_(p, r, 0x1.a019b90d2ae7ap-16);
  p = fma_(p, r, 0x1.a01a01a7c41d5p-13);
  p = fma_(p, r, 0x1.6c16c1788bd90p-10);
  p = fma_(p, r, 0x1.11111111109b3p-7);
  p = fma_(p, r, 0x1.5555555553d63p-5);
  p = fma_(p, r, 0x1.5555555555556p-3);
  p = fma_(p, r, 0x1.0000000000001p-1);
  p = fma_(p, r, 1.0);
  p = fma_(p, r, 1.0);
  // scale by 2^k through the exponent field (integer pipe); clamp keeps the result a normal number
  int32_t kc = k < -1021 ? -1021 : (k > 1023 ? 1023 : k);
  double res = make_double(hi_int(p) + (kc << 20), lo_int(p));
  // hi word of -744.0 is 0xC0874000; anything more negative (as unsigned compare) underflows to zero
  return ((uint32_t)hi_int(x) > 0xC0874000u) ? 0.0 : res;
}

// K independent exps evaluated in lock-step (in place).
//  * A dependent DFMA issues 8 cycles after its producer while the FP64 pipe accepts a warp instruction every 2 (3 when
//    all three operands are distinct registers), so independent chains must be interleaved; ptxas only does that when
//    the chains share operands or the source is written this way (checked in SASS).
//  * The polynomial is split into even and odd halves, p(r) = E(r^2) + r O(r^2): two depth-5 Horner chains per value.
//  * Coefficients live in the constant bank on the device (DFMA takes a c[bank][offset] operand directly; as literals
//    ptxas re-materialises them into uniform registers inside the loop, ~13 extra issue slots per evaluation).
//  * Valid for x <= 709 (kernel expectations are bounded by the product of the kernel variances, so there is no
//    overflow path); x < -744 returns exactly 0.  The integer tail (clamp, exponent insert, select) is free: it
//    issues in the shadow of the FP64 pipe (measured: removing it does not change the loop time).
#if defined(__CUDA_ARCH__)
#define GPP_EXP_TABLE static __constant__ double
#else
#define GPP_EXP_TABLE static const double
#endif
GPP_EXP_TABLE kExpC[16] = {
    1.4426950408889634074,          // 0  log2(e)
    -6.93147180369123816490e-01,    // 1  -ln2 hi (fdlibm split)
    -1.90821492927058770002e-10,    // 2  -ln2 lo
    0x1.0000000000001p-1,           // 3  c2
    0x1.5555555555556p-3,           // 4  c3
    0x1.5555555553d63p-5,           // 5  c4
    0x1.11111111109b3p-7,           // 6  c5
    0x1.6c16c1788bd90p-10,          // 7  c6
    0x1.a01a01a7c41d5p-13,          // 8  c7
    0x1.a019b90d2ae7ap-16,          // 9  c8
    0x1.71de0dae63bb3p-19,          // 10 c9
    0x1.289185613a3d6p-22,          // 11 c10
    0x1.af38a9b0ec855p-26,          // 12 c11
    0.0, 0.0, 0.0};

template <int K>
GPP_HD void fast_exp_n(double (&x)[K]) {
  const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: the low word of t holds rint(x * log2 e)
  double t[K], r[K], p[K], s2[K], q[K];
#pragma unroll
  for (int k = 0; k < K; ++k) t[k] = fma_(x[k], kExpC[0], MAGIC);
#pragma unroll
  for (int k = 0; k < K; ++k) r[k] = t[k] - MAGIC;
#pragma unroll
  for (int k = 0; k < K; ++k) p[k] = fma_(r[k], kExpC[1], x[k]);
#pragma unroll
  for (int k = 0; k < K; ++k) r[k] = fma_(r[k], kExpC[2], p[k]);
#pragma unroll
  for (int k = 0; k < K; ++k) s2[k] = r[k] * r[k];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    p[k] = fma_(kExpC[11], s2[k], kExpC[9]);    // even: c10, c8
    q[k] = fma_(kExpC[12], s2[k], kExpC[10]);   // odd : c11, c9
  }
#define GPP_EXP_STEP(ce, co)                      \
  _Pragma("unroll") for (int k = 0; k < K; ++k) { \
    p[k] = fma_(p[k], s2[k], ce);                 \
    q[k] = fma_(q[k], s2[k], co);                 \
  }
  GPP_EXP_STEP(kExpC[7], kExpC[8])   // c6, c7
  GPP_EXP_STEP(kExpC[5], kExpC[6])   // c4, c5
  GPP_EXP_STEP(kExpC[3], kExpC[4])   // c2, c3
  GPP_EXP_STEP(1.0, 1.0)             // c0, c1
#undef GPP_EXP_STEP
#pragma unroll
  for (int k = 0; k < K; ++k) p[k] = fma_(q[k], r[k], p[k]);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    int32_t e = lo_int(t[k]);
    e = e < -1021 ? -1021 : e;
    double res = make_double(hi_int(p[k]) + (e << 20), lo_int(p[k]));
    x[k] = ((uint32_t)hi_int(x[k]) > 0xC0874000u) ? 0.0 : res;   // x < -744 (incl. -inf, huge negatives): exactly 0
  }
}

// Table-driven exp: exp(x) = 2^k T[j] P(r),  n = rint(x 64/ln 2), k = n >> 6, j = n & 63, r = x - n ln2/64 (|r| <= 0.0055),
// P = degree-5 Taylor (truncation 3.5e-17).  10 FP64-pipe ops (vs 16 for fast_exp_n) + one table load.
// `tab` points at a copy of kExp2Tab replicated GPP_EXP_TAB_REP times with entry j of replica c at tab[j*REP + c]; on the
// device the kernels keep it in shared memory and pass tab + (lane & 15): the 64-bit loads of a half-warp then hit 16
// distinct bank pairs whatever the j's are (conflict-free).  Same domain contract as fast_exp_n.
#define GPP_EXP_TAB_REP 16
GPP_EXP_TABLE kExp2Tab[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};
GPP_EXP_TABLE kExpT[8] = {
    0x1.71547652b82fep+6,     // 0  64 log2(e)
    -0x1.62e42fee00000p-7,    // 1  -ln2/64 hi (fdlibm split / 64: 32 significant bits, n*hi exact for |n| < 2^20)
    -0x1.a39ef35793c76p-39,   // 2  -ln2/64 lo
    0.5,                      // 3  1/2!
    0x1.5555555555555p-3,     // 4  1/3!
    0x1.5555555555555p-5,     // 5  1/4!
    0x1.1111111111111p-7,     // 6  1/5!
    0.0};

// core: p = T_j exp(r) (unscaled, in [0.99, 2.02)), n = rint(x 64 log2 e)
template <int K>
GPP_HD void exp_tab_core(const double (&x)[K], const double* tab, double (&p)[K], int32_t (&n)[K]) {
  const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: the low word of t holds n = rint(x * 64 log2 e)
  double t[K], r[K], s2[K], tj[K];
#pragma unroll
  for (int k = 0; k < K; ++k) t[k] = fma_(x[k], kExpT[0], MAGIC);
#pragma unroll
  for (int k = 0; k < K; ++k) { n[k] = lo_int(t[k]); tj[k] = tab[(n[k] & 63) * GPP_EXP_TAB_REP]; }
#pragma unroll
  for (int k = 0; k < K; ++k) r[k] = t[k] - MAGIC;
#pragma unroll
  for (int k = 0; k < K; ++k) p[k] = fma_(r[k], kExpT[1], x[k]);
#pragma unroll
  for (int k = 0; k < K; ++k) r[k] = fma_(r[k], kExpT[2], p[k]);
#pragma unroll
  for (int k = 0; k < K; ++k) s2[k] = r[k] * r[k];
#pragma unroll
  for (int k = 0; k < K; ++k) p[k] = fma_(r[k], kExpT[6], kExpT[5]);
#pragma unroll
  for (int k = 0; k < K; ++k) p[k] = fma_(p[k], r[k], kExpT[4]);
#pragma unroll
  for (int k = 0; k < K; ++k) p[k] = fma_(p[k], r[k], kExpT[3]);
#pragma unroll
  for (int k = 0; k < K; ++k) p[k] = fma_(p[k], s2[k], r[k]);      // exp(r) - 1
#pragma unroll
  for (int k = 0; k < K; ++k) p[k] = fma_(tj[k], p[k], tj[k]);     // T_j exp(r)
}

// scale by 2^(n >> 6) through the exponent field (no clamp: callers only use the result when x >= -707, for which
// n >> 6 >= -1020 and p >= 0.99, so the result is a normal number)
GPP_HD double exp_tab_scale(double p, int32_t n) { return make_double(hi_int(p) + ((n >> 6) << 20), lo_int(p)); }
// x >= -707 ?  (hi word of -707.0 is 0xC0861800; the unsigned compare also rejects -inf, NaN-with-sign and huge negatives)
GPP_HD bool exp_tab_in_range(double x) { return !((uint32_t)hi_int(x) > 0xC0861800u); }

template <int K>
GPP_HD void fast_exp_tab_n(double (&x)[K], const double* tab) {
  double p[K];
  int32_t n[K];
  exp_tab_core<K>(x, tab, p, n);
#pragma unroll
  for (int k = 0; k < K; ++k) x[k] = exp_tab_in_range(x[k]) ? exp_tab_scale(p[k], n[k]) : 0.0;   // x < -707: exactly 0
}

// 256-entry variant for the contraction kernel: n = rint(x 256/ln 2), |r| <= ln2/512 = 0.00135, degree-4 Taylor (truncation
// 3.8e-17), and a single full-precision ln2/256 constant in the reduction: r = fma(n, -ln2/256, x) is exact up to one rounding,
// the constant's own error shifts r by |n| 2.4e-19 <= |x| 9e-17, i.e. by less than the rounding error x already carries.
// 8 FP64-pipe ops.  exp_tab256_core returns p = T_j exp(r) in [1, 2) and n; the caller scales by 2^(n >> 8).
GPP_EXP_TABLE kExp2Tab256[256] = {
#include "exp2_tab256.inc"
};
GPP_EXP_TABLE kExpT256[8] = {
    0x1.71547652b82fep+8,     // 0  256 log2(e)
    -0x1.62e42fefa39efp-9,    // 1  -ln2/256
    0.5,                      // 2  1/2!
    0x1.5555555555555p-3,     // 3  1/3!
    0x1.5555555555555p-5,     // 4  1/4!
    0.0, 0.0, 0.0};

// host-testable scalar reference of the 256-entry algorithm (same operation order as the device loop)
GPP_HD double exp_tab256_ref(double x) {
  const double MAGIC = 6755399441055744.0;
  double t = fma_(x, kExpT256[0], MAGIC);
  int32_t n = lo_int(t);
  double nd = t - MAGIC;
  double r = fma_(nd, kExpT256[1], x);
  double q = fma_(r, kExpT256[4], kExpT256[3]);
  q = fma_(q, r, kExpT256[2]);
  q = fma_(q, r, 1.0);
  double em1 = q * r;
  double tj = kExp2Tab256[n & 255];
  double p = fma_(tj, em1, tj);
  return make_double(hi_int(p) + ((n >> 8) << 20), lo_int(p));
}

// ---------------------------------------------------------------------------------------------
// D x D helpers, row-major in registers.  All loops fully unrolled (D is compile-time).
// ---------------------------------------------------------------------------------------------
template <int D>
struct Mat {
  double a[D * D];
  GPP_HD double& operator()(int i, int j) { return a[i * D + j]; }
  GPP_HD const double& operator()(int i, int j) const { return a[i * D + j]; }
};

// In-place lower Cholesky A = L L^T (upper triangle left untouched/garbage). Returns false if not PD.
template <int D>
GPP_HD bool cholesky(Mat<D>& A) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    double d = A(j, j);
#pragma unroll
    for (int k = 0; k < j; ++k) d = fma_(-A(j, k), A(j, k), d);
    ok = ok && (d > 0.0);
    double s = sqrt(d);
    A(j, j) = s;
    double inv = 1.0 / s;
#pragma unroll
    for (int i = j + 1; i < D; ++i) {
      double v = A(i, j);
#pragma unroll
      for (int k = 0; k < j; ++k) v = fma_(-A(i, k), A(j, k), v);
      A(i, j) = v * inv;
    }
  }
  return ok;
}

// Li = L^{-1} (lower), from lower-triangular L.
template <int D>
GPP_HD void tri_inverse(const Mat<D>& L, Mat<D>& Li) {
#pragma unroll
  for (int j = 0; j < D; ++j) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      if (i < j) { Li(i, j) = 0.0; continue; }
      double v = (i == j) ? 1.0 : 0.0;
#pragma unroll
      for (int k = j; k < i; ++k) v = fma_(-L(i, k), Li(k, j), v);
      Li(i, j) = v / L(i, i);
    }
  }
}

// G = Li^T Li  (= (L L^T)^{-1}), full symmetric.
template <int D>
GPP_HD void gram_inverse(const Mat<D>& Li, Mat<D>& G) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) {
      double v = 0.0;
#pragma unroll
      for (int k = j; k < D; ++k) v = fma_(Li(k, i), Li(k, j), v);
      G(i, j) = v;
      G(j, i) = v;
    }
}

}  // namespace gpp
