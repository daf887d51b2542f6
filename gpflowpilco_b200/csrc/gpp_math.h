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

// Table-driven exp for the Psi2 kernels: exp(x) = 2^k T[j] P(r) with a 2^(j/256) table kept in shared memory, replicated
// GPP_EXP_TAB_REP times (entry j of replica c at tab[j*REP + c]; a lane uses replica lane & (REP-1), so the 64-bit loads of a
// half-warp hit 16 distinct bank pairs whatever the j's are).  Device loop: mma_exp.cuh (exp_tab_contract).
#define GPP_EXP_TAB_REP 16
// n = rint(x 256/ln 2), |r| <= ln2/512 = 0.00135, degree-4 Taylor (truncation
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
  // D reciprocals up front (independent, so they pipeline) instead of one dependent FP64 division per entry: these factorisations
  // sit on single-thread critical paths of the rollout stages, where a division is ~100 cycles of latency
  double rd[D];
#pragma unroll
  for (int i = 0; i < D; ++i) rd[i] = 1.0 / L(i, i);
#pragma unroll
  for (int j = 0; j < D; ++j) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      if (i < j) { Li(i, j) = 0.0; continue; }
      double v = (i == j) ? 1.0 : 0.0;
#pragma unroll
      for (int k = j; k < i; ++k) v = fma_(-L(i, k), Li(k, j), v);
      Li(i, j) = v * rd[i];
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
