// Forward-mode dual numbers (value + one tangent) for the small moment-matching rules of mm_small.cuh: the backward of
// the rollout glue (encoder rule, squashing link, expected cost) evaluates directional derivatives with the SAME templated
// code that computes the forward values, one direction per thread.  The heavy Psi-statistic sums have closed-form adjoints
// instead (mm_predict_bwd.cu, rollout_mm_bwd.cu).
#pragma once
#include "mm_small.cuh"

namespace gpp {

struct Dual {
  double v, d;
  __device__ __forceinline__ Dual() : v(0.0), d(0.0) {}
  __device__ __forceinline__ Dual(double x) : v(x), d(0.0) {}
  __device__ __forceinline__ Dual(double x, double t) : v(x), d(t) {}
};

__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
__device__ __forceinline__ Dual operator+(Dual a, double b) { return Dual(a.v + b, a.d); }
__device__ __forceinline__ Dual operator+(double a, Dual b) { return Dual(a + b.v, b.d); }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
__device__ __forceinline__ Dual operator-(Dual a, double b) { return Dual(a.v - b, a.d); }
__device__ __forceinline__ Dual operator-(double a, Dual b) { return Dual(a - b.v, -b.d); }
__device__ __forceinline__ Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, fma(a.v, b.d, a.d * b.v)); }
__device__ __forceinline__ Dual operator*(Dual a, double b) { return Dual(a.v * b, a.d * b); }
__device__ __forceinline__ Dual operator*(double a, Dual b) { return Dual(a * b.v, a * b.d); }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
  double q = a.v / b.v;
  return Dual(q, (a.d - q * b.d) / b.v);
}
__device__ __forceinline__ Dual operator/(Dual a, double b) { return Dual(a.v / b, a.d / b); }
__device__ __forceinline__ Dual operator/(double a, Dual b) {
  double q = a / b.v;
  return Dual(q, -q * b.d / b.v);
}

__device__ __forceinline__ Dual s_exp(Dual x) { double e = exp(x.v); return Dual(e, e * x.d); }
__device__ __forceinline__ Dual s_sin(Dual x) { double s, c; sincos(x.v, &s, &c); return Dual(s, c * x.d); }
__device__ __forceinline__ Dual s_cos(Dual x) { double s, c; sincos(x.v, &s, &c); return Dual(c, -s * x.d); }
__device__ __forceinline__ Dual s_sqrt(Dual x) { double r = sqrt(x.v); return Dual(r, 0.5 * x.d / r); }
__device__ __forceinline__ Dual s_erfc(Dual x) { return Dual(erfc(x.v), -1.1283791670955125739 * exp(-x.v * x.v) * x.d); }
__device__ __forceinline__ double s_value(Dual x) { return x.v; }

// Owen's T with known value: tangent from the closed-form partials (mm_small.cuh)
__device__ __forceinline__ Dual owens_t_given(Dual h, Dual a, double t0) {
  const double dh = -0.5 * 0.39894228040143267794 * exp(-0.5 * h.v * h.v) * erf(a.v * h.v * 0.70710678118654752440);
  const double da = exp(-0.5 * h.v * h.v * (1.0 + a.v * a.v)) * 0.15915494309189535 / (1.0 + a.v * a.v);
  return Dual(t0, dh * h.d + da * a.d);
}

}  // namespace gpp
