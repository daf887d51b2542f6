// Bivariate normal upper-tail probability (Genz) on the device:  bvnu(h, k, r) = P(x > h, y > k), corr(x, y) = r.
// Replaces upstream gpflow_pilco/utils/bvn.py:88-232 (a TensorFlow transcription of Genz's BVNU; thresholds r = 0.925,
// hk = asr = -100), which the multi-dimensional NormalCDF rule calls (gpflow_pilco/moment_matching/bijectors.py:59-63).
// The Gauss-Legendre order (6 / 12 / 20 points) is chosen per element from |r| as in Genz's original; upstream picks it from
// the largest |r| of the whole batch (bvn.py:221-228) — the two differ below 1e-15.
#pragma once
#include "common.cuh"

namespace gpp {

__device__ const double kBvnX6[3] = {0.9324695142031522, 0.6612093864662647, 0.2386191860831970};
__device__ const double kBvnW6[3] = {0.1713244923791705, 0.3607615730481384, 0.4679139345726904};
__device__ const double kBvnX12[6] = {0.9815606342467191, 0.9041172563704750, 0.7699026741943050,
                                      0.5873179542866171, 0.3678314989981802, 0.1252334085114692};
__device__ const double kBvnW12[6] = {0.04717533638651177, 0.1069393259953183, 0.1600783285433464,
                                      0.2031674267230659, 0.2334925365383547, 0.2491470458134029};
__device__ const double kBvnX20[10] = {0.9931285991850949, 0.9639719272779138, 0.9122344282513259, 0.8391169718222188, 0.7463319064601508,
                                       0.6360536807265150, 0.5108670019508271, 0.3737060887154196, 0.2277858511416451, 0.07652652113349733};
__device__ const double kBvnW20[10] = {0.01761400713915212, 0.04060142980038694, 0.06267204833410906, 0.08327674157670475, 0.1019301198172404,
                                       0.1181945319615184, 0.1316886384491766, 0.1420961093183821, 0.1491729864726037, 0.1527533871307259};

__device__ __forceinline__ double bvn_ndtr(double x) { return 0.5 * erfc(-0.70710678118654752440 * x); }

// finite h, k
__device__ double bvnu(double h, double k, double r) {
  if (r == 0.0) return bvn_ndtr(-h) * bvn_ndtr(-k);
  const double tp = 6.283185307179586477;
  const double ar = fabs(r);
  const double *X, *W;
  int ng;
  if (ar < 0.3) { X = kBvnX6; W = kBvnW6; ng = 3; }
  else if (ar < 0.75) { X = kBvnX12; W = kBvnW12; ng = 6; }
  else { X = kBvnX20; W = kBvnW20; ng = 10; }
  const double hk = h * k;
  double res;
  if (ar < 0.925) {
    const double hs = 0.5 * (h * h + k * k), asr = 0.5 * asin(r);
    double acc = 0.0;
    for (int i = 0; i < ng; ++i)
#pragma unroll
      for (int sg = -1; sg <= 1; sg += 2) {
        const double sn = sin(asr * (1.0 + sg * X[i]));
        acc = fma(W[i], exp((sn * hk - hs) / (1.0 - sn * sn)), acc);
      }
    res = acc * asr / tp + bvn_ndtr(-h) * bvn_ndtr(-k);
  } else {
    const double sgn = r > 0.0 ? 1.0 : -1.0;
    const double kk = k * sgn, hkk = hk * sgn;
    double part = 0.0;
    if (ar < 1.0) {
      const double as = 1.0 - r * r;
      double a = sqrt(as);
      const double bs = (h - kk) * (h - kk);
      const double asr = -0.5 * (bs / as + hkk);
      const double c = 0.125 * (4.0 - hkk), d = 0.0125 * (12.0 - hkk);
      double t = 0.0;
      if (asr > -100.0) t = a * exp(asr) * (1.0 - c * (bs - as) * (1.0 - d * bs) / 3.0 + c * d * as * as);
      if (hkk > -100.0) {
        const double b = sqrt(bs);
        const double sp = sqrt(tp) * bvn_ndtr(-b / a);
        t -= exp(-0.5 * hkk) * sp * b * (1.0 - c * bs * (1.0 - d * bs) / 3.0);
      }
      a *= 0.5;
      double acc = 0.0;
      for (int i = 0; i < ng; ++i)
#pragma unroll
        for (int sg = -1; sg <= 1; sg += 2) {
          const double x = a * (1.0 + sg * X[i]);
          const double xs = x * x;
          const double asr2 = -0.5 * (bs / xs + hkk);
          if (asr2 > -100.0) {
            const double sp = 1.0 + c * xs * (1.0 + 5.0 * d * xs);
            const double rs = sqrt(1.0 - xs);
            const double ep = exp(-0.5 * hkk * xs / ((1.0 + rs) * (1.0 + rs))) / rs;
            acc = fma(W[i], exp(asr2) * (sp - ep), acc);
          }
        }
      part = (a * acc - t) / tp;
    }
    if (r > 0.0) res = part + bvn_ndtr(-fmax(h, kk));
    else if (h >= kk) res = -part;
    else if (h < 0.0) res = bvn_ndtr(kk) - bvn_ndtr(h) - part;
    else res = bvn_ndtr(-h) - bvn_ndtr(-kk) - part;
  }
  return fmin(fmax(res, 0.0), 1.0);
}

// P(xl < x < xu, yl < y < yu)   (upstream bvn.py:67-85)
__device__ __forceinline__ double bvn_box(double xl, double xu, double yl, double yu, double r) {
  const double p = bvnu(xl, yl, r) - bvnu(xu, yl, r) - bvnu(xl, yu, r) + bvnu(xu, yu, r);
  return fmin(fmax(p, 0.0), 1.0);
}

}  // namespace gpp
