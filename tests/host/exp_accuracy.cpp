// Host check of the scalar FP64 building blocks of the kernels (gpp_math.h is host+device): prints the worst relative error of
// exp_tab256_ref (the 8-op exp of the Psi2 loops) and fast_exp (polynomial exp of the small kernels) against long double expl.
#include <cmath>
#include <cstdio>
#include <random>

#include "../../gpflowpilco_b200/csrc/gpp_math.h"

int main() {
  std::mt19937_64 g(1);
  double worst_scaled = 0, worst_mid = 0, worst_poly = 0;
  for (int it = 0; it < 6000000; ++it) {
    const int c = it % 3;
    const double lo = c == 0 ? -700 : (c == 1 ? -40 : -2), hi = c == 0 ? -40 : (c == 1 ? 5 : 2);
    const double x = lo + std::uniform_real_distribution<double>(0, 1)(g) * (hi - lo);
    const long double ref = expl((long double)x);
    const double e1 = std::fabs((double)((gpp::exp_tab256_ref(x) - ref) / ref));
    const double e2 = std::fabs((double)((gpp::fast_exp(x) - ref) / ref));
    worst_scaled = std::fmax(worst_scaled, e1 / std::fmax(1.0, std::fabs(x)));   // error grows like |x| eps by design
    if (c) worst_mid = std::fmax(worst_mid, e1);
    worst_poly = std::fmax(worst_poly, e2);
  }
  std::printf("%.6e %.6e %.6e %.17g %.17g\n", worst_scaled, worst_mid, worst_poly, gpp::exp_tab256_ref(0.0), gpp::fast_exp(-800.0));
  return 0;
}
