/* TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the GSL special functions the
 * reference calls (GSL is an un-vendored, version-unpinned dependency of the
 * reference, CMakeLists.txt:13, and is not installed in this image).
 * Call sites: common/fastbessel.cc:47, distribution/expsin2distribution.cc:15,
 * distribution/besselproductdistribution.hh:54,133-134, .cc:9-10,
 * common/auxilliary.cc:163.
 * Implemented with C++17 std::cyl_bessel_i (agrees with scipy.special.i0/i0e/ive
 * to <= 3e-15 relative, see tests/test_oracle_cpu.py) plus the Hankel asymptotic
 * series where exp(-x)*I_n(x) would overflow. */
#ifndef MLMCPI_ORACLE_SHIM_GSL_SF_BESSEL_H
#define MLMCPI_ORACLE_SHIM_GSL_SF_BESSEL_H
#include <cmath>

static inline double shim_bessel_In_scaled_asym(const int n, const double x) {
  /* DLMF 10.40.1: e^{-x} I_n(x) ~ (2 pi x)^{-1/2} sum_k (-1)^k a_k(n) / x^k */
  const double mu = 4.0 * n * n;
  double term = 1.0, sum = 1.0;
  for (int k = 1; k < 60; ++k) {
    const double f = (mu - (2.0 * k - 1.0) * (2.0 * k - 1.0)) / (8.0 * k * x);
    term *= -f;
    sum += term;
    if (std::fabs(term) < 1e-17 * std::fabs(sum))
      break;
  }
  return sum / std::sqrt(2.0 * M_PI * x);
}
static inline double gsl_sf_bessel_I0(const double x) {
  return std::cyl_bessel_i(0.0, std::fabs(x));
}
static inline double gsl_sf_bessel_I0_scaled(const double x) {
  const double ax = std::fabs(x);
  if (ax > 600.0)
    return shim_bessel_In_scaled_asym(0, ax);
  return std::exp(-ax) * std::cyl_bessel_i(0.0, ax);
}
static inline double gsl_sf_bessel_In_scaled(const int n, const double x) {
  const double ax = std::fabs(x);
  double r;
  if (ax > 600.0)
    r = shim_bessel_In_scaled_asym(n, ax);
  else
    r = std::exp(-ax) * std::cyl_bessel_i((double)(n < 0 ? -n : n), ax);
  if (x < 0.0 && (n & 1))
    r = -r;
  return r;
}
#endif
