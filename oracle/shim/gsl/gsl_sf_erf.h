/* TEST INFRASTRUCTURE ONLY (oracle/): gsl_sf_erf -> std::erf.
 * Call sites: distribution/besselproductdistribution.hh:98-99. */
#ifndef MLMCPI_ORACLE_SHIM_GSL_SF_ERF_H
#define MLMCPI_ORACLE_SHIM_GSL_SF_ERF_H
#include <cmath>
static inline double gsl_sf_erf(const double x) { return std::erf(x); }
#endif
