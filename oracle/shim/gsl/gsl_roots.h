/* TEST INFRASTRUCTURE ONLY (oracle/): bisection root bracketing solver with the
 * gsl_root_fsolver interface used by
 * action/qft/quenchedschwingerrenormalisation.cc:20-62. */
#ifndef MLMCPI_ORACLE_SHIM_GSL_ROOTS_H
#define MLMCPI_ORACLE_SHIM_GSL_ROOTS_H
#include "gsl_errno.h"
#include "gsl_math.h"
#include <cmath>
#include <cstdlib>
struct gsl_root_fsolver_type {
  int dummy;
};
struct gsl_root_fsolver {
  gsl_function *f;
  double x_lo, x_hi, root, f_lo, f_hi;
};
static const gsl_root_fsolver_type shim_bisection_type = {0};
static const gsl_root_fsolver_type *const gsl_root_fsolver_bisection =
    &shim_bisection_type;
static inline gsl_root_fsolver *
gsl_root_fsolver_alloc(const gsl_root_fsolver_type *) {
  return (gsl_root_fsolver *)std::calloc(1, sizeof(gsl_root_fsolver));
}
static inline void gsl_root_fsolver_free(gsl_root_fsolver *s) { std::free(s); }
static inline int gsl_root_fsolver_set(gsl_root_fsolver *s, gsl_function *f,
                                       double x_lo, double x_hi) {
  s->f = f;
  s->x_lo = x_lo;
  s->x_hi = x_hi;
  s->root = 0.5 * (x_lo + x_hi);
  s->f_lo = GSL_FN_EVAL(f, x_lo);
  s->f_hi = GSL_FN_EVAL(f, x_hi);
  return GSL_SUCCESS;
}
static inline int gsl_root_fsolver_iterate(gsl_root_fsolver *s) {
  /* same update rule as GSL's roots/bisection.c */
  if (s->f_lo == 0.0) {
    s->root = s->x_lo;
    s->x_hi = s->x_lo;
    return GSL_SUCCESS;
  }
  if (s->f_hi == 0.0) {
    s->root = s->x_hi;
    s->x_lo = s->x_hi;
    return GSL_SUCCESS;
  }
  const double x_bis = 0.5 * (s->x_lo + s->x_hi);
  const double f_bis = GSL_FN_EVAL(s->f, x_bis);
  if (f_bis == 0.0) {
    s->root = x_bis;
    s->x_lo = x_bis;
    s->x_hi = x_bis;
    return GSL_SUCCESS;
  }
  if ((s->f_lo > 0.0 && f_bis < 0.0) || (s->f_lo < 0.0 && f_bis > 0.0)) {
    s->root = 0.5 * (s->x_lo + x_bis);
    s->x_hi = x_bis;
    s->f_hi = f_bis;
  } else {
    s->root = 0.5 * (x_bis + s->x_hi);
    s->x_lo = x_bis;
    s->f_lo = f_bis;
  }
  return GSL_SUCCESS;
}
static inline double gsl_root_fsolver_root(const gsl_root_fsolver *s) {
  return s->root;
}
static inline double gsl_root_fsolver_x_lower(const gsl_root_fsolver *s) {
  return s->x_lo;
}
static inline double gsl_root_fsolver_x_upper(const gsl_root_fsolver *s) {
  return s->x_hi;
}
static inline int gsl_root_test_interval(double x_lo, double x_hi, double epsabs,
                                         double epsrel) {
  const double abs_lo = std::fabs(x_lo), abs_hi = std::fabs(x_hi);
  double min_abs;
  if ((x_lo > 0.0 && x_hi > 0.0) || (x_lo < 0.0 && x_hi < 0.0))
    min_abs = abs_lo < abs_hi ? abs_lo : abs_hi;
  else
    min_abs = 0.0;
  const double tol = epsabs + epsrel * min_abs;
  return (std::fabs(x_hi - x_lo) < tol) ? GSL_SUCCESS : GSL_CONTINUE;
}
#endif
