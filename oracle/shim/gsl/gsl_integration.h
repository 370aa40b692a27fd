/* TEST INFRASTRUCTURE ONLY (oracle/): the slice of gsl_integration used by
 * common/auxilliary.cc:134-193 (QAWO with a sin/cos weight over one period and
 * QAG).  The integrands there are entire functions on [-pi, pi]; a composite
 * 32-panel x 20-point Gauss-Legendre rule reaches ~1e-15 absolute for them, so
 * no adaptivity is needed.  Weight for QAWO: sin(omega x) or cos(omega x). */
#ifndef MLMCPI_ORACLE_SHIM_GSL_INTEGRATION_H
#define MLMCPI_ORACLE_SHIM_GSL_INTEGRATION_H
#include "gsl_math.h"
#include <cmath>
#include <cstddef>
#include <cstdlib>
enum gsl_integration_qawo_enum { GSL_INTEG_COSINE, GSL_INTEG_SINE };
struct gsl_integration_workspace {
  size_t limit;
};
struct gsl_integration_qawo_table {
  double omega, L;
  enum gsl_integration_qawo_enum sine;
};
static inline gsl_integration_workspace *
gsl_integration_workspace_alloc(const size_t n) {
  gsl_integration_workspace *w =
      (gsl_integration_workspace *)std::malloc(sizeof(gsl_integration_workspace));
  w->limit = n;
  return w;
}
static inline void gsl_integration_workspace_free(gsl_integration_workspace *w) {
  std::free(w);
}
static inline gsl_integration_qawo_table *
gsl_integration_qawo_table_alloc(double omega, double L,
                                 enum gsl_integration_qawo_enum sine, size_t) {
  gsl_integration_qawo_table *t = (gsl_integration_qawo_table *)std::malloc(
      sizeof(gsl_integration_qawo_table));
  t->omega = omega;
  t->L = L;
  t->sine = sine;
  return t;
}
static inline int
gsl_integration_qawo_table_set(gsl_integration_qawo_table *t, double omega,
                               double L, enum gsl_integration_qawo_enum sine) {
  t->omega = omega;
  t->L = L;
  t->sine = sine;
  return 0;
}
static inline void
gsl_integration_qawo_table_free(gsl_integration_qawo_table *t) {
  std::free(t);
}
/* 20-point Gauss-Legendre nodes/weights on [-1,1] (positive half) */
static const double shim_gl20_x[10] = {
    0.0765265211334973337546404, 0.2277858511416450780804962,
    0.3737060887154195606725482, 0.5108670019508270980043641,
    0.6360536807265150254528367, 0.7463319064601507926143051,
    0.8391169718222188233945291, 0.9122344282513259058677524,
    0.9639719272779137912676661, 0.9931285991850949247861224};
static const double shim_gl20_w[10] = {
    0.1527533871307258506980843, 0.1491729864726037467878287,
    0.1420961093183820513292983, 0.1316886384491766268984945,
    0.1181945319615184173123774, 0.1019301198172404350367501,
    0.0832767415767047487247581, 0.0626720483341090635695065,
    0.0406014298003869413310400, 0.0176140071391521183118620};
static inline double shim_integrate(const gsl_function *f, double a, double b,
                                    int weight, double omega) {
  const int n_panel = 64;
  const double h = (b - a) / n_panel;
  double total = 0.0;
  for (int p = 0; p < n_panel; ++p) {
    const double c = a + (p + 0.5) * h;
    double s = 0.0;
    for (int k = 0; k < 10; ++k) {
      for (int sg = -1; sg <= 1; sg += 2) {
        const double x = c + sg * 0.5 * h * shim_gl20_x[k];
        double v = GSL_FN_EVAL(f, x);
        if (weight == 1)
          v *= std::cos(omega * x);
        else if (weight == 2)
          v *= std::sin(omega * x);
        s += shim_gl20_w[k] * v;
      }
    }
    total += 0.5 * h * s;
  }
  return total;
}
static inline int gsl_integration_qawo(gsl_function *f, const double a,
                                       const double, const double, const size_t,
                                       gsl_integration_workspace *,
                                       gsl_integration_qawo_table *t,
                                       double *result, double *abserr) {
  *result = shim_integrate(f, a, a + t->L,
                           t->sine == GSL_INTEG_SINE ? 2 : 1, t->omega);
  *abserr = 0.0;
  return 0;
}
static inline int gsl_integration_qag(const gsl_function *f, double a, double b,
                                      double, double, size_t, int,
                                      gsl_integration_workspace *,
                                      double *result, double *abserr) {
  *result = shim_integrate(f, a, b, 0, 0.0);
  *abserr = 0.0;
  return 0;
}
#endif
