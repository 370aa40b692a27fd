// analytic.cu -- host-side analytic results and the non-perturbative coupling matching.
//
// Setup-time scalars only; nothing here runs per sample.  The reference
// evaluates these with GSL (QAWO / QAG quadrature, bisection root solver, scaled Bessel
// functions); GSL is not a dependency of this library, so the integrals are done with a
// composite Gauss-Legendre rule and the root with a plain bisection.  Reference citations
// relative to /root/reference/src.
#include <cmath>
#include <mutex>
#include <vector>

#include "../../include/mlmcpi.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace {

// ---- composite Gauss-Legendre rule on [-pi, pi]: NP panels of NG nodes -------------------
constexpr int NG = 16, NP = 512, NMAX = 20;

struct Rule {
  std::vector<double> phi, w;              // nodes and weights on [-pi, pi]
  std::vector<double> cosphi;              // cos(phi) - 1
  std::vector<double> f1[NMAX], f2[NMAX];  // phi sin(n phi), phi^2 cos(n phi)
  std::vector<double> cn[NMAX];            // cos(n phi)
  Rule() {
    // nodes of P_NG by Newton iteration on the three-term recurrence
    double x[NG], wt[NG];
    for (int i = 0; i < NG; ++i) {
      double z = std::cos(M_PI * (i + 0.75) / (NG + 0.5)), pp = 1.0;
      for (int it = 0; it < 100; ++it) {
        double p0 = 1.0, p1 = z;
        for (int k = 2; k <= NG; ++k) {
          const double p2 = ((2.0 * k - 1.0) * z * p1 - (k - 1.0) * p0) / k;
          p0 = p1;
          p1 = p2;
        }
        pp = NG * (z * p1 - p0) / (z * z - 1.0);
        const double dz = p1 / pp;
        z -= dz;
        if (std::fabs(dz) < 1e-16)
          break;
      }
      x[i] = z;
      wt[i] = 2.0 / ((1.0 - z * z) * pp * pp);
    }
    const double h = 2.0 * M_PI / NP;
    for (int p = 0; p < NP; ++p)
      for (int i = 0; i < NG; ++i) {
        phi.push_back(-M_PI + h * (p + 0.5 * (x[i] + 1.0)));
        w.push_back(0.5 * h * wt[i]);
      }
    for (size_t k = 0; k < phi.size(); ++k)
      cosphi.push_back(std::cos(phi[k]) - 1.0);
    for (int n = 0; n < NMAX; ++n)
      for (size_t k = 0; k < phi.size(); ++k) {
        f1[n].push_back(phi[k] * std::sin(n * phi[k]));
        f2[n].push_back(phi[k] * phi[k] * std::cos(n * phi[k]));
        cn[n].push_back(std::cos(n * phi[k]));
      }
  }
};

const Rule &rule() {
  static Rule r;
  return r;
}

// common/auxilliary.cc:7-29
double sigma_hat(double xi, unsigned p) {
  if (p == 0)
    return 1.0;
  if (p & 1u)
    return 0.0;
  double num = 0.0, den = 1.0;
  for (unsigned m = 1; m < 100; ++m) {
    const double e = std::exp(-0.5 * xi * m * m);
    num += 2. * std::pow((double)m, (double)p) * e;
    den += 2. * e;
  }
  return num / den;
}

// common/auxilliary.cc:99-194: e^{-x} I_n(x) and the two phi-weighted integrals
//   dIn  = -1/(4 pi^2) int phi   sin(n phi) e^{x (cos phi - 1)} dphi
//   ddIn =  1/(8 pi^3) int phi^2 cos(n phi) e^{x (cos phi - 1)} dphi      over [-pi, pi]
void compute_In(double x, double In[NMAX], double dIn[NMAX], double ddIn[NMAX]) {
  const Rule &r = rule();
  const size_t N = r.phi.size();
  std::vector<double> e(N);
  for (size_t k = 0; k < N; ++k)
    e[k] = r.w[k] * std::exp(x * r.cosphi[k]);
  for (int n = 0; n < NMAX; ++n) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (size_t k = 0; k < N; ++k) {
      s0 += e[k] * r.cn[n][k];
      s1 += e[k] * r.f1[n][k];
      s2 += e[k] * r.f2[n][k];
    }
    // I_n: the integral representation has absolute accuracy only; where I_n is tiny
    // (small x, large n) use the relatively accurate library function
    In[n] = (x < 500.0) ? std::exp(-x) * std::cyl_bessel_i((double)n, x) : s0 / (2.0 * M_PI);
    dIn[n] = -s1 / (4. * M_PI * M_PI);
    ddIn[n] = s2 / (8. * M_PI * M_PI * M_PI);
  }
}

// common/auxilliary.cc:44-82
double Phi_chit(double beta, double n_plaq) {
  double In[NMAX], dIn[NMAX], ddIn[NMAX];
  compute_In(beta, In, dIn, ddIn);
  double weight[NMAX], weight_sum = 0.0;
  for (int n = 0; n < NMAX; ++n) {
    weight[n] = (1 + (n > 0)) * std::pow(In[n] / In[0], n_plaq);
    weight_sum += weight[n];
  }
  double phi = 0.0;
  for (int n = 0; n < NMAX; ++n) {
    if (weight[n] == 0.0)
      continue; // (its bracket may be 0/0 when I_n underflows)
    phi += beta * weight[n] / weight_sum *
           (ddIn[n] / In[n] - (n_plaq - 1) * (dIn[n] * dIn[n]) / (In[n] * In[n]));
  }
  return phi;
}

// common/auxilliary.cc:85-97
double Phi_chit_perturbative(double beta, double n_plaq) {
  const double xi = n_plaq / beta, z = 1. / beta;
  const double s2 = sigma_hat(xi, 2), s4 = sigma_hat(xi, 4);
  const double lo = 1.0 - xi * s2;
  const double nlo = 0.5 - xi * s2 + 0.25 * xi * xi * (s4 - s2 * s2);
  return (lo + z * nlo) / (4. * M_PI * M_PI);
}

} // namespace

extern "C" {

double mlmcpi_sigma_hat(double xi, unsigned int p) { return sigma_hat(xi, p); }

/* qoi/qft/qoi2dsusceptibility.cc:30-50 */
double mlmcpi_schwinger_chit_analytical(double beta, unsigned int n_plaq) {
  if (!(beta > 0.0) || beta > 2000.0)
    return NAN; // the reference refuses beta > 2000 (auxilliary.cc:45-51)
  return n_plaq / beta * Phi_chit(beta, n_plaq);
}
double mlmcpi_schwinger_chit_perturbative(double beta, unsigned int n_plaq) {
  return n_plaq / beta * Phi_chit_perturbative(beta, n_plaq);
}
double mlmcpi_schwinger_var_chit_continuum(double beta, unsigned int n_plaq) {
  const double zeta = 4 * M_PI * M_PI * beta / n_plaq;
  const double s2 = sigma_hat(zeta, 2), s4 = sigma_hat(zeta, 4);
  return s4 - s2 * s2;
}

/* qm/rotoraction.cc:92-115; which = 0 exact, 1 perturbative, 2 continuum */
double mlmcpi_rotor_chit(double m0, double a_lat, double T_final, int which) {
  const double xi = T_final / m0;
  if (which == 0)
    return 1. / m0 * Phi_chit(m0 / a_lat, T_final / a_lat);
  const double s2 = sigma_hat(xi, 2);
  if (which == 2)
    return 1. / (4. * M_PI * M_PI * m0) * (1. - xi * s2);
  const double s4 = sigma_hat(xi, 4), z = a_lat / m0;
  return 1. / (4. * M_PI * M_PI * m0) *
         (1. - xi * s2 + (0.5 - xi * s2 + 0.25 * xi * xi * (s4 - s2 * s2)) * z);
}

/* common/auxilliary.cc:197-209 */
double mlmcpi_gff_phi_squared_analytical(double mass, int Mt_lat, int Mx_lat) {
  const double mu2 = mass * mass / ((double)Mt_lat * Mx_lat);
  double s = 0.0;
  for (int k1 = 0; k1 < Mt_lat; ++k1)
    for (int k2 = 0; k2 < Mx_lat; ++k2) {
      const double s1 = std::sin(M_PI * k1 / Mt_lat), s2 = std::sin(M_PI * k2 / Mx_lat);
      s += 1. / (4. * (s1 * s1 + s2 * s2) + mu2);
    }
  return s / ((double)Mt_lat * Mx_lat);
}

/* qm/harmonicoscillatoraction.cc:69-80 */
double mlmcpi_ho_xsquared_analytical(double m0, double mu2, double a_lat, int M_lat, int continuum) {
  if (continuum) {
    const double T = a_lat * M_lat, e = std::exp(-std::sqrt(mu2) * T);
    return 1. / (2. * m0 * std::sqrt(mu2)) * (1. + e) / (1. - e);
  }
  const double R = 1. + 0.5 * a_lat * a_lat * mu2 -
                   a_lat * std::sqrt(mu2) * std::sqrt(1. + 0.25 * a_lat * a_lat * mu2);
  const double RM = std::pow(R, (double)M_lat);
  return 1. / (2. * m0 * std::sqrt(mu2) * std::sqrt(1. + 0.25 * a_lat * a_lat * mu2)) * (1. + RM) /
         (1. - RM);
}

/* qft/quenchedschwingerrenormalisation.cc:7-64: beta_coarse = x beta with x the root of
 * chi_t(x beta, P / rho) - chi_t(beta, P) on [0.01, 2] (bisection, relative tolerance 1e-12,
 * at most 100 iterations); rho = 4 (coarsening both) or 2; fallback x = 1 / rho when the
 * bracket holds no sign change */
double mlmcpi_schwinger_betacoarse_nonperturbative(double beta, unsigned int n_plaq, int rho_refine) {
  const unsigned int n_coarse = n_plaq / rho_refine;
  const double target = mlmcpi_schwinger_chit_analytical(beta, n_plaq);
  auto f = [&](double x) { return mlmcpi_schwinger_chit_analytical(x * beta, n_coarse) - target; };
  double lo = 0.01, hi = 2.0;
  if (hi * beta > 2000.0)
    hi = 2000.0 / beta; // Phi_chit is only defined up to beta = 2000
  double f_lo = f(lo), f_hi = f(hi);
  if (!(f_lo * f_hi < 0.0))
    return beta / rho_refine;
  double x = 0.5 * (lo + hi);
  for (int k = 0; k < 100; ++k) {
    x = 0.5 * (lo + hi);
    const double fx = f(x);
    if ((fx < 0.0) == (f_lo < 0.0)) {
      lo = x;
      f_lo = fx;
    } else {
      hi = x;
    }
    if (std::fabs(hi - lo) < 1e-12 * std::fmin(std::fabs(lo), std::fabs(hi)))
      break;
  }
  return 0.5 * (lo + hi) * beta;
}

} // extern "C"
