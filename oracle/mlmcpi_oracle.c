/* TEST INFRASTRUCTURE ONLY (oracle/).  See mlmcpi_oracle.h.
 *
 * Plain-C CPU restatement of the sampler inner loop of
 * eikehmueller/mlmcpathintegral.  Each function cites the reference file:line
 * (relative to /root/reference/src) whose arithmetic it follows, in the same
 * order of operations, so that the deterministic functions agree with the
 * reference's own translation units (oracle/_ref) to the last bit wherever no
 * Bessel function is involved (compiled with -ffp-contract=off).
 *
 * Pinned by tests/test_oracle_cpu.py against tests/golden/ (recorded from
 * oracle/_ref by tools/make_golden.py) and, when oracle/_ref is present,
 * against the reference library directly on random inputs.
 */
#include "mlmcpi_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ====================================================================== RNG */

/* Philox4x32-10 (Salmon et al., SC'11); the counter-based generator the CUDA
 * kernels use in place of the reference's std::mt19937_64. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2],
                       uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

/* Stream convention shared with the product (include/mlmcpi.h, "Random
 * streams"): counter = (index, chain, draw_lo, stream<<24 | call number),
 * key = (seed_lo, seed_hi ^ draw_hi). */
void orc_rng_init(orc_rng *r, uint64_t seed, int stream, uint64_t draw,
                  uint32_t chain, uint32_t index) {
  r->c0 = index;
  r->c1 = chain;
  r->c2 = (uint32_t)draw;
  r->a = ((uint32_t)stream) << 24;
  r->k0 = (uint32_t)seed;
  r->k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(draw >> 32);
}

/* two uniforms in [0,1) with 53 random bits each */
void orc_rng_uniform2(orc_rng *r, double *u0, double *u1) {
  uint32_t ctr[4] = {r->c0, r->c1, r->c2, r->a};
  uint32_t key[2] = {r->k0, r->k1};
  uint32_t o[4];
  orc_philox4x32_10(ctr, key, o);
  r->a += 1;
  *u0 = ((double)(o[0] >> 5) * 67108864.0 + (double)(o[1] >> 6)) *
        (1.0 / 9007199254740992.0);
  *u1 = ((double)(o[2] >> 5) * 67108864.0 + (double)(o[3] >> 6)) *
        (1.0 / 9007199254740992.0);
}

/* two independent N(0,1) variates (Box-Muller) */
void orc_rng_normal2(orc_rng *r, double *z0, double *z1) {
  double u0, u1;
  orc_rng_uniform2(r, &u0, &u1);
  const double rad = sqrt(-2.0 * log(1.0 - u0));
  *z0 = rad * cos(2.0 * M_PI * u1);
  *z1 = rad * sin(2.0 * M_PI * u1);
}

/* ============================================================ scalar maths */

/* common/auxilliary.hh:42-44 */
double orc_mod_2pi(double x) {
  return x - 2. * M_PI * floor(0.5 * (x + M_PI) / M_PI);
}

/* I0(x): power series for |x| <= 20, Hankel asymptotic series (DLMF 10.40.1)
 * above.  Stands in for gsl_sf_bessel_I0 (GSL: un-vendored, version-unpinned
 * dependency of the reference, CMakeLists.txt:13); both branches are accurate to
 * a few ulp, checked against scipy.special.i0/i0e in tests/test_oracle_cpu.py. */
static double i0_series(double x) {
  const double q = 0.25 * x * x;
  double term = 1.0, sum = 1.0;
  for (int k = 1; k < 200; ++k) {
    term *= q / ((double)k * (double)k);
    sum += term;
    if (term < 1e-17 * sum)
      break;
  }
  return sum;
}
static double i0_scaled_asym(double x) {
  double term = 1.0, sum = 1.0;
  for (int k = 1; k < 60; ++k) {
    const double f = (2.0 * k - 1.0) * (2.0 * k - 1.0) / (8.0 * k * x);
    if (f >= 1.0)
      break;
    term *= f;
    sum += term;
    if (term < 1e-17 * sum)
      break;
  }
  return sum / sqrt(2.0 * M_PI * x);
}
double orc_bessel_I0(double x) {
  x = fabs(x);
  if (x <= 20.0)
    return i0_series(x);
  return exp(x) * i0_scaled_asym(x);
}
double orc_bessel_I0_scaled(double x) {
  x = fabs(x);
  if (x <= 20.0)
    return exp(-x) * i0_series(x);
  return i0_scaled_asym(x);
}

/* common/fastbessel.cc:7-49, coefficients common/fastbessel.hh:38-50 */
double orc_fast_bessel_I0_scaled(double z) {
  double a[8];
  a[0] = 1.0;
  for (int n = 1; n < 8; ++n)
    a[n] = 0.125 * (2.0 * n - 1.0) * (2.0 * n - 1.0) / n * a[n - 1];
  int kmax;
  if (z > 1100.)
    kmax = 4;
  else if (z > 400.)
    kmax = 5;
  else if (z > 200.)
    kmax = 6;
  else if (z > 100.)
    kmax = 7;
  else
    return orc_bessel_I0_scaled(z);
  const double z_inv = 1. / z;
  double p = a[kmax];
  for (int k = kmax - 1; k >= 0; --k)
    p = z_inv * p + a[k];
  return p / sqrt(2. * M_PI * z);
}

/* common/auxilliary.cc:7-29 */
static double Sigma_hat(const double xi, const unsigned int p) {
  if (p % 2 == 0) {
    if (p == 0)
      return 1.0;
    const unsigned int mmax = 100;
    double num = 0.0, denom = 1.0;
    for (unsigned int m = 1; m < mmax; ++m) {
      const double exp_factor = exp(-0.5 * xi * m * m);
      num += 2. * pow(m, p) * exp_factor;
      denom += 2. * exp_factor;
    }
    return num / denom;
  }
  return 0.0;
}
/* common/auxilliary.cc:32-43 */
static double log_factorial(unsigned int n) {
  double s = 0.0;
  for (unsigned int k = 2; k <= n; ++k)
    s += log(k);
  return s;
}
static double log_nCk(unsigned int n, unsigned int k) {
  return log_factorial(n) - log_factorial(k) - log_factorial(n - k);
}

/* ===================================================== lattice index maps */

/* lattice/lattice2d.hh:230-245 */
uint32_t orc_vertex_cart2lin(int Mt, int Mx, int rotated, int i, int j) {
  if (rotated) {
    const int Mt_half = Mt / 2, Mx_half = Mx / 2;
    const int i_shifted = ((i + Mt) - (i & 1)) / 2;
    const int j_shifted = ((j + Mx) - (j & 1)) / 2;
    const int offset = (Mt * Mx / 4) * (i & 1);
    return (uint32_t)(Mt_half * (j_shifted % Mx_half) + i_shifted % Mt_half +
                      offset);
  }
  return (uint32_t)(Mt * ((j + Mx) % Mx) + ((i + Mt) % Mt));
}
/* lattice/lattice2d.hh:255-268 */
void orc_vertex_lin2cart(int Mt, int Mx, int rotated, uint32_t ell, int *i,
                         int *j) {
  if (rotated) {
    const int Mt_half = Mt / 2;
    const int parity = ell / (Mt * Mx / 4);
    const uint32_t ell_half = ell - (Mt * Mx / 4) * parity;
    const int j_half = ell_half / Mt_half;
    *j = 2 * j_half + parity;
    *i = 2 * (ell_half - Mt_half * j_half) + parity;
  } else {
    *j = ell / Mt;
    *i = ell - Mt * (*j);
  }
}
/* lattice/lattice2d.hh:348-353 */
uint32_t orc_link_cart2lin(int Mt, int Mx, int i, int j, int mu) {
  return (uint32_t)(2 * Mt * ((j + Mx) % Mx) + 2 * ((i + Mt) % Mt) + mu);
}
/* lattice/lattice2d.hh:367-375 */
void orc_link_lin2cart(int Mt, int Mx, uint32_t ell, int *i, int *j, int *mu) {
  (void)Mx;
  *j = ell / (2 * Mt);
  const uint32_t r = ell - (2 * Mt) * (*j);
  *i = r >> 1;
  *mu = r & 1;
}
/* lattice/lattice2d.hh:196-202 */
int orc_n_vertices(int Mt, int Mx, int rotated) {
  return rotated ? Mt * Mx / 2 : Mt * Mx;
}
/* lattice/lattice2d.cc:135-155 */
void orc_neighbours(int Mt, int Mx, int rotated, uint32_t ell, uint32_t nb[8]) {
  static const int oi_r[8] = {+1, +1, -1, -1, +2, -2, 0, 0};
  static const int oj_r[8] = {+1, -1, +1, -1, 0, 0, +2, -2};
  static const int oi_u[8] = {+1, -1, 0, 0, +1, +1, -1, -1};
  static const int oj_u[8] = {0, 0, +1, -1, +1, -1, +1, -1};
  int i, j;
  orc_vertex_lin2cart(Mt, Mx, rotated, ell, &i, &j);
  for (int k = 0; k < 8; ++k) {
    if (rotated)
      nb[k] = orc_vertex_cart2lin(Mt, Mx, 1, i + oi_r[k], j + oj_r[k]);
    else
      nb[k] = orc_vertex_cart2lin(Mt, Mx, 0, i + oi_u[k], j + oj_u[k]);
  }
}

/* coarsening factors of a level: lattice/lattice2d.cc:20-60 */
static int coarsen_factors(int Mt, int Mx, int ctype, int level, int *rho_t,
                           int *rho_x) {
  const int rotated = (ctype == ORC_COARSEN_ROTATE) && (level % 2);
  int allowed = 1;
  switch (ctype) {
  case ORC_COARSEN_BOTH:
    *rho_t = 2;
    *rho_x = 2;
    break;
  case ORC_COARSEN_TEMPORAL:
    *rho_t = 2;
    *rho_x = 1;
    break;
  case ORC_COARSEN_SPATIAL:
    *rho_t = 1;
    *rho_x = 2;
    break;
  case ORC_COARSEN_ALTERNATE:
    if (level % 2 == 0) {
      *rho_t = 2;
      *rho_x = 1;
    } else {
      *rho_t = 1;
      *rho_x = 2;
    }
    break;
  case ORC_COARSEN_ROTATE:
    if (rotated) {
      if ((Mt % 2) || (Mx % 2))
        allowed = 0;
      *rho_t = 2;
      *rho_x = 2;
    } else {
      *rho_t = 1;
      *rho_x = 1;
    }
    break;
  default:
    return 0;
  }
  return allowed;
}

/* lattice/lattice2d.cc:61-81 */
int orc_coarse_shape(int Mt, int Mx, int ctype, int level, int *Mt_c, int *Mx_c,
                     int *rot_c) {
  int rho_t = 1, rho_x = 1;
  int allowed = coarsen_factors(Mt, Mx, ctype, level, &rho_t, &rho_x);
  int Mt_coarse = Mt, Mx_coarse = Mx;
  if (rho_t > 1) {
    if (Mt % rho_t)
      allowed = 0;
    Mt_coarse = Mt / rho_t;
  }
  if (rho_x > 1) {
    if (Mx % rho_x)
      allowed = 0;
    Mx_coarse = Mx / rho_x;
  }
  allowed = allowed && (Mt_coarse > 1) && (Mx_coarse > 1);
  *Mt_c = Mt_coarse;
  *Mx_c = Mx_coarse;
  *rot_c = (ctype == ORC_COARSEN_ROTATE) && ((level + 1) % 2);
  return allowed;
}

static int cmp_u32(const void *a, const void *b) {
  const uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
  return (x > y) - (x < y);
}

/* lattice/lattice2d.cc:82-130.  coarse[] and fineonly[] sorted ascending;
 * map_vals[k] is the coarse-lattice index of fine vertex coarse[k] (std::map
 * iterates in key order, i.e. the order of the sorted coarse list). */
int orc_coarsening_lists(int Mt, int Mx, int ctype, int level, uint32_t *coarse,
                         uint32_t *fineonly, uint32_t *map_vals, int *counts) {
  int rho_t = 1, rho_x = 1, Mt_c, Mx_c, rot_c;
  if (!orc_coarse_shape(Mt, Mx, ctype, level, &Mt_c, &Mx_c, &rot_c))
    return -1;
  coarsen_factors(Mt, Mx, ctype, level, &rho_t, &rho_x);
  const int rotated = (ctype == ORC_COARSEN_ROTATE) && (level % 2);
  int nc = 0, nf = 0;
  for (int i = 0; i < Mt; ++i)
    for (int j = 0; j < Mx; ++j) {
      if (ctype == ORC_COARSEN_ROTATE) {
        if (rotated) {
          if ((i + j) % 2 == 0) {
            const uint32_t ell = orc_vertex_cart2lin(Mt, Mx, 1, i, j);
            if ((i % 2 == 0) && (j % 2 == 0))
              coarse[nc++] = ell;
            else
              fineonly[nf++] = ell;
          }
        } else {
          const uint32_t ell = orc_vertex_cart2lin(Mt, Mx, 0, i, j);
          if ((i + j) % 2 == 0)
            coarse[nc++] = ell;
          else
            fineonly[nf++] = ell;
        }
      } else {
        const uint32_t ell = orc_vertex_cart2lin(Mt, Mx, 0, i, j);
        if ((i % rho_t == 0) && (j % rho_x == 0))
          coarse[nc++] = ell;
        else
          fineonly[nf++] = ell;
      }
    }
  qsort(coarse, nc, sizeof(uint32_t), cmp_u32);
  qsort(fineonly, nf, sizeof(uint32_t), cmp_u32);
  for (int k = 0; k < nc; ++k) {
    int i, j;
    orc_vertex_lin2cart(Mt, Mx, rotated, coarse[k], &i, &j);
    map_vals[k] = orc_vertex_cart2lin(Mt_c, Mx_c, rot_c, i / rho_t, j / rho_x);
  }
  counts[0] = nc;
  counts[1] = nf;
  return 0;
}

/* =========================================== deterministic hot-path pieces */

#define LNK(i, j, mu) orc_link_cart2lin(Mt, Mx, (i), (j), (mu))

int orc_sample_size(const orc_model *m) {
  switch (m->model) {
  case ORC_HO:
  case ORC_QUARTIC:
  case ORC_ROTOR:
    return m->M_lat; /* qm/qmaction.hh:93 */
  case ORC_SCHWINGER:
    return 2 * m->Mt_lat * m->Mx_lat; /* qft/quenchedschwingeraction.hh:137 */
  case ORC_GFF:
    return orc_n_vertices(m->Mt_lat, m->Mx_lat, m->rotated);
  }
  return 0;
}

/* plaquette angle; qft/quenchedschwingeraction.cc:14-17 */
static double plaq(const double *x, int Mt, int Mx, int i, int j) {
  return x[LNK(i, j, 0)] + x[LNK(i + 1, j, 1)] - x[LNK(i, j + 1, 0)] -
         x[LNK(i, j, 1)];
}

static double gff_nn_sum(const orc_model *m, const double *x, uint32_t ell) {
  uint32_t nb[8];
  orc_neighbours(m->Mt_lat, m->Mx_lat, m->rotated, ell, nb);
  double Delta = 0.0;
  for (int k = 0; k < 4; ++k)
    Delta += x[nb[k]];
  return Delta;
}

double orc_action(const orc_model *m, const double *x) {
  const int M = m->M_lat;
  const double a = m->a_lat;
  switch (m->model) {
  case ORC_HO: { /* qm/harmonicoscillatoraction.cc:8-18 */
    const double ainv2 = 1. / (a * a);
    double x_diff = x[0] - x[M - 1];
    double S = ainv2 * x_diff * x_diff + m->mu2 * x[0] * x[0];
    for (int j = 1; j < M; ++j) {
      x_diff = x[j] - x[j - 1];
      S += ainv2 * x_diff * x_diff + m->mu2 * x[j] * x[j];
    }
    return 0.5 * a * m->m0 * S;
  }
  case ORC_QUARTIC: { /* qm/quarticoscillatoraction.cc:7-28 */
    const double ainv2 = 1. / (a * a);
    double x_j = x[0];
    double x_j_squared = x_j * x_j;
    double x_j_shifted = x_j - m->x0;
    double x_j_shifted_squared = x_j_shifted * x_j_shifted;
    double x_diff = x[0] - x[M - 1];
    double S = m->m0 * (ainv2 * x_diff * x_diff + m->mu2 * x_j_squared) +
               0.5 * m->lambda * x_j_shifted_squared * x_j_shifted_squared;
    for (int j = 1; j < M; ++j) {
      x_j = x[j];
      x_j_squared = x_j * x_j;
      x_j_shifted = x_j - m->x0;
      x_j_shifted_squared = x_j_shifted * x_j_shifted;
      x_diff = x_j - x[j - 1];
      S += m->m0 * (ainv2 * x_diff * x_diff + m->mu2 * x_j_squared) +
           0.5 * m->lambda * x_j_shifted_squared * x_j_shifted_squared;
    }
    return 0.5 * a * S;
  }
  case ORC_ROTOR: { /* qm/rotoraction.cc:9-18 */
    double x_diff = x[0] - x[M - 1];
    double S = 1. - cos(x_diff);
    for (int j = 1; j < M; ++j) {
      x_diff = x[j] - x[j - 1];
      S += 1. - cos(x_diff);
    }
    return m->m0 / a * S;
  }
  case ORC_SCHWINGER: { /* qft/quenchedschwingeraction.cc:7-22 */
    const int Mt = m->Mt_lat, Mx = m->Mx_lat;
    double S = 0;
    for (int i = 0; i < Mt; ++i)
      for (int j = 0; j < Mx; ++j)
        S += (1. - cos(plaq(x, Mt, Mx, i, j)));
    return m->beta * S;
  }
  case ORC_GFF: { /* qft/gffaction.cc:7-24 (n_gibbs_smooth == 0 branch) */
    const double kappa = 4. + m->gff_mu2;
    const int N = orc_sample_size(m);
    double S = 0.0;
    for (int ell = 0; ell < N; ++ell) {
      uint32_t nb[8];
      orc_neighbours(m->Mt_lat, m->Mx_lat, m->rotated, ell, nb);
      const double phi_n = x[ell];
      double S_local = kappa * phi_n;
      for (int k = 0; k < 4; ++k)
        S_local -= x[nb[k]];
      S += phi_n * S_local;
    }
    return 0.5 * S;
  }
  }
  return NAN;
}

void orc_force(const orc_model *m, const double *x, double *p) {
  const int M = m->M_lat;
  const double a = m->a_lat;
  switch (m->model) {
  case ORC_HO: { /* qm/harmonicoscillatoraction.cc:21-35 */
    const double tmp_1 = m->m0 / a;
    const double tmp_2 = 2. + a * a * m->mu2;
    p[0] = tmp_1 * (tmp_2 * x[0] - x[M - 1] - x[1]);
    for (int j = 1; j < M - 1; ++j)
      p[j] = tmp_1 * (tmp_2 * x[j] - x[j - 1] - x[j + 1]);
    p[M - 1] = tmp_1 * (tmp_2 * x[M - 1] - x[M - 2] - x[0]);
    return;
  }
  case ORC_QUARTIC: { /* qm/quarticoscillatoraction.cc:31-53 */
    const double tmp_1 = m->m0 / a;
    const double tmp_2 = 2. + a * a * m->mu2;
    const double tmp_3 = a * m->lambda;
    for (int j = 0; j < M; ++j) {
      const double X_j = x[j];
      const double X_j_shifted = X_j - m->x0;
      const double xm = x[(j + M - 1) % M], xp = x[(j + 1) % M];
      p[j] = tmp_1 * (tmp_2 * X_j - xm - xp) +
             tmp_3 * X_j_shifted * X_j_shifted * X_j_shifted;
    }
    return;
  }
  case ORC_ROTOR: { /* qm/rotoraction.cc:58-79 */
    const double tmp = m->m0 / a;
    for (int j = 0; j < M; ++j) {
      const double x_m = x[(j + M - 1) % M], xx = x[j], x_p = x[(j + 1) % M];
      p[j] = tmp * (sin(xx - x_m) + sin(xx - x_p));
    }
    return;
  }
  case ORC_SCHWINGER: { /* qft/quenchedschwingeraction.cc:68-89 */
    const int Mt = m->Mt_lat, Mx = m->Mx_lat;
    for (int ell = 0; ell < 2 * Mt * Mx; ++ell)
      p[ell] = 0.0;
    for (int i = 0; i < Mt; ++i)
      for (int j = 0; j < Mx; ++j) {
        const double F = m->beta * sin(plaq(x, Mt, Mx, i, j));
        p[LNK(i, j, 0)] += F;
        p[LNK(i + 1, j, 1)] += F;
        p[LNK(i, j + 1, 0)] -= F;
        p[LNK(i, j, 1)] -= F;
      }
    return;
  }
  case ORC_GFF: { /* qft/gffaction.cc:82-94 */
    const double kappa = 4. + m->gff_mu2;
    const int N = orc_sample_size(m);
    for (int ell = 0; ell < N; ++ell) {
      uint32_t nb[8];
      orc_neighbours(m->Mt_lat, m->Mx_lat, m->rotated, ell, nb);
      double momentum = kappa * x[ell];
      for (int k = 0; k < 4; ++k)
        momentum -= x[nb[k]];
      p[ell] = momentum;
    }
    return;
  }
  }
}

/* getWminimum / getWcurvature of the three QM actions */
void orc_W(const orc_model *m, double x_m, double x_p, double *Wmin,
           double *Wcurv) {
  const double a = m->a_lat;
  switch (m->model) {
  case ORC_HO: /* qm/harmonicoscillatoraction.hh:97-98,171-189 */
    *Wcurv = (2. / a + a * m->mu2) * m->m0;
    *Wmin = (0.5 / (1. + 0.5 * a * a * m->mu2)) * (x_m + x_p);
    return;
  case ORC_QUARTIC: { /* qm/quarticoscillatoraction.hh:160-194 */
    const double xc = 0.5 * (x_m + x_p);
    *Wcurv = (2. / a + a * m->mu2) * m->m0 +
             3. * m->lambda * a * (xc - m->x0) * (xc - m->x0);
    const double xbar = 0.5 * (x_m + x_p);
    const double rho = 1. / (1. + 0.5 * a * a * m->mu2);
    double x = xbar;
    for (int i = 0; i < 4; ++i) {
      const double x_shifted = x - m->x0;
      x = rho * (xbar - 0.5 * a * a * m->lambda / m->m0 * x_shifted *
                            x_shifted * x_shifted);
    }
    *Wmin = x;
    return;
  }
  case ORC_ROTOR: /* qm/rotoraction.hh:195-213 */
    *Wcurv = 2.0 * m->m0 / a * fabs(cos(0.5 * (x_p - x_m)));
    *Wmin = atan2(sin(x_p) + sin(x_m), cos(x_p) + cos(x_m));
    return;
  }
  *Wmin = *Wcurv = NAN;
}

/* qft/quenchedschwingeraction.cc:25-43 */
static void staple_angles(const double *x, int Mt, int Mx, int i, int j, int mu,
                          double *theta_p, double *theta_m) {
  if (mu == 0) {
    *theta_p =
        orc_mod_2pi(x[LNK(i, j + 1, 0)] + x[LNK(i, j, 1)] - x[LNK(i + 1, j, 1)]);
    *theta_m = orc_mod_2pi(x[LNK(i, j - 1, 0)] + x[LNK(i + 1, j - 1, 1)] -
                           x[LNK(i, j - 1, 1)]);
  } else {
    *theta_p =
        orc_mod_2pi(x[LNK(i, j, 0)] + x[LNK(i + 1, j, 1)] - x[LNK(i, j + 1, 0)]);
    *theta_m = orc_mod_2pi(x[LNK(i - 1, j + 1, 0)] + x[LNK(i - 1, j, 1)] -
                           x[LNK(i - 1, j, 0)]);
  }
}

void orc_overrelax_update(const orc_model *m, double *x, uint32_t ell) {
  switch (m->model) {
  case ORC_ROTOR: { /* qm/rotoraction.cc:40-56 */
    const int M = m->M_lat;
    const double x_m = x[(ell + M - 1) % M], x_p = x[(ell + 1) % M];
    double x0, c;
    orc_W(m, x_m, x_p, &x0, &c);
    x[ell] = orc_mod_2pi(2.0 * x0 - x[ell]);
    return;
  }
  case ORC_GFF: { /* qft/gffaction.cc:68-79 */
    const double Delta = gff_nn_sum(m, x, ell);
    x[ell] = 2. * Delta / (4. + m->gff_mu2) - x[ell];
    return;
  }
  case ORC_SCHWINGER: { /* qft/quenchedschwingeraction.cc:57-65 */
    const int Mt = m->Mt_lat, Mx = m->Mx_lat;
    int i, j, mu;
    double theta_p, theta_m;
    orc_link_lin2cart(Mt, Mx, ell, &i, &j, &mu);
    staple_angles(x, Mt, Mx, i, j, mu, &theta_p, &theta_m);
    x[ell] = orc_mod_2pi((theta_p + theta_m) - x[ell]);
    return;
  }
  }
}

/* sampler/overrelaxedheatbathsampler.cc:10-18 with random_order = false */
void orc_overrelax_sweep_lex(const orc_model *m, double *x) {
  const int n = orc_sample_size(m);
  for (int ell = 0; ell < n; ++ell)
    orc_overrelax_update(m, x, ell);
}

/* Colouring used by the CUDA sweeps: dofs of one colour do not
 * interact, so they can be updated simultaneously.  Same per-dof update and the
 * same stationary distribution as the reference's sequential sweep. */
int orc_n_colours(const orc_model *m) {
  switch (m->model) {
  case ORC_ROTOR:
  case ORC_GFF:
    return 2;
  case ORC_SCHWINGER:
    return 4;
  }
  return 0;
}
int orc_colour_of(const orc_model *m, uint32_t ell) {
  switch (m->model) {
  case ORC_ROTOR:
    return ell & 1;
  case ORC_GFF: {
    int i, j;
    orc_vertex_lin2cart(m->Mt_lat, m->Mx_lat, m->rotated, ell, &i, &j);
    return m->rotated ? (i & 1) : ((i + j) & 1);
  }
  case ORC_SCHWINGER: {
    int i, j, mu;
    orc_link_lin2cart(m->Mt_lat, m->Mx_lat, ell, &i, &j, &mu);
    return mu == 0 ? (j & 1) : 2 + (i & 1);
  }
  }
  return -1;
}
void orc_overrelax_sweep_coloured(const orc_model *m, double *x) {
  const int n = orc_sample_size(m);
  const int nc = orc_n_colours(m);
  for (int c = 0; c < nc; ++c)
    for (int ell = 0; ell < n; ++ell)
      if (orc_colour_of(m, ell) == c)
        orc_overrelax_update(m, x, ell);
}

/* Action::copy_from_coarse of the FINE action */
void orc_prolong(const orc_model *fine, const double *xc, double *x) {
  switch (fine->model) {
  case ORC_HO:
  case ORC_QUARTIC:
  case ORC_ROTOR: /* qm/qmaction.cc:7-15 */
    for (int j = 0; j < fine->M_lat / 2; ++j)
      x[2 * j] = xc[j];
    return;
  case ORC_SCHWINGER: { /* qft/quenchedschwingeraction.cc:92-147 */
    const int Mt = fine->Mt_lat, Mx = fine->Mx_lat;
    if (fine->coarsening == ORC_COARSEN_BOTH) {
      const int Mtc = Mt / 2, Mxc = Mx / 2;
      for (int i = 0; i < Mtc; ++i)
        for (int j = 0; j < Mxc; ++j) {
          double theta_c = xc[orc_link_cart2lin(Mtc, Mxc, i, j, 0)];
          x[LNK(2 * i, 2 * j, 0)] = 0.5 * theta_c;
          x[LNK(2 * i + 1, 2 * j, 0)] = 0.5 * theta_c;
          theta_c = xc[orc_link_cart2lin(Mtc, Mxc, i, j, 1)];
          x[LNK(2 * i, 2 * j, 1)] = 0.5 * theta_c;
          x[LNK(2 * i, 2 * j + 1, 1)] = 0.5 * theta_c;
        }
    } else if (fine->coarsening == ORC_COARSEN_TEMPORAL) {
      const int Mtc = Mt / 2, Mxc = Mx;
      for (int i = 0; i < Mtc; ++i)
        for (int j = 0; j < Mx; ++j) {
          double theta_c = xc[orc_link_cart2lin(Mtc, Mxc, i, j, 0)];
          x[LNK(2 * i, j, 0)] = 0.5 * theta_c;
          x[LNK(2 * i + 1, j, 0)] = 0.5 * theta_c;
          theta_c = xc[orc_link_cart2lin(Mtc, Mxc, i, j, 1)];
          x[LNK(2 * i, j, 1)] = theta_c;
        }
    } else if (fine->coarsening == ORC_COARSEN_SPATIAL) {
      const int Mtc = Mt, Mxc = Mx / 2;
      for (int i = 0; i < Mt; ++i)
        for (int j = 0; j < Mxc; ++j) {
          double theta_c = xc[orc_link_cart2lin(Mtc, Mxc, i, j, 0)];
          x[LNK(i, 2 * j, 0)] = theta_c;
          theta_c = xc[orc_link_cart2lin(Mtc, Mxc, i, j, 1)];
          x[LNK(i, 2 * j, 1)] = 0.5 * theta_c;
          x[LNK(i, 2 * j + 1, 1)] = 0.5 * theta_c;
        }
    }
    return;
  }
  case ORC_GFF: { /* qft/gffaction.cc:97-106 */
    const int N = orc_sample_size(fine);
    uint32_t *buf = (uint32_t *)malloc(3 * (size_t)N * sizeof(uint32_t));
    int counts[2];
    /* level parity only matters for ROTATE: rotated <=> odd level */
    const int level = fine->rotated ? 1 : 0;
    if (orc_coarsening_lists(fine->Mt_lat, fine->Mx_lat, fine->coarsening, level,
                             buf, buf + N, buf + 2 * N, counts) == 0)
      for (int k = 0; k < counts[0]; ++k)
        x[buf[k]] = xc[buf[2 * N + k]];
    free(buf);
    return;
  }
  }
}

/* Action::copy_from_fine of the COARSE action */
void orc_restrict(const orc_model *fine, const double *xf, double *xc) {
  switch (fine->model) {
  case ORC_HO:
  case ORC_QUARTIC:
  case ORC_ROTOR: /* qm/qmaction.cc:18-26 */
    for (int j = 0; j < fine->M_lat / 2; ++j)
      xc[j] = xf[2 * j];
    return;
  case ORC_SCHWINGER: { /* qft/quenchedschwingeraction.cc:150-195 */
    const int Mt = fine->Mt_lat, Mx = fine->Mx_lat;
    if (fine->coarsening == ORC_COARSEN_BOTH) {
      const int Mtc = Mt / 2, Mxc = Mx / 2;
      for (int i = 0; i < Mtc; ++i)
        for (int j = 0; j < Mxc; ++j) {
          xc[orc_link_cart2lin(Mtc, Mxc, i, j, 0)] =
              orc_mod_2pi(xf[LNK(2 * i, 2 * j, 0)] + xf[LNK(2 * i + 1, 2 * j, 0)]);
          xc[orc_link_cart2lin(Mtc, Mxc, i, j, 1)] =
              orc_mod_2pi(xf[LNK(2 * i, 2 * j, 1)] + xf[LNK(2 * i, 2 * j + 1, 1)]);
        }
    } else if (fine->coarsening == ORC_COARSEN_TEMPORAL) {
      const int Mtc = Mt / 2, Mxc = Mx;
      for (int i = 0; i < Mtc; ++i)
        for (int j = 0; j < Mxc; ++j) {
          xc[orc_link_cart2lin(Mtc, Mxc, i, j, 0)] =
              orc_mod_2pi(xf[LNK(2 * i, j, 0)] + xf[LNK(2 * i + 1, j, 0)]);
          xc[orc_link_cart2lin(Mtc, Mxc, i, j, 1)] =
              orc_mod_2pi(xf[LNK(2 * i, j, 1)]);
        }
    } else if (fine->coarsening == ORC_COARSEN_SPATIAL) {
      const int Mtc = Mt, Mxc = Mx / 2;
      for (int i = 0; i < Mtc; ++i)
        for (int j = 0; j < Mxc; ++j) {
          xc[orc_link_cart2lin(Mtc, Mxc, i, j, 0)] =
              orc_mod_2pi(xf[LNK(i, 2 * j, 0)]);
          xc[orc_link_cart2lin(Mtc, Mxc, i, j, 1)] =
              orc_mod_2pi(xf[LNK(i, 2 * j, 1)] + xf[LNK(i, 2 * j + 1, 1)]);
        }
    }
    return;
  }
  case ORC_GFF: { /* qft/gffaction.cc:109-118 */
    const int N = orc_sample_size(fine);
    uint32_t *buf = (uint32_t *)malloc(3 * (size_t)N * sizeof(uint32_t));
    int counts[2];
    const int level = fine->rotated ? 1 : 0;
    if (orc_coarsening_lists(fine->Mt_lat, fine->Mx_lat, fine->coarsening, level,
                             buf, buf + N, buf + 2 * N, counts) == 0)
      for (int k = 0; k < counts[0]; ++k)
        xc[buf[2 * N + k]] = xf[buf[k]];
    free(buf);
    return;
  }
  }
}

/* ========================================================== distributions */

/* distribution/expsin2distribution.cc:7-24 */
double orc_expsin2_pdf(double x, double sigma) {
  const double sin_x_half = sin(0.5 * x);
  const double z = 0.5 * sigma;
  double besselI0;
  if (z > 100.) {
    const double z_inv = 1. / z;
    besselI0 =
        sqrt(2. * M_PI * z_inv) * (1. + 0.125 * z_inv + 0.0703125 * z_inv * z_inv);
  } else {
    besselI0 = 2. * M_PI * orc_bessel_I0_scaled(z);
  }
  return exp(-sigma * sin_x_half * sin_x_half) / besselI0;
}

/* distribution/expcosdistribution.cc:7-21 */
double orc_expcos_pdf(double beta, double x, double x_p, double x_m) {
  double dx = x_p - x_m;
  double z = x - x_m;
  int sign_flip = (dx < 0.0) ? -1 : +1;
  dx *= sign_flip;
  if (dx > M_PI) {
    sign_flip *= -1;
    dx = 2. * M_PI - dx;
  }
  z *= sign_flip;
  const double sigma = 2. * beta * fabs(cos(0.5 * dx));
  const double Z_norm = 2. * M_PI * orc_fast_bessel_I0_scaled(sigma);
  return 1. / Z_norm * exp(sigma * (cos(z - 0.5 * dx) - 1.0));
}

/* distribution/besselproductdistribution.hh:52-80: alphaZ[0] unscaled,
 * alphaZ[k>=1] divided by alphaZ[0] */
void orc_besselproduct_alpha(double beta, double alphaZ[17]) {
  const unsigned int kmax = 16, nmax = 32;
  double alpha0 = 0.0;
  for (unsigned int k = 0; k <= kmax; ++k) {
    double s = 0.0;
    for (unsigned int n = k; n <= nmax; ++n)
      for (unsigned int mm = k; mm <= nmax; ++mm) {
        const double log_comb = log_nCk(2 * n, n - k) + log_nCk(2 * mm, mm - k) -
                                2 * (log_factorial(n) + log_factorial(mm));
        s += pow(0.5 * beta, 2 * (n + mm)) * exp(log_comb);
      }
    double alpha = ((k == 0) ? 2 : 4) * M_PI * s;
    if (k == 0)
      alpha0 = alpha;
    else
      alpha /= alpha0;
    alphaZ[k] = alpha;
  }
}

/* distribution/besselproductdistribution.cc:15-25 */
double orc_besselproduct_Znorm_inv(const double alphaZ[17], double phi,
                                   int rescaled) {
  double s = 1.0;
  for (unsigned int k = 1; k <= 16; ++k)
    s += alphaZ[k] * cos(k * phi);
  if (!rescaled)
    s *= alphaZ[0];
  return 1.0 / s;
}

/* distribution/besselproductdistribution.cc:7-12 */
double orc_besselproduct_pdf(double beta, double x, double x_p, double x_m) {
  double alphaZ[17];
  orc_besselproduct_alpha(beta, alphaZ);
  const double I0_p = orc_bessel_I0(2 * beta * cos(0.5 * (x - x_p)));
  const double I0_m = orc_bessel_I0(2 * beta * cos(0.5 * (x - x_m)));
  return orc_besselproduct_Znorm_inv(alphaZ, x_p - x_m, 0) * I0_p * I0_m;
}

/* distribution/approximatebesselproductdistribution.cc:38-54 (keeps the
 * reference's rho = (s_p/s_m)^{3/2} exp(-4 (s_p - s_m))) */
static void approx_N_p_sigma2inv(double beta, double x0, double *N_p,
                                 double *sigma2_p_inv, double *sigma2_m_inv) {
  const double epsilon = 0.125 * M_PI;
  if (x0 < epsilon) {
    *sigma2_p_inv = beta;
    *sigma2_m_inv = 0.0;
    *N_p = 1.0;
  } else {
    *sigma2_p_inv = beta * cos(0.25 * x0);
    *sigma2_m_inv = beta * sin(0.25 * x0);
    const double rho = pow(*sigma2_p_inv / *sigma2_m_inv, 1.5) *
                       exp(-4.0 * (*sigma2_p_inv - *sigma2_m_inv));
    *N_p = 1.0 / (1.0 + rho);
  }
}

/* distribution/approximatebesselproductdistribution.cc:7-35 */
double orc_approxbessel_pdf(double beta, double x, double x_p, double x_m) {
  double x0 = x_p - x_m;
  double z = x - x_m;
  double sign_flip = (x0 < 0) ? -1 : +1;
  x0 *= sign_flip;
  if (x0 > M_PI) {
    x0 = 2. * M_PI - x0;
    sign_flip *= -1;
  }
  z *= sign_flip;
  double N_p, sigma2_p_inv, sigma2_m_inv;
  approx_N_p_sigma2inv(beta, x0, &N_p, &sigma2_p_inv, &sigma2_m_inv);
  const double N_m = 1. - N_p;
  double s_p = 0.0, s_m = 0.0;
  for (int k = -4; k <= 4; ++k) {
    double z_shifted = z - 0.5 * x0 + 2 * k * M_PI;
    s_p += sqrt(sigma2_p_inv) * exp(-0.5 * sigma2_p_inv * z_shifted * z_shifted);
    z_shifted += M_PI;
    s_m += sqrt(sigma2_m_inv) * exp(-0.5 * sigma2_m_inv * z_shifted * z_shifted);
  }
  return sqrt(0.5 / M_PI) * (N_p * s_p + N_m * s_m);
}

/* The draw() routines below restate the reference's sampling algorithms
 * (proposal, envelope, acceptance test, output map) on the Philox stream.
 * Random numbers are consumed in blocks: one orc_rng_normal2 + one
 * orc_rng_uniform2 serve two consecutive attempts (z0,u0) then (z1,u1). */

/* distribution/expsin2distribution.hh:44-58 */
double orc_expsin2_draw(orc_rng *r, double sigma) {
  for (;;) {
    double z[2], u[2];
    orc_rng_normal2(r, &z[0], &z[1]);
    orc_rng_uniform2(r, &u[0], &u[1]);
    for (int t = 0; t < 2; ++t) {
      const double r_x = M_PI / sqrt(2. * sigma) * z[t];
      if (fabs(r_x) < M_PI) {
        const double sin_psi_half = sin(0.5 * r_x);
        if (u[t] <
            exp(-sigma * (sin_psi_half * sin_psi_half - r_x * r_x / (M_PI * M_PI))))
          return r_x;
      }
    }
  }
}

/* Proposal of the ExpCos rejection sampler.
 *   0: the reference's own envelope (distribution/expcosdistribution.hh:50-65).
 *   1: NOT the reference's algorithm: the tighter chord-bound envelope
 *      1 - cos x >= 2 x^2 / pi^2 (uniform proposal for tau < 1/2).
 *   2: NOT the reference's algorithm: as 1 for tau < 64; for tau >= 64 the Taylor bound
 *      1 - cos x >= x^2/2 (1 - x^2/12) on x^2 <= 160/tau, i.e. the Gaussian envelope
 *      exp(-(tau - 40/3) x^2 / 2), with the squeeze u <= 1 - (20/3) x^2; the target is
 *      truncated to x^2 <= 160/tau (tail mass < 1e-33).  The product's default
 *      (MLMCPI_OPT_EXPCOS_ENVELOPE = 2).
 *   The target pdf ~ exp(tau cos x) is the reference's; tests/test_oracle_cpu.py checks all
 *   variants against the reference's own draw() by a two-sample KS test. */
static int g_expcos_envelope = 0;
void orc_set_expcos_envelope(int envelope) { g_expcos_envelope = envelope; }

/* first != NULL: the first Gaussian attempt uses the variates (first[0] = z, first[1] = u) handed in by the
 * caller instead of a block of the stream r (the fill-in of CoarsenBoth shares one normal pair and one
 * uniform pair between the two horizontal interior links of a coarse cell); further attempts consume blocks
 * of r: (z0, u0), (z1, u1), next block, ...  The uniform-proposal branch (tau < 1/2) ignores `first`. */
static double expcos_draw_impl(orc_rng *r, double beta, double x_p, double x_m, const double *first) {
  const double dx = x_m - x_p;
  const double tau = 2. * beta * fabs(cos(0.5 * dx));
  double x = 0.0;
  int accepted = 0;
  if (g_expcos_envelope >= 1 && tau < 0.5) {
    while (!accepted) {
      double a[2], u[2];
      orc_rng_uniform2(r, &a[0], &a[1]);
      orc_rng_uniform2(r, &u[0], &u[1]);
      for (int t = 0; t < 2 && !accepted; ++t) {
        x = -M_PI + 2. * M_PI * a[t];
        accepted = (u[t] <= exp(tau * (cos(x) - 1.)));
      }
    }
  } else {
    const int tight = (g_expcos_envelope == 2 && tau >= 64.0);
    const double a = tau - 40. / 3.;
    /* envelope 0: expcosdistribution.hh:53-61 (sigma = pi sqrt(2/tau), fourpi2_inv = 1/(4 pi^2)) */
    const double sigma = tight ? 1.0 / sqrt(a)
                               : (g_expcos_envelope >= 1 ? 0.5 * M_PI / sqrt(tau) : M_PI * sqrt(2. / tau));
    const double quad = g_expcos_envelope >= 1 ? 2. / (M_PI * M_PI) : 1. / (4. * M_PI * M_PI);
    double z[2] = {0, 0}, u[2] = {0, 0};
    int t = first ? -1 : 0;
    while (!accepted) {
      double zz, uu;
      if (t < 0) {
        zz = first[0];
        uu = first[1];
      } else {
        if ((t & 1) == 0) {
          orc_rng_normal2(r, &z[0], &z[1]);
          orc_rng_uniform2(r, &u[0], &u[1]);
        }
        zz = z[t & 1];
        uu = u[t & 1];
      }
      ++t;
      x = sigma * zz;
      if (tight) {
        const double x2 = x * x;
        if (tau * x2 <= 160.)
          accepted = (uu <= 1. - (20. / 3.) * x2) || (uu <= exp(tau * (cos(x) - 1.) + 0.5 * a * x2));
      } else if ((-M_PI <= x) && (x < M_PI)) {
        accepted = (uu <= exp(tau * (cos(x) - 1. + quad * x * x)));
      }
    }
  }
  return orc_mod_2pi(x + 0.5 * (x_p + x_m) + (fabs(dx) > M_PI) * M_PI);
}
double orc_expcos_draw(orc_rng *r, double beta, double x_p, double x_m) {
  return expcos_draw_impl(r, beta, x_p, x_m, NULL);
}

/* distribution/besselproductdistribution.hh:82-142.  Per outer attempt one
 * orc_rng_uniform2 gives (side selector, acceptance variate); the inner
 * "redraw until inside the interval" loop consumes orc_rng_normal2 blocks. */
double orc_besselproduct_draw(orc_rng *r, double beta, double x_p, double x_m) {
  const double I0_twobeta = orc_bessel_I0(2 * beta);
  const double sigma_beta = M_PI / sqrt(2 * log(I0_twobeta));
  double a_min, a_max, mu;
  double dx = x_m - x_p;
  const double sign_flip = (dx < 0) ? -1 : +1;
  dx *= sign_flip;
  const double N_p = erf((M_PI - 0.5 * dx) / sigma_beta);
  const double N_m =
      erf(0.5 * dx / sigma_beta) * pow(I0_twobeta, 2. * (dx / M_PI - 1.));
  const double C_gauss_p =
      pow(I0_twobeta, 2. * (1. - dx * dx / (4. * M_PI * M_PI)));
  const double C_gauss_m = pow(
      I0_twobeta,
      2. * (1. - (dx - 2. * M_PI) * (dx - 2. * M_PI) / (4. * M_PI * M_PI)));
  const double sigma = sigma_beta / sqrt(2.);
  double x = 0.0, C_gauss;
  for (;;) {
    double xi, xi_acc;
    orc_rng_uniform2(r, &xi, &xi_acc);
    if (xi >= N_m / (N_p + N_m)) {
      a_min = -M_PI + dx;
      a_max = +M_PI;
      mu = 0.5 * dx;
      C_gauss = C_gauss_p;
    } else {
      a_min = -M_PI;
      a_max = -M_PI + dx;
      mu = 0.5 * (dx - 2. * M_PI);
      C_gauss = C_gauss_m;
    }
    int inside = 0;
    while (!inside) {
      double z[2];
      orc_rng_normal2(r, &z[0], &z[1]);
      for (int t = 0; t < 2 && !inside; ++t) {
        x = sigma * z[t] + mu;
        inside = ((x >= a_min) && (x < a_max));
      }
    }
    const double I0 = orc_bessel_I0(2. * beta * cos(0.5 * x));
    const double I0_dx = orc_bessel_I0(2. * beta * cos(0.5 * (x - dx)));
    const double x_shifted = (x - mu) / sigma_beta;
    const double rho_accept = I0 * I0_dx / C_gauss * exp(x_shifted * x_shifted);
    if (xi_acc <= rho_accept)
      break;
  }
  return orc_mod_2pi(sign_flip * x + x_p);
}

/* distribution/approximatebesselproductdistribution.hh:81-106 with the mode-selecting
 * uniform variate xi supplied by the caller; the normal is the first variate of one
 * orc_rng_normal2 */
static double approxbessel_draw_xi(orc_rng *r, double beta, double x_p, double x_m,
                                   double xi) {
  double x0 = x_p - x_m;
  double sign_flip = (x0 < 0) ? -1 : +1;
  x0 *= sign_flip;
  if (x0 > M_PI) {
    x0 = 2. * M_PI - x0;
    sign_flip *= -1;
  }
  double N_p, sigma2_p_inv, sigma2_m_inv;
  approx_N_p_sigma2inv(beta, x0, &N_p, &sigma2_p_inv, &sigma2_m_inv);
  double z0, z1;
  orc_rng_normal2(r, &z0, &z1);
  double sigma, xshift;
  if (xi <= N_p) {
    sigma = 1. / sqrt(sigma2_p_inv);
    xshift = 0.0;
  } else {
    sigma = 1. / sqrt(sigma2_m_inv);
    xshift = M_PI;
  }
  const double x = sigma * z0 + 0.5 * x0 - xshift;
  return orc_mod_2pi(sign_flip * x + x_m);
}

/* stand-alone draw: xi = first variate of one orc_rng_uniform2 */
double orc_approxbessel_draw(orc_rng *r, double beta, double x_p, double x_m) {
  double xi, unused;
  orc_rng_uniform2(r, &xi, &unused);
  return approxbessel_draw_xi(r, beta, x_p, x_m, xi);
}

/* ================================================ conditioned fine actions */

static double cond_action_1d(const orc_model *m, const double *x) {
  /* qm/gaussianconditionedfineaction.cc:27-43 / qm/rotorconditionedfineaction.cc:27-43
   * (wrap-around point first, then j = 0 .. M/2-2) */
  const int M = m->M_lat;
  double S = 0.0;
  for (int jj = -1; jj < M / 2 - 1; ++jj) {
    double x_m, x_p, xf;
    if (jj < 0) {
      x_m = x[M - 2];
      x_p = x[0];
      xf = x[M - 1];
    } else {
      x_m = x[2 * jj];
      x_p = x[2 * jj + 2];
      xf = x[2 * jj + 1];
    }
    double Wmin, Wcurv;
    orc_W(m, x_m, x_p, &Wmin, &Wcurv);
    const double dx = xf - Wmin;
    double term;
    if (m->model == ORC_ROTOR) {
      const double sigma = 2.0 * Wcurv;
      term = -log(orc_expsin2_pdf(dx, sigma));
    } else {
      term = 0.5 * Wcurv * dx * dx - 0.5 * log(Wcurv);
    }
    if (jj < 0)
      S = term;
    else
      S += term;
  }
  return S;
}

double orc_cond_action(const orc_model *fine, const double *x) {
  switch (fine->model) {
  case ORC_HO:
  case ORC_QUARTIC:
  case ORC_ROTOR:
    return cond_action_1d(fine, x);
  case ORC_GFF: { /* qft/gffconditionedfineaction.cc:28-50 */
    const int N = orc_sample_size(fine);
    uint32_t *buf = (uint32_t *)malloc(3 * (size_t)N * sizeof(uint32_t));
    int counts[2];
    const int level = fine->rotated ? 1 : 0;
    double S = 0;
    if (orc_coarsening_lists(fine->Mt_lat, fine->Mx_lat, fine->coarsening, level,
                             buf, buf + N, buf + 2 * N, counts) == 0) {
      const double sigma2 = 1. / (4. + fine->gff_mu2);
      const double sigma2_inv = 1. / sigma2;
      for (int k = 0; k < counts[1]; ++k) {
        const uint32_t ell = buf[N + k];
        const double Delta = gff_nn_sum(fine, x, ell);
        const double dphi = x[ell] - sigma2 * Delta;
        S += 0.5 * sigma2_inv * dphi * dphi;
      }
    } else {
      S = NAN;
    }
    free(buf);
    return S;
  }
  case ORC_SCHWINGER: {
    const int Mt = fine->Mt_lat, Mx = fine->Mx_lat;
    const double beta = fine->beta;
    double S = 0.0;
    if (fine->coarsening == ORC_COARSEN_BOTH) {
      /* qft/quenchedschwingerconditionedfineaction.cc:212-290 */
      if (beta <= 8.0) {
        double alphaZ[17];
        orc_besselproduct_alpha(beta, alphaZ);
        for (int i = 0; i < Mt / 2; ++i)
          for (int j = 0; j < Mx / 2; ++j) {
            const double phi_12 =
                +x[LNK(2 * i, 2 * j + 1, 1)] + x[LNK(2 * i, 2 * j + 2, 0)];
            const double phi_23 =
                +x[LNK(2 * i + 1, 2 * j + 2, 0)] - x[LNK(2 * i + 2, 2 * j + 1, 1)];
            const double phi_34 =
                -x[LNK(2 * i + 1, 2 * j, 0)] - x[LNK(2 * i + 2, 2 * j, 1)];
            const double phi_41 = -x[LNK(2 * i, 2 * j, 0)] + x[LNK(2 * i, 2 * j, 1)];
            const double theta_1 = +x[LNK(2 * i, 2 * j + 1, 0)];
            const double theta_2 = -x[LNK(2 * i + 1, 2 * j + 1, 1)];
            const double theta_3 = -x[LNK(2 * i + 1, 2 * j + 1, 0)];
            const double theta_4 = +x[LNK(2 * i + 1, 2 * j, 1)];
            const double Phi = phi_12 + phi_23 + phi_34 + phi_41;
            S -= beta * (cos(theta_1 - theta_2 - phi_12) +
                         cos(theta_2 - theta_3 - phi_23) +
                         cos(theta_3 - theta_4 - phi_34) +
                         cos(theta_4 - theta_1 - phi_41));
            S -= log(orc_besselproduct_Znorm_inv(alphaZ, Phi, 1));
          }
      } else {
        for (int i = 0; i < Mt / 2; ++i)
          for (int j = 0; j < Mx / 2; ++j) {
            const double phi_p = orc_mod_2pi(
                +x[LNK(2 * i + 1, 2 * j, 0)] + x[LNK(2 * i + 2, 2 * j, 1)] +
                x[LNK(2 * i + 2, 2 * j + 1, 1)] - x[LNK(2 * i + 1, 2 * j + 2, 0)]);
            const double phi_m = orc_mod_2pi(
                -x[LNK(2 * i, 2 * j, 0)] + x[LNK(2 * i, 2 * j, 1)] +
                x[LNK(2 * i, 2 * j + 1, 1)] + x[LNK(2 * i, 2 * j + 2, 0)]);
            const double theta = orc_mod_2pi(+x[LNK(2 * i + 1, 2 * j, 1)] +
                                             x[LNK(2 * i + 1, 2 * j + 1, 1)]);
            S -= log(orc_approxbessel_pdf(beta, theta, phi_p, phi_m));
          }
        for (int i = 0; i < Mt; ++i)
          for (int j = 0; j < Mx / 2; ++j) {
            const double phi_p = orc_mod_2pi(-x[LNK(i, 2 * j, 1)] +
                                             x[LNK(i, 2 * j, 0)] +
                                             x[LNK(i + 1, 2 * j, 1)]);
            const double phi_m =
                orc_mod_2pi(+x[LNK(i, 2 * j + 1, 1)] + x[LNK(i, 2 * j + 2, 0)] -
                            x[LNK(i + 1, 2 * j + 1, 1)]);
            const double theta = orc_mod_2pi(+x[LNK(i, 2 * j + 1, 0)]);
            S -= log(orc_expcos_pdf(beta, theta, phi_p, phi_m));
          }
      }
    } else if (fine->coarsening == ORC_COARSEN_TEMPORAL) {
      /* qft/quenchedschwingerconditionedfineaction.cc:338-355 */
      for (int i = 0; i < Mt / 2; ++i)
        for (int j = 0; j < Mx; ++j) {
          const double phi_p = orc_mod_2pi(
              -x[LNK(2 * i, j, 0)] + x[LNK(2 * i, j, 1)] + x[LNK(2 * i, j + 1, 0)]);
          const double phi_m =
              orc_mod_2pi(+x[LNK(2 * i + 1, j, 0)] + x[LNK(2 * i + 2, j, 1)] -
                          x[LNK(2 * i + 1, j + 1, 0)]);
          const double theta = orc_mod_2pi(+x[LNK(2 * i + 1, j, 1)]);
          S -= log(orc_expcos_pdf(beta, theta, phi_p, phi_m));
        }
    } else if (fine->coarsening == ORC_COARSEN_SPATIAL) {
      /* qft/quenchedschwingerconditionedfineaction.cc:356-372 */
      for (int i = 0; i < Mt; ++i)
        for (int j = 0; j < Mx / 2; ++j) {
          const double phi_p = orc_mod_2pi(
              -x[LNK(i, 2 * j, 1)] + x[LNK(i, 2 * j, 0)] + x[LNK(i + 1, 2 * j, 1)]);
          const double phi_m =
              orc_mod_2pi(+x[LNK(i, 2 * j + 1, 1)] + x[LNK(i, 2 * j + 2, 0)] -
                          x[LNK(i + 1, 2 * j + 1, 1)]);
          const double theta = orc_mod_2pi(+x[LNK(i, 2 * j + 1, 0)]);
          S -= log(orc_expcos_pdf(beta, theta, phi_p, phi_m));
        }
    } else {
      S = NAN;
    }
    return S;
  }
  }
  return NAN;
}

/* ==================================================================== QoIs */

double orc_qoi(const orc_model *m, int qoi, const double *x, int64_t *Qint) {
  const double four_pi2_inv = 0.25 / (M_PI * M_PI);
  if (Qint)
    *Qint = 0;
  switch (qoi) {
  case ORC_QOI_X2: { /* qoi/qm/qoixsquared.cc:7-20 */
    const int M = m->M_lat;
    double X2 = 0.0;
    for (int i = 0; i < M; ++i)
      X2 += x[i] * x[i];
    return X2 / M;
  }
  case ORC_QOI_ROTOR_CHI: { /* qoi/qm/qoisusceptibility.cc:7-23 */
    const int M = m->M_lat;
    double dx = x[0] - x[M - 1];
    double Q = orc_mod_2pi(dx);
    for (int i = 1; i < M; ++i) {
      dx = x[i] - x[i - 1];
      Q += orc_mod_2pi(dx);
    }
    if (Qint)
      *Qint = llround(Q / (2. * M_PI));
    return four_pi2_inv * (Q * Q) / m->T_final;
  }
  case ORC_QOI_SCHWINGER_CHI: { /* qoi/qft/qoi2dsusceptibility.cc:7-27 */
    const int Mt = m->Mt_lat, Mx = m->Mx_lat;
    double Q = 0.0;
    for (int i = 0; i < Mt; ++i)
      for (int j = 0; j < Mx; ++j)
        Q += orc_mod_2pi(plaq(x, Mt, Mx, i, j));
    if (Qint)
      *Qint = llround(Q / (2. * M_PI));
    return four_pi2_inv * Q * Q;
  }
  case ORC_QOI_AVG_PLAQUETTE: { /* qoi/qft/qoiavgplaquette.cc:7-27 */
    const int Mt = m->Mt_lat, Mx = m->Mx_lat;
    double S_plaq = 0.0;
    for (int i = 0; i < Mt; ++i)
      for (int j = 0; j < Mx; ++j)
        S_plaq += cos(plaq(x, Mt, Mx, i, j));
    return S_plaq / (Mx * Mt);
  }
  case ORC_QOI_PHI2: { /* qoi/qft/qoi2dphisquared.cc:7-15 */
    const int N = orc_sample_size(m);
    double phi_squared = 0.0;
    for (int ell = 0; ell < N; ++ell)
      phi_squared += x[ell] * x[ell];
    return phi_squared / N;
  }
  }
  return NAN;
}

/* ===================================================================== HMC */

/* sampler/hmcsampler.cc:31-46 */
void orc_leapfrog(const orc_model *m, int nt, double dt, double *x, double *p) {
  const int n = orc_sample_size(m);
  double *dp = (double *)malloc((size_t)n * sizeof(double));
  for (int k = 0; k <= nt; ++k) {
    double dt_p = dt, dt_x = dt;
    if (k == 0)
      dt_p = 0.5 * dt;
    if (k == nt) {
      dt_p = 0.5 * dt;
      dt_x = 0.0;
    }
    orc_force(m, x, dp);
    for (int l = 0; l < n; ++l)
      p[l] -= dt_p * dp[l];
    for (int l = 0; l < n; ++l)
      x[l] += dt_x * p[l];
  }
  free(dp);
}

/* coarse-level model: Action::coarse_action() of each model */
int orc_coarse_model(const orc_model *fine, int renorm, int level, int ctype,
                     double T_final, orc_model *coarse) {
  *coarse = *fine;
  switch (fine->model) {
  case ORC_HO: /* qm/harmonicoscillatorrenormalisation.hh:46-79 */
  case ORC_QUARTIC: /* qm/quarticoscillatoraction.hh:105-110 (no renormalisation) */
  case ORC_ROTOR: { /* qm/rotorrenormalisation.hh:38-57, .cc:8-14 */
    if (fine->M_lat % 2)
      return -1;
    const double a = fine->a_lat;
    coarse->M_lat = fine->M_lat / 2;
    coarse->a_lat = T_final / coarse->M_lat; /* lattice/lattice1d.cc:6-9 */
    coarse->T_final = T_final;
    if (fine->model == ORC_HO) {
      if (renorm == 1) {
        coarse->m0 = fine->m0 * (1. - 0.5 * a * a * fine->mu2);
        coarse->mu2 = fine->mu2 * (1. + 0.25 * a * a * fine->mu2);
      } else if (renorm == 2) {
        coarse->m0 = fine->m0 / (1. + 0.5 * a * a * fine->mu2);
        coarse->mu2 = fine->mu2 * (1. + 0.25 * a * a * fine->mu2);
      }
    } else if (fine->model == ORC_ROTOR) {
      if (renorm == 1) {
        const double xi = T_final / fine->m0;
        const double S_hat2 = Sigma_hat(xi, 2);
        const double S_hat4 = Sigma_hat(xi, 4);
        const double deltaI =
            0.5 *
            (1. - 2. * xi * S_hat2 + 0.5 * xi * xi * (S_hat4 - S_hat2 * S_hat2)) /
            (1. - 2. * xi * S_hat2 + xi * xi * (S_hat4 - S_hat2 * S_hat2));
        coarse->m0 = (1. + deltaI * a / fine->m0) * fine->m0;
      } else if (renorm == 2) {
        return -1;
      }
    }
    return 0;
  }
  case ORC_SCHWINGER: { /* qft/quenchedschwingerrenormalisation.hh:45-105 */
    int Mt_c, Mx_c, rot_c;
    if (!orc_coarse_shape(fine->Mt_lat, fine->Mx_lat, ctype, level, &Mt_c, &Mx_c,
                          &rot_c))
      return -1;
    coarse->Mt_lat = Mt_c;
    coarse->Mx_lat = Mx_c;
    const double beta = fine->beta;
    const double raw = (ctype == ORC_COARSEN_BOTH) ? 0.25 * beta : 0.5 * beta;
    coarse->beta = raw;
    if (renorm == 1 && beta > 4.0) {
      const double rho = (ctype == ORC_COARSEN_BOTH) ? 0.25 : 0.5;
      const double delta = (ctype == ORC_COARSEN_BOTH) ? 1.5 : 0.5;
      coarse->beta = rho * (1. + delta / beta) * beta;
    } else if (renorm == 2 && beta > 4.0) {
      return -1; /* nonperturbative: host-side root find, not restated */
    }
    /* how the coarse level is itself coarsened */
    if (ctype == ORC_COARSEN_ALTERNATE)
      coarse->coarsening =
          ((level + 1) % 2 == 0) ? ORC_COARSEN_TEMPORAL : ORC_COARSEN_SPATIAL;
    return 0;
  }
  case ORC_GFF: { /* qft/gffaction.hh:174-181,201-208 (5-point part only) */
    int Mt_c, Mx_c, rot_c;
    if (!orc_coarse_shape(fine->Mt_lat, fine->Mx_lat, ctype, level, &Mt_c, &Mx_c,
                          &rot_c))
      return -1;
    const double a_f =
        fine->rotated ? sqrt(2.) / fine->Mt_lat : 1. / fine->Mt_lat;
    const double a_c = rot_c ? sqrt(2.) / Mt_c : 1. / Mt_c;
    const double mass2 = fine->gff_mu2 / (a_f * a_f);
    coarse->Mt_lat = Mt_c;
    coarse->Mx_lat = Mx_c;
    coarse->rotated = rot_c;
    coarse->gff_mu2 = a_c * a_c * mass2;
    return 0;
  }
  }
  return -1;
}

/* ================================================= stochastic hot-path pieces */

void orc_init_state(const orc_model *m, uint64_t seed, uint64_t draw,
                    uint32_t chain, double *x) {
  const int n = orc_sample_size(m);
  for (int k = 0; 2 * k < n; ++k) {
    orc_rng r;
    orc_rng_init(&r, seed, ORC_STREAM_INIT, draw, chain, k);
    double v0, v1;
    switch (m->model) {
    case ORC_ROTOR:     /* qm/rotoraction.cc:82-85 */
    case ORC_SCHWINGER: /* qft/quenchedschwingeraction.cc:198-204 */
      orc_rng_uniform2(&r, &v0, &v1);
      v0 = -M_PI + 2. * M_PI * v0;
      v1 = -M_PI + 2. * M_PI * v1;
      break;
    case ORC_GFF: /* i.i.d. N(0, 1/(4+mu2)) instead of the exact sparse-Cholesky
                     draw of qft/gffaction.cc:121-123 (start state only) */
      orc_rng_normal2(&r, &v0, &v1);
      v0 *= 1. / sqrt(4. + m->gff_mu2);
      v1 *= 1. / sqrt(4. + m->gff_mu2);
      break;
    default: /* qm/harmonicoscillatoraction.hh:155-158 */
      v0 = v1 = 0.0;
    }
    x[2 * k] = v0;
    if (2 * k + 1 < n)
      x[2 * k + 1] = v1;
  }
}

/* sampler/hmcsampler.cc:24-26: one Box-Muller pair per two dofs */
void orc_hmc_momentum(const orc_model *m, uint64_t seed, uint64_t draw,
                      uint32_t chain, double *p) {
  const int n = orc_sample_size(m);
  for (int k = 0; 2 * k < n; ++k) {
    orc_rng r;
    orc_rng_init(&r, seed, ORC_STREAM_HMC_MOMENTUM, draw, chain, k);
    double z0, z1;
    orc_rng_normal2(&r, &z0, &z1);
    p[2 * k] = z0;
    if (2 * k + 1 < n)
      p[2 * k + 1] = z1;
  }
}

/* qm/harmonicoscillatoraction.cc:38-56: L_cov = Cholesky factor (lower, row-major [M][M]) of the
 * covariance Sigma^{-1}, Sigma the cyclic tridiagonal precision matrix of the action.  The
 * reference indexes the sub-diagonal with `(i - 1) % M_lat` on unsigned integers, which is the
 * periodic neighbour only when M_lat divides 2^32; the periodic neighbour is meant and used here. */
int orc_ho_exact_factor(const orc_model *m, double *L) {
  const int M = m->M_lat;
  const double d = m->a_lat * m->m0 * m->mu2 + 2.0 * m->m0 / m->a_lat, c = -m->m0 / m->a_lat;
  double *P = (double *)calloc((size_t)M * M, sizeof(double));
  double *C = (double *)calloc((size_t)M * M, sizeof(double));
  if (!P || !C)
    return -1;
  for (int i = 0; i < M; ++i) {
    P[(size_t)i * M + i] = d;
    P[(size_t)i * M + (i + 1) % M] += c;
    P[(size_t)i * M + (i + M - 1) % M] += c;
  }
  /* Gauss-Jordan inverse (the precision matrix is symmetric positive definite: no pivoting) */
  for (int i = 0; i < M; ++i)
    C[(size_t)i * M + i] = 1.0;
  for (int k = 0; k < M; ++k) {
    const double piv = 1.0 / P[(size_t)k * M + k];
    for (int j = 0; j < M; ++j) {
      P[(size_t)k * M + j] *= piv;
      C[(size_t)k * M + j] *= piv;
    }
    for (int i = 0; i < M; ++i) {
      if (i == k)
        continue;
      const double f = P[(size_t)i * M + k];
      if (f == 0.0)
        continue;
      for (int j = 0; j < M; ++j) {
        P[(size_t)i * M + j] -= f * P[(size_t)k * M + j];
        C[(size_t)i * M + j] -= f * C[(size_t)k * M + j];
      }
    }
  }
  /* Cholesky C = L L^T */
  memset(L, 0, (size_t)M * M * sizeof(double));
  for (int j = 0; j < M; ++j) {
    double s = C[(size_t)j * M + j];
    for (int k = 0; k < j; ++k)
      s -= L[(size_t)j * M + k] * L[(size_t)j * M + k];
    if (!(s > 0.0)) {
      free(P);
      free(C);
      return -2;
    }
    const double ljj = sqrt(s);
    L[(size_t)j * M + j] = ljj;
    for (int i = j + 1; i < M; ++i) {
      double t = C[(size_t)i * M + j];
      for (int k = 0; k < j; ++k)
        t -= L[(size_t)i * M + k] * L[(size_t)j * M + k];
      L[(size_t)i * M + j] = t / ljj;
    }
  }
  free(P);
  free(C);
  return 0;
}

/* HarmonicOscillatorAction::draw (qm/harmonicoscillatoraction.cc:59-66): x = L_cov y with y
 * i.i.d. standard normal (one Box-Muller pair per two entries, stream ORC_STREAM_EXACT) */
int orc_ho_exact_draw(const orc_model *m, uint64_t seed, uint64_t draw, uint32_t chain, double *x) {
  const int M = m->M_lat;
  double *L = (double *)malloc(((size_t)M * M + M) * sizeof(double));
  if (!L)
    return -1;
  double *y = L + (size_t)M * M;
  const int rc = orc_ho_exact_factor(m, L);
  if (rc) {
    free(L);
    return rc;
  }
  for (int k = 0; 2 * k < M; ++k) {
    orc_rng r;
    orc_rng_init(&r, seed, ORC_STREAM_EXACT, draw, chain, k);
    double z0, z1;
    orc_rng_normal2(&r, &z0, &z1);
    y[2 * k] = z0;
    if (2 * k + 1 < M)
      y[2 * k + 1] = z1;
  }
  for (int i = 0; i < M; ++i) {
    double s = 0.0;
    for (int j = 0; j <= i; ++j)
      s += L[(size_t)i * M + j] * y[j];
    x[i] = s;
  }
  free(L);
  return 0;
}

/* sampler/hmcsampler.cc:22-69 */
int orc_hmc_step(const orc_model *m, int nt, double dt, uint64_t seed,
                 uint64_t draw, uint32_t chain, double *x, double *out) {
  const int n = orc_sample_size(m);
  double *p = (double *)malloc(2 * (size_t)n * sizeof(double));
  double *xt = p + n;
  orc_hmc_momentum(m, seed, draw, chain, p);
  double T_kin_cur = 0.0;
  for (int l = 0; l < n; ++l)
    T_kin_cur += p[l] * p[l];
  T_kin_cur *= 0.5;
  memcpy(xt, x, (size_t)n * sizeof(double));
  orc_leapfrog(m, nt, dt, xt, p);
  double T_kin_trial = 0.0;
  for (int l = 0; l < n; ++l)
    T_kin_trial += p[l] * p[l];
  T_kin_trial *= 0.5;
  const double S_trial = orc_action(m, xt), S_cur = orc_action(m, x);
  const double deltaH = (S_trial - S_cur) + (T_kin_trial - T_kin_cur);
  int accept = (deltaH < 0.0);
  if (!accept) {
    orc_rng r;
    double u0, u1;
    orc_rng_init(&r, seed, ORC_STREAM_HMC_ACCEPT, draw, chain, 0);
    orc_rng_uniform2(&r, &u0, &u1);
    accept = (u0 < exp(-deltaH));
  }
  if (accept)
    memcpy(x, xt, (size_t)n * sizeof(double));
  if (out) {
    out[0] = deltaH;
    out[1] = S_cur;
    out[2] = S_trial;
    out[3] = T_kin_cur;
    out[4] = T_kin_trial;
  }
  free(p);
  return accept;
}

static void heatbath_update(const orc_model *m, uint64_t seed, uint64_t draw,
                            uint32_t chain, double *x, uint32_t ell) {
  orc_rng r;
  orc_rng_init(&r, seed, ORC_STREAM_HEATBATH, draw, chain, ell);
  switch (m->model) {
  case ORC_ROTOR: { /* qm/rotoraction.cc:21-37 */
    const int M = m->M_lat;
    const double x_m = x[(ell + M - 1) % M], x_p = x[(ell + 1) % M];
    double x0, c;
    orc_W(m, x_m, x_p, &x0, &c);
    const double sigma = 2. * c;
    x[ell] = orc_mod_2pi(x0 + orc_expsin2_draw(&r, sigma));
    return;
  }
  case ORC_GFF: { /* qft/gffaction.cc:32-42 */
    const double Delta = gff_nn_sum(m, x, ell);
    const double sigma = 1. / sqrt(4. + m->gff_mu2);
    /* the vertices 2q, 2q + 1 share the block with index 2q: z0 for the even, z1 for the odd one (include/mlmcpi.h) */
    double z0, z1;
    orc_rng_init(&r, seed, ORC_STREAM_HEATBATH, draw, chain, ell & ~1u);
    orc_rng_normal2(&r, &z0, &z1);
    x[ell] = sigma * ((ell & 1u) ? z1 : z0) + Delta / (4. + m->gff_mu2);
    return;
  }
  case ORC_SCHWINGER: { /* qft/quenchedschwingeraction.cc:46-54 */
    const int Mt = m->Mt_lat, Mx = m->Mx_lat;
    int i, j, mu;
    double theta_p, theta_m;
    orc_link_lin2cart(Mt, Mx, ell, &i, &j, &mu);
    staple_angles(x, Mt, Mx, i, j, mu, &theta_p, &theta_m);
    /* paired variates (include/mlmcpi.h, stream convention): the links of a colour in a row are numbered
     * n = i (mu = 0) or i / 2 (mu = 1); the block of the even link of the pair (2p, 2p + 1) gives the first
     * attempt of both, further attempts come from the link's own stream (even link: from call 2) */
    {
      const int n = (mu == 0) ? i : i / 2;
      const int is_first = ((n & 1) == 0);
      const int i_first = is_first ? i : (mu == 0 ? i - 1 : i - 2);
      orc_rng rA;
      double z[2], u[2], first[2];
      orc_rng_init(&rA, seed, ORC_STREAM_HEATBATH, draw, chain, (uint32_t)(2 * (Mt * j + i_first) + mu));
      orc_rng_normal2(&rA, &z[0], &z[1]);
      orc_rng_uniform2(&rA, &u[0], &u[1]);
      first[0] = z[is_first ? 0 : 1];
      first[1] = u[is_first ? 0 : 1];
      x[ell] = expcos_draw_impl(is_first ? &rA : &r, m->beta, theta_p, theta_m, first);
    }
    return;
  }
  }
}

void orc_heatbath_sweep_coloured(const orc_model *m, uint64_t seed, uint64_t draw,
                                 uint32_t chain, double *x) {
  const int n = orc_sample_size(m);
  const int nc = orc_n_colours(m);
  for (int c = 0; c < nc; ++c)
    for (int ell = 0; ell < n; ++ell)
      if (orc_colour_of(m, ell) == c)
        heatbath_update(m, seed, draw, chain, x, ell);
}

static double uniform_angle(orc_rng *r, double *second) {
  double u0, u1;
  orc_rng_uniform2(r, &u0, &u1);
  if (second)
    *second = -M_PI + 2. * M_PI * u1;
  return -M_PI + 2. * M_PI * u0;
}

/* ConditionedFineAction::fill_fine_points of each model */
void orc_fill(const orc_model *fine, uint64_t seed, uint64_t draw, uint32_t chain,
              double *x) {
  switch (fine->model) {
  case ORC_HO:
  case ORC_QUARTIC:
  case ORC_ROTOR: {
    /* qm/gaussianconditionedfineaction.cc:7-24, qm/rotorconditionedfineaction.cc:7-24 */
    const int M = fine->M_lat;
    for (int j = 0; j < M / 2; ++j) {
      const double x_m = x[2 * j], x_p = x[(2 * j + 2) % M];
      double x0, c;
      orc_W(fine, x_m, x_p, &x0, &c);
      orc_rng r;
      orc_rng_init(&r, seed, ORC_STREAM_FILL1, draw, chain, j);
      if (fine->model == ORC_ROTOR) {
        const double sigma = 2. * c;
        x[2 * j + 1] = orc_mod_2pi(x0 + orc_expsin2_draw(&r, sigma));
      } else {
        const double sigma = 1. / sqrt(c);
        double z0, z1;
        orc_rng_normal2(&r, &z0, &z1);
        x[2 * j + 1] = x0 + z0 * sigma;
      }
    }
    return;
  }
  case ORC_GFF: { /* qft/gffconditionedfineaction.cc:7-25 */
    const int N = orc_sample_size(fine);
    uint32_t *buf = (uint32_t *)malloc(3 * (size_t)N * sizeof(uint32_t));
    int counts[2];
    const int level = fine->rotated ? 1 : 0;
    if (orc_coarsening_lists(fine->Mt_lat, fine->Mx_lat, fine->coarsening, level,
                             buf, buf + N, buf + 2 * N, counts) == 0) {
      const double sigma = 1. / sqrt(4. + fine->gff_mu2);
      for (int k = 0; k < counts[1]; ++k) {
        const uint32_t ell = buf[N + k];
        const double Delta = gff_nn_sum(fine, x, ell);
        orc_rng r;
        double z0, z1;
        orc_rng_init(&r, seed, ORC_STREAM_FILL1, draw, chain, ell);
        orc_rng_normal2(&r, &z0, &z1);
        x[ell] = sigma * (z0 + sigma * Delta);
      }
    }
    free(buf);
    return;
  }
  case ORC_SCHWINGER: {
    const int Mt = fine->Mt_lat, Mx = fine->Mx_lat;
    const double beta = fine->beta;
    if (fine->coarsening == ORC_COARSEN_BOTH) {
      /* qft/quenchedschwingerconditionedfineaction.cc:7-78; cell index
       * c = (Mt/2) j + i numbers the Philox streams */
      for (int i = 0; i < Mt / 2; ++i) /* STEP 1 (:14-31) */
        for (int j = 0; j < Mx / 2; ++j) {
          orc_rng r;
          orc_rng_init(&r, seed, ORC_STREAM_FILL1, draw, chain, (Mt / 2) * j + i);
          double dtheta_s;
          const double dtheta_t = uniform_angle(&r, &dtheta_s);
          x[LNK(2 * i, 2 * j, 0)] = orc_mod_2pi(x[LNK(2 * i, 2 * j, 0)] + dtheta_t);
          x[LNK(2 * i + 1, 2 * j, 0)] =
              orc_mod_2pi(x[LNK(2 * i + 1, 2 * j, 0)] - dtheta_t);
          x[LNK(2 * i, 2 * j, 1)] = orc_mod_2pi(x[LNK(2 * i, 2 * j, 1)] + dtheta_s);
          x[LNK(2 * i, 2 * j + 1, 1)] =
              orc_mod_2pi(x[LNK(2 * i, 2 * j + 1, 1)] - dtheta_s);
        }
      for (int i = 0; i < Mt / 2; ++i) /* STEP 2 (:33-61) */
        for (int j = 0; j < Mx / 2; ++j) {
          const double theta_p = orc_mod_2pi(
              x[LNK(2 * i + 1, 2 * j, 0)] + x[LNK(2 * i + 2, 2 * j, 1)] +
              x[LNK(2 * i + 2, 2 * j + 1, 1)] - x[LNK(2 * i + 1, 2 * j + 2, 0)]);
          const double theta_m = orc_mod_2pi(
              x[LNK(2 * i, 2 * j, 1)] + x[LNK(2 * i, 2 * j + 1, 1)] +
              x[LNK(2 * i, 2 * j + 2, 0)] - x[LNK(2 * i, 2 * j, 0)]);
          orc_rng r;
          orc_rng_init(&r, seed, ORC_STREAM_FILL2, draw, chain, (Mt / 2) * j + i);
          /* first call of the stream: (split angle, mode selector of the approximate
           * distribution) */
          double u0, u1;
          orc_rng_uniform2(&r, &u0, &u1);
          const double dtheta = -M_PI + 2. * M_PI * u0;
          double theta_tilde;
          if (beta <= 8.0)
            theta_tilde = orc_besselproduct_draw(&r, beta, theta_p, theta_m);
          else
            theta_tilde = approxbessel_draw_xi(&r, beta, theta_p, theta_m, u1);
          x[LNK(2 * i + 1, 2 * j, 1)] = orc_mod_2pi(0.5 * theta_tilde + dtheta);
          x[LNK(2 * i + 1, 2 * j + 1, 1)] = orc_mod_2pi(0.5 * theta_tilde - dtheta);
        }
      /* STEP 3 (:63-77).  Variates: the two horizontal interior links (2ic, 2j+1, 0), (2ic+1, 2j+1, 0) of a
       * coarse cell share one normal pair and one uniform pair -- calls 0, 1 of the stream of the first link --
       * for their FIRST attempts, (z0, u0) and (z1, u1); further attempts continue on the link's own stream
       * (first link: calls 2, 3, ...; second link: its stream from call 0). */
      for (int ic = 0; ic < Mt / 2; ++ic)
        for (int j = 0; j < Mx / 2; ++j) {
          orc_rng r0, r1;
          double z[2], u[2];
          orc_rng_init(&r0, seed, ORC_STREAM_FILL3, draw, chain, Mt * j + 2 * ic);
          orc_rng_normal2(&r0, &z[0], &z[1]);
          orc_rng_uniform2(&r0, &u[0], &u[1]);
          orc_rng_init(&r1, seed, ORC_STREAM_FILL3, draw, chain, Mt * j + 2 * ic + 1);
          for (int h = 0; h < 2; ++h) {
            const int i = 2 * ic + h;
            const double theta_p = orc_mod_2pi(
                x[LNK(i, 2 * j, 0)] + x[LNK(i + 1, 2 * j, 1)] - x[LNK(i, 2 * j, 1)]);
            const double theta_m =
                orc_mod_2pi(x[LNK(i, 2 * j + 1, 1)] + x[LNK(i, 2 * j + 2, 0)] -
                            x[LNK(i + 1, 2 * j + 1, 1)]);
            const double first[2] = {z[h], u[h]};
            x[LNK(i, 2 * j + 1, 0)] = expcos_draw_impl(h == 0 ? &r0 : &r1, beta, theta_p, theta_m, first);
          }
        }
    } else if (fine->coarsening == ORC_COARSEN_TEMPORAL) {
      /* qft/quenchedschwingerconditionedfineaction.cc:147-174 */
      for (int i = 0; i < Mt / 2; ++i)
        for (int j = 0; j < Mx; ++j) {
          orc_rng r;
          orc_rng_init(&r, seed, ORC_STREAM_FILL1, draw, chain, (Mt / 2) * j + i);
          const double dtheta = uniform_angle(&r, NULL);
          x[LNK(2 * i, j, 0)] = orc_mod_2pi(x[LNK(2 * i, j, 0)] + dtheta);
          x[LNK(2 * i + 1, j, 0)] = orc_mod_2pi(x[LNK(2 * i + 1, j, 0)] - dtheta);
        }
      for (int i = 0; i < Mt / 2; ++i)
        for (int j = 0; j < Mx; ++j) {
          const double theta_p = orc_mod_2pi(
              x[LNK(2 * i, j, 1)] + x[LNK(2 * i, j + 1, 0)] - x[LNK(2 * i, j, 0)]);
          const double theta_m =
              orc_mod_2pi(x[LNK(2 * i + 1, j, 0)] + x[LNK(2 * i + 2, j, 1)] -
                          x[LNK(2 * i + 1, j + 1, 0)]);
          orc_rng r;
          orc_rng_init(&r, seed, ORC_STREAM_FILL3, draw, chain, (Mt / 2) * j + i);
          x[LNK(2 * i + 1, j, 1)] = orc_expcos_draw(&r, beta, theta_p, theta_m);
        }
    } else if (fine->coarsening == ORC_COARSEN_SPATIAL) {
      /* qft/quenchedschwingerconditionedfineaction.cc:175-203 */
      for (int i = 0; i < Mt; ++i)
        for (int j = 0; j < Mx / 2; ++j) {
          orc_rng r;
          orc_rng_init(&r, seed, ORC_STREAM_FILL1, draw, chain, Mt * j + i);
          const double dtheta = uniform_angle(&r, NULL);
          x[LNK(i, 2 * j, 1)] = orc_mod_2pi(x[LNK(i, 2 * j, 1)] + dtheta);
          x[LNK(i, 2 * j + 1, 1)] = orc_mod_2pi(x[LNK(i, 2 * j + 1, 1)] - dtheta);
        }
      for (int i = 0; i < Mt; ++i)
        for (int j = 0; j < Mx / 2; ++j) {
          const double theta_p = orc_mod_2pi(
              x[LNK(i, 2 * j, 0)] + x[LNK(i + 1, 2 * j, 1)] - x[LNK(i, 2 * j, 1)]);
          const double theta_m =
              orc_mod_2pi(x[LNK(i, 2 * j + 1, 1)] + x[LNK(i, 2 * j + 2, 0)] -
                          x[LNK(i + 1, 2 * j + 1, 1)]);
          orc_rng r;
          orc_rng_init(&r, seed, ORC_STREAM_FILL3, draw, chain, Mt * j + i);
          x[LNK(i, 2 * j + 1, 0)] = orc_expcos_draw(&r, beta, theta_p, theta_m);
        }
    }
    return;
  }
  }
}

/* montecarlo/twolevelmetropolisstep.cc:35-89 */
int orc_twolevel_step(const orc_model *fine, const orc_model *coarse,
                      uint64_t seed, uint64_t draw, uint32_t chain,
                      const double *x_coarse, double *x_fine, double *S_fine,
                      double *S_cond, double *out) {
  const int nf = orc_sample_size(fine), nc = orc_sample_size(coarse);
  double *theta_prime = (double *)malloc(((size_t)nf + nc) * sizeof(double));
  double *theta_fine_C = theta_prime + nf;
  /* theta_prime persists between draws in the reference; for the models here
   * prolongation + fill-in overwrite every entry, so its history is immaterial */
  memcpy(theta_prime, x_fine, (size_t)nf * sizeof(double));
  orc_prolong(fine, x_coarse, theta_prime);
  orc_fill(fine, seed, draw, chain, theta_prime);
  const double fine_action_theta_prime = orc_action(fine, theta_prime);
  const double deltaS_fine = fine_action_theta_prime - *S_fine;
  orc_restrict(fine, x_fine, theta_fine_C);
  const double deltaS_coarse =
      orc_action(coarse, theta_fine_C) - orc_action(coarse, x_coarse);
  const double cond_theta_prime = orc_cond_action(fine, theta_prime);
  const double deltaS_trial = *S_cond - cond_theta_prime;
  const double deltaS = deltaS_fine + deltaS_coarse + deltaS_trial;
  int accept = (deltaS < 0.0);
  if (!accept) {
    orc_rng r;
    double u0, u1;
    orc_rng_init(&r, seed, ORC_STREAM_TWOLEVEL_ACCEPT, draw, chain, 0);
    orc_rng_uniform2(&r, &u0, &u1);
    accept = (u0 < exp(-deltaS));
  }
  if (accept) {
    memcpy(x_fine, theta_prime, (size_t)nf * sizeof(double));
    *S_fine = fine_action_theta_prime;
    *S_cond = cond_theta_prime;
  }
  if (out) {
    out[0] = deltaS_fine;
    out[1] = deltaS_coarse;
    out[2] = deltaS_trial;
  }
  free(theta_prime);
  return accept;
}

/* ================================================================ cluster */

/* ClusterSampler::single_cluster_update1d + process_link1d, sampler/clustersampler.cc:88-132,
 * with RotorAction::S_ell / new_reflection / flip, action/qm/rotoraction.hh:226-253.
 * Philox stream CLUSTER, draw = update: index 0 = (xbar, start site); index 1 + k = the two
 * uniforms of link k (sites k, k+1 mod M): the first is used when the link is processed in the
 * forward walk, the second in the backward walk. */
static int process_link1d(const orc_model *m, double *x, double xbar, uint64_t seed, uint64_t update,
                          uint32_t chain, int i, int direction, int *i_next) {
  const int M = m->M_lat;
  const int i_neighbour = (i + direction + M) % M;
  const int link = direction > 0 ? i : i_neighbour;
  orc_rng r;
  double uf, ub;
  orc_rng_init(&r, seed, ORC_STREAM_CLUSTER, update, chain, 1u + (uint32_t)link);
  orc_rng_uniform2(&r, &uf, &ub);
  const double Sell =
      -2.0 * m->m0 / m->a_lat * cos(x[i] - xbar) * cos(x[i_neighbour] - xbar);
  const double p_connect = 1. - exp(fmin(0, -Sell));
  const int bonded = ((direction > 0 ? uf : ub) < p_connect);
  if (bonded)
    x[i_neighbour] = orc_mod_2pi(M_PI + 2. * xbar - x[i_neighbour]);
  *i_next = i_neighbour;
  return bonded;
}
void orc_cluster_update(const orc_model *m, uint64_t seed, uint64_t update0, int n_updates,
                        uint32_t chain, double *x) {
  const int M = m->M_lat;
  for (int u = 0; u < n_updates; ++u) {
    orc_rng r;
    orc_rng_init(&r, seed, ORC_STREAM_CLUSTER, update0 + u, chain, 0);
    double u0, u1;
    orc_rng_uniform2(&r, &u0, &u1);
    const double xbar = -M_PI + 2. * M_PI * u0;
    int i0 = (int)(u1 * M);
    if (i0 >= M)
      i0 = M - 1;
    x[i0] = orc_mod_2pi(M_PI + 2. * xbar - x[i0]);
    int i_p = i0, i_last_p, bonded;
    do {
      i_last_p = i_p;
      bonded = process_link1d(m, x, xbar, seed, update0 + u, chain, i_p, +1, &i_p);
    } while ((i_p != i0) && bonded);
    int i_m = i0;
    do {
      bonded = process_link1d(m, x, xbar, seed, update0 + u, chain, i_m, -1, &i_m);
    } while ((i_m != i_last_p) && bonded);
  }
}

/* QuenchedSchwingerClusterSampler::draw, sampler/quenchedschwingerclustersampler.cc:50-82;
 * gauge angle of site (i,j): Philox stream GAUGE, index Mt*j + i */
void orc_schwinger_from_cluster(const orc_model *m, uint64_t seed, uint64_t draw, uint32_t chain,
                                const double *psi, double *x) {
  const int Mt = m->Mt_lat, Mx = m->Mx_lat;
  for (int l = 0; l < 2 * Mt * Mx; ++l)
    x[l] = 0.0;
  int i_lin = 0;
  for (int i = 0; i < Mt - 1; ++i)
    for (int j = 0; j < Mx; ++j) {
      x[LNK(i + 1, j, 1)] = x[LNK(i, j, 1)] + psi[i_lin + 1] - psi[i_lin];
      i_lin++;
    }
  for (int j = 0; j < Mx - 1; ++j) {
    x[LNK(Mt - 1, j + 1, 0)] =
        x[LNK(Mt - 1, j, 0)] - x[LNK(Mt - 1, j, 1)] - psi[i_lin + 1] + psi[i_lin];
    i_lin++;
  }
  for (int i = 0; i < Mt; ++i)
    for (int j = 0; j < Mx; ++j) {
      orc_rng r;
      orc_rng_init(&r, seed, ORC_STREAM_GAUGE, draw, chain, (uint32_t)(Mt * j + i));
      const double theta = uniform_angle(&r, NULL);
      x[LNK(i, j, 0)] = orc_mod_2pi(x[LNK(i, j, 0)] + theta);
      x[LNK(i - 1, j, 0)] = orc_mod_2pi(x[LNK(i - 1, j, 0)] - theta);
      x[LNK(i, j, 1)] = orc_mod_2pi(x[LNK(i, j, 1)] + theta);
      x[LNK(i, j - 1, 1)] = orc_mod_2pi(x[LNK(i, j - 1, 1)] - theta);
    }
}

/* ============================================================== statistics */

/* common/statistics.cc:4-97 (single rank): out6 = {average, variance,
 * variance_error, tau_int, error, samples} */
void orc_statistics(int k_max, int n, const double *q, double *out6) {
  double *S_k = (double *)calloc((size_t)k_max, sizeof(double));
  double *Q_k = (double *)calloc((size_t)k_max, sizeof(double));
  int n_window = 0;
  unsigned int n_samples = 0;
  double avg = 0, avg1 = 0, avg2 = 0, avg3 = 0, avg4 = 0;
  for (int s = 0; s < n; ++s) {
    const double Q = q[s];
    n_samples++;
    /* push_front / pop_back */
    for (int k = (n_window < k_max ? n_window : k_max - 1); k > 0; --k)
      Q_k[k] = Q_k[k - 1];
    Q_k[0] = Q;
    if (n_window < k_max)
      n_window++;
    avg = ((n_samples - 1.0) * avg + Q) / (1.0 * n_samples);
    avg1 = ((n_samples - 1.0) * avg1 + Q) / (1.0 * n_samples);
    avg2 = ((n_samples - 1.0) * avg2 + Q * Q) / (1.0 * n_samples);
    avg3 = ((n_samples - 1.0) * avg3 + Q * Q * Q) / (1.0 * n_samples);
    avg4 = ((n_samples - 1.0) * avg4 + Q * Q * Q * Q) / (1.0 * n_samples);
    for (int k = 0; k < n_window; ++k) {
      const unsigned int N_k = n_samples - k;
      S_k[k] = ((N_k - 1.0) * S_k[k] + Q_k[0] * Q_k[k]) / (1.0 * N_k);
    }
  }
  const double variance =
      1.0 * n_samples / (n_samples - 1.0) * (S_k[0] - avg1 * avg1);
  const double variance_error =
      sqrt(1.0 / n_samples *
           (avg4 - 4 * avg1 * avg3 + 8 * avg1 * avg1 * avg2 - avg2 * avg2 -
            4 * avg1 * avg1 * avg1 * avg1));
  double tau_int_tmp = 0.0;
  const double C0 = S_k[0] - avg1 * avg1;
  for (int k = 1; k < k_max; ++k)
    tau_int_tmp += (1. - k / (1.0 * n_samples)) * (S_k[k] - avg1 * avg1);
  const double tau_int = fmax(1.0, 1.0 + 2.0 * tau_int_tmp / C0);
  out6[0] = avg;
  out6[1] = variance;
  out6[2] = variance_error;
  out6[3] = tau_int;
  out6[4] = sqrt(tau_int * variance / (1.0 * n_samples));
  out6[5] = n_samples;
  free(S_k);
  free(Q_k);
}
