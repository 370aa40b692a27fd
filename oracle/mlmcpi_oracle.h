/* TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * Plain-C, single-chain, CPU restatement of the sampler inner loop of
 * eikehmueller/mlmcpathintegral.  Every function cites
 * the reference file:line it follows.  All states are in the REFERENCE's own
 * layout (links: ell = 2*Mt*j + 2*i + mu, lattice2d.hh:348-353; vertices:
 * lattice2d.hh:230-245; 1-D paths contiguous).
 *
 * Pinning: tests/test_oracle_cpu.py checks these functions against the
 * golden vectors in tests/golden/, which tools/make_golden.py recorded from the
 * reference's own translation units (oracle/_ref, built by oracle/Makefile).
 *
 * Random numbers: the reference uses std::mt19937_64 + libstdc++ distributions
 * with hard-coded seeds, which a GPU cannot reproduce.  The
 * stochastic functions here restate the reference's sampling ALGORITHMS
 * (proposal, envelope, acceptance test, output map) on top of the counter-based
 * Philox4x32-10 stream the CUDA kernels use, so that CUDA and oracle can be
 * compared draw by draw; the distributions themselves are validated
 * statistically against the reference's pdfs and its own draw() routines.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use
 * this library.  The product never links it.
 */
#ifndef MLMCPI_ORACLE_H
#define MLMCPI_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* model kinds (same numbering as include/mlmcpi.h) */
enum { ORC_HO = 0, ORC_QUARTIC = 1, ORC_ROTOR = 2, ORC_SCHWINGER = 3, ORC_GFF = 4 };
/* coarsening types: lattice/lattice2d.hh:18-26 */
enum { ORC_COARSEN_BOTH = 0, ORC_COARSEN_TEMPORAL = 1, ORC_COARSEN_SPATIAL = 2,
       ORC_COARSEN_ALTERNATE = 3, ORC_COARSEN_ROTATE = 4 };
/* QoIs */
enum { ORC_QOI_X2 = 0, ORC_QOI_ROTOR_CHI = 1, ORC_QOI_SCHWINGER_CHI = 2,
       ORC_QOI_AVG_PLAQUETTE = 3, ORC_QOI_PHI2 = 4 };
/* random streams (must equal the MLMCPI_STREAM_* constants of the product) */
enum { ORC_STREAM_INIT = 1, ORC_STREAM_HMC_MOMENTUM = 2, ORC_STREAM_HMC_ACCEPT = 3,
       ORC_STREAM_HEATBATH = 4, ORC_STREAM_FILL1 = 5, ORC_STREAM_FILL2 = 6,
       ORC_STREAM_FILL3 = 7, ORC_STREAM_TWOLEVEL_ACCEPT = 8, ORC_STREAM_CLUSTER = 9,
       ORC_STREAM_GAUGE = 10, ORC_STREAM_EXACT = 11 };

typedef struct {
  int model;
  int M_lat;          /* 1-D: number of sites */
  int Mt_lat, Mx_lat; /* 2-D */
  int rotated;        /* 2-D vertex lattices: rotated level of CoarsenRotate */
  int coarsening;     /* 2-D: how THIS level is coarsened to the next one:
                         BOTH / TEMPORAL / SPATIAL (ALTERNATE resolved by level),
                         ROTATE */
  double a_lat;       /* 1-D lattice spacing T/M */
  double T_final;     /* 1-D: total time T (qoi/qm/qoisusceptibility.hh:34) */
  double m0, mu2, lambda, x0; /* QM couplings */
  double beta;        /* Schwinger */
  double gff_mu2;     /* GFF: a^2 m^2 */
} orc_model;

/* ---- RNG ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
typedef struct { uint32_t c0, c1, c2, a, k0, k1; } orc_rng;
void orc_rng_init(orc_rng *r, uint64_t seed, int stream, uint64_t draw,
                  uint32_t chain, uint32_t index);
void orc_rng_uniform2(orc_rng *r, double *u0, double *u1);
void orc_rng_normal2(orc_rng *r, double *z0, double *z1);

/* ---- scalar maths ---- */
double orc_mod_2pi(double x);
double orc_bessel_I0(double x);
double orc_bessel_I0_scaled(double x);
double orc_fast_bessel_I0_scaled(double z);

/* ---- lattice index maps ---- */
uint32_t orc_vertex_cart2lin(int Mt, int Mx, int rotated, int i, int j);
void orc_vertex_lin2cart(int Mt, int Mx, int rotated, uint32_t ell, int *i, int *j);
uint32_t orc_link_cart2lin(int Mt, int Mx, int i, int j, int mu);
void orc_link_lin2cart(int Mt, int Mx, uint32_t ell, int *i, int *j, int *mu);
int orc_n_vertices(int Mt, int Mx, int rotated);
void orc_neighbours(int Mt, int Mx, int rotated, uint32_t ell, uint32_t nb[8]);
/* coarse lattice shape of a level; returns 0 if it cannot be coarsened */
int orc_coarse_shape(int Mt, int Mx, int ctype, int level, int *Mt_c, int *Mx_c,
                     int *rot_c);
/* coarse/fine-only vertex lists and fine->coarse map (sorted, as the reference) */
int orc_coarsening_lists(int Mt, int Mx, int ctype, int level, uint32_t *coarse,
                         uint32_t *fineonly, uint32_t *map_vals, int *counts);

/* ---- deterministic hot-path functions ---- */
int orc_sample_size(const orc_model *m);
double orc_action(const orc_model *m, const double *x);
void orc_force(const orc_model *m, const double *x, double *p);
void orc_W(const orc_model *m, double x_m, double x_p, double *Wmin, double *Wcurv);
void orc_overrelax_update(const orc_model *m, double *x, uint32_t ell);
void orc_overrelax_sweep_lex(const orc_model *m, double *x);
/* coloured sweep: the order the CUDA kernels use (same per-dof update) */
int orc_n_colours(const orc_model *m);
int orc_colour_of(const orc_model *m, uint32_t ell);
void orc_overrelax_sweep_coloured(const orc_model *m, double *x);
void orc_prolong(const orc_model *fine, const double *xc, double *x);
void orc_restrict(const orc_model *fine, const double *xf, double *xc);
double orc_cond_action(const orc_model *fine, const double *x);
double orc_qoi(const orc_model *m, int qoi, const double *x, int64_t *Qint);
void orc_leapfrog(const orc_model *m, int nt, double dt, double *x, double *p);
/* coarse-level couplings: renorm 0 none, 1 perturbative (2 nonperturbative: host only) */
int orc_coarse_model(const orc_model *fine, int renorm, int level, int ctype,
                     double T_final, orc_model *coarse);

/* ---- distributions (pdfs follow the reference evaluate() methods) ---- */
double orc_expsin2_pdf(double x, double sigma);
double orc_expcos_pdf(double beta, double x, double x_p, double x_m);
void orc_besselproduct_alpha(double beta, double alphaZ[17]);
double orc_besselproduct_Znorm_inv(const double alphaZ[17], double phi, int rescaled);
double orc_besselproduct_pdf(double beta, double x, double x_p, double x_m);
double orc_approxbessel_pdf(double beta, double x, double x_p, double x_m);
double orc_expsin2_draw(orc_rng *r, double sigma);
/* exact (Cholesky) sampler of the harmonic oscillator, qm/harmonicoscillatoraction.cc:38-66 */
int orc_ho_exact_factor(const orc_model *m, double *L);
int orc_ho_exact_draw(const orc_model *m, uint64_t seed, uint64_t draw, uint32_t chain, double *x);
/* 0: the reference's envelope (default); 1, 2: the product's tighter envelopes (same pdf) */
void orc_set_expcos_envelope(int envelope);
double orc_expcos_draw(orc_rng *r, double beta, double x_p, double x_m);
double orc_besselproduct_draw(orc_rng *r, double beta, double x_p, double x_m);
double orc_approxbessel_draw(orc_rng *r, double beta, double x_p, double x_m);

/* ---- stochastic hot-path functions (Philox streams of the product) ---- */
void orc_init_state(const orc_model *m, uint64_t seed, uint64_t draw, uint32_t chain, double *x);
void orc_hmc_momentum(const orc_model *m, uint64_t seed, uint64_t draw, uint32_t chain, double *p);
/* one HMC single_step; returns accept flag; out = {deltaH, S_cur, S_trial, T_cur, T_trial} */
int orc_hmc_step(const orc_model *m, int nt, double dt, uint64_t seed, uint64_t draw,
                 uint32_t chain, double *x, double *out);
void orc_heatbath_sweep_coloured(const orc_model *m, uint64_t seed, uint64_t draw,
                                 uint32_t chain, double *x);
void orc_fill(const orc_model *fine, uint64_t seed, uint64_t draw, uint32_t chain, double *x);
/* one TwoLevelMetropolisStep::draw; S_fine/S_cond are the cached values for x_fine
 * (updated on accept); out = {dS_fine, dS_coarse, dS_trial}; returns accept */
int orc_twolevel_step(const orc_model *fine, const orc_model *coarse, uint64_t seed,
                      uint64_t draw, uint32_t chain, const double *x_coarse,
                      double *x_fine, double *S_fine, double *S_cond, double *out);

/* ---- cluster samplers (sampler/clustersampler.cc, quenchedschwingerclustersampler.cc) ---- */
void orc_cluster_update(const orc_model *rotor, uint64_t seed, uint64_t update0, int n_updates,
                        uint32_t chain, double *x);
void orc_schwinger_from_cluster(const orc_model *m, uint64_t seed, uint64_t draw, uint32_t chain,
                                const double *psi, double *x);

/* ---- statistics (common/statistics.cc) ---- */
void orc_statistics(int k_max, int n, const double *q, double *out6);

#ifdef __cplusplus
}
#endif
#endif
