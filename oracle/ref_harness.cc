/* TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into the product.
 *
 * extern "C" harness around the REFERENCE's own, unmodified translation units
 * (compiled from /root/reference/src by oracle/Makefile against the header
 * shims in oracle/shim/).  It exposes the reference classes on the hot path
 *  through plain pointers so that
 *   - tools/make_golden.py can record known-answer vectors in tests/golden/
 *   - tests/ can validate the C restatement (oracle/mlmcpi_oracle.c) against it
 *   - bench.py --impl reference can time the reference CPU path.
 * This file contains no algorithm of its own apart from the loop bodies that
 * drive the reference's per-site virtuals in the order the reference samplers
 * do (each cites the reference lines it mirrors).
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "action/action.hh"
#include "action/conditionedfineaction.hh"
#include "action/qft/gffaction.hh"
#include "action/qft/gffconditionedfineaction.hh"
#include "action/qft/quenchedschwingeraction.hh"
#include "action/qft/quenchedschwingerconditionedfineaction.hh"
#include "action/qm/gaussianconditionedfineaction.hh"
#include "action/qm/harmonicoscillatoraction.hh"
#include "action/qm/quarticoscillatoraction.hh"
#include "action/qm/rotoraction.hh"
#include "action/qm/rotorconditionedfineaction.hh"
#include "common/auxilliary.hh"
#include "common/fastbessel.hh"
#include "common/samplestate.hh"
#include "common/statistics.hh"
#include "distribution/approximatebesselproductdistribution.hh"
#include "distribution/besselproductdistribution.hh"
#include "distribution/expcosdistribution.hh"
#include "distribution/expsin2distribution.hh"
#include "lattice/lattice1d.hh"
#include "lattice/lattice2d.hh"
#include "montecarlo/twolevelmetropolisstep.hh"
#include "qoi/qft/qoi2dphisquared.hh"
#include "qoi/qft/qoi2dsusceptibility.hh"
#include "qoi/qft/qoiavgplaquette.hh"
#include "qoi/qm/qoisusceptibility.hh"
#include "qoi/qm/qoixsquared.hh"

namespace {

typedef std::shared_ptr<SampleState> StatePtr;

StatePtr make_state(const double *x, const unsigned int n) {
  StatePtr s = std::make_shared<SampleState>(n);
  if (x)
    std::memcpy(s->data.data(), x, n * sizeof(double));
  return s;
}
void read_state(const StatePtr s, double *x) {
  std::memcpy(x, s->data.data(), s->data.size() * sizeof(double));
}

enum Kind { HO = 0, QUARTIC = 1, ROTOR = 2, SCHWINGER = 3, GFF = 4 };

struct ActionHolder {
  int kind;
  std::shared_ptr<Action> action;
  std::shared_ptr<Lattice1D> lattice1d;
  std::shared_ptr<Lattice2D> lattice2d;
};

struct CondHolder {
  std::shared_ptr<ConditionedFineAction> cond;
};

std::shared_ptr<Lattice2D> lattice2d_at_level(const unsigned int Mt,
                                              const unsigned int Mx,
                                              const int ctype, const int level) {
  std::shared_ptr<Lattice2D> lat =
      std::make_shared<Lattice2D>(Mt, Mx, (CoarseningType)ctype);
  for (int l = 0; l < level; ++l) {
    lat = lat->get_coarse_lattice();
    if (lat == nullptr)
      return nullptr;
  }
  return lat;
}

} // namespace

extern "C" {

/* ---------------------------------------------------------------- lattice */

/* out = {Mt, Mx, rotated, Nvertices, Nedges, has_coarse}; returns 0 or -1 */
int ref_lattice2d_info(unsigned Mt, unsigned Mx, int ctype, int level,
                       int *out) {
  auto lat = lattice2d_at_level(Mt, Mx, ctype, level);
  if (lat == nullptr)
    return -1;
  out[0] = lat->getMt_lat();
  out[1] = lat->getMx_lat();
  out[2] = lat->is_rotated();
  out[3] = lat->getNvertices();
  out[4] = lat->getNedges();
  out[5] = (lat->get_coarse_lattice() != nullptr);
  return 0;
}

/* cart2lin over i in [imin,imax), j in [jmin,jmax) (row-major in (i,j));
 * lin2cart as (i,j) pairs; neighbours as 8 entries per vertex */
int ref_lattice2d_vertex_maps(unsigned Mt, unsigned Mx, int ctype, int level,
                              int imin, int imax, int jmin, int jmax,
                              unsigned *cart2lin, int *lin2cart,
                              unsigned *neighbours) {
  auto lat = lattice2d_at_level(Mt, Mx, ctype, level);
  if (lat == nullptr)
    return -1;
  size_t k = 0;
  for (int i = imin; i < imax; ++i)
    for (int j = jmin; j < jmax; ++j, ++k) {
      if (lat->is_rotated() && (((i + j) % 2) != 0))
        cart2lin[k] = 0xFFFFFFFFu;
      else
        cart2lin[k] = lat->vertex_cart2lin(i, j);
    }
  const auto &nb = lat->get_neighbour_vertices();
  for (unsigned ell = 0; ell < lat->getNvertices(); ++ell) {
    int i, j;
    lat->vertex_lin2cart(ell, i, j);
    lin2cart[2 * ell] = i;
    lin2cart[2 * ell + 1] = j;
    for (int q = 0; q < 8; ++q)
      neighbours[8 * ell + q] = nb[ell][q];
  }
  return 0;
}

int ref_lattice2d_link_maps(unsigned Mt, unsigned Mx, int ctype, int level,
                            int imin, int imax, int jmin, int jmax,
                            unsigned *cart2lin, int *lin2cart) {
  auto lat = lattice2d_at_level(Mt, Mx, ctype, level);
  if (lat == nullptr || lat->is_rotated())
    return -1;
  size_t k = 0;
  for (int i = imin; i < imax; ++i)
    for (int j = jmin; j < jmax; ++j)
      for (int mu = 0; mu < 2; ++mu, ++k)
        cart2lin[k] = lat->link_cart2lin(i, j, mu);
  for (unsigned ell = 0; ell < lat->getNedges(); ++ell) {
    int i, j, mu;
    lat->link_lin2cart(ell, i, j, mu);
    lin2cart[3 * ell] = i;
    lin2cart[3 * ell + 1] = j;
    lin2cart[3 * ell + 2] = mu;
  }
  return 0;
}

/* counts = {n_coarse, n_fineonly, n_map}; buffers must hold Nvertices each */
int ref_lattice2d_coarsening(unsigned Mt, unsigned Mx, int ctype, int level,
                             unsigned *coarse, unsigned *fineonly,
                             unsigned *map_keys, unsigned *map_vals,
                             int *counts) {
  auto lat = lattice2d_at_level(Mt, Mx, ctype, level);
  if (lat == nullptr)
    return -1;
  const auto &c = lat->get_coarse_vertices();
  const auto &f = lat->get_fineonly_vertices();
  const auto &m = lat->get_fine2coarse_map();
  std::copy(c.begin(), c.end(), coarse);
  std::copy(f.begin(), f.end(), fineonly);
  size_t k = 0;
  for (auto it = m.begin(); it != m.end(); ++it, ++k) {
    map_keys[k] = it->first;
    map_vals[k] = it->second;
  }
  counts[0] = c.size();
  counts[1] = f.size();
  counts[2] = m.size();
  return 0;
}

/* out_d = {a_lat}; neighbours 2 per vertex; returns M of the coarse lattice */
int ref_lattice1d(unsigned M, double T, double *out_d, unsigned *neighbours) {
  Lattice1D lat(M, T);
  out_d[0] = lat.geta_lat();
  const auto &nb = lat.get_neighbour_vertices();
  for (unsigned ell = 0; ell < M; ++ell) {
    neighbours[2 * ell] = nb[ell][0];
    neighbours[2 * ell + 1] = nb[ell][1];
  }
  if (M % 2)
    return -1;
  return lat.coarse_lattice()->getM_lat();
}

/* ---------------------------------------------------------------- actions */

/* kind HO:        ip = {M, renorm}          dp = {T, m0, mu2}
 * kind QUARTIC:   ip = {M, renorm}          dp = {T, m0, mu2, lambda, x0}
 * kind ROTOR:     ip = {M, renorm}          dp = {T, m0}
 * kind SCHWINGER: ip = {Mt, Mx, ctype, renorm}   dp = {beta}
 * kind GFF:       ip = {Mt, Mx, ctype}      dp = {mass}            */
void *ref_action_create(int kind, const int *ip, const double *dp) {
  ActionHolder *h = new ActionHolder;
  h->kind = kind;
  try {
    if (kind == HO || kind == QUARTIC || kind == ROTOR) {
      h->lattice1d = std::make_shared<Lattice1D>(ip[0], dp[0]);
      RenormalisationType rn = (RenormalisationType)ip[1];
      if (kind == HO)
        h->action = std::make_shared<HarmonicOscillatorAction>(
            h->lattice1d, rn, dp[1], dp[2]);
      else if (kind == QUARTIC)
        h->action = std::make_shared<QuarticOscillatorAction>(
            h->lattice1d, rn, dp[1], dp[2], dp[3], dp[4]);
      else
        h->action = std::make_shared<RotorAction>(h->lattice1d, rn, dp[1]);
    } else if (kind == SCHWINGER) {
      h->lattice2d =
          std::make_shared<Lattice2D>(ip[0], ip[1], (CoarseningType)ip[2]);
      h->action = std::make_shared<QuenchedSchwingerAction>(
          h->lattice2d, nullptr, (RenormalisationType)ip[3], dp[0]);
    } else if (kind == GFF) {
      h->lattice2d =
          std::make_shared<Lattice2D>(ip[0], ip[1], (CoarseningType)ip[2]);
      h->action = std::make_shared<GFFAction>(h->lattice2d, nullptr, dp[0]);
    } else {
      delete h;
      return nullptr;
    }
  } catch (...) {
    delete h;
    return nullptr;
  }
  return h;
}

void ref_action_destroy(void *a) { delete (ActionHolder *)a; }

void *ref_action_coarse(void *a) {
  ActionHolder *h = (ActionHolder *)a;
  ActionHolder *c = new ActionHolder;
  c->kind = h->kind;
  try {
    c->action = h->action->coarse_action();
  } catch (...) {
    delete c;
    return nullptr;
  }
  if (h->kind == SCHWINGER || h->kind == GFF)
    c->lattice2d =
        std::dynamic_pointer_cast<QFTAction>(c->action)->get_lattice();
  else
    c->lattice1d =
        std::dynamic_pointer_cast<QMAction>(c->action)->get_lattice();
  return c;
}

int ref_action_sample_size(void *a) {
  return ((ActionHolder *)a)->action->sample_size();
}

/* which: 0 = m0 / beta / mu2(GFF); 1 = mu2 (HO) */
double ref_action_param(void *a, int which) {
  ActionHolder *h = (ActionHolder *)a;
  switch (h->kind) {
  case HO:
  case QUARTIC:
  case ROTOR:
    if (which == 0)
      return std::dynamic_pointer_cast<QMAction>(h->action)->getm0();
    break;
  case SCHWINGER:
    return std::dynamic_pointer_cast<QuenchedSchwingerAction>(h->action)
        ->getbeta();
  case GFF:
    return std::dynamic_pointer_cast<GFFAction>(h->action)->getmu2();
  }
  return NAN;
}

double ref_action_evaluate(void *a, const double *x) {
  ActionHolder *h = (ActionHolder *)a;
  return h->action->evaluate(make_state(x, h->action->sample_size()));
}

void ref_action_force(void *a, const double *x, double *p) {
  ActionHolder *h = (ActionHolder *)a;
  StatePtr ps = make_state(nullptr, h->action->sample_size());
  h->action->force(make_state(x, h->action->sample_size()), ps);
  read_state(ps, p);
}

/* n_sweeps lexicographic sweeps ell = 0..n-1 (the order
 * OverrelaxedHeatBathSampler::draw uses with random_order = false,
 * sampler/overrelaxedheatbathsampler.cc:8-31), or over idx[] if given */
void ref_action_overrelax_sweep(void *a, double *x, int n_sweeps,
                                const unsigned *idx, int n_idx) {
  ActionHolder *h = (ActionHolder *)a;
  const unsigned n = h->action->sample_size();
  StatePtr s = make_state(x, n);
  for (int sw = 0; sw < n_sweeps; ++sw) {
    if (idx)
      for (int k = 0; k < n_idx; ++k)
        h->action->overrelaxation_update(s, idx[k]);
    else
      for (unsigned ell = 0; ell < n; ++ell)
        h->action->overrelaxation_update(s, ell);
  }
  read_state(s, x);
}

/* stochastic: uses the action's own mt19937_64 engine (reference seeds) */
void ref_action_heatbath_sweep(void *a, double *x, int n_sweeps) {
  ActionHolder *h = (ActionHolder *)a;
  const unsigned n = h->action->sample_size();
  StatePtr s = make_state(x, n);
  for (int sw = 0; sw < n_sweeps; ++sw)
    for (unsigned ell = 0; ell < n; ++ell)
      h->action->heatbath_update(s, ell);
  read_state(s, x);
}

/* fine action `a` prolongs coarse state xc into x (entries not written by the
 * reference keep the input values of x) */
void ref_action_copy_from_coarse(void *a, const double *xc, int nc, double *x) {
  ActionHolder *h = (ActionHolder *)a;
  StatePtr s = make_state(x, h->action->sample_size());
  h->action->copy_from_coarse(make_state(xc, nc), s);
  read_state(s, x);
}

/* coarse action `a` (obtained from ref_action_coarse) restricts xf into x */
void ref_action_copy_from_fine(void *a, const double *xf, int nf, double *x) {
  ActionHolder *h = (ActionHolder *)a;
  StatePtr s = make_state(nullptr, h->action->sample_size());
  h->action->copy_from_fine(make_state(xf, nf), s);
  read_state(s, x);
}

void ref_action_W(void *a, double x_m, double x_p, double *out) {
  ActionHolder *h = (ActionHolder *)a;
  auto qm = std::dynamic_pointer_cast<QMAction>(h->action);
  out[0] = qm->getWminimum(x_m, x_p);
  out[1] = qm->getWcurvature(x_m, x_p);
}

void ref_action_initialise_state(void *a, double *x) {
  ActionHolder *h = (ActionHolder *)a;
  StatePtr s = make_state(nullptr, h->action->sample_size());
  h->action->initialise_state(s);
  read_state(s, x);
}

/* ------------------------------------------------ conditioned fine actions */

void *ref_cond_create(void *a) {
  ActionHolder *h = (ActionHolder *)a;
  CondHolder *c = new CondHolder;
  try {
    switch (h->kind) {
    case HO:
    case QUARTIC:
      c->cond = std::make_shared<GaussianConditionedFineAction>(
          std::dynamic_pointer_cast<QMAction>(h->action));
      break;
    case ROTOR:
      c->cond = RotorConditionedFineActionFactory().get(h->action);
      break;
    case SCHWINGER:
      c->cond = QuenchedSchwingerConditionedFineActionFactory().get(h->action);
      break;
    case GFF:
      c->cond = std::make_shared<GFFConditionedFineAction>(
          std::dynamic_pointer_cast<GFFAction>(h->action));
      break;
    }
  } catch (...) {
    delete c;
    return nullptr;
  }
  return c;
}
void ref_cond_destroy(void *c) { delete (CondHolder *)c; }

double ref_cond_evaluate(void *c, const double *x, int n) {
  return ((CondHolder *)c)->cond->evaluate(make_state(x, n));
}
/* stochastic: reference engine and seeds */
void ref_cond_fill(void *c, double *x, int n) {
  StatePtr s = make_state(x, n);
  ((CondHolder *)c)->cond->fill_fine_points(s);
  read_state(s, x);
}

/* -------------------------------------------------------------------- QoI */

/* qoi: 0 x^2, 1 rotor susceptibility, 2 2d susceptibility, 3 avg plaquette,
 * 4 phi^2 */
double ref_qoi_evaluate(int qoi, void *a, const double *x) {
  ActionHolder *h = (ActionHolder *)a;
  StatePtr s = make_state(x, h->action->sample_size());
  switch (qoi) {
  case 0:
    return QoIXsquared(h->lattice1d).evaluate(s);
  case 1:
    return QoISusceptibility(h->lattice1d).evaluate(s);
  case 2:
    return QoI2DSusceptibility(h->lattice2d).evaluate(s);
  case 3:
    return QoIAvgPlaquette(h->lattice2d).evaluate(s);
  case 4:
    return QoI2DPhiSquared(h->lattice2d).evaluate(s);
  }
  return NAN;
}

/* -------------------------------------------------------------------- HMC */

/* The leapfrog loop of HMCSampler::single_step (sampler/hmcsampler.cc:31-46)
 * driven through the reference's Action::force and the same two axpy
 * statements.  HMCSampler itself cannot be instantiated for timing at large
 * lattices: its constructor runs n_burnin draws plus 100 x 1000 autotune
 * trajectories (hmcsampler.hh:99-106, hmcsampler.cc:89-103). */
void ref_hmc_leapfrog(void *a, unsigned nt_hmc, double dt_hmc, double *x,
                      double *p) {
  ActionHolder *h = (ActionHolder *)a;
  const unsigned n = h->action->sample_size();
  StatePtr phi_state_trial = make_state(x, n);
  StatePtr p_state_cur = make_state(p, n);
  StatePtr dp_state = make_state(nullptr, n);
  for (unsigned int k = 0; k <= nt_hmc; ++k) {
    double dt_p = dt_hmc;
    double dt_x = dt_hmc;
    if (k == 0)
      dt_p = 0.5 * dt_hmc;
    if (k == nt_hmc) {
      dt_p = 0.5 * dt_hmc;
      dt_x = 0.0;
    }
    h->action->force(phi_state_trial, dp_state);
    p_state_cur->data -= dt_p * dp_state->data;
    phi_state_trial->data += dt_x * p_state_cur->data;
  }
  read_state(phi_state_trial, x);
  read_state(p_state_cur, p);
}

/* Full HMC draws (hmcsampler.cc:8-69 with n_rep = 1) with an mt19937_64 engine
 * seeded by `seed`; records QoI `qoi` after every draw; returns #accepted.
 * Also returns wall seconds in *seconds. */
int ref_hmc_draws(void *a, unsigned nt_hmc, double dt_hmc, int n_draws,
                  uint64_t seed, double *x, int qoi, double *q_out,
                  double *seconds) {
  ActionHolder *h = (ActionHolder *)a;
  const unsigned n = h->action->sample_size();
  std::mt19937_64 engine(seed);
  std::normal_distribution<double> normal_dist(0.0, 1.0);
  std::uniform_real_distribution<double> uniform_dist(0.0, 1.0);
  StatePtr phi_state_cur = make_state(x, n);
  StatePtr phi_state_trial = make_state(nullptr, n);
  StatePtr p_state_cur = make_state(nullptr, n);
  StatePtr dp_state = make_state(nullptr, n);
  int n_acc = 0;
  auto t0 = std::chrono::steady_clock::now();
  for (int d = 0; d < n_draws; ++d) {
    std::generate(p_state_cur->data.data(), p_state_cur->data.data() + n,
                  [&]() { return normal_dist(engine); });
    double T_kin_cur = 0.5 * p_state_cur->data.squaredNorm();
    phi_state_trial->data = phi_state_cur->data;
    for (unsigned int k = 0; k <= nt_hmc; ++k) {
      double dt_p = dt_hmc;
      double dt_x = dt_hmc;
      if (k == 0)
        dt_p = 0.5 * dt_hmc;
      if (k == nt_hmc) {
        dt_p = 0.5 * dt_hmc;
        dt_x = 0.0;
      }
      h->action->force(phi_state_trial, dp_state);
      p_state_cur->data -= dt_p * dp_state->data;
      phi_state_trial->data += dt_x * p_state_cur->data;
    }
    double T_kin_trial = 0.5 * p_state_cur->data.squaredNorm();
    double deltaS = h->action->evaluate(phi_state_trial) -
                    h->action->evaluate(phi_state_cur);
    double deltaH = deltaS + (T_kin_trial - T_kin_cur);
    bool accept_step = (deltaH < 0.0);
    if (!accept_step)
      accept_step = (uniform_dist(engine) < exp(-deltaH));
    if (accept_step)
      phi_state_cur->data = phi_state_trial->data;
    n_acc += (int)accept_step;
    if (q_out)
      q_out[d] = ref_qoi_evaluate(qoi, a, phi_state_cur->data.data());
  }
  auto t1 = std::chrono::steady_clock::now();
  if (seconds)
    *seconds = std::chrono::duration<double>(t1 - t0).count();
  read_state(phi_state_cur, x);
  return n_acc;
}

/* ------------------------------------------------------- two-level MH step */

/* One TwoLevelMetropolisStep::draw (montecarlo/twolevelmetropolisstep.cc:35-89)
 * with the stochastic fill-in replaced by a supplied trial state theta_prime
 * (so that the deterministic part, the three action differences, can be
 * compared exactly).  out = {dS_fine, dS_coarse, dS_trial}. */
void ref_twolevel_deltas(void *a_coarse, void *a_fine, void *cond,
                         const double *theta_fine, const double *theta_prime,
                         const double *phi_coarse, double *out) {
  ActionHolder *hc = (ActionHolder *)a_coarse;
  ActionHolder *hf = (ActionHolder *)a_fine;
  CondHolder *cd = (CondHolder *)cond;
  const unsigned nf = hf->action->sample_size();
  const unsigned nc = hc->action->sample_size();
  StatePtr tf = make_state(theta_fine, nf);
  StatePtr tp = make_state(theta_prime, nf);
  StatePtr pc = make_state(phi_coarse, nc);
  StatePtr tfc = make_state(nullptr, nc);
  out[0] = hf->action->evaluate(tp) - hf->action->evaluate(tf);
  hc->action->copy_from_fine(tf, tfc);
  out[1] = hc->action->evaluate(tfc) - hc->action->evaluate(pc);
  out[2] = cd->cond->evaluate(tf) - cd->cond->evaluate(tp);
}

/* n_draws of the reference's own TwoLevelMetropolisStep (its constructor runs
 * 10 000 timed draws: small lattices only).  Coarse states come from nothing:
 * phi_coarse is held fixed; used only for timing and acceptance statistics. */
double ref_twolevel_draws(void *a_coarse, void *a_fine, void *cond, int n_draws,
                          const double *phi_coarse, double *phi_fine,
                          double *seconds) {
  ActionHolder *hc = (ActionHolder *)a_coarse;
  ActionHolder *hf = (ActionHolder *)a_fine;
  CondHolder *cd = (CondHolder *)cond;
  TwoLevelMetropolisStep step(hc->action, hf->action, cd->cond);
  StatePtr pc = make_state(phi_coarse, hc->action->sample_size());
  StatePtr pf = make_state(phi_fine, hf->action->sample_size());
  step.set_state(pf);
  step.reset_stats();
  auto t0 = std::chrono::steady_clock::now();
  for (int d = 0; d < n_draws; ++d)
    step.draw(pc, pf);
  auto t1 = std::chrono::steady_clock::now();
  if (seconds)
    *seconds = std::chrono::duration<double>(t1 - t0).count();
  read_state(pf, phi_fine);
  return step.p_accept();
}

/* ---------------------------------------------------------- distributions */

/* dist: 0 ExpSin2 (param = sigma; x_p, x_m ignored)
 *       1 ExpCos (param = beta)    2 BesselProduct (param = beta <= 8)
 *       3 ApproximateBesselProduct (param = beta)                       */
double ref_dist_evaluate(int dist, double param, double x, double x_p,
                         double x_m) {
  switch (dist) {
  case 0:
    return ExpSin2Distribution().evaluate(x, param);
  case 1:
    return ExpCosDistribution(param).evaluate(x, x_p, x_m);
  case 2:
    return BesselProductDistribution(param).evaluate(x, x_p, x_m);
  case 3:
    return ApproximateBesselProductDistribution(param).evaluate(x, x_p, x_m);
  }
  return NAN;
}

void ref_dist_draw(int dist, double param, double x_p, double x_m,
                   uint64_t seed, int n, double *out) {
  std::mt19937_64 engine(seed);
  switch (dist) {
  case 0: {
    ExpSin2Distribution d;
    for (int k = 0; k < n; ++k)
      out[k] = d.draw(engine, param);
  } break;
  case 1: {
    ExpCosDistribution d(param);
    for (int k = 0; k < n; ++k)
      out[k] = d.draw(engine, x_p, x_m);
  } break;
  case 2: {
    BesselProductDistribution d(param);
    for (int k = 0; k < n; ++k)
      out[k] = d.draw(engine, x_p, x_m);
  } break;
  case 3: {
    ApproximateBesselProductDistribution d(param);
    for (int k = 0; k < n; ++k)
      out[k] = d.draw(engine, x_p, x_m);
  } break;
  }
}

double ref_besselproduct_Znorm_inv(double beta, double phi, int rescaled) {
  return BesselProductDistribution(beta).Znorm_inv(phi, rescaled != 0);
}

/* ---------------------------------------------------------------- scalars */

double ref_mod_2pi(double x) { return mod_2pi(x); }
double ref_mod_pi(double x) { return mod_pi(x); }
double ref_fast_bessel_I0_scaled(double z) { return fast_bessel_I0_scaled(z); }
double ref_Sigma_hat(double xi, unsigned p) { return Sigma_hat(xi, p); }
double ref_log_nCk(unsigned n, unsigned k) { return log_nCk(n, k); }
double ref_schwinger_chit_analytical(double beta, unsigned n_plaq) {
  return quenchedschwinger_chit_analytical(beta, n_plaq);
}
double ref_schwinger_chit_perturbative(double beta, unsigned n_plaq) {
  return quenchedschwinger_chit_perturbative(beta, n_plaq);
}
double ref_schwinger_var_chit_continuum(double beta, unsigned n_plaq) {
  return quenchedschwinger_var_chit_continuum_analytical(beta, n_plaq);
}
double ref_gff_phi_squared_analytical(double mass, double Mt, double Mx) {
  return gff_phi_squared_analytical(mass, Mt, Mx);
}
/* which: 0 exact, 1 perturbative, 2 continuum */
double ref_rotor_chit(void *a, int which) {
  auto r = std::dynamic_pointer_cast<RotorAction>(((ActionHolder *)a)->action);
  return which == 0 ? r->chit_exact()
                    : (which == 1 ? r->chit_perturbative() : r->chit_continuum());
}
double ref_ho_xsquared_analytical(void *a, int continuum) {
  auto r = std::dynamic_pointer_cast<HarmonicOscillatorAction>(
      ((ActionHolder *)a)->action);
  return continuum ? r->Xsquared_analytical_continuum()
                   : r->Xsquared_analytical();
}
/* HO exact sampler (Cholesky of the dense covariance,
 * qm/harmonicoscillatoraction.cc:38-66): n draws, reference engine */
void ref_ho_exact_draws(void *a, int n, double *out) {
  auto r = std::dynamic_pointer_cast<HarmonicOscillatorAction>(
      ((ActionHolder *)a)->action);
  const unsigned M = r->sample_size();
  StatePtr s = make_state(nullptr, M);
  for (int k = 0; k < n; ++k) {
    r->draw(s);
    std::memcpy(out + (size_t)k * M, s->data.data(), M * sizeof(double));
  }
}
/* GFF exact sampler (gffaction.cc:200-213) */
void ref_gff_exact_draws(void *a, int n, double *out) {
  auto r = std::dynamic_pointer_cast<GFFAction>(((ActionHolder *)a)->action);
  const unsigned M = r->sample_size();
  StatePtr s = make_state(nullptr, M);
  for (int k = 0; k < n; ++k) {
    r->draw(s);
    std::memcpy(out + (size_t)k * M, s->data.data(), M * sizeof(double));
  }
}

/* ------------------------------------------------------------- statistics */

/* out = {average, variance, variance_error, tau_int, error, samples} */
void ref_statistics(unsigned k_max, int n, const double *q, double *out) {
  Statistics stats("ref", k_max);
  for (int k = 0; k < n; ++k)
    stats.record_sample(q[k]);
  out[0] = stats.average();
  out[1] = stats.variance();
  out[2] = stats.variance_error();
  out[3] = stats.tau_int();
  out[4] = stats.error();
  out[5] = stats.samples();
}

} /* extern "C" */
