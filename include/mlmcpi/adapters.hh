// adapters.hh -- the reference's C++ class names on top of the C-ABI (include/mlmcpi.h).
//
// Header-only, link with -lmlmcpi.  Each class mirrors the public interface of the
// reference class of the same name (file:line relative to /root/reference/src) so that the
// callers -- MonteCarloSingleLevel::evaluate (montecarlo/montecarlosinglelevel.cc:23-94),
// the drivers, or user code -- keep their source: one state per call, std::shared_ptr
// ownership, `state->data[i]` element access.  A call uploads the state, launches the
// kernels for a batch of one chain and downloads the result; production code that wants
// throughput uses the batched C-ABI directly (thousands of chains per call).
//
// Errors follow the reference's convention (action/action.hh:48-52): message on stderr and
// std::runtime_error.
#ifndef MLMCPI_ADAPTERS_HH
#define MLMCPI_ADAPTERS_HH
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../mlmcpi.h"

namespace mlmcpi {

// ------------------------------------------------------------------ device context
/** process-wide context (device $MLMCPI_DEVICE, else $MLMCPI_RANK -- one process per GPU --, else 0);
 * there is no CPU fallback */
class Device {
public:
  static mlmcpi_ctx *ctx() {
    static Device d;
    return d.ctx_;
  }
  static void check(int rc, const char *what) {
    if (rc != 0) {
      std::string msg = std::string("ERROR: ") + what + ": " + mlmcpi_last_error(ctx());
      std::cerr << msg << std::endl;
      throw std::runtime_error(msg);
    }
  }

private:
  Device() {
    const char *dev = std::getenv("MLMCPI_DEVICE");
    if (!dev)
      dev = std::getenv("MLMCPI_RANK");
    if (mlmcpi_create(&ctx_, dev ? std::atoi(dev) : 0, 0x5EED0001ull, nullptr) != 0) {
      std::cerr << "ERROR: no CUDA device (mlmcpi has no CPU fallback)" << std::endl;
      throw std::runtime_error("mlmcpi_create failed");
    }
  }
  ~Device() { mlmcpi_destroy(ctx_); }
  mlmcpi_ctx *ctx_ = nullptr;
};

/** device buffer of doubles */
class DeviceVector {
public:
  explicit DeviceVector(size_t n) : n_(n) { Device::check(mlmcpi_alloc(Device::ctx(), n, &d_), "alloc"); }
  ~DeviceVector() { mlmcpi_free(Device::ctx(), d_); }
  DeviceVector(const DeviceVector &) = delete;
  DeviceVector &operator=(const DeviceVector &) = delete;
  double *ptr() const { return d_; }
  size_t size() const { return n_; }
  void upload(const double *h) { Device::check(mlmcpi_upload(Device::ctx(), d_, h, n_), "upload"); }
  void download(double *h) const { Device::check(mlmcpi_download(Device::ctx(), h, d_, n_), "download"); }

private:
  double *d_ = nullptr;
  size_t n_;
};

// ---------------------------------------------------- common/samplestate.hh:19-53
/** minimal stand-in for the Eigen::VectorXd member of the reference's SampleState */
class StateVector {
public:
  explicit StateVector(size_t n) : v_(n, 0.0) {}
  double &operator[](size_t i) { return v_[i]; }
  const double &operator[](size_t i) const { return v_[i]; }
  size_t size() const { return v_.size(); }
  double *data() { return v_.data(); }
  const double *data() const { return v_.data(); }
  double squaredNorm() const {
    double s = 0;
    for (double x : v_)
      s += x * x;
    return s;
  }

private:
  std::vector<double> v_;
};

class SampleState {
public:
  explicit SampleState(const unsigned int n_data) : data(n_data) {}
  void save_to_disk(const std::string filename) const { // common/samplestate.cc:7-16
    FILE *f = std::fopen(filename.c_str(), "w");
    if (!f)
      return;
    for (size_t i = 0; i < data.size(); ++i)
      std::fprintf(f, "%20.12e ", data[i]);
    std::fprintf(f, "\n");
    std::fclose(f);
  }
  StateVector data;
};

// --------------------------------------------------------------- lattices
enum CoarseningType { // lattice/lattice2d.hh:18-26
  CoarsenUnspecified = -1,
  CoarsenBoth = 0,
  CoarsenTemporal = 1,
  CoarsenSpatial = 2,
  CoarsenAlternate = 3,
  CoarsenRotate = 4
};
enum RenormalisationType { // action/renormalisation.hh:17-21
  RenormalisationNone = 0,
  RenormalisationPerturbative = 1,
  RenormalisationNonperturbative = 2
};

class Lattice1D { // lattice/lattice1d.hh:60-101
public:
  Lattice1D(const unsigned int M_lat_, const double T_final_, const int coarsening_level_ = 0)
      : M_lat(M_lat_), T_final(T_final_), a_lat(T_final_ / M_lat_), level(coarsening_level_) {}
  unsigned int getM_lat() const { return M_lat; }
  double getT_final() const { return T_final; }
  double geta_lat() const { return a_lat; }
  int get_coarsening_level() const { return level; }
  std::shared_ptr<Lattice1D> coarse_lattice() const {
    if (M_lat % 2)
      throw std::runtime_error("ERROR: cannot coarsen lattice with odd number of sites");
    return std::make_shared<Lattice1D>(M_lat / 2, T_final, level + 1);
  }

private:
  unsigned int M_lat;
  double T_final, a_lat;
  int level;
};

class Lattice2D { // lattice/lattice2d.hh:98-437
public:
  Lattice2D(const unsigned int Mt_lat_, const unsigned int Mx_lat_, const CoarseningType coarsening_type_,
            const int coarsening_level_ = 0)
      : Mt_lat(Mt_lat_), Mx_lat(Mx_lat_), ctype(coarsening_type_), level(coarsening_level_),
        rotated((coarsening_type_ == CoarsenRotate) && (coarsening_level_ % 2)) {}
  unsigned int getMt_lat() const { return Mt_lat; }
  unsigned int getMx_lat() const { return Mx_lat; }
  unsigned int getNedges() const { return rotated ? Mt_lat * Mx_lat : 2 * Mt_lat * Mx_lat; }
  unsigned int getNvertices() const { return rotated ? Mt_lat * Mx_lat / 2 : Mt_lat * Mx_lat; }
  unsigned int getNcells() const { return Mt_lat * Mx_lat; } // lattice/lattice2d.hh (unrotated levels)
  bool is_rotated() const { return rotated; }
  CoarseningType get_coarsening_type() const { return ctype; }
  int get_coarsening_level() const { return level; }
  unsigned int vertex_cart2lin(const int i, const int j) const {
    return mlmcpi_vertex_cart2lin(Mt_lat, Mx_lat, rotated, i, j);
  }
  void vertex_lin2cart(const unsigned int ell, int &i, int &j) const {
    mlmcpi_vertex_lin2cart(Mt_lat, Mx_lat, rotated, ell, &i, &j);
  }
  unsigned int link_cart2lin(const int i, const int j, const int mu) const {
    return mlmcpi_link_cart2lin(Mt_lat, Mx_lat, i, j, mu);
  }
  void link_lin2cart(const unsigned int ell, int &i, int &j, int &mu) const {
    mlmcpi_link_lin2cart(Mt_lat, Mx_lat, ell, &i, &j, &mu);
  }
  std::shared_ptr<Lattice2D> get_coarse_lattice() const {
    int mt, mx, rot;
    if (!mlmcpi_coarse_shape(Mt_lat, Mx_lat, ctype, level, &mt, &mx, &rot))
      return nullptr;
    return std::make_shared<Lattice2D>(mt, mx, ctype, level + 1);
  }

private:
  unsigned int Mt_lat, Mx_lat;
  CoarseningType ctype;
  int level;
  bool rotated;
};

// ------------------------------------------------------- action/action.hh:28-163
class Action {
public:
  Action(const mlmcpi_model &m, RenormalisationType renormalisation_, int level_, int ctype_)
      : model_(m), renormalisation(renormalisation_), level(level_), ctype(ctype_) {}
  virtual ~Action() {}
  virtual unsigned int sample_size() const { return mlmcpi_sample_size(&model_); }
  virtual double evaluation_cost() const { return sample_size(); }
  virtual int get_coarsening_level() const { return level; }
  int get_renormalisation() const { return renormalisation; }
  int get_coarsening_type() const { return ctype; }
  const mlmcpi_model &model() const { return model_; }

  /** Action::coarse_action(): the renormalised action on the next-coarser lattice */
  virtual std::shared_ptr<Action> coarse_action() {
    mlmcpi_model c;
    if (mlmcpi_coarse_model(&model_, renormalisation, level, ctype, model_.T_final, &c) != 0)
      fail("cannot coarsen the action");
    std::shared_ptr<Action> coarse = std::make_shared<Action>(c, renormalisation, level + 1, ctype);
    coarse->fine_model_ = model_; // the coarse action knows its fine lattice (qftaction.hh:79-120)
    coarse->has_fine_ = true;
    return coarse;
  }
  virtual const double evaluate(const std::shared_ptr<SampleState> state) const {
    DeviceVector x(check_size(state)), S(1);
    x.upload(state->data.data());
    Device::check(mlmcpi_action(Device::ctx(), &model_, x.ptr(), 1, S.ptr()), "Action::evaluate");
    double s;
    S.download(&s);
    return s;
  }
  virtual void force(const std::shared_ptr<SampleState> state, std::shared_ptr<SampleState> p_state) const {
    DeviceVector x(check_size(state)), f(check_size(p_state));
    x.upload(state->data.data());
    Device::check(mlmcpi_force(Device::ctx(), &model_, x.ptr(), f.ptr(), 1), "Action::force");
    f.download(p_state->data.data());
  }
  virtual void initialise_state(std::shared_ptr<SampleState> state) const {
    DeviceVector x(check_size(state));
    Device::check(mlmcpi_init_state(Device::ctx(), &model_, x.ptr(), 1, 0, init_counter++), "initialise_state");
    x.download(state->data.data());
  }
  virtual void copy_from_coarse(const std::shared_ptr<SampleState> coarse, std::shared_ptr<SampleState> state) {
    DeviceVector xc(coarse->data.size()), x(check_size(state));
    xc.upload(coarse->data.data());
    x.upload(state->data.data());
    Device::check(mlmcpi_prolong(Device::ctx(), &model_, xc.ptr(), x.ptr(), 1), "copy_from_coarse");
    x.download(state->data.data());
  }
  /** called on the COARSE action (one obtained from coarse_action()), as in the reference */
  virtual void copy_from_fine(const std::shared_ptr<SampleState> fine, std::shared_ptr<SampleState> state) {
    if (!has_fine_)
      fail("cannot copy from fine lattice.");
    DeviceVector xf(fine->data.size()), x(check_size(state));
    xf.upload(fine->data.data());
    Device::check(mlmcpi_restrict(Device::ctx(), &fine_model_, xf.ptr(), x.ptr(), 1), "copy_from_fine");
    x.download(state->data.data());
  }
  /** one coloured sweep of overrelaxation_update / heatbath_update over all dofs (the
   * reference's per-dof virtuals are replaced by sweeps: a single-dof kernel launch per
   * link would defeat the device) */
  virtual void overrelaxation_sweep(std::shared_ptr<SampleState> state) {
    DeviceVector x(check_size(state));
    x.upload(state->data.data());
    Device::check(mlmcpi_overrelax_sweep(Device::ctx(), &model_, x.ptr(), 1), "overrelaxation sweep");
    x.download(state->data.data());
  }
  virtual void heatbath_sweep(std::shared_ptr<SampleState> state) {
    DeviceVector x(check_size(state));
    x.upload(state->data.data());
    Device::check(mlmcpi_heatbath_sweep(Device::ctx(), &model_, x.ptr(), 1, 0, sweep_counter++), "heat bath sweep");
    x.download(state->data.data());
  }
  /** per-dof interface, action/action.hh:85-110: the updates OverrelaxedHeatBathSampler::draw loops over
   * (one single-dof kernel launch per call -- interface parity; the sweeps above are the fast path) */
  virtual void heatbath_update(std::shared_ptr<SampleState> state, const unsigned int ell) { dof_update(state, ell, 1); }
  virtual void overrelaxation_update(std::shared_ptr<SampleState> state, const unsigned int ell) {
    dof_update(state, ell, 0);
  }
  /** action/action.hh:104-110: the dofs a heat-bath sweep visits (all of them for the actions here) */
  virtual const std::vector<unsigned int> &get_heatbath_indexset() {
    if (heatbath_indexset.empty())
      for (unsigned int ell = 0; ell < sample_size(); ++ell)
        heatbath_indexset.push_back(ell);
    return heatbath_indexset;
  }
  virtual std::string info_string() const {
    std::stringstream s;
    if (model_.model <= MLMCPI_ROTOR)
      s << "lattice = " << model_.M_lat << ", m0 = " << model_.m0;
    else
      s << "lattice = " << model_.Mt_lat << " x " << model_.Mx_lat
        << (model_.model == MLMCPI_SCHWINGER ? ", beta = " : ", mu2 = ")
        << (model_.model == MLMCPI_SCHWINGER ? model_.beta : model_.gff_mu2);
    return s.str();
  }

protected:
  size_t check_size(const std::shared_ptr<SampleState> &s) const {
    if (s->data.size() != sample_size())
      fail("state of wrong size");
    return s->data.size();
  }
  [[noreturn]] static void fail(const char *msg) {
    std::cerr << "ERROR: " << msg << std::endl;
    throw std::runtime_error(msg);
  }
  mlmcpi_model model_;
  void dof_update(std::shared_ptr<SampleState> state, const unsigned int ell, int heatbath) {
    DeviceVector x(check_size(state));
    x.upload(state->data.data());
    Device::check(mlmcpi_dof_update(Device::ctx(), &model_, x.ptr(), 1, (int)ell, heatbath, 0, sweep_counter++),
                  heatbath ? "heatbath_update" : "overrelaxation_update");
    x.download(state->data.data());
  }
  std::vector<unsigned int> heatbath_indexset;
  mlmcpi_model fine_model_ = {};
  bool has_fine_ = false;
  RenormalisationType renormalisation;
  int level, ctype;
  mutable uint64_t init_counter = 0, sweep_counter = 0;
};

static inline mlmcpi_model qm_model(int kind, const Lattice1D &lat, double m0, double mu2 = 0, double lambda = 0,
                                    double x0 = 0) {
  mlmcpi_model m = {};
  m.model = kind;
  m.M_lat = lat.getM_lat();
  m.a_lat = lat.geta_lat();
  m.T_final = lat.getT_final();
  m.m0 = m0;
  m.mu2 = mu2;
  m.lambda = lambda;
  m.x0 = x0;
  return m;
}

/** action/qm/qmaction.hh: base of the 1-D actions (the reference's drivers hold a shared_ptr<QMAction>) */
class QMAction : public Action {
public:
  using Action::Action;
  double getm0() const { return model_.m0; }
};

/** action/qm/harmonicoscillatoraction.hh */
class HarmonicOscillatorAction : public QMAction {
public:
  HarmonicOscillatorAction(const std::shared_ptr<Lattice1D> lattice, const RenormalisationType renormalisation_,
                           const double m0_, const double mu2_)
      : QMAction(qm_model(MLMCPI_HO, *lattice, m0_, mu2_), renormalisation_, lattice->get_coarsening_level(), 0) {}
  /** qm/harmonicoscillatoraction.cc:76-80 */
  double Xsquared_analytical_continuum() const {
    return mlmcpi_ho_xsquared_analytical(model_.m0, model_.mu2, model_.a_lat, model_.M_lat, 1);
  }
  /** qm/harmonicoscillatoraction.cc:69-74 */
  double Xsquared_analytical() const {
    const double a = model_.a_lat, mu2 = model_.mu2;
    const double R = 1. + 0.5 * a * a * mu2 - a * std::sqrt(mu2) * std::sqrt(1. + 0.25 * a * a * mu2);
    const double RM = std::pow(R, (double)model_.M_lat);
    return 1. / (2. * model_.m0 * std::sqrt(mu2) * std::sqrt(1. + 0.25 * a * a * mu2)) * (1. + RM) / (1. - RM);
  }
};
/** action/qm/quarticoscillatoraction.hh */
class QuarticOscillatorAction : public QMAction {
public:
  QuarticOscillatorAction(const std::shared_ptr<Lattice1D> lattice, const RenormalisationType renormalisation_,
                          const double m0_, const double mu2_, const double lambda_, const double x0_)
      : QMAction(qm_model(MLMCPI_QUARTIC, *lattice, m0_, mu2_, lambda_, x0_), renormalisation_,
               lattice->get_coarsening_level(), 0) {}
};
/** action/clusteraction.hh:25-72: what the generic ClusterSampler needs from an action.  The batched
 * cluster kernels (csrc/qm.cu) have S_ell / new_reflection / flip of the rotor built in; this host
 * interface exists so that code written against the reference's ClusterAction keeps compiling. */
class ClusterAction {
public:
  virtual ~ClusterAction() {}
  virtual double S_ell(const std::shared_ptr<SampleState> x_path, const unsigned int i, const unsigned int j) const = 0;
  virtual void new_reflection() const = 0;
  virtual void flip(std::shared_ptr<SampleState> x_path, const unsigned int ell) const = 0;
  virtual void initialise_state(std::shared_ptr<SampleState> x_path) const = 0;
  virtual unsigned int sample_size() const = 0;
};

/** action/qm/rotoraction.hh */
class RotorAction : public QMAction, public ClusterAction {
public:
  /** rotoraction.hh:226-253 */
  virtual double S_ell(const std::shared_ptr<SampleState> x_path, const unsigned int i, const unsigned int j) const {
    return -2. * model_.m0 / model_.a_lat * std::cos(x_path->data[i] - xbar) * std::cos(x_path->data[j] - xbar);
  }
  virtual void new_reflection() const {
    // splitmix64 stream: the host-side reflection angle of the interface; the batched kernels draw
    // theirs from Philox (stream CLUSTER)
    reflection_state += 0x9E3779B97F4A7C15ull;
    uint64_t z = reflection_state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    xbar = -M_PI + 2. * M_PI * ((double)(z >> 11) * (1.0 / 9007199254740992.0));
  }
  virtual void flip(std::shared_ptr<SampleState> x_path, const unsigned int ell) const {
    const double x = M_PI + 2. * xbar - x_path->data[ell];
    x_path->data[ell] = x - 2. * M_PI * std::floor(0.5 * (x + M_PI) / M_PI); // mod_2pi, auxilliary.hh:42-44
  }
  virtual void initialise_state(std::shared_ptr<SampleState> x_path) const { Action::initialise_state(x_path); }
  virtual unsigned int sample_size() const { return Action::sample_size(); }
  mutable double xbar = 0.0;
  mutable uint64_t reflection_state = 21172817ull; // the reference's rotor seed, rotoraction.hh:106

  RotorAction(const std::shared_ptr<Lattice1D> lattice, const RenormalisationType renormalisation_, const double m0_)
      : QMAction(qm_model(MLMCPI_ROTOR, *lattice, m0_), renormalisation_, lattice->get_coarsening_level(), 0) {}
  /** qm/rotoraction.cc:92-115 */
  double chit_exact() const { return mlmcpi_rotor_chit(model_.m0, model_.a_lat, model_.T_final, 0); }
  double chit_perturbative() const { return mlmcpi_rotor_chit(model_.m0, model_.a_lat, model_.T_final, 1); }
  double chit_continuum() const { return mlmcpi_rotor_chit(model_.m0, model_.a_lat, model_.T_final, 2); }
};

static inline int level_coarsening(int ctype, int level) {
  if (ctype == CoarsenAlternate)
    return (level % 2 == 0) ? CoarsenTemporal : CoarsenSpatial;
  return ctype;
}

/** action/qft/quenchedschwingeraction.hh */
class QuenchedSchwingerAction : public Action {
public:
  QuenchedSchwingerAction(const std::shared_ptr<Lattice2D> lattice, const std::shared_ptr<Lattice2D> /*fine_lattice*/,
                          const RenormalisationType renormalisation_, const double beta_)
      : Action(make(*lattice, beta_), renormalisation_, lattice->get_coarsening_level(),
               lattice->get_coarsening_type()) {
    const int c = lattice->get_coarsening_type();
    if (c == CoarsenRotate || c == CoarsenUnspecified) // quenchedschwingeraction.hh:118-131
      fail("invalid coarsening for quenched Schwinger model. Has to be 'both', 'temporal', 'spatial' or "
           "'alternate'.");
  }
  double getbeta() const { return model_.beta; }

private:
  static mlmcpi_model make(const Lattice2D &lat, double beta) {
    mlmcpi_model m = {};
    m.model = MLMCPI_SCHWINGER;
    m.Mt_lat = lat.getMt_lat();
    m.Mx_lat = lat.getMx_lat();
    m.beta = beta;
    m.coarsening = level_coarsening(lat.get_coarsening_type(), lat.get_coarsening_level());
    return m;
  }
};

/** action/qft/gffaction.hh (fine-level 5-point action) */
class GFFAction : public Action {
public:
  GFFAction(const std::shared_ptr<Lattice2D> lattice, const std::shared_ptr<Lattice2D> /*fine_lattice*/,
            const double mass_)
      : Action(make(*lattice, mass_), RenormalisationNone, lattice->get_coarsening_level(),
               lattice->get_coarsening_type()) {
    if (lattice->getMt_lat() != lattice->getMx_lat()) // gffaction.hh:169-173
      fail("Lattice has to be squared for GFF action");
  }
  double getmu2() const { return model_.gff_mu2; }

private:
  static mlmcpi_model make(const Lattice2D &lat, double mass) {
    mlmcpi_model m = {};
    m.model = MLMCPI_GFF;
    m.Mt_lat = lat.getMt_lat();
    m.Mx_lat = lat.getMx_lat();
    m.rotated = lat.is_rotated();
    m.coarsening = lat.get_coarsening_type();
    const double a = (lat.is_rotated() ? std::sqrt(2.) : 1.) / lat.getMt_lat();
    m.gff_mu2 = a * a * mass * mass;
    return m;
  }
};

// --------------------------------------- action/conditionedfineaction.hh:38-67
class ConditionedFineAction {
public:
  explicit ConditionedFineAction(const std::shared_ptr<Action> action_) : action(action_) {}
  virtual ~ConditionedFineAction() {}
  virtual void fill_fine_points(std::shared_ptr<SampleState> state) const {
    DeviceVector x(state->data.size());
    x.upload(state->data.data());
    Device::check(mlmcpi_fill(Device::ctx(), &action->model(), x.ptr(), 1, 0, counter++), "fill_fine_points");
    x.download(state->data.data());
  }
  virtual double evaluate(const std::shared_ptr<SampleState> state) const {
    DeviceVector x(state->data.size()), S(1);
    x.upload(state->data.data());
    Device::check(mlmcpi_cond_action(Device::ctx(), &action->model(), x.ptr(), 1, S.ptr()),
                  "ConditionedFineAction::evaluate");
    double s;
    S.download(&s);
    return s;
  }

protected:
  const std::shared_ptr<Action> action;
  mutable uint64_t counter = 0;
};
class ConditionedFineActionFactory {
public:
  virtual ~ConditionedFineActionFactory() {}
  virtual std::shared_ptr<ConditionedFineAction> get(std::shared_ptr<Action> action) {
    return std::make_shared<ConditionedFineAction>(action);
  }
};
typedef ConditionedFineActionFactory QuenchedSchwingerConditionedFineActionFactory;
typedef ConditionedFineActionFactory RotorConditionedFineActionFactory;
typedef ConditionedFineActionFactory GaussianConditionedFineActionFactory;
typedef ConditionedFineActionFactory GFFConditionedFineActionFactory;

// ------------------------------------------- qoi/quantityofinterest.hh:16-36
class QoI {
public:
  QoI(const std::shared_ptr<Action> action_, int which_) : model_(action_->model()), which(which_) {}
  /** the reference constructs its QoIs from the lattice (qoi/qft/qoi2dsusceptibility.hh:35-40, ...) */
  QoI(const mlmcpi_model &m, int which_) : model_(m), which(which_) {}
  virtual ~QoI() {}
  int id() const { return which; } // MLMCPI_QOI_*
  virtual const double evaluate(const std::shared_ptr<SampleState> state) {
    DeviceVector x(state->data.size()), q(1);
    x.upload(state->data.data());
    Device::check(mlmcpi_qoi(Device::ctx(), &model_, which, x.ptr(), 1, q.ptr(), nullptr), "QoI::evaluate");
    double v;
    q.download(&v);
    return v;
  }

protected:
  static mlmcpi_model lattice_model(int kind, const Lattice1D &lat) { return qm_model(kind, lat, 1.0); }
  static mlmcpi_model lattice_model(int kind, const Lattice2D &lat) {
    mlmcpi_model m = {};
    m.model = kind;
    m.Mt_lat = lat.getMt_lat();
    m.Mx_lat = lat.getMx_lat();
    m.rotated = lat.is_rotated();
    m.coarsening = level_coarsening(lat.get_coarsening_type(), lat.get_coarsening_level());
    m.beta = 1.0;
    return m;
  }
  const mlmcpi_model model_;
  const int which;
};
struct QoIXsquared : QoI {
  explicit QoIXsquared(std::shared_ptr<Action> a) : QoI(a, MLMCPI_QOI_X2) {}
  explicit QoIXsquared(std::shared_ptr<Lattice1D> l) : QoI(lattice_model(MLMCPI_HO, *l), MLMCPI_QOI_X2) {}
};
struct QoISusceptibility : QoI {
  explicit QoISusceptibility(std::shared_ptr<Action> a) : QoI(a, MLMCPI_QOI_ROTOR_CHI) {}
  explicit QoISusceptibility(std::shared_ptr<Lattice1D> l) : QoI(lattice_model(MLMCPI_ROTOR, *l), MLMCPI_QOI_ROTOR_CHI) {}
};
struct QoI2DSusceptibility : QoI {
  explicit QoI2DSusceptibility(std::shared_ptr<Action> a) : QoI(a, MLMCPI_QOI_SCHWINGER_CHI) {}
  explicit QoI2DSusceptibility(std::shared_ptr<Lattice2D> l)
      : QoI(lattice_model(MLMCPI_SCHWINGER, *l), MLMCPI_QOI_SCHWINGER_CHI) {}
};
struct QoIAvgPlaquette : QoI {
  explicit QoIAvgPlaquette(std::shared_ptr<Action> a) : QoI(a, MLMCPI_QOI_AVG_PLAQUETTE) {}
  explicit QoIAvgPlaquette(std::shared_ptr<Lattice2D> l)
      : QoI(lattice_model(MLMCPI_SCHWINGER, *l), MLMCPI_QOI_AVG_PLAQUETTE) {}
};
struct QoI2DPhiSquared : QoI {
  explicit QoI2DPhiSquared(std::shared_ptr<Action> a) : QoI(a, MLMCPI_QOI_PHI2) {}
  explicit QoI2DPhiSquared(std::shared_ptr<Lattice2D> l) : QoI(lattice_model(MLMCPI_GFF, *l), MLMCPI_QOI_PHI2) {}
};

// ---------------------------------------------- montecarlo/mcmcstep.hh:21-72
class MCMCStep {
public:
  MCMCStep() : accept(false) { reset_stats(); }
  virtual ~MCMCStep() {}
  void reset_stats() { n_total_samples = n_accepted_samples = 0; }
  double p_accept() { return n_accepted_samples / (1. * n_total_samples); }
  bool accepted() const { return accept; }
  virtual void show_stats() { std::cout << "  acceptance rate = " << p_accept() << std::endl; }
  virtual void set_state(std::shared_ptr<SampleState> x_state) = 0;

protected:
  unsigned int n_accepted_samples, n_total_samples;
  bool accept;
};

/** sampler/sampler.hh:20-43 and its concrete samplers, backed by one mlmcpi_sampler
 * object with a batch of one chain (the state stays on the device between draws) */
class Sampler : public MCMCStep {
public:
  struct HMCParameters { // sampler/hmcsampler.hh:21-65
    unsigned int nt = 100;
    double dt = 0.1;
    unsigned int n_burnin = 100;
    unsigned int n_rep = 1;
  };
  struct HeatBathParameters { // sampler/overrelaxedheatbathsampler.hh
    unsigned int n_sweep_overrelax = 10, n_sweep_heatbath = 1, n_burnin = 100;
  };
  virtual ~Sampler() { mlmcpi_sampler_destroy(s_); }
  virtual void draw(std::shared_ptr<SampleState> state) {
    Device::check(mlmcpi_sampler_draw(s_, out_->ptr(), acc_), "Sampler::draw");
    double raw;
    Device::check(mlmcpi_download(Device::ctx(), &raw, accbuf_->ptr(), 1), "download");
    int32_t acc;
    std::memcpy(&acc, &raw, sizeof(acc)); // the flag is the first 4 bytes of the 8-byte buffer
    accept = acc != 0;
    n_total_samples++;
    n_accepted_samples += (int)accept;
    if (accept) // copy_if_rejected == false in the reference (mcmcstep.hh:27)
      out_->download(state->data.data());
  }
  virtual void set_state(std::shared_ptr<SampleState> state) {
    out_->upload(state->data.data());
    Device::check(mlmcpi_sampler_set_state(s_, out_->ptr()), "Sampler::set_state");
  }

protected:
  Sampler(const std::shared_ptr<Action> action, const mlmcpi_sampler_params &prm) {
    Device::check(mlmcpi_sampler_create(Device::ctx(), &action->model(), &prm, 1, 0, &s_), "sampler create");
    out_.reset(new DeviceVector(action->sample_size()));
    accbuf_.reset(new DeviceVector(1));
    acc_ = reinterpret_cast<int32_t *>(accbuf_->ptr());
  }
  void burn_in(unsigned int n, unsigned int n_dof) {
    std::shared_ptr<SampleState> tmp = std::make_shared<SampleState>(n_dof);
    for (unsigned int k = 0; k < n; ++k)
      draw(tmp);
    reset_stats();
  }
  mlmcpi_sampler *s_ = nullptr;
  std::unique_ptr<DeviceVector> out_, accbuf_;
  int32_t *acc_ = nullptr;
};

static inline mlmcpi_sampler_params make_params(int kind, int n_levels, int renorm, int ctype, unsigned nt, double dt,
                                                unsigned n_rep, unsigned n_or, unsigned n_hb) {
  mlmcpi_sampler_params p = {};
  p.kind = kind;
  p.n_levels = n_levels;
  p.renorm = renorm;
  p.ctype = ctype;
  p.nt = (int)nt;
  p.dt = dt;
  p.n_rep = (int)n_rep;
  p.n_sweep_overrelax = (int)n_or;
  p.n_sweep_heatbath = (int)n_hb;
  return p;
}

/** sampler/hmcsampler.hh:84-109: burn-in and step-size autotuning in the constructor */
class HMCSampler : public Sampler {
public:
  HMCSampler(const std::shared_ptr<Action> action, const HMCParameters p, bool autotune = true)
      : Sampler(action, make_params(MLMCPI_SAMPLER_HMC, 1, 0, 0, p.nt, p.dt, p.n_rep, 0, 0)) {
    burn_in(p.n_burnin, action->sample_size());
    if (autotune) {
      double dt, pa;
      const int rc = mlmcpi_sampler_autotune(s_, 0.8, 100, 1000, &dt, &pa);
      std::cout << (rc == 0 ? "  Tuned         dt_{HMC} = " : "  FAILED to tune, reverting to ") << dt << std::endl;
    }
  }
};

/** sampler/overrelaxedheatbathsampler.hh:102-129 */
class OverrelaxedHeatBathSampler : public Sampler {
public:
  OverrelaxedHeatBathSampler(const std::shared_ptr<Action> action, const HeatBathParameters p)
      : Sampler(action, make_params(MLMCPI_SAMPLER_HEATBATH, 1, 0, 0, 0, 0, 1, p.n_sweep_overrelax, p.n_sweep_heatbath)) {
    burn_in(p.n_burnin, action->sample_size());
  }
};

/** sampler/clustersampler.hh:66-160 (1-D rotor) and sampler/quenchedschwingerclustersampler.hh:24-98:
 * n_updates Wolff single-cluster updates per draw */
class ClusterSampler : public Sampler {
public:
  ClusterSampler(const std::shared_ptr<Action> action, const unsigned int n_updates = 10,
                 const unsigned int n_burnin = 100)
      : Sampler(action, cluster_params(n_updates)) {
    burn_in(n_burnin, action->sample_size());
  }

private:
  static mlmcpi_sampler_params cluster_params(unsigned int n_updates) {
    mlmcpi_sampler_params q = make_params(MLMCPI_SAMPLER_CLUSTER, 1, 0, 0, 0, 0, 1, 0, 0);
    q.n_updates = (int)n_updates;
    return q;
  }
};
typedef ClusterSampler QuenchedSchwingerClusterSampler;

/** sampler/hierarchicalsampler.hh: HMC or heat bath on the coarsest of n_max_level levels */
class HierarchicalSampler : public Sampler {
public:
  HierarchicalSampler(const std::shared_ptr<Action> fine_action, unsigned int n_max_level, int coarse_kind,
                      RenormalisationType renorm, CoarseningType ctype, const HMCParameters hmc = HMCParameters(),
                      const HeatBathParameters hb = HeatBathParameters())
      : Sampler(fine_action,
                make_params(coarse_kind, (int)n_max_level - fine_action->get_coarsening_level(), renorm, ctype, hmc.nt,
                            hmc.dt, hmc.n_rep, hb.n_sweep_overrelax, hb.n_sweep_heatbath)),
        n_level(n_max_level - fine_action->get_coarsening_level()) {}
  virtual void show_stats() {
    std::vector<double> p(n_level);
    mlmcpi_sampler_stats(s_, p.data());
    std::cout << "  acceptance rate = " << p_accept() << std::endl;
    for (unsigned int l = 0; l < n_level; ++l)
      std::cout << "  level " << l << (l == 0 ? " [finest]  " : (l == n_level - 1 ? " [coarsest]" : "           "))
                << " :  p = " << p[l] << std::endl;
  }

private:
  unsigned int n_level;
};

// ---------------------------------- montecarlo/twolevelmetropolisstep.hh/.cc
class TwoLevelMetropolisStep : public MCMCStep {
public:
  TwoLevelMetropolisStep(const std::shared_ptr<Action> coarse_action_, const std::shared_ptr<Action> fine_action_,
                         const std::shared_ptr<ConditionedFineAction> /*conditioned_fine_action*/)
      : coarse_action(coarse_action_), fine_action(fine_action_), theta(fine_action_->sample_size()),
        phi_c(coarse_action_->sample_size()), cache(3) {}
  virtual void set_state(std::shared_ptr<SampleState> state) { // twolevelmetropolisstep.cc:92-97
    theta.upload(state->data.data());
    Device::check(mlmcpi_action(Device::ctx(), &fine_action->model(), theta.ptr(), 1, cache.ptr()), "set_state");
    Device::check(mlmcpi_cond_action(Device::ctx(), &fine_action->model(), theta.ptr(), 1, cache.ptr() + 1), "set_state");
  }
  void draw(const std::shared_ptr<SampleState> coarse_state, std::shared_ptr<SampleState> state) { // :35-89
    phi_c.upload(coarse_state->data.data());
    int32_t *acc = reinterpret_cast<int32_t *>(cache.ptr() + 2);
    Device::check(mlmcpi_twolevel_step(Device::ctx(), &fine_action->model(), &coarse_action->model(), phi_c.ptr(),
                                       theta.ptr(), cache.ptr(), cache.ptr() + 1, 1, 0, counter++, acc, nullptr),
                  "TwoLevelMetropolisStep::draw");
    double raw;
    Device::check(mlmcpi_download(Device::ctx(), &raw, cache.ptr() + 2, 1), "download");
    int32_t a;
    std::memcpy(&a, &raw, sizeof(a));
    accept = a != 0;
    n_total_samples++;
    n_accepted_samples += (int)accept;
    if (accept)
      theta.download(state->data.data());
  }

private:
  const std::shared_ptr<Action> coarse_action, fine_action;
  DeviceVector theta, phi_c, cache; // cache = {S_f(theta), S_cond(theta), accept flag}
  uint64_t counter = 0;
};

// -------------------------------------------- common/statistics.hh/.cc (host)
class Statistics {
public:
  Statistics(const std::string label_, const unsigned int k_max_) : obj_label(label_), k_max(k_max_) { hard_reset(); }
  std::string label() const { return obj_label; }
  void reset() {
    n_samples = 0;
    avg = 0.0;
  }
  void hard_reset() {
    reset();
    Q_k.clear();
    S_k.assign(k_max, 0.0);
    avg_longterm = avg2_longterm = avg3_longterm = avg4_longterm = 0.0;
    n_samples_longterm = 0;
  }
  void record_sample(const double Q) { // statistics.cc:4-27
    n_samples++;
    n_samples_longterm++;
    Q_k.push_front(Q);
    if (Q_k.size() > k_max)
      Q_k.pop_back();
    avg = ((n_samples - 1.0) * avg + Q) / (1.0 * n_samples);
    const double w = n_samples_longterm - 1.0, n = 1.0 * n_samples_longterm;
    avg_longterm = (w * avg_longterm + Q) / n;
    avg2_longterm = (w * avg2_longterm + Q * Q) / n;
    avg3_longterm = (w * avg3_longterm + Q * Q * Q) / n;
    avg4_longterm = (w * avg4_longterm + Q * Q * Q * Q) / n;
    for (unsigned int k = 0; k < Q_k.size(); ++k) {
      const unsigned int N_k = n_samples_longterm - k;
      S_k[k] = ((N_k - 1.0) * S_k[k] + Q_k[0] * Q_k[k]) / (1.0 * N_k);
    }
  }
  double average() const { return avg; }
  double variance() const {
    return 1.0 * n_samples_longterm / (n_samples_longterm - 1.0) * (S_k[0] - avg_longterm * avg_longterm);
  }
  double tau_int() const {
    double t = 0.0;
    const double C0 = S_k[0] - avg_longterm * avg_longterm;
    for (unsigned int k = 1; k < S_k.size(); ++k)
      t += (1. - k / (1.0 * n_samples_longterm)) * (S_k[k] - avg_longterm * avg_longterm);
    return std::fmax(1.0, 1.0 + 2.0 * t / C0);
  }
  double error() const { return std::sqrt(tau_int() * variance() / (1.0 * samples())); }
  unsigned int samples() const { return n_samples; }
  unsigned int autocorr_window() const { return k_max; }

private:
  const std::string obj_label;
  const unsigned int k_max;
  unsigned int n_samples_longterm, n_samples;
  std::deque<double> Q_k;
  std::vector<double> S_k;
  double avg, avg_longterm, avg2_longterm, avg3_longterm, avg4_longterm;
};

} // namespace mlmcpi
#endif // MLMCPI_ADAPTERS_HH
