// montecarlo.hh -- the reference's Monte Carlo drivers (MonteCarloSingleLevel, MonteCarloTwoLevel,
// MonteCarloMultiLevel), its sampler / QoI factories and its Statistics print-out, run on a
// BATCH of independent chains through the C-ABI (include/mlmcpi.h).
//
// The reference runs one chain per MPI rank and averages the ranks' moments
// (montecarlo/montecarlosinglelevel.cc:54-86, common/statistics.cc:30-35).  Here one process
// drives B chains on one GPU; every `n_samples` / `n_target` of the reference is a total over
// the chains, and a "local" count is ceil(n / B) -- exactly what distribute_n() does over ranks.
// Constructor signatures follow the reference so that driver_qm / driver_qft keep their source.
#ifndef MLMCPI_MONTECARLO_HH
#define MLMCPI_MONTECARLO_HH
#include <chrono>
#include <iomanip>

#include "../mlmcpi_comm.h"
#include "adapters.hh"
#include "parameters.hh"

namespace mlmcpi {

/** the processes of one run (one per GPU): rank, world size and the NCCL communicator from the
 * environment (MLMCPI_RANK, MLMCPI_WORLD_SIZE, MLMCPI_COMM_FILE; examples/run_multi_gpu.sh sets
 * them).  The role of mpi/mpi_wrapper.hh: chains are sharded over the processes, Statistics
 * all-reduce their packed moments, everything else is process-local. */
class Parallel {
public:
  static mlmcpi_comm *comm() {
    static Parallel p;
    return p.comm_;
  }
  /** rank and world size come straight from the environment, so that they are known (e.g. for
   * master-only output) before a device is touched */
  static int rank() { return std::max(0, env_int("MLMCPI_RANK", 0)); }
  static int world_size() { return std::max(1, env_int("MLMCPI_WORLD_SIZE", 1)); }
  static bool master() { return rank() == 0; }

private:
  static int env_int(const char *name, int fallback) {
    const char *v = std::getenv(name);
    return v ? std::atoi(v) : fallback;
  }
  Parallel() {
    if (mlmcpi_comm_create_from_env(Device::ctx(), &comm_) != 0) {
      std::cerr << "ERROR: cannot set up the NCCL communicator (MLMCPI_RANK / MLMCPI_WORLD_SIZE / MLMCPI_COMM_FILE)"
                << std::endl;
      throw std::runtime_error("mlmcpi_comm_create_from_env failed");
    }
    // the library's own Statistics queries (MultilevelSampler, MonteCarloMultiLevel, HMC autotune)
    // run over the chains of all processes from here on
    Device::check(mlmcpi_comm_attach(comm_), "mlmcpi_comm_attach");
  }
  ~Parallel() { mlmcpi_comm_destroy(comm_); }
  mlmcpi_comm *comm_ = nullptr;
};

/** std::cout on the master process, a sink elsewhere (mpi_parallel::cout of the reference) */
inline std::ostream &pcout() {
  static std::ostream sink(nullptr);
  return Parallel::master() ? std::cout : sink;
}

/** number of chains per device of this process: set once by the driver, or -- for the reference's own
 * drivers, whose command line has no room for it -- from the environment variable MLMCPI_CHAINS */
inline unsigned int &batch_size() {
  static unsigned int B = [] {
    const char *v = std::getenv("MLMCPI_CHAINS");
    const int n = v ? std::atoi(v) : 0;
    return (unsigned int)(n > 0 ? n : 256);
  }();
  return B;
}

// ----------------------------------------------------------- device-side Statistics
/** Statistics (common/statistics.hh) with one accumulator per chain on the device; getters
 * return the chain-averaged estimators the reference obtains by MPI reduction */
class BatchedStatistics {
public:
  BatchedStatistics(const std::string &label, unsigned int k_max, unsigned int B) : label_(label), k_max_(k_max) {
    Device::check(mlmcpi_stats_create(Device::ctx(), (int)k_max, (int)B, &st_), "stats create");
  }
  ~BatchedStatistics() { mlmcpi_stats_destroy(st_); }
  BatchedStatistics(const BatchedStatistics &) = delete;
  void reset() { Device::check(mlmcpi_stats_reset(st_), "stats reset"); }
  void hard_reset() { Device::check(mlmcpi_stats_hard_reset(st_), "stats hard_reset"); }
  void record_sample(const double *d_q) { Device::check(mlmcpi_stats_record(st_, d_q), "stats record"); }
  double average() const { return query()[0]; }
  double variance() const { return query()[1]; }
  double variance_error() const { return query()[2]; }
  double tau_int() const { return query()[3]; }
  double error() const { return query()[4]; }
  unsigned long samples() const { return (unsigned long)query()[5]; }
  unsigned int autocorr_window() const { return k_max_; }
  const std::string &label() const { return label_; }
  /** common/statistics.cc:101-116 */
  friend std::ostream &operator<<(std::ostream &os, const BatchedStatistics &s) {
    const std::vector<double> q = s.query();
    os << " " << std::setprecision(6) << std::fixed;
    os << s.label_ << ": Avg +/- Err = " << q[0] << " +/- " << q[4] << std::endl;
    os << " " << s.label_ << ": Var +/- Err = " << q[1] << " +/- " << q[2] << std::endl;
    os << std::setprecision(3) << std::fixed;
    os << " " << s.label_ << ": tau_{int}   = " << q[3] << std::endl;
    os << " " << s.label_ << ": window      = " << s.k_max_ << std::endl;
    os << " " << s.label_ << ": # samples   = " << (unsigned long)q[5] << std::endl;
    return os;
  }

private:
  std::vector<double> query() const {
    std::vector<double> out(6); // moments summed over the chains of ALL processes
    Device::check(mlmcpi_comm_stats(Parallel::comm(), st_, (int)k_max_, out.data()), "stats all-reduce");
    return out;
  }
  const std::string label_;
  const unsigned int k_max_;
  mlmcpi_stats *st_ = nullptr;
};

// ------------------------------------------------------------------- QoI factories
/** qoi/quantityofinterest.hh:29-36 */
class QoIFactory {
public:
  explicit QoIFactory(int which_) : which(which_) {}
  virtual ~QoIFactory() {}
  virtual std::shared_ptr<QoI> get(std::shared_ptr<Action> action) { return std::make_shared<QoI>(action, which); }
  int id() const { return which; }

protected:
  const int which;
};
struct QoIXsquaredFactory : QoIFactory { QoIXsquaredFactory() : QoIFactory(MLMCPI_QOI_X2) {} };
struct QoISusceptibilityFactory : QoIFactory { QoISusceptibilityFactory() : QoIFactory(MLMCPI_QOI_ROTOR_CHI) {} };
struct QoI2DSusceptibilityFactory : QoIFactory { QoI2DSusceptibilityFactory() : QoIFactory(MLMCPI_QOI_SCHWINGER_CHI) {} };
struct QoIAvgPlaquetteFactory : QoIFactory { QoIAvgPlaquetteFactory() : QoIFactory(MLMCPI_QOI_AVG_PLAQUETTE) {} };
struct QoI2DPhiSquaredFactory : QoIFactory { QoI2DPhiSquaredFactory() : QoIFactory(MLMCPI_QOI_PHI2) {} };

// ------------------------------------------------------------------ batched samplers
/** Sampler (sampler/sampler.hh:20-43) for B chains: one mlmcpi_sampler object; the current
 * states live on the device ([B][n], reference dof order) */
class BatchedSampler {
public:
  BatchedSampler(const std::shared_ptr<Action> action, const mlmcpi_sampler_params &prm, unsigned int n_burnin,
                 bool autotune_hmc)
      : action_(action), B_(batch_size()), n_levels_(prm.n_levels),
        x_((size_t)action->sample_size() * batch_size()) {
    Device::check(mlmcpi_sampler_create(Device::ctx(), &action->model(), &prm, (int)B_,
                                        (uint32_t)(Parallel::rank() * B_), &s_),
                  "sampler create");
    // draw() leaves the output of a rejected chain untouched: start from the chains' states, not from zeros
    Device::check(mlmcpi_sampler_get_state(s_, x_.ptr()), "sampler state");
    // the reference's sampler constructors burn in and, for HMC, tune the step size
    // (sampler/hmcsampler.hh:99-108, overrelaxedheatbathsampler.hh:118-126, clustersampler.cc:30-35)
    for (unsigned int k = 0; k < n_burnin; ++k)
      draw();
    if (autotune_hmc && prm.kind == MLMCPI_SAMPLER_HMC) {
      double dt = 0, pa = 0;
      const int rc = mlmcpi_sampler_autotune(s_, 0.8, 100, 1000, &dt, &pa);
      pcout() << std::setprecision(6) << std::fixed;
      if (rc == 0)
        pcout() << "  Tuned         dt_{HMC} = " << dt << "  [ acceptance probability = " << pa << " ]" << std::endl;
      else
        pcout() << "  FAILED to tune HMC step size, reverting to dt_{HMC} = " << dt << std::endl;
    }
  }
  ~BatchedSampler() { mlmcpi_sampler_destroy(s_); }
  BatchedSampler(const BatchedSampler &) = delete;
  /** Sampler::draw for every chain; the new states are in states() */
  void draw() { Device::check(mlmcpi_sampler_draw(s_, x_.ptr(), nullptr), "Sampler::draw"); }
  double *states() { return x_.ptr(); }
  /** QoI::evaluate of every chain's state (mlmcpi_sampler_qoi: the susceptibility of a hierarchical Schwinger sampler
   *  comes out of the draw itself) */
  void evaluate(int qoi, double *d_q) { Device::check(mlmcpi_sampler_qoi(s_, qoi, d_q), "QoI::evaluate"); }
  unsigned int chains() const { return B_; }
  const std::shared_ptr<Action> &action() const { return action_; }
  double cost_per_sample() {
    double usec = 0;
    Device::check(mlmcpi_sampler_cost(s_, 10, &usec), "cost_per_sample");
    return usec;
  }
  /** sampler/hierarchicalsampler.cc:89-118 / montecarlo/mcmcstep.cc:8-14 */
  void show_stats() {
    std::vector<double> p(n_levels_);
    Device::check(mlmcpi_sampler_stats(s_, p.data()), "sampler stats");
    pcout() << std::setprecision(4) << std::fixed;
    pcout() << "  cost per sample = " << cost_per_sample() << " mu s per chain (" << B_ << " chains side by side)"
              << std::endl
              << std::endl;
    if (n_levels_ == 1) {
      pcout() << std::setprecision(5) << std::fixed;
      pcout() << "  acceptance probability  p = " << p[0] << std::endl;
      pcout() << "  rejection probability 1-p = " << 1. - p[0] << std::endl;
      return;
    }
    pcout() << "  acceptance rate = " << p[0] << std::endl;
    for (int l = 0; l < n_levels_; ++l)
      pcout() << "  level " << l << " "
                << (l == 0 ? "[finest]  " : (l == n_levels_ - 1 ? "[coarsest]" : "          ")) << " :  p = " << p[l]
                << std::endl;
  }

private:
  const std::shared_ptr<Action> action_;
  const unsigned int B_;
  const int n_levels_;
  DeviceVector x_;
  mlmcpi_sampler *s_ = nullptr;
};

/** sampler/sampler.hh:46-58; a factory is a recipe (mlmcpi_sampler_params) applied to an action */
class SamplerFactory {
public:
  virtual ~SamplerFactory() {}
  virtual mlmcpi_sampler_params params(const Action &action) const = 0;
  virtual unsigned int n_burnin() const { return 0; }
  virtual bool autotune() const { return false; }
  virtual std::shared_ptr<BatchedSampler> get(std::shared_ptr<Action> action) {
    return std::make_shared<BatchedSampler>(action, params(*action), n_burnin(), autotune());
  }

protected:
  static mlmcpi_sampler_params base(const Action &a, int kind) {
    mlmcpi_sampler_params p = {};
    p.kind = kind;
    p.n_levels = 1;
    p.renorm = a.get_renormalisation();
    p.ctype = a.get_coarsening_type();
    p.nt = 100;
    p.dt = 0.1;
    p.n_rep = 1;
    p.n_sweep_overrelax = 10;
    p.n_sweep_heatbath = 1;
    p.n_autocorr_window = 20;
    p.n_updates = 10;
    return p;
  }
};

class HMCSamplerFactory : public SamplerFactory { // sampler/hmcsampler.hh:167-190
public:
  explicit HMCSamplerFactory(const HMCParameters p) : p_(p) {}
  mlmcpi_sampler_params params(const Action &a) const {
    mlmcpi_sampler_params q = base(a, MLMCPI_SAMPLER_HMC);
    q.nt = (int)p_.nt();
    q.dt = p_.dt();
    q.n_rep = (int)p_.n_rep();
    return q;
  }
  unsigned int n_burnin() const { return p_.n_burnin(); }
  bool autotune() const { return true; }

private:
  const HMCParameters p_;
};

class OverrelaxedHeatBathSamplerFactory : public SamplerFactory { // overrelaxedheatbathsampler.hh:132-155
public:
  explicit OverrelaxedHeatBathSamplerFactory(const OverrelaxedHeatBathParameters p) : p_(p) {}
  mlmcpi_sampler_params params(const Action &a) const {
    mlmcpi_sampler_params q = base(a, MLMCPI_SAMPLER_HEATBATH);
    q.n_sweep_overrelax = (int)p_.n_sweep_overrelax();
    q.n_sweep_heatbath = (int)p_.n_sweep_heatbath();
    return q;
  }
  unsigned int n_burnin() const { return p_.n_burnin(); }

private:
  const OverrelaxedHeatBathParameters p_;
};

/** ClusterSampler (rotor) / QuenchedSchwingerClusterSampler: sampler/clustersampler.hh,
 * sampler/quenchedschwingerclustersampler.hh */
class ClusterSamplerFactory : public SamplerFactory {
public:
  explicit ClusterSamplerFactory(const ClusterParameters p) : p_(p) {}
  mlmcpi_sampler_params params(const Action &a) const {
    mlmcpi_sampler_params q = base(a, MLMCPI_SAMPLER_CLUSTER);
    q.n_updates = (int)p_.n_updates();
    return q;
  }
  unsigned int n_burnin() const { return p_.n_burnin(); }

private:
  const ClusterParameters p_;
};
typedef ClusterSamplerFactory QuenchedSchwingerClusterSamplerFactory;

/** sampler = 'exact': the action's own draw() (harmonic oscillator: Cholesky sampler,
 * qm/harmonicoscillatoraction.hh HarmonicOscillatorSamplerFactory) */
class ExactSamplerFactory : public SamplerFactory {
public:
  mlmcpi_sampler_params params(const Action &a) const { return base(a, MLMCPI_SAMPLER_EXACT); }
};

/** HierarchicalSampler (sampler/hierarchicalsampler.hh:141-170): the coarse factory's recipe
 * on the coarsest of n_max_level levels, two-level Metropolis steps above it */
class HierarchicalSamplerFactory : public SamplerFactory {
public:
  HierarchicalSamplerFactory(const std::shared_ptr<SamplerFactory> coarse_sampler_factory_,
                             const std::shared_ptr<ConditionedFineActionFactory> /*conditioned_fine_action_factory*/,
                             const HierarchicalParameters p)
      : coarse(coarse_sampler_factory_), p_(p) {}
  mlmcpi_sampler_params params(const Action &a) const {
    mlmcpi_sampler_params q = coarse->params(a);
    q.n_levels = std::max(1, (int)p_.n_max_level() - a.get_coarsening_level());
    return q;
  }
  /** the reference burns in and (HMC) tunes the coarsest-level sampler in its constructor
   * (hierarchicalsampler.cc:41, hmcsampler.hh:99-108) */
  std::shared_ptr<BatchedSampler> get(std::shared_ptr<Action> action) {
    return std::make_shared<BatchedSampler>(action, params(*action), coarse->n_burnin(), coarse->autotune());
  }

protected:
  const std::shared_ptr<SamplerFactory> coarse;
  const HierarchicalParameters p_;
};

/** MultilevelSampler (sampler/multilevelsampler.hh:160-200) */
class MultilevelSamplerFactory : public HierarchicalSamplerFactory {
public:
  MultilevelSamplerFactory(const std::shared_ptr<QoIFactory> qoi_factory_,
                           const std::shared_ptr<SamplerFactory> coarse_sampler_factory_,
                           const std::shared_ptr<ConditionedFineActionFactory> cfa, const StatisticsParameters param_stats,
                           const HierarchicalParameters p)
      : HierarchicalSamplerFactory(coarse_sampler_factory_, cfa, p), qoi(qoi_factory_->id()),
        window(param_stats.n_autocorr_window()) {}
  mlmcpi_sampler_params params(const Action &a) const {
    mlmcpi_sampler_params q = HierarchicalSamplerFactory::params(a);
    q.multilevel = 1;
    q.qoi = qoi;
    q.n_autocorr_window = (int)window;
    return q;
  }

private:
  const int qoi;
  const unsigned int window;
};

// ---------------------------------------------------------------------- timers
class Timer { // common/timer.hh
public:
  explicit Timer(const std::string &label_) : label(label_) {}
  void reset() { t = 0; }
  void start() { t0 = std::chrono::steady_clock::now(); }
  void stop() { t += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
  double elapsed() const { return t; }
  friend std::ostream &operator<<(std::ostream &os, const Timer &tm) {
    os << std::setprecision(3) << std::fixed << "[timer " << tm.label << "] : " << tm.t << " s";
    return os;
  }

private:
  const std::string label;
  double t = 0;
  std::chrono::steady_clock::time_point t0;
};

// ---------------------------------------------- montecarlo/montecarlosinglelevel.hh/.cc
class MonteCarloSingleLevel {
public:
  MonteCarloSingleLevel(std::shared_ptr<Action> action_, std::shared_ptr<QoI> qoi_,
                        std::shared_ptr<SamplerFactory> sampler_factory, const StatisticsParameters param_stats,
                        const SingleLevelMCParameters param_singlelevelmc)
      : action(action_), qoi(qoi_), sampler(sampler_factory->get(action_)), n_burnin(param_singlelevelmc.n_burnin()),
        n_min_samples_qoi(param_stats.n_min_samples_qoi()), n_samples(param_singlelevelmc.n_samples()),
        epsilon(param_singlelevelmc.epsilon()), B(batch_size()),
        stats_Q("Q", param_stats.n_autocorr_window(), batch_size()), q_(batch_size()), timer("SinglevelMC") {}

  /** montecarlosinglelevel.cc:23-94 */
  void evaluate() {
    stats_Q.hard_reset();
    for (unsigned int i = 0; i < n_burnin; ++i)
      sample();
    pcout() << "Burnin completed" << std::endl;
    const double two_epsilon_inv2 = 2. / (epsilon * epsilon);
    stats_Q.reset();
    unsigned long n_target = (n_samples > 0) ? n_samples : n_min_samples_qoi;
    unsigned long n_local_target = distribute_n(n_target), n_local = 0;
    timer.reset();
    timer.start();
    do {
      for (; n_local < n_local_target; ++n_local)
        sample();
      if (n_samples == 0)
        n_target = (unsigned long)std::ceil(stats_Q.tau_int() * two_epsilon_inv2 * stats_Q.variance());
      n_local_target = distribute_n(n_target);
    } while (n_local < n_local_target);
    Device::check(mlmcpi_sync(Device::ctx()), "sync");
    timer.stop();
    n_draws = n_local;
  }
  void show_statistics() {
    pcout() << stats_Q << std::endl;
    pcout() << timer << std::endl;
    const double sites = (action->model().model == MLMCPI_SCHWINGER) ? 0.5 * action->sample_size() : action->sample_size();
    const double all = (double)B * Parallel::world_size();
    pcout() << std::setprecision(3) << std::scientific << " throughput: " << (double)n_draws * all / timer.elapsed()
              << " samples/s, " << (double)n_draws * all / stats_Q.tau_int() / timer.elapsed()
              << " effective samples/s on " << Parallel::world_size() << " x " << B << " chains x "
              << (unsigned long)sites << " sites" << std::endl
              << std::endl;
  }
  double numerical_result() const { return stats_Q.average(); }
  double statistical_error() const { return stats_Q.error(); }
  std::shared_ptr<BatchedSampler> get_sampler() { return sampler; }

private:
  unsigned long distribute_n(unsigned long n) const { // mpi_wrapper.cc distribute_n over all chains
    const unsigned long all = (unsigned long)B * Parallel::world_size();
    return (n + all - 1) / all;
  }
  void sample() {
    sampler->draw();
    sampler->evaluate(qoi->id(), q_.ptr());
    stats_Q.record_sample(q_.ptr());
  }
  const std::shared_ptr<Action> action;
  const std::shared_ptr<QoI> qoi;
  const std::shared_ptr<BatchedSampler> sampler;
  const unsigned int n_burnin, n_min_samples_qoi, n_samples;
  const double epsilon;
  const unsigned int B;
  BatchedStatistics stats_Q;
  DeviceVector q_;
  Timer timer;
  unsigned long n_draws = 0;
};

// ------------------------------------------------- montecarlo/montecarlotwolevel.hh/.cc
class MonteCarloTwoLevel {
public:
  MonteCarloTwoLevel(const std::shared_ptr<Action> fine_action_, const std::shared_ptr<QoIFactory> qoi_factory_,
                     const std::shared_ptr<SamplerFactory> sampler_factory,
                     const std::shared_ptr<ConditionedFineActionFactory> /*conditioned_fine_action_factory*/,
                     const StatisticsParameters param_stats, const TwoLevelMCParameters param_twolevelmc)
      : n_burnin(param_twolevelmc.n_burnin()), n_samples(param_twolevelmc.n_samples()), B(batch_size()),
        fine_action(fine_action_), coarse_action(fine_action_->coarse_action()), qoi(qoi_factory_->id()),
        stats_fine("QoI[fine]", param_twolevelmc.n_fine_autocorr_window(), batch_size()),
        stats_coarse("QoI[coarse]", param_twolevelmc.n_coarse_autocorr_window(), batch_size()),
        stats_diff("delta QoI", param_twolevelmc.n_delta_autocorr_window(), batch_size()),
        stats_coarse_sampler("QoI[coarsesampler]", param_stats.n_autocorr_window(), batch_size()),
        theta((size_t)fine_action_->sample_size() * batch_size()), cache(6 * (size_t)batch_size()) {
    pcout() << "Twolevel Monte Carlo:" << std::endl;
    pcout() << "  fine action   : " << fine_action->info_string() << std::endl;
    pcout() << "  coarse action : " << coarse_action->info_string() << std::endl;
    coarse_sampler = sampler_factory->get(coarse_action);
    // TwoLevelMetropolisStep constructor: zero state and its cached actions (twolevelmetropolisstep.cc:11-22);
    // every one of the B chains has to forget its start: thermalised zero state (mlmcpi_thermal_state)
    Device::check(mlmcpi_thermal_state(Device::ctx(), &fine_action->model(), theta.ptr(), (int)B,
                                       (uint32_t)(Parallel::rank() * B)),
                  "thermal start");
    Device::check(mlmcpi_action(Device::ctx(), &fine_action->model(), theta.ptr(), (int)B, Sf()), "S_f");
    Device::check(mlmcpi_cond_action(Device::ctx(), &fine_action->model(), theta.ptr(), (int)B, Scond()), "S_cond");
  }

  /** montecarlotwolevel.cc:38-79 */
  void evaluate_difference() {
    stats_coarse.hard_reset();
    stats_coarse_sampler.hard_reset();
    stats_fine.hard_reset();
    stats_diff.hard_reset();
    for (unsigned int k = 0; k < n_burnin; ++k)
      sample();
    pcout() << "Burnin completed" << std::endl;
    stats_coarse_sampler.reset();
    const unsigned long all_chains = (unsigned long)B * Parallel::world_size();
    const unsigned long n_local_samples = (n_samples + all_chains - 1) / all_chains;
    stats_coarse.hard_reset();
    stats_fine.hard_reset();
    stats_diff.hard_reset();
    for (unsigned long k = 0; k < n_local_samples; ++k)
      sample();
  }
  /** montecarlotwolevel.cc:95-108 */
  void show_statistics() {
    pcout() << stats_fine << std::endl;
    pcout() << stats_coarse << std::endl;
    pcout() << stats_diff << std::endl;
    pcout() << std::endl;
    pcout() << "=== Coarse level sampler statistics === " << std::endl;
    pcout() << stats_coarse_sampler << std::endl;
    coarse_sampler->show_stats();
    pcout() << std::endl;
    pcout() << "=== Two level sampler statistics === " << std::endl;
    pcout() << std::setprecision(5) << std::fixed;
    const double p = n_total ? (double)n_accepted / ((double)n_total * B) : 0.0;
    pcout() << "  acceptance probability  p = " << p << std::endl;
    pcout() << "  rejection probability 1-p = " << 1. - p << std::endl;
  }
  const BatchedStatistics &get_stats_diff() const { return stats_diff; }
  const BatchedStatistics &get_stats_fine() const { return stats_fine; }
  const BatchedStatistics &get_stats_coarse() const { return stats_coarse; }

private:
  double *Sf() { return cache.ptr(); }
  double *Scond() { return cache.ptr() + B; }
  double *q_fine() { return cache.ptr() + 2 * B; }
  double *q_coarse() { return cache.ptr() + 3 * B; }
  double *q_diff() { return cache.ptr() + 4 * B; }
  int32_t *accept() { return reinterpret_cast<int32_t *>(cache.ptr() + 5 * B); }
  /** montecarlotwolevel.cc:82-93: skip ceil(2 tau_int) (at most 100) coarse draws between samples */
  void draw_coarse_sample() {
    const double two_tau_int = std::fmin(100., std::ceil(2. * stats_coarse_sampler.tau_int()));
    do {
      coarse_sampler->draw();
      Device::check(mlmcpi_qoi(Device::ctx(), &coarse_action->model(), qoi, coarse_sampler->states(), (int)B,
                               q_coarse(), nullptr),
                    "QoI::evaluate");
      stats_coarse_sampler.record_sample(q_coarse());
      t_sampler++;
    } while (t_sampler < two_tau_int);
    t_sampler = 0;
  }
  void sample() {
    draw_coarse_sample();
    Device::check(mlmcpi_twolevel_step(Device::ctx(), &fine_action->model(), &coarse_action->model(),
                                       coarse_sampler->states(), theta.ptr(), Sf(), Scond(), (int)B,
                                       (uint32_t)(Parallel::rank() * B), draw_counter++,
                                       accept(), nullptr),
                  "TwoLevelMetropolisStep::draw");
    Device::check(mlmcpi_qoi(Device::ctx(), &fine_action->model(), qoi, theta.ptr(), (int)B, q_fine(), nullptr),
                  "QoI::evaluate");
    Device::check(mlmcpi_qoi(Device::ctx(), &coarse_action->model(), qoi, coarse_sampler->states(), (int)B, q_coarse(),
                             nullptr),
                  "QoI::evaluate");
    Device::check(mlmcpi_axpy(Device::ctx(), q_diff(), q_fine(), -1.0, q_coarse(), B), "difference");
    stats_fine.record_sample(q_fine());
    stats_coarse.record_sample(q_coarse());
    stats_diff.record_sample(q_diff());
    std::vector<int32_t> acc(B + 1);
    Device::check(mlmcpi_download(Device::ctx(), reinterpret_cast<double *>(acc.data()),
                                  reinterpret_cast<const double *>(accept()), (B + 1) / 2),
                  "download");
    for (unsigned int c = 0; c < B; ++c)
      n_accepted += acc[c];
    n_total++;
  }
  const unsigned int n_burnin, n_samples, B;
  const std::shared_ptr<Action> fine_action, coarse_action;
  const int qoi;
  std::shared_ptr<BatchedSampler> coarse_sampler;
  BatchedStatistics stats_fine, stats_coarse, stats_diff, stats_coarse_sampler;
  DeviceVector theta, cache;
  uint64_t draw_counter = 0;
  unsigned long n_accepted = 0, n_total = 0;
  unsigned int t_sampler = 0;
};

// ---------------------------------------------- montecarlo/montecarlomultilevel.hh/.cc
class MonteCarloMultiLevel {
public:
  MonteCarloMultiLevel(std::shared_ptr<Action> fine_action_, std::shared_ptr<QoIFactory> qoi_factory_,
                       std::shared_ptr<SamplerFactory> sampler_factory,
                       std::shared_ptr<ConditionedFineActionFactory> /*conditioned_fine_action_factory*/,
                       const StatisticsParameters param_stats, const MultiLevelMCParameters param_multilevelmc)
      : n_level(param_multilevelmc.n_level()), epsilon(param_multilevelmc.epsilon()), timer("MultilevelMC") {
    // (the reference refuses to run this method on more than one rank, driver_qft.cc:409-414; here
    // the chains are sharded over the processes and every Statistics query of the allocation loop is
    // taken over all of them -- SURVEY 8e -- so all processes walk through the same loop)
    (void)Parallel::comm();
    mlmcpi_mlmc_params p = {};
    p.n_level = (int)n_level;
    p.n_burnin = (int)param_multilevelmc.n_burnin();
    p.epsilon = epsilon;
    p.n_autocorr_window = (int)param_stats.n_autocorr_window();
    p.n_min_samples_qoi = (int)param_stats.n_min_samples_qoi();
    p.qoi = qoi_factory_->id();
    p.max_iterations = 0;
    p.sampler = sampler_factory->params(*fine_action_);
    Device::check(mlmcpi_mlmc_create(Device::ctx(), &fine_action_->model(), &p, (int)batch_size(),
                                     (uint32_t)(Parallel::rank() * batch_size()), &m_),
                  "mlmc create");
  }
  ~MonteCarloMultiLevel() { mlmcpi_mlmc_destroy(m_); }
  /** montecarlomultilevel.cc:71-167 */
  void evaluate() {
    timer.reset();
    timer.start();
    const int rc = mlmcpi_mlmc_evaluate(m_);
    if (rc < 0)
      Device::check(rc, "MonteCarloMultiLevel::evaluate");
    timer.stop();
    level_.assign(6 * n_level, 0.0);
    Device::check(mlmcpi_mlmc_result(m_, &value_, &error_, level_.data()), "mlmc result");
  }
  double numerical_result() const { return value_; }
  double statistical_error() const { return error_; }
  /** montecarlomultilevel.cc:207-240 */
  void show_statistics() {
    pcout() << std::setprecision(6) << std::fixed;
    pcout() << " Q: Avg +/- Err = " << value_ << " +/- " << error_ << std::endl;
    pcout() << " tolerance epsilon = " << epsilon << std::endl;
    pcout() << timer << std::endl << std::endl;
  }
  /** per-level table: samples, mean and variance of Y_l, tau_int, effective cost, target */
  void show_detailed_statistics() {
    pcout() << " level      samples        E[Y_l]        Var[Y_l]     tau_int    cost_eff [mu s]    n_target" << std::endl;
    for (unsigned int l = 0; l < n_level; ++l) {
      const double *r = &level_[6 * l];
      pcout() << std::setw(6) << l << std::setw(13) << (unsigned long)r[0] << std::scientific << std::setprecision(4)
                << std::setw(14) << r[1] << std::setw(16) << r[2] << std::fixed << std::setprecision(3) << std::setw(12)
                << r[3] << std::setw(19) << r[4] << std::setw(12) << (unsigned long)r[5] << std::endl;
    }
    pcout() << std::endl;
  }

private:
  const unsigned int n_level;
  const double epsilon;
  Timer timer;
  mlmcpi_mlmc *m_ = nullptr;
  double value_ = 0, error_ = 0;
  std::vector<double> level_;
};

} // namespace mlmcpi
#endif // MLMCPI_MONTECARLO_HH
