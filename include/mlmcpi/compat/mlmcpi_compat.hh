// mlmcpi_compat.hh -- the reference's class names in the GLOBAL namespace, under the reference's
// header paths (every header of this directory tree is a one-line include of this file), so that
// the reference's own drivers compile UNCHANGED against the device library:
//
//   g++ -std=c++17 -Iinclude/mlmcpi/compat -Iinclude /root/reference/src/driver_qft.cc
//       -Lmlmcpathintegral_b200 -lmlmcpi -lmlmcpi_comm -o driver_qft
//
// (tests/test_reference_drivers.py does exactly this with src/driver_qft.cc and src/driver_qm.cc,
// byte for byte, and runs the binaries on the GPU).  What the drivers touch:
//   * parameter classes + readFile + operator<<      (common/parameters.hh and the headers that
//     define HMCParameters, ClusterParameters, ...)  -> mlmcpi/parameters.hh
//   * Lattice1D / Lattice2D, the Action subclasses, QoIs and their factories, the sampler and
//     conditioned-fine-action factories             -> mlmcpi/adapters.hh, mlmcpi/montecarlo.hh
//   * MonteCarloSingleLevel / TwoLevel / MultiLevel  -> mlmcpi/montecarlo.hh (B chains side by side)
//   * mpi_init / mpi_finalize / mpi_parallel::cout / mpi_exit / mpi_comm_size (mpi/mpi_wrapper.hh):
//     one process per GPU, NCCL all-reduce of the statistics moments -> mlmcpi::Parallel
//   * Timer, current_time, the analytic-result functions of common/auxilliary.hh.
// The number of chains per GPU comes from the optional `device: chains = N` section of the
// parameter file or the environment variable MLMCPI_CHAINS (the reference's drivers take exactly one
// command-line argument).
#ifndef MLMCPI_COMPAT_HH
#define MLMCPI_COMPAT_HH
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <iomanip>
#include <iostream>
#include <memory>
#include <string>

#include "mlmcpi/montecarlo.hh"

using namespace mlmcpi;

// ------------------------------------------------------------- mpi/mpi_wrapper.hh
// One process per GPU (MLMCPI_RANK / MLMCPI_WORLD_SIZE / MLMCPI_COMM_FILE, examples/run_multi_gpu.sh);
// the chains are sharded over the processes and Statistics all-reduce their moments over NCCL
// (mlmcpi::Parallel), which is the role MPI plays in the reference.
namespace mpi_parallel {
/** output on the master process only (MPIMasterStream, mpi/mpi_wrapper.hh:36-71) */
inline std::ostream &cout = mlmcpi::pcout();
inline std::ostream &master_cerr() {
  static std::ostream sink(nullptr);
  return mlmcpi::Parallel::master() ? std::cerr : sink;
}
inline std::ostream &cerr = master_cerr();
} // namespace mpi_parallel
inline void mpi_init() {}     // the communicator is set up on first use (mlmcpi::Parallel)
inline void mpi_finalize() {} // ... and torn down with the process
inline void mpi_exit(const int exit_code) { std::exit(exit_code); }
inline int mpi_comm_size() { return mlmcpi::Parallel::world_size(); }
inline int mpi_comm_rank() { return mlmcpi::Parallel::rank(); }
inline bool mpi_master() { return mlmcpi::Parallel::master(); }

// ------------------------------------------------------------------ common/timer.hh
inline std::string current_time() {
  std::time_t t = std::time(nullptr);
  char buf[64];
  std::strftime(buf, sizeof(buf), "%Y-%m-%d %H:%M:%S", std::localtime(&t));
  return buf;
}

// ------------------------------------------------ common/auxilliary.hh, qoi/qft/*.hh: analytic results
inline double quenchedschwinger_chit_analytical(const double beta, const unsigned int n_plaq) {
  return mlmcpi_schwinger_chit_analytical(beta, n_plaq); // qoi/qft/qoi2dsusceptibility.cc:30-39
}
inline double quenchedschwinger_chit_perturbative(const double beta, const unsigned int n_plaq) {
  return mlmcpi_schwinger_chit_perturbative(beta, n_plaq); // :42-45
}
inline double quenchedschwinger_var_chit_continuum_analytical(const double beta, const unsigned int n_plaq) {
  return mlmcpi_schwinger_var_chit_continuum(beta, n_plaq); // :47-50
}
inline double gff_phi_squared_analytical(const double mass, const unsigned int Mt_lat, const unsigned int Mx_lat) {
  return mlmcpi_gff_phi_squared_analytical(mass, (int)Mt_lat, (int)Mx_lat); // common/auxilliary.cc:197-209
}

// -------------------------------------------------------------- factories by their reference names
typedef mlmcpi::ExactSamplerFactory GFFSamplerFactory;                 // action/qft/gffaction.hh
typedef mlmcpi::ExactSamplerFactory HarmonicOscillatorSamplerFactory;  // action/qm/harmonicoscillatoraction.hh
typedef mlmcpi::ConditionedFineActionFactory QuenchedSchwingerConditionedFineActionFactory;
typedef mlmcpi::ConditionedFineActionFactory RotorConditionedFineActionFactory;

// ------------------------------------------------------------- O(3) sigma model: off the hot path
// (out of scope).  The reference's driver_qft.cc names these types, so they exist; selecting
// action = 'nonlinearsigma' ends the run with the reference's error convention (action/action.hh:48-52).
class NonlinearSigmaParameters : public mlmcpi::Parameters {
public:
  NonlinearSigmaParameters() : Parameters("nonlinearsigma"), beta_(1.0), renormalisation_(RenormalisationNone) {}
  double beta() const { return beta_; }
  RenormalisationType renormalisation() const { return renormalisation_; }

protected:
  void parse() {
    beta_ = getDouble("beta", Positive);
    renormalisation_ = (RenormalisationType)getChoice("renormalisation", renormalisation_choices());
  }

private:
  double beta_;
  RenormalisationType renormalisation_;
};
[[noreturn]] inline void nonlinearsigma_unsupported() {
  mpi_parallel::cerr << " ERROR: the nonlinear sigma model is not part of the device library." << std::endl;
  mpi_exit(EXIT_FAILURE);
  throw std::runtime_error("the nonlinear sigma model is not part of the device library");
}
class NonlinearSigmaAction : public mlmcpi::Action {
public:
  NonlinearSigmaAction(const std::shared_ptr<Lattice2D>, const std::shared_ptr<Lattice2D>, const RenormalisationType,
                       const double)
      : Action(mlmcpi_model{}, RenormalisationNone, 0, 0) {
    nonlinearsigma_unsupported();
  }
};
struct NonlinearSigmaConditionedFineActionFactory : mlmcpi::ConditionedFineActionFactory {};
struct QoI2DMagneticSusceptibility : mlmcpi::QoI {
  explicit QoI2DMagneticSusceptibility(std::shared_ptr<Lattice2D>) : QoI(mlmcpi_model{}, -1) {
    nonlinearsigma_unsupported();
  }
};
struct QoI2DMagneticSusceptibilityFactory : mlmcpi::QoIFactory {
  QoI2DMagneticSusceptibilityFactory() : QoIFactory(-1) {}
};

#endif
