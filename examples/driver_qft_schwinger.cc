// Single-level Monte Carlo for the quenched Schwinger model with the hierarchical sampler,
// written against include/mlmcpi/adapters.hh exactly the way the reference's driver_qft.cc
// (lines 227-262, 345-372) and MonteCarloSingleLevel::evaluate
// (montecarlo/montecarlosinglelevel.cc:23-94) use the reference classes.
//
//   g++ -std=c++17 -O2 -Iinclude examples/driver_qft_schwinger.cc \
//       -Lmlmcpathintegral_b200 -lmlmcpi -Wl,-rpath,$PWD/mlmcpathintegral_b200 -o driver_qft_schwinger
#include <cstdio>
#include <cstdlib>

#include "mlmcpi/adapters.hh"

using namespace mlmcpi;

int main(int argc, char *argv[]) {
  const unsigned int M = argc > 1 ? std::atoi(argv[1]) : 16;
  const double beta = argc > 2 ? std::atof(argv[2]) : 4.0;
  const unsigned int n_burnin = 100, n_samples = argc > 3 ? std::atoi(argv[3]) : 2000;

  std::shared_ptr<Lattice2D> lattice = std::make_shared<Lattice2D>(M, M, CoarsenBoth);
  std::shared_ptr<Action> action =
      std::make_shared<QuenchedSchwingerAction>(lattice, nullptr, RenormalisationPerturbative, beta);
  std::shared_ptr<QoI> qoi = std::make_shared<QoI2DSusceptibility>(action);
  Sampler::HMCParameters hmc;
  hmc.nt = 20;
  hmc.dt = 0.1;
  std::shared_ptr<Sampler> sampler = std::make_shared<HierarchicalSampler>(
      action, 2, MLMCPI_SAMPLER_HMC, RenormalisationPerturbative, CoarsenBoth, hmc);
  std::cout << "action: " << action->info_string() << std::endl;

  // --- the reference's Action / ConditionedFineAction interface on a single state
  std::shared_ptr<SampleState> phi = std::make_shared<SampleState>(action->sample_size());
  action->initialise_state(phi);
  std::shared_ptr<SampleState> p = std::make_shared<SampleState>(action->sample_size());
  action->force(phi, p);
  double psum = 0;
  for (size_t l = 0; l < p->data.size(); ++l)
    psum += p->data[l];
  std::printf("S = %.12f  sum(force) = %.3e (gauge invariance)\n", action->evaluate(phi), psum);
  std::shared_ptr<Action> coarse = action->coarse_action();
  std::shared_ptr<SampleState> phi_c = std::make_shared<SampleState>(coarse->sample_size());
  coarse->copy_from_fine(phi, phi_c);
  std::shared_ptr<ConditionedFineAction> cond = QuenchedSchwingerConditionedFineActionFactory().get(action);
  action->copy_from_coarse(phi_c, phi);
  cond->fill_fine_points(phi);
  std::printf("coarse beta = %.6f  S_c = %.6f  S_cond = %.6f\n", coarse->model().beta, coarse->evaluate(phi_c),
              cond->evaluate(phi));

  // --- MonteCarloSingleLevel::evaluate
  Statistics stats("QoI", 20);
  sampler->set_state(phi);
  for (unsigned int k = 0; k < n_burnin; ++k)
    sampler->draw(phi);
  sampler->reset_stats();
  for (unsigned int k = 0; k < n_samples; ++k) {
    sampler->draw(phi);
    stats.record_sample(qoi->evaluate(phi));
  }
  sampler->show_stats();
  std::printf("chi_t: Avg +/- Err = %.6f +/- %.6f  tau_int = %.3f  samples = %u\n", stats.average(), stats.error(),
              stats.tau_int(), stats.samples());
  return 0;
}
