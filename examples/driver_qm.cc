// driver_qm -- the reference's quantum mechanics driver (src/driver_qm.cc) on the device
// library: harmonic oscillator, quartic oscillator and topological rotor; reads the same
// parameters_qm_*.in file and runs the single-, two- or multilevel method on a batch of
// independent chains.
//
//   g++ -std=c++17 -O2 -Iinclude examples/driver_qm.cc -Lmlmcpathintegral_b200 -lmlmcpi -lmlmcpi_comm
//       -Wl,-rpath,$PWD/mlmcpathintegral_b200 -o driver_qm
//   ./driver_qm PARAMETERFILE [CHAINS]
#include "driver_common.hh"

using namespace mlmcpi;

int main(int argc, char *argv[]) {
  Timer total_time("total");
  total_time.start();
  pcout() << "++===================================++" << std::endl;
  pcout() << "!!   Path integral multilevel MCMC   !!" << std::endl;
  pcout() << "!!   quantum mechanics, B200         !!" << std::endl;
  pcout() << "++===================================++" << std::endl << std::endl;
  pcout() << "Starting run at " << current_time() << std::endl;
  if (argc < 2 || argc > 3) {
    pcout() << "Usage: " << argv[0] << " PARAMETERFILE [CHAINS]" << std::endl << std::endl;
    return 0;
  }
  const std::string filename = argv[1];
  pcout() << " Reading parameter from file '" << filename << "'" << std::endl << std::endl;

  /* ====== Read parameters ====== */
  GeneralParameters param_general;
  QMParameters param_qm;
  Lattice1DParameters param_lattice;
  StatisticsParameters param_stats;
  if (!read_section(param_general, filename) || !read_section(param_qm, filename) ||
      !read_section(param_lattice, filename) || !read_section(param_stats, filename))
    return 1;
  HarmonicOscillatorParameters param_ho;
  QuarticOscillatorParameters param_qo;
  RotorParameters param_rotor;
  switch (param_qm.action()) {
  case ActionHarmonicOscillator:
    if (!read_section(param_ho, filename))
      return 1;
    break;
  case ActionQuarticOscillator:
    if (!read_section(param_qo, filename))
      return 1;
    break;
  case ActionRotor:
    if (!read_section(param_rotor, filename))
      return 1;
    break;
  }
  HMCParameters param_hmc;
  ClusterParameters param_cluster;
  OverrelaxedHeatBathParameters param_heatbath;
  SingleLevelMCParameters param_singlelevelmc;
  HierarchicalParameters param_hierarchical;
  TwoLevelMCParameters param_twolevelmc;
  MultiLevelMCParameters param_multilevelmc;
  DeviceParameters param_device;
  if (!read_section(param_hmc, filename) || !read_section(param_cluster, filename) ||
      !read_section(param_heatbath, filename) || !read_section(param_singlelevelmc, filename) ||
      !read_section(param_hierarchical, filename) || !read_section(param_twolevelmc, filename) ||
      !read_section(param_multilevelmc, filename) || !read_section(param_device, filename))
    return 1;
  batch_size() = (argc == 3) ? std::max(1, std::atoi(argv[2])) : param_device.chains();
  pcout() << "Running " << batch_size() << " independent chains side by side on the device." << std::endl;

  try {
    /* ====== Lattice, quantity of interest, action (driver_qm.cc:222-262) ====== */
    std::shared_ptr<Lattice1D> lattice = std::make_shared<Lattice1D>(param_lattice.M_lat(), param_lattice.T_final());
    const bool rotor = (param_qm.action() == ActionRotor);
    std::shared_ptr<Action> action;
    std::shared_ptr<QoIFactory> qoi_factory;
    if (rotor) {
      action = std::make_shared<RotorAction>(lattice, param_rotor.renormalisation(), param_rotor.m0());
      qoi_factory = std::make_shared<QoISusceptibilityFactory>();
      pcout() << "QoI = Susceptibility Q[X]^2/T " << std::endl;
    } else {
      if (param_qm.action() == ActionHarmonicOscillator)
        action = std::make_shared<HarmonicOscillatorAction>(lattice, param_ho.renormalisation(), param_ho.m0(),
                                                            param_ho.mu2());
      else
        action = std::make_shared<QuarticOscillatorAction>(lattice, RenormalisationNone, param_qo.m0(), param_qo.mu2(),
                                                           param_qo.lambda(), param_qo.x0());
      qoi_factory = std::make_shared<QoIXsquaredFactory>();
      pcout() << "QoI = X^2 " << std::endl;
    }
    std::shared_ptr<QoI> qoi = qoi_factory->get(action);
    pcout() << std::endl;

    /* ====== Analytical results (driver_qm.cc:268-303) ====== */
    const bool has_analytical = (param_qm.action() != ActionQuarticOscillator);
    const bool estimates_mean =
        (param_general.method() == MethodSingleLevel || param_general.method() == MethodMultiLevel);
    double analytical_result = 0.0, numerical_result = 0.0, statistical_error = 1.0;
    const double a_lat = lattice->geta_lat();
    if (estimates_mean && param_qm.action() == ActionHarmonicOscillator) {
      analytical_result = mlmcpi_ho_xsquared_analytical(param_ho.m0(), param_ho.mu2(), a_lat, param_lattice.M_lat(), 0);
      pcout() << std::endl << std::setprecision(6) << std::fixed;
      pcout() << " Analytical result        <x^2> = " << analytical_result << std::endl;
      pcout() << " Continuum limit [a -> 0] <x^2> = "
                << mlmcpi_ho_xsquared_analytical(param_ho.m0(), param_ho.mu2(), a_lat, param_lattice.M_lat(), 1)
                << std::endl
                << std::endl;
    }
    if (estimates_mean && rotor) {
      const double m0 = param_rotor.m0(), T = param_lattice.T_final();
      analytical_result = mlmcpi_rotor_chit(m0, a_lat, T, 0);
      pcout() << std::endl << std::setprecision(6) << std::fixed;
      pcout() << " Analytical result        <chi_t> = " << analytical_result << std::endl;
      pcout() << " Perturbative expansion   <chi_t> = " << mlmcpi_rotor_chit(m0, a_lat, T, 1)
                << " + O((a/I)^2), a/I = " << a_lat / m0 << std::endl;
      pcout() << " Continuum limit [a -> 0] <chi_t> = " << mlmcpi_rotor_chit(m0, a_lat, T, 2) << std::endl
                << std::endl;
    }

    std::shared_ptr<ConditionedFineActionFactory> conditioned_fine_action_factory =
        std::make_shared<ConditionedFineActionFactory>();
    std::shared_ptr<SamplerFactory> coarse_sampler_factory = construct_sampler_factory(
        param_hierarchical.coarsesampler(), rotor, param_qm.action() == ActionHarmonicOscillator, nullptr, nullptr, nullptr, param_hmc, param_cluster, param_heatbath,
        param_hierarchical, param_stats);
    if (!coarse_sampler_factory)
      return 1;
    auto factory_for = [&](int samplerid) {
      return construct_sampler_factory(samplerid, rotor, param_qm.action() == ActionHarmonicOscillator, qoi_factory, coarse_sampler_factory,
                                       conditioned_fine_action_factory, param_hmc, param_cluster, param_heatbath,
                                       param_hierarchical, param_stats);
    };

    if (param_general.method() == MethodSingleLevel) {
      pcout() << "+--------------------------------+" << std::endl;
      pcout() << "! Single level MC                !" << std::endl;
      pcout() << "+--------------------------------+" << std::endl << std::endl;
      std::shared_ptr<SamplerFactory> sampler_factory = factory_for(param_singlelevelmc.sampler());
      if (!sampler_factory)
        return 1;
      MonteCarloSingleLevel montecarlo_singlelevel(action, qoi, sampler_factory, param_stats, param_singlelevelmc);
      montecarlo_singlelevel.evaluate();
      pcout() << std::endl;
      montecarlo_singlelevel.show_statistics();
      numerical_result = montecarlo_singlelevel.numerical_result();
      statistical_error = montecarlo_singlelevel.statistical_error();
      pcout() << "=== Sampler statistics === " << std::endl;
      montecarlo_singlelevel.get_sampler()->show_stats();
      pcout() << std::endl;
    }
    if (param_general.method() == MethodTwoLevel) {
      pcout() << "+--------------------------------+" << std::endl;
      pcout() << "! Two level MC                   !" << std::endl;
      pcout() << "+--------------------------------+" << std::endl << std::endl;
      std::shared_ptr<SamplerFactory> sampler_factory = factory_for(param_twolevelmc.sampler());
      if (!sampler_factory)
        return 1;
      MonteCarloTwoLevel montecarlo_twolevel(action, qoi_factory, sampler_factory, conditioned_fine_action_factory,
                                             param_stats, param_twolevelmc);
      montecarlo_twolevel.evaluate_difference();
      montecarlo_twolevel.show_statistics();
      pcout() << std::endl;
    }
    if (param_general.method() == MethodMultiLevel) {
      pcout() << "+--------------------------------+" << std::endl;
      pcout() << "! Multilevel MC                  !" << std::endl;
      pcout() << "+--------------------------------+" << std::endl << std::endl;
      std::shared_ptr<SamplerFactory> sampler_factory = factory_for(param_multilevelmc.sampler());
      if (!sampler_factory)
        return 1;
      MonteCarloMultiLevel montecarlo_multilevel(action, qoi_factory, sampler_factory,
                                                 conditioned_fine_action_factory, param_stats, param_multilevelmc);
      montecarlo_multilevel.evaluate();
      montecarlo_multilevel.show_statistics();
      if (param_multilevelmc.show_detailed_stats())
        montecarlo_multilevel.show_detailed_statistics();
      numerical_result = montecarlo_multilevel.numerical_result();
      statistical_error = montecarlo_multilevel.statistical_error();
    }
    if (has_analytical && estimates_mean)
      print_comparison(numerical_result, statistical_error, analytical_result);
  } catch (const std::exception &e) {
    return 1; // the message has been printed where the error was raised (action/action.hh:48-52)
  }
  total_time.stop();
  pcout() << total_time << std::endl;
  return 0;
}
