// driver_qft -- the reference's 2-D lattice field theory driver (src/driver_qft.cc) on the
// device library: reads the same parameters_qft_*.in file, builds lattice, action, QoI and the
// requested sampler, and runs the single-, two- or multilevel Monte Carlo method on a batch of
// independent chains.  Supported actions: quenchedschwinger, gff (the O(3) sigma model is outside
// the device library).
//
//   g++ -std=c++17 -O2 -Iinclude examples/driver_qft.cc -Lmlmcpathintegral_b200 -lmlmcpi -lmlmcpi_comm
//       -Wl,-rpath,$PWD/mlmcpathintegral_b200 -o driver_qft
//   ./driver_qft PARAMETERFILE [CHAINS]
#include "driver_common.hh"

using namespace mlmcpi;

int main(int argc, char *argv[]) {
  Timer total_time("total");
  total_time.start();
  pcout() << "++===================================++" << std::endl;
  pcout() << "!!   Path integral multilevel MCMC   !!" << std::endl;
  pcout() << "!!   2D lattice field theories, B200 !!" << std::endl;
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
  QFTParameters param_qft;
  Lattice2DParameters param_lattice;
  StatisticsParameters param_stats;
  if (!read_section(param_general, filename) || !read_section(param_qft, filename) ||
      !read_section(param_lattice, filename) || !read_section(param_stats, filename))
    return 1;
  SchwingerParameters param_schwinger;
  GFFParameters param_gff;
  switch (param_qft.action()) {
  case ActionQuenchedSchwinger:
    if (!read_section(param_schwinger, filename))
      return 1;
    break;
  case ActionGFF:
    if (!read_section(param_gff, filename))
      return 1;
    break;
  default:
    std::cerr << " ERROR: the nonlinear sigma model is not part of the device library." << std::endl;
    return 1;
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
    /* ====== Lattice, quantity of interest, action ====== */
    std::shared_ptr<Lattice2D> lattice =
        std::make_shared<Lattice2D>(param_lattice.Mt_lat(), param_lattice.Mx_lat(), param_lattice.coarsening_type());
    const bool schwinger = (param_qft.action() == ActionQuenchedSchwinger);
    std::shared_ptr<Action> action;
    std::shared_ptr<QoIFactory> qoi_factory;
    pcout() << std::endl;
    if (schwinger) {
      action = std::make_shared<QuenchedSchwingerAction>(lattice, nullptr, param_schwinger.renormalisation(),
                                                         param_schwinger.beta());
      qoi_factory = std::make_shared<QoI2DSusceptibilityFactory>();
      pcout() << "QoI = Susceptibility Q[phi]^2 " << std::endl;
    } else {
      action = std::make_shared<GFFAction>(lattice, nullptr, param_gff.mass());
      qoi_factory = std::make_shared<QoI2DPhiSquaredFactory>();
      pcout() << "QoI = Mean squared field 1/M*sum phi^2 " << std::endl;
    }
    std::shared_ptr<QoI> qoi = qoi_factory->get(action);

    /* ====== Analytical results (driver_qft.cc:280-318) ====== */
    const unsigned int n_cells = param_lattice.Mt_lat() * param_lattice.Mx_lat();
    double analytical_result = 0.0, numerical_result = 0.0, statistical_error = 1.0;
    pcout() << std::endl << std::setprecision(8) << std::fixed;
    if (schwinger) {
      const double beta = param_schwinger.beta();
      analytical_result = (beta > 2000.0) ? mlmcpi_schwinger_chit_perturbative(beta, n_cells)
                                          : mlmcpi_schwinger_chit_analytical(beta, n_cells);
      pcout() << " Analytical results" << std::endl;
      pcout() << "      E[V*chi_t]              = " << analytical_result;
      if (beta > 2000.0)
        pcout() << " + O(beta^{-2}) = O(" << std::pow(beta, -2) << ")";
      pcout() << std::endl;
      pcout() << "      lim_{a->0} Var[V*chi_t] = " << mlmcpi_schwinger_var_chit_continuum(beta, n_cells) << std::endl
                << std::endl;
    } else {
      analytical_result =
          mlmcpi_gff_phi_squared_analytical(param_gff.mass(), param_lattice.Mt_lat(), param_lattice.Mx_lat());
      pcout() << " Analytical result" << std::endl;
      pcout() << "      E[Q^2]              = " << analytical_result << std::endl;
    }

    std::shared_ptr<ConditionedFineActionFactory> conditioned_fine_action_factory =
        std::make_shared<ConditionedFineActionFactory>();
    std::shared_ptr<SamplerFactory> coarse_sampler_factory = construct_sampler_factory(
        param_hierarchical.coarsesampler(), schwinger, !schwinger, nullptr, nullptr, nullptr, param_hmc, param_cluster,
        param_heatbath, param_hierarchical, param_stats);
    if (!coarse_sampler_factory)
      return 1;
    auto factory_for = [&](int samplerid) {
      return construct_sampler_factory(samplerid, schwinger, !schwinger, qoi_factory, coarse_sampler_factory,
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
      // (the reference refuses to run this method on more than one MPI rank; here the chains of
      // the batch step through the levels in lockstep)
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
    if (param_general.method() == MethodSingleLevel || param_general.method() == MethodMultiLevel)
      print_comparison(numerical_result, statistical_error, analytical_result);
  } catch (const std::exception &e) {
    return 1; // the message has been printed where the error was raised (action/action.hh:48-52)
  }
  total_time.stop();
  pcout() << total_time << std::endl;
  return 0;
}
