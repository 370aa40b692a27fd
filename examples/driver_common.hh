// driver_common.hh -- pieces shared by driver_qm.cc and driver_qft.cc: reading a parameter
// section with echo, and the sampler-factory selection of the reference's
// construct_sampler_factory() helpers (driver_qm.cc:40-120, driver_qft.cc:41-105).
#ifndef MLMCPI_DRIVER_COMMON_HH
#define MLMCPI_DRIVER_COMMON_HH
#include <ctime>

#include "mlmcpi/montecarlo.hh"

namespace mlmcpi {

/** read one section, echo it (as the reference drivers do) */
template <class P> bool read_section(P &param, const std::string &filename) {
  if (param.readFile(filename))
    return false;
  pcout() << param << std::endl;
  return true;
}

inline std::string current_time() {
  std::time_t t = std::time(nullptr);
  char buf[64];
  std::strftime(buf, sizeof(buf), "%Y-%m-%d %H:%M:%S", std::localtime(&t));
  return buf;
}

/** the sampler factory for a sampler id of the parameter file; nullptr (after a message) if
 * the combination is not supported */
inline std::shared_ptr<SamplerFactory>
construct_sampler_factory(const int samplerid, const bool cluster_supported, const bool exact_supported,
                          const std::shared_ptr<QoIFactory> qoi_factory,
                          const std::shared_ptr<SamplerFactory> coarse_sampler_factory,
                          const std::shared_ptr<ConditionedFineActionFactory> conditioned_fine_action_factory,
                          const HMCParameters param_hmc, const ClusterParameters param_cluster,
                          const OverrelaxedHeatBathParameters param_heatbath,
                          const HierarchicalParameters param_hierarchical, const StatisticsParameters param_stats) {
  switch (samplerid) {
  case SamplerHMC:
    return std::make_shared<HMCSamplerFactory>(param_hmc);
  case SamplerOverrelaxedHeatBath:
    return std::make_shared<OverrelaxedHeatBathSamplerFactory>(param_heatbath);
  case SamplerHierarchical:
    return std::make_shared<HierarchicalSamplerFactory>(coarse_sampler_factory, conditioned_fine_action_factory,
                                                        param_hierarchical);
  case SamplerMultilevel:
    return std::make_shared<MultilevelSamplerFactory>(qoi_factory, coarse_sampler_factory,
                                                      conditioned_fine_action_factory, param_stats, param_hierarchical);
  case SamplerCluster:
    if (cluster_supported)
      return std::make_shared<ClusterSamplerFactory>(param_cluster);
    std::cerr << " ERROR: cluster not supported for chosen action." << std::endl;
    return nullptr;
  case SamplerExact:
    if (exact_supported)
      return std::make_shared<ExactSamplerFactory>();
    std::cerr << " ERROR: exact sampler not supported for chosen action." << std::endl;
    return nullptr;
  }
  std::cerr << " ERROR: Unsupported sampler." << std::endl;
  return nullptr;
}

inline void print_comparison(double numerical_result, double statistical_error, double analytical_result) {
  const double diff = std::fabs(numerical_result - analytical_result);
  pcout() << std::setprecision(8) << std::fixed;
  pcout() << "Comparison to analytical result " << std::endl;
  pcout() << "  (analytical - numerical) = " << diff;
  pcout() << std::setprecision(3) << std::fixed;
  pcout() << " = " << diff / statistical_error << " * (statistical error) " << std::endl << std::endl;
}

} // namespace mlmcpi
#endif
