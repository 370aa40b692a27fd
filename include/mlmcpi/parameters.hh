// parameters.hh -- reader for the reference's parameters_*.in files and the typed parameter
// classes of its drivers (same class and accessor names, same sections and keys).
//
// File format (common/parameters.cc of the reference): sections introduced by `name:`,
// entries `key = value` with integer, floating point, boolean (true/false) or quoted string
// values, `#` starts a comment.  readFile() returns 0 on success and non-zero after printing a
// message, which is what the drivers test (`if (param.readFile(filename)) return 1;`,
// driver_qft.cc:133-199).
//
// One tokenising reader (ParameterFile) replaces the reference's per-class regex machinery;
// every typed class below names its section, its keys with their constraints
// (common/parameters.hh NumConstraintFlag) and converts the string-valued options to the enums
// the drivers switch on.
#ifndef MLMCPI_PARAMETERS_HH
#define MLMCPI_PARAMETERS_HH
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "adapters.hh"

namespace mlmcpi {

enum MethodType { MethodSingleLevel = 0, MethodTwoLevel = 1, MethodMultiLevel = 2 }; // common/parameters.hh
enum SamplerType {
  SamplerHMC = 0,
  SamplerOverrelaxedHeatBath = 1,
  SamplerCluster = 2,
  SamplerExact = 3,
  SamplerHierarchical = 4,
  SamplerMultilevel = 5
};
enum QMActionType { ActionHarmonicOscillator = 0, ActionQuarticOscillator = 1, ActionRotor = 2 }; // qm/qmaction.hh
enum QFTActionType { ActionQuenchedSchwinger = 0, ActionNonlinearSigma = 1, ActionGFF = 2 };      // qft/qftaction.hh
enum NumConstraintFlag { AnyValue = 0, Positive = 1, NonNegative = 2, Negative = 3, NonPositive = 4 };

/** all `section: key = value` entries of one file */
class ParameterFile {
public:
  /** returns 0 on success */
  int load(const std::string &filename) {
    std::ifstream is(filename.c_str());
    if (!is) {
      std::cerr << " ERROR: cannot open parameter file '" << filename << "'" << std::endl;
      return 1;
    }
    entries_.clear();
    std::string line, section;
    int lineno = 0;
    while (std::getline(is, line)) {
      ++lineno;
      const std::string t = strip(uncomment(line));
      if (t.empty())
        continue;
      const size_t eq = t.find('=');
      if (eq == std::string::npos) {
        if (t.size() > 1 && t[t.size() - 1] == ':' && is_identifier(t.substr(0, t.size() - 1))) {
          section = t.substr(0, t.size() - 1);
          continue;
        }
        std::cerr << " ERROR: cannot parse line " << lineno << " of '" << filename << "': " << line << std::endl;
        return 1;
      }
      const std::string key = strip(t.substr(0, eq)), value = strip(t.substr(eq + 1));
      if (section.empty() || !is_identifier(key) || value.empty()) {
        std::cerr << " ERROR: cannot parse line " << lineno << " of '" << filename << "': " << line << std::endl;
        return 1;
      }
      entries_[section + ":" + key] = value;
    }
    return 0;
  }
  bool has(const std::string &section, const std::string &key) const {
    return entries_.count(section + ":" + key) != 0;
  }
  const std::string &raw(const std::string &section, const std::string &key) const {
    return entries_.find(section + ":" + key)->second;
  }

private:
  static std::string uncomment(const std::string &s) { // '#' outside quotes starts a comment
    char quote = 0;
    for (size_t i = 0; i < s.size(); ++i) {
      if (quote) {
        if (s[i] == quote)
          quote = 0;
      } else if (s[i] == '\'' || s[i] == '"') {
        quote = s[i];
      } else if (s[i] == '#') {
        return s.substr(0, i);
      }
    }
    return s;
  }
  static std::string strip(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a]))
      ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1]))
      --b;
    return s.substr(a, b - a);
  }
  static bool is_identifier(const std::string &s) {
    if (s.empty())
      return false;
    for (char c : s)
      if (!(std::isalnum((unsigned char)c) || c == '_'))
        return false;
    return true;
  }
  std::map<std::string, std::string> entries_;
};

/** base of the typed parameter classes (common/parameters.hh `class Parameters`) */
class Parameters {
public:
  explicit Parameters(const std::string &section) : section_(section) {}
  virtual ~Parameters() {}
  /** read this class's section; 0 on success */
  int readFile(const std::string &filename) {
    ParameterFile f;
    if (f.load(filename))
      return 1;
    file_ = &f;
    error_ = 0;
    shown_.clear();
    parse();
    file_ = nullptr;
    return error_;
  }
  friend std::ostream &operator<<(std::ostream &os, const Parameters &p) {
    os << " " << p.section_ << ":" << std::endl;
    for (const auto &kv : p.shown_)
      os << "    " << kv.first << " = " << kv.second << std::endl;
    return os;
  }

protected:
  virtual void parse() = 0;
  int getInt(const std::string &key, NumConstraintFlag c = AnyValue) {
    const std::string *v = lookup(key);
    if (!v)
      return 0;
    char *end = nullptr;
    const long x = std::strtol(v->c_str(), &end, 10);
    if (end == v->c_str() || *end != 0)
      return fail(key, "is not an integer"), 0;
    check((double)x, key, c);
    return (int)x;
  }
  double getDouble(const std::string &key, NumConstraintFlag c = AnyValue) {
    const std::string *v = lookup(key);
    if (!v)
      return 0.0;
    char *end = nullptr;
    const double x = std::strtod(v->c_str(), &end);
    if (end == v->c_str() || *end != 0)
      return fail(key, "is not a number"), 0.0;
    check(x, key, c);
    return x;
  }
  bool getBool(const std::string &key) {
    const std::string *v = lookup(key);
    if (!v)
      return false;
    if (*v == "true" || *v == "True" || *v == "TRUE")
      return true;
    if (*v == "false" || *v == "False" || *v == "FALSE")
      return false;
    return fail(key, "is not a boolean"), false;
  }
  std::string getString(const std::string &key) {
    const std::string *v = lookup(key);
    if (!v)
      return "";
    if (v->size() < 2 || (v->front() != '\'' && v->front() != '"') || v->back() != v->front())
      return fail(key, "is not a quoted string"), "";
    return v->substr(1, v->size() - 2);
  }
  /** optional integer key (tolerates files written for an older set of keys) */
  int getIntOr(const std::string &key, int fallback, NumConstraintFlag c = AnyValue) {
    return file_->has(section_, key) ? getInt(key, c) : fallback;
  }
  /** string option -> enum; unknown strings are an error listing the allowed values */
  int getChoice(const std::string &key, const std::vector<std::pair<std::string, int>> &allowed) {
    const std::string s = getString(key);
    if (error_)
      return 0;
    for (const auto &a : allowed)
      if (a.first == s)
        return a.second;
    std::string list;
    for (const auto &a : allowed)
      list += (list.empty() ? "" : ", ") + a.first;
    fail(key, ("= '" + s + "' is invalid, allowed values: [" + list + "]").c_str());
    return 0;
  }
  void fail(const std::string &key, const char *what) {
    std::cerr << " ERROR: parameter '" << key << "' in section '" << section_ << "' " << what << std::endl;
    error_ = 1;
  }

private:
  const std::string *lookup(const std::string &key) {
    if (!file_->has(section_, key)) {
      fail(key, "is missing");
      return nullptr;
    }
    const std::string &v = file_->raw(section_, key);
    shown_.push_back(std::make_pair(key, v));
    return &v;
  }
  void check(double x, const std::string &key, NumConstraintFlag c) {
    const bool ok = (c == AnyValue) || (c == Positive && x > 0) || (c == NonNegative && x >= 0) ||
                    (c == Negative && x < 0) || (c == NonPositive && x <= 0);
    if (!ok)
      fail(key, "violates its sign constraint");
  }
  const std::string section_;
  const ParameterFile *file_ = nullptr;
  int error_ = 0;
  std::vector<std::pair<std::string, std::string>> shown_;
};

static inline const std::vector<std::pair<std::string, int>> &renormalisation_choices() {
  static const std::vector<std::pair<std::string, int>> c = {{"none", RenormalisationNone},
                                                             {"perturbative", RenormalisationPerturbative},
                                                             {"nonperturbative", RenormalisationNonperturbative}};
  return c;
}

#define MLMCPI_PARAM_GETTER(type, name)                                                             \
private:                                                                                            \
  type name##_ = 0;                                                                                 \
                                                                                                    \
public:                                                                                             \
  type name() const { return name##_; }

/** general: method (common/parameters.hh GeneralParameters) */
class GeneralParameters : public Parameters {
public:
  GeneralParameters() : Parameters("general") {}
  MLMCPI_PARAM_GETTER(int, method)

protected:
  void parse() {
    method_ = getChoice("method", {{"singlelevel", MethodSingleLevel},
                                   {"twolevel", MethodTwoLevel},
                                   {"multilevel", MethodMultiLevel}});
  }
};

/** quantummechanics: action (action/qm/qmaction.hh) */
class QMParameters : public Parameters {
public:
  QMParameters() : Parameters("quantummechanics") {}
  MLMCPI_PARAM_GETTER(int, action)

protected:
  void parse() {
    action_ = getChoice("action", {{"harmonicoscillator", ActionHarmonicOscillator},
                                   {"quarticoscillator", ActionQuarticOscillator},
                                   {"rotor", ActionRotor}});
  }
};

/** quantumfieldtheory: action (action/qft/qftaction.hh) */
class QFTParameters : public Parameters {
public:
  QFTParameters() : Parameters("quantumfieldtheory") {}
  MLMCPI_PARAM_GETTER(int, action)

protected:
  void parse() {
    action_ = getChoice("action", {{"quenchedschwinger", ActionQuenchedSchwinger},
                                   {"nonlinearsigma", ActionNonlinearSigma},
                                   {"gff", ActionGFF}});
  }
};

/** lattice: M_lat, T_final (lattice/lattice1d.hh) */
class Lattice1DParameters : public Parameters {
public:
  Lattice1DParameters() : Parameters("lattice") {}
  MLMCPI_PARAM_GETTER(unsigned int, M_lat)
  MLMCPI_PARAM_GETTER(double, T_final)

protected:
  void parse() {
    M_lat_ = getInt("M_lat", Positive);
    T_final_ = getDouble("T_final", Positive);
  }
};

/** lattice: Mt_lat, Mx_lat, coarsening (lattice/lattice2d.hh) */
class Lattice2DParameters : public Parameters {
public:
  Lattice2DParameters() : Parameters("lattice") {}
  MLMCPI_PARAM_GETTER(unsigned int, Mt_lat)
  MLMCPI_PARAM_GETTER(unsigned int, Mx_lat)
  CoarseningType coarsening_type() const { return (CoarseningType)coarsening_; }

protected:
  void parse() {
    Mt_lat_ = getInt("Mt_lat", Positive);
    Mx_lat_ = getInt("Mx_lat", Positive);
    coarsening_ = getChoice("coarsening", {{"both", CoarsenBoth},
                                           {"temporal", CoarsenTemporal},
                                           {"spatial", CoarsenSpatial},
                                           {"alternate", CoarsenAlternate},
                                           {"rotate", CoarsenRotate}});
  }

private:
  int coarsening_ = CoarsenBoth;
};

/** statistics: n_autocorr_window, n_min_samples_qoi (common/statistics.hh) */
class StatisticsParameters : public Parameters {
public:
  StatisticsParameters() : Parameters("statistics") {}
  MLMCPI_PARAM_GETTER(unsigned int, n_autocorr_window)
  MLMCPI_PARAM_GETTER(unsigned int, n_min_samples_qoi)

protected:
  void parse() {
    n_autocorr_window_ = getInt("n_autocorr_window", Positive);
    n_min_samples_qoi_ = getInt("n_min_samples_qoi", Positive);
  }
};

/** harmonicoscillator: m0, mu2, renormalisation (action/qm/harmonicoscillatoraction.hh) */
class HarmonicOscillatorParameters : public Parameters {
public:
  HarmonicOscillatorParameters() : Parameters("harmonicoscillator") {}
  MLMCPI_PARAM_GETTER(double, m0)
  MLMCPI_PARAM_GETTER(double, mu2)
  RenormalisationType renormalisation() const { return (RenormalisationType)renormalisation_; }

protected:
  void parse() {
    m0_ = getDouble("m0", Positive);
    mu2_ = getDouble("mu2");
    renormalisation_ = getChoice("renormalisation", renormalisation_choices());
  }

private:
  int renormalisation_ = 0;
};

/** quarticoscillator: m0, mu2, lambda, x0 (action/qm/quarticoscillatoraction.hh) */
class QuarticOscillatorParameters : public Parameters {
public:
  QuarticOscillatorParameters() : Parameters("quarticoscillator") {}
  MLMCPI_PARAM_GETTER(double, m0)
  MLMCPI_PARAM_GETTER(double, mu2)
  MLMCPI_PARAM_GETTER(double, lambda)
  MLMCPI_PARAM_GETTER(double, x0)

protected:
  void parse() {
    m0_ = getDouble("m0", Positive);
    mu2_ = getDouble("mu2");
    lambda_ = getDouble("lambda", NonNegative);
    x0_ = getDouble("x0");
  }
};

/** rotor: m0, renormalisation (action/qm/rotoraction.hh) */
class RotorParameters : public Parameters {
public:
  RotorParameters() : Parameters("rotor") {}
  MLMCPI_PARAM_GETTER(double, m0)
  RenormalisationType renormalisation() const { return (RenormalisationType)renormalisation_; }

protected:
  void parse() {
    m0_ = getDouble("m0", Positive);
    renormalisation_ = getChoice("renormalisation", renormalisation_choices());
  }

private:
  int renormalisation_ = 0;
};

/** schwinger: beta, renormalisation (action/qft/quenchedschwingeraction.hh) */
class SchwingerParameters : public Parameters {
public:
  SchwingerParameters() : Parameters("schwinger") {}
  MLMCPI_PARAM_GETTER(double, beta)
  RenormalisationType renormalisation() const { return (RenormalisationType)renormalisation_; }

protected:
  void parse() {
    beta_ = getDouble("beta", Positive);
    renormalisation_ = getChoice("renormalisation", renormalisation_choices());
  }

private:
  int renormalisation_ = 0;
};

/** gff: mass, renormalisation (action/qft/gffaction.hh) */
class GFFParameters : public Parameters {
public:
  GFFParameters() : Parameters("gff") {}
  MLMCPI_PARAM_GETTER(double, mass)
  RenormalisationType renormalisation() const { return (RenormalisationType)renormalisation_; }

protected:
  void parse() {
    mass_ = getDouble("mass", Positive);
    renormalisation_ = getChoice("renormalisation", renormalisation_choices());
  }

private:
  int renormalisation_ = 0;
};

static inline const std::vector<std::pair<std::string, int>> &base_sampler_choices() {
  static const std::vector<std::pair<std::string, int>> c = {
      {"HMC", SamplerHMC}, {"heatbath", SamplerOverrelaxedHeatBath}, {"cluster", SamplerCluster}, {"exact", SamplerExact}};
  return c;
}

/** singlelevelmc: n_burnin, n_samples, epsilon, sampler (montecarlo/montecarlosinglelevel.hh) */
class SingleLevelMCParameters : public Parameters {
public:
  SingleLevelMCParameters() : Parameters("singlelevelmc") {}
  MLMCPI_PARAM_GETTER(unsigned int, n_burnin)
  MLMCPI_PARAM_GETTER(unsigned int, n_samples)
  MLMCPI_PARAM_GETTER(double, epsilon)
  MLMCPI_PARAM_GETTER(int, sampler)

protected:
  void parse() {
    n_burnin_ = getInt("n_burnin", Positive);
    n_samples_ = getInt("n_samples", NonNegative);
    epsilon_ = getDouble("epsilon", Positive);
    std::vector<std::pair<std::string, int>> c = base_sampler_choices();
    c.push_back({"hierarchical", SamplerHierarchical});
    c.push_back({"multilevel", SamplerMultilevel});
    sampler_ = getChoice("sampler", c);
  }
};

/** twolevelmc (montecarlo/montecarlotwolevel.hh); the three window keys default to 10 when
 * absent, which lets parameters_qm_template.in be used unchanged */
class TwoLevelMCParameters : public Parameters {
public:
  TwoLevelMCParameters() : Parameters("twolevelmc") {}
  MLMCPI_PARAM_GETTER(unsigned int, n_burnin)
  MLMCPI_PARAM_GETTER(unsigned int, n_samples)
  MLMCPI_PARAM_GETTER(unsigned int, n_coarse_autocorr_window)
  MLMCPI_PARAM_GETTER(unsigned int, n_fine_autocorr_window)
  MLMCPI_PARAM_GETTER(unsigned int, n_delta_autocorr_window)
  MLMCPI_PARAM_GETTER(int, sampler)

protected:
  void parse() {
    n_burnin_ = getInt("n_burnin", Positive);
    n_samples_ = getInt("n_samples", Positive);
    n_coarse_autocorr_window_ = getIntOr("n_coarse_autocorr_window", 10, Positive);
    n_fine_autocorr_window_ = getIntOr("n_fine_autocorr_window", 10, Positive);
    n_delta_autocorr_window_ = getIntOr("n_delta_autocorr_window", 10, Positive);
    std::vector<std::pair<std::string, int>> c = base_sampler_choices();
    c.push_back({"hierarchical", SamplerHierarchical});
    sampler_ = getChoice("sampler", c);
  }
};

/** multilevelmc (montecarlo/montecarlomultilevel.hh) */
class MultiLevelMCParameters : public Parameters {
public:
  MultiLevelMCParameters() : Parameters("multilevelmc") {}
  MLMCPI_PARAM_GETTER(unsigned int, n_level)
  MLMCPI_PARAM_GETTER(unsigned int, n_burnin)
  MLMCPI_PARAM_GETTER(double, epsilon)
  MLMCPI_PARAM_GETTER(bool, show_detailed_stats)
  MLMCPI_PARAM_GETTER(int, sampler)

protected:
  void parse() {
    n_level_ = getInt("n_level", Positive);
    n_burnin_ = getInt("n_burnin", Positive);
    epsilon_ = getDouble("epsilon", Positive);
    show_detailed_stats_ = getBool("show_detailed_stats");
    sampler_ = getChoice("sampler", {{"hierarchical", SamplerHierarchical},
                                     {"multilevel", SamplerMultilevel},
                                     {"cluster", SamplerCluster}});
  }
};

/** hierarchical: n_max_level, coarsesampler (sampler/hierarchicalsampler.hh) */
class HierarchicalParameters : public Parameters {
public:
  HierarchicalParameters() : Parameters("hierarchical") {}
  MLMCPI_PARAM_GETTER(unsigned int, n_max_level)
  MLMCPI_PARAM_GETTER(int, coarsesampler)

protected:
  void parse() {
    n_max_level_ = getInt("n_max_level", Positive);
    coarsesampler_ = getChoice("coarsesampler", base_sampler_choices());
  }
};

/** hmc: nt, dt, n_burnin, n_rep (sampler/hmcsampler.hh:21-65) */
class HMCParameters : public Parameters {
public:
  HMCParameters() : Parameters("hmc") {}
  MLMCPI_PARAM_GETTER(unsigned int, nt)
  MLMCPI_PARAM_GETTER(double, dt)
  MLMCPI_PARAM_GETTER(unsigned int, n_burnin)
  MLMCPI_PARAM_GETTER(unsigned int, n_rep)

protected:
  void parse() {
    nt_ = getInt("nt", Positive);
    dt_ = getDouble("dt", Positive);
    n_burnin_ = getInt("n_burnin", Positive);
    n_rep_ = getInt("n_rep", Positive);
  }
};

/** heatbath (sampler/overrelaxedheatbathsampler.hh).  random_order is read for compatibility:
 * the device sweeps are coloured (all sites of a colour at once), which has no visiting order */
class OverrelaxedHeatBathParameters : public Parameters {
public:
  OverrelaxedHeatBathParameters() : Parameters("heatbath") {}
  MLMCPI_PARAM_GETTER(unsigned int, n_sweep_overrelax)
  MLMCPI_PARAM_GETTER(unsigned int, n_sweep_heatbath)
  MLMCPI_PARAM_GETTER(bool, random_order)
  MLMCPI_PARAM_GETTER(unsigned int, n_burnin)

protected:
  void parse() {
    n_sweep_overrelax_ = getInt("n_sweep_overrelax", Positive);
    n_sweep_heatbath_ = getInt("n_sweep_heatbath", Positive);
    random_order_ = getBool("random_order");
    n_burnin_ = getInt("n_burnin", Positive);
  }
};

/** clusteralgorithm: n_burnin, n_updates (sampler/clustersampler.hh) */
class ClusterParameters : public Parameters {
public:
  ClusterParameters() : Parameters("clusteralgorithm") {}
  MLMCPI_PARAM_GETTER(unsigned int, n_burnin)
  MLMCPI_PARAM_GETTER(unsigned int, n_updates)

protected:
  void parse() {
    n_burnin_ = getInt("n_burnin", Positive);
    n_updates_ = getInt("n_updates", Positive);
  }
};

/** gpu: chains -- OPTIONAL section that only this implementation reads: the number of
 * independent chains run side by side on the device (the role MPI ranks play in the reference).
 * Order of precedence: this key, the environment variable MLMCPI_CHAINS, 256. */
class DeviceParameters : public Parameters {
public:
  DeviceParameters() : Parameters("gpu") {}
  MLMCPI_PARAM_GETTER(unsigned int, chains)

protected:
  void parse() {
    const char *env = std::getenv("MLMCPI_CHAINS");
    chains_ = getIntOr("chains", env ? std::atoi(env) : 256, Positive);
    if (chains_ < 1)
      chains_ = 1;
  }
};

} // namespace mlmcpi
#endif // MLMCPI_PARAMETERS_HH
