#!/usr/bin/env python
"""Record statistical-parity fixtures from the REFERENCE'S OWN DRIVERS.

oracle/_ref/driver_qft and oracle/_ref/driver_qm are the reference's src/driver_qft.cc and
src/driver_qm.cc compiled byte for byte (oracle/Makefile; Eigen/GSL header shims only).  This
script writes parameter files derived from the reference's own templates
(/root/reference/parameters_q{ft,m}_template.in), runs the stock drivers on this container's CPU
and stores what they print -- estimator, error, variance, tau_int (common/statistics.cc:82-90),
sample count, per-level acceptance -- in tests/golden/stats.json.  tests/test_gpu_parity.py runs the
same samplers through the CUDA library and compares within the combined statistical errors
(north_star: "sampled observables match the reference within combined statistical error bars at the
same integrated autocorrelation").

    python tools/make_golden_stats.py [--only NAME] [--jobs N]

/root/reference is only needed HERE; the fixtures travel.
"""
import argparse
import concurrent.futures as cf
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "stats.json")

# every case: driver, {section: {key: value}} overrides of the template
CASES = {
    # --- quenched Schwinger, hierarchical sampler, cluster coarse sampler (the ergodic choice)
    "schwinger16_b4_hier2_cluster": dict(driver="qft", set={
        "quantumfieldtheory": {"action": "'quenchedschwinger'"},
        "lattice": {"Mt_lat": 16, "Mx_lat": 16},
        "schwinger": {"beta": 4.0, "renormalisation": "'none'"},
        "singlelevelmc": {"n_burnin": 1000, "n_samples": 400000, "sampler": "'hierarchical'"},
        "hierarchical": {"n_max_level": 2, "coarsesampler": "'cluster'"},
        "clusteralgorithm": {"n_updates": 10}}),
    "schwinger32_b16_hier2_cluster": dict(driver="qft", set={
        "quantumfieldtheory": {"action": "'quenchedschwinger'"},
        "lattice": {"Mt_lat": 32, "Mx_lat": 32},
        "schwinger": {"beta": 16.0, "renormalisation": "'perturbative'"},
        "singlelevelmc": {"n_burnin": 1000, "n_samples": 100000, "sampler": "'hierarchical'"},
        "hierarchical": {"n_max_level": 2, "coarsesampler": "'cluster'"},
        "clusteralgorithm": {"n_updates": 10}}),
    "schwinger64_b64_hier2_cluster": dict(driver="qft", set={
        "quantumfieldtheory": {"action": "'quenchedschwinger'"},
        "lattice": {"Mt_lat": 64, "Mx_lat": 64},
        "schwinger": {"beta": 64.0, "renormalisation": "'perturbative'"},
        "singlelevelmc": {"n_burnin": 1000, "n_samples": 40000, "sampler": "'hierarchical'"},
        "hierarchical": {"n_max_level": 2, "coarsesampler": "'cluster'"},
        "clusteralgorithm": {"n_updates": 10}}),
    # --- the same with the HMC coarse sampler (BASELINE configs[3] at a size the CPU can run)
    "schwinger16_b4_hier2_hmc": dict(driver="qft", set={
        "quantumfieldtheory": {"action": "'quenchedschwinger'"},
        "lattice": {"Mt_lat": 16, "Mx_lat": 16},
        "schwinger": {"beta": 4.0, "renormalisation": "'perturbative'"},
        "singlelevelmc": {"n_burnin": 1000, "n_samples": 200000, "sampler": "'hierarchical'"},
        "hierarchical": {"n_max_level": 2, "coarsesampler": "'HMC'"},
        "hmc": {"nt": 20, "dt": 0.1}}),
    # --- single-level samplers
    "schwinger16_b4_cluster": dict(driver="qft", set={
        "quantumfieldtheory": {"action": "'quenchedschwinger'"},
        "lattice": {"Mt_lat": 16, "Mx_lat": 16},
        "schwinger": {"beta": 4.0},
        "singlelevelmc": {"n_burnin": 1000, "n_samples": 200000, "sampler": "'cluster'"},
        "clusteralgorithm": {"n_updates": 10}}),
    "schwinger16_b4_heatbath": dict(driver="qft", set={
        "quantumfieldtheory": {"action": "'quenchedschwinger'"},
        "lattice": {"Mt_lat": 16, "Mx_lat": 16},
        "schwinger": {"beta": 4.0},
        "singlelevelmc": {"n_burnin": 1000, "n_samples": 200000, "sampler": "'heatbath'"},
        "heatbath": {"n_sweep_overrelax": 10, "n_sweep_heatbath": 1, "random_order": "false"}}),
    # --- Gaussian free field: hierarchical sampler, heat-bath coarse sampler (coarsening rotate)
    "gff16_hier2_heatbath": dict(driver="qft", set={
        "quantumfieldtheory": {"action": "'gff'"},
        "lattice": {"Mt_lat": 16, "Mx_lat": 16, "coarsening": "'rotate'"},
        "gff": {"mass": 10.0},
        "singlelevelmc": {"n_burnin": 1000, "n_samples": 200000, "sampler": "'hierarchical'"},
        "hierarchical": {"n_max_level": 2, "coarsesampler": "'heatbath'"},
        "heatbath": {"n_sweep_overrelax": 10, "n_sweep_heatbath": 1, "random_order": "false"}}),
    "gff16_hier3_heatbath": dict(driver="qft", set={
        "quantumfieldtheory": {"action": "'gff'"},
        "lattice": {"Mt_lat": 16, "Mx_lat": 16, "coarsening": "'rotate'"},
        "gff": {"mass": 10.0},
        "singlelevelmc": {"n_burnin": 1000, "n_samples": 200000, "sampler": "'hierarchical'"},
        "hierarchical": {"n_max_level": 3, "coarsesampler": "'heatbath'"},
        "heatbath": {"n_sweep_overrelax": 10, "n_sweep_heatbath": 1, "random_order": "false"}}),
    "gff32_hier4_heatbath": dict(driver="qft", set={
        "quantumfieldtheory": {"action": "'gff'"},
        "lattice": {"Mt_lat": 32, "Mx_lat": 32, "coarsening": "'rotate'"},
        "gff": {"mass": 10.0},
        "singlelevelmc": {"n_burnin": 1000, "n_samples": 50000, "sampler": "'hierarchical'"},
        "hierarchical": {"n_max_level": 4, "coarsesampler": "'heatbath'"},
        "heatbath": {"n_sweep_overrelax": 10, "n_sweep_heatbath": 1, "random_order": "false"}}),
    # --- the same hierarchies with the HMC coarse sampler: HMC integrates the 5-point force but accepts with
    #     Action::evaluate = Q_hat (hmcsampler.cc:21-69), so it samples Q_hat exactly: a CONSISTENT
    #     algorithm, whose estimator does not depend on sweep orders and must agree within errors.
    #     (coarsesampler = 'exact' cannot serve: GFFAction::draw never sets the `accept` flag of MCMCStep,
    #     so HierarchicalSampler::draw breaks out at hierarchicalsampler.cc:73 on every draw and the
    #     reference driver prints avg = 0, p = nan.)
    "gff16_hier2_hmc": dict(driver="qft", set={
        "quantumfieldtheory": {"action": "'gff'"},
        "lattice": {"Mt_lat": 16, "Mx_lat": 16, "coarsening": "'rotate'"},
        "gff": {"mass": 10.0},
        "singlelevelmc": {"n_burnin": 1000, "n_samples": 200000, "sampler": "'hierarchical'"},
        "hierarchical": {"n_max_level": 2, "coarsesampler": "'HMC'"},
        "hmc": {"nt": 20, "dt": 0.2}}),
    "gff16_hier3_hmc": dict(driver="qft", set={
        "quantumfieldtheory": {"action": "'gff'"},
        "lattice": {"Mt_lat": 16, "Mx_lat": 16, "coarsening": "'rotate'"},
        "gff": {"mass": 10.0},
        "singlelevelmc": {"n_burnin": 1000, "n_samples": 400000, "sampler": "'hierarchical'"},
        "hierarchical": {"n_max_level": 3, "coarsesampler": "'HMC'"},
        "hmc": {"nt": 20, "dt": 0.2}}),
    # --- topological rotor (driver_qm): hierarchical sampler with HMC and with cluster coarse sampler
    "rotor32_hier3_hmc": dict(driver="qm", set={
        "quantummechanics": {"action": "'rotor'"},
        "lattice": {"M_lat": 32, "T_final": 4.0},
        "rotor": {"m0": 0.25, "renormalisation": "'perturbative'"},
        "singlelevelmc": {"n_burnin": 1000, "n_samples": 400000, "sampler": "'hierarchical'"},
        "hierarchical": {"n_max_level": 3, "coarsesampler": "'HMC'"},
        "hmc": {"nt": 20, "dt": 0.1}}),
    "rotor64_cluster": dict(driver="qm", set={
        "quantummechanics": {"action": "'rotor'"},
        "lattice": {"M_lat": 64, "T_final": 4.0},
        "rotor": {"m0": 0.25},
        "singlelevelmc": {"n_burnin": 1000, "n_samples": 400000, "sampler": "'cluster'"},
        "clusteralgorithm": {"n_updates": 1}}),
}


def render(template, overrides):
    """apply {section: {key: value}} to the text of a reference parameter template; keys the template
    lacks (parameters_qm_template.in has no twolevelmc autocorrelation windows) are
    added to their section"""
    lines, section, seen = [], None, set()
    for line in template.splitlines():
        m = re.match(r"^([A-Za-z0-9_]+):\s*$", line)
        if m:
            section = m.group(1)
        else:
            m2 = re.match(r"^(\s+)([A-Za-z0-9_]+)\s*=\s*([^#]*)(#.*)?$", line)
            if m2 and section in overrides and m2.group(2) in overrides[section]:
                line = f"{m2.group(1)}{m2.group(2)} = {overrides[section][m2.group(2)]}"
                seen.add((section, m2.group(2)))
        lines.append(line)
    out = []
    for line in lines:
        out.append(line)
        m = re.match(r"^([A-Za-z0-9_]+):\s*$", line)
        if m and m.group(1) in overrides:
            for k, v in overrides[m.group(1)].items():
                if (m.group(1), k) not in seen:
                    out.append(f"  {k} = {v}")
                    seen.add((m.group(1), k))
    missing = [(s, k) for s, kv in overrides.items() for k in kv if (s, k) not in seen]
    if missing:
        raise KeyError(f"sections not in the template: {missing}")
    return "\n".join(out) + "\n"


def parse_output(text):
    r = {}
    m = re.search(r"Q: Avg \+/- Err = (\S+) \+/- (\S+)", text)
    r["average"], r["error"] = float(m.group(1)), float(m.group(2))
    m = re.search(r"Q: Var \+/- Err = (\S+) \+/- (\S+)", text)
    r["variance"], r["variance_error"] = float(m.group(1)), float(m.group(2))
    r["tau_int"] = float(re.search(r"Q: tau_\{int\}\s+= (\S+)", text).group(1))
    r["window"] = int(re.search(r"Q: window\s+= (\S+)", text).group(1))
    r["samples"] = int(re.search(r"Q: # samples\s+= (\S+)", text).group(1))
    r["acceptance"] = [float(x) for x in re.findall(r"level \d+ .*: +p = (\S+)", text)]
    m = re.search(r"acceptance rate = (\S+)", text)
    if m:
        r["acceptance_rate"] = float(m.group(1))
    m = re.search(r"E\[V\*chi_t\]\s+= (\S+)", text) or re.search(r"E\[Q\^2\]\s+= (\S+)", text)
    if m:
        r["analytical"] = float(m.group(1))
    m = re.search(r"Tuned\s+dt_\{HMC\} = (\S+)", text)
    if m:
        r["hmc_dt_tuned"] = float(m.group(1))
    if "FAILED to tune" in text:
        r["hmc_dt_tuned"] = None
    m = re.search(r"cost per sample = (\S+) mu s", text)
    if m:
        r["cost_per_sample_usec"] = float(m.group(1))
    return r


def run_case(name):
    c = CASES[name]
    template = open(os.path.join(REF, f"parameters_{c['driver']}_template.in")).read()
    over = {k: dict(v) for k, v in c["set"].items()}
    if c["driver"] == "qm":
        tl = over.setdefault("twolevelmc", {})
        for k in ("n_coarse_autocorr_window", "n_fine_autocorr_window", "n_delta_autocorr_window"):
            tl.setdefault(k, 10)
    text = render(template, over)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "parameters.in")
        open(p, "w").write(text)
        t0 = time.time()
        out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", f"driver_{c['driver']}"), p],
                             capture_output=True, text=True, cwd=d)
        wall = time.time() - t0
    if out.returncode != 0:
        raise RuntimeError(f"{name}: driver failed\n{out.stdout[-2000:]}\n{out.stderr[-2000:]}")
    r = parse_output(out.stdout)
    r.update(driver=f"driver_{c['driver']}", overrides=c["set"], wall_s=round(wall, 1), parameters=text)
    return name, r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--jobs", type=int, default=max(1, len(os.sched_getaffinity(0)) - 1))
    a = ap.parse_args()
    names = [a.only] if a.only else list(CASES)
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    with cf.ThreadPoolExecutor(a.jobs) as ex:
        for name, r in ex.map(run_case, names):
            res[name] = r
            print(name, {k: r[k] for k in ("average", "error", "tau_int", "acceptance", "wall_s")}, flush=True)
            json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
