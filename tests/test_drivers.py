"""driver_qm / driver_qft (examples/): the reference's drivers on the device library.
CPU: they compile, parse the reference's parameter-file format (sections, comments, typed
values, constraints) and fail loudly without a GPU.  GPU: single-, two- and multilevel runs
reproduce the analytic expectation values within the reported statistical error."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "mlmcpathintegral_b200")

QFT = """
general:
  method = '{method}'
quantumfieldtheory:
  action = '{action}'
lattice:
  Mt_lat = {M}      # trailing comment
  Mx_lat = {M}
  coarsening = '{coarsening}'
statistics:
  n_autocorr_window = 20
  n_min_samples_qoi = 200
schwinger:
  beta = {beta}
  renormalisation = '{renorm}'
gff:
  mass = 10.0
  renormalisation = 'none'
singlelevelmc:
  n_burnin = 50
  n_samples = {n_samples}
  epsilon = 1.0E-2
  sampler = '{sampler}'
hierarchical:
  n_max_level = {n_max_level}
  coarsesampler = '{coarsesampler}'
twolevelmc:
  n_burnin = 20
  n_samples = {n_samples}
  n_coarse_autocorr_window = 10
  n_fine_autocorr_window = 10
  n_delta_autocorr_window = 10
  sampler = '{coarsesampler}'
multilevelmc:
  n_level = {n_max_level}
  n_burnin = 50
  epsilon = {epsilon}
  show_detailed_stats = true
  sampler = 'hierarchical'
hmc:
  nt = 20
  dt = 0.10
  n_burnin = 20
  n_rep = 1
heatbath:
  n_sweep_overrelax = 2
  n_sweep_heatbath = 1
  random_order = true
  n_burnin = 20
clusteralgorithm:
  n_burnin = 20
  n_updates = 10
"""

QM = """
general:
  method = '{method}'
quantummechanics:
  action = '{action}'
lattice:
  M_lat = {M}
  T_final = 4.0
statistics:
  n_autocorr_window = 20
  n_min_samples_qoi = 100
harmonicoscillator:
  m0 = 1.0
  mu2 = 1.0
  renormalisation = 'perturbative'
quarticoscillator:
  m0 = 1.0
  mu2 = 1.0
  lambda = 1.0
  x0 = 1.0
rotor:
  m0 = 0.25
  renormalisation = 'perturbative'
singlelevelmc:
  n_burnin = 100
  n_samples = {n_samples}
  epsilon = 1.0E-2
  sampler = '{sampler}'
twolevelmc:
  n_burnin = 20
  n_samples = {n_samples}
  sampler = 'heatbath'
multilevelmc:
  n_level = 3
  n_burnin = 100
  epsilon = {epsilon}
  show_detailed_stats = true
  sampler = 'hierarchical'
hierarchical:
  n_max_level = 3
  coarsesampler = '{coarsesampler}'
hmc:
  nt = 100
  dt = 0.10
  n_burnin = 20
  n_rep = 1
heatbath:
  n_sweep_overrelax = 2
  n_sweep_heatbath = 1
  random_order = true
  n_burnin = 20
clusteralgorithm:
  n_burnin = 20
  n_updates = 10
"""

QFT_DEFAULTS = dict(method="singlelevel", action="quenchedschwinger", M=16, coarsening="both", beta=4.0,
                    renorm="perturbative", n_samples=100000, sampler="hierarchical", n_max_level=2,
                    coarsesampler="HMC", epsilon=0.05)
QM_DEFAULTS = dict(method="singlelevel", action="harmonicoscillator", M=32, n_samples=200000, sampler="HMC",
                   coarsesampler="HMC", epsilon=0.01)


@pytest.fixture(scope="module")
def drivers(tmp_path_factory):
    out = tmp_path_factory.mktemp("drivers")
    exes = {}
    for name in ("driver_qm", "driver_qft"):
        exe = str(out / name)
        subprocess.run(["g++", "-std=c++17", "-O2", "-w", f"-I{ROOT}/include", f"{ROOT}/examples/{name}.cc",
                        f"-L{LIBDIR}", "-lmlmcpi", "-lmlmcpi_comm", f"-Wl,-rpath,{LIBDIR}", "-o", exe], check=True)
        exes[name] = exe
    return exes


def run(exe, text, tmp_path, chains=64, timeout=900):
    p = tmp_path / "parameters.in"
    p.write_text(text)
    return subprocess.run([exe, str(p), str(chains)], capture_output=True, text=True, timeout=timeout)


def sigma_ratio(out):
    m = re.search(r"\(analytical - numerical\) = ([0-9.eE+-]+) = ([0-9.eE+-]+) \* \(statistical error\)", out)
    assert m, out
    return float(m.group(2))


def test_drivers_parse_reference_format_and_need_a_gpu(drivers, tmp_path):
    import torch
    r = run(drivers["driver_qft"], QFT.format(**QFT_DEFAULTS), tmp_path)
    assert "Mt_lat = 16" in r.stdout and "coarsening = 'both'" in r.stdout and "beta = 4.0" in r.stdout
    assert "E[V*chi_t]" in r.stdout  # analytic result printed before any device work
    if not torch.cuda.is_available():
        assert r.returncode != 0 and "no CPU fallback" in r.stderr
    # the example files shipped with the repo parse as well
    for exe, f in (("driver_qft", "parameters_qft_schwinger.in"), ("driver_qm", "parameters_qm_rotor.in")):
        r = subprocess.run([drivers[exe], os.path.join(ROOT, "examples", f)], capture_output=True, text=True)
        assert "ERROR: parameter" not in r.stderr and "cannot parse" not in r.stderr, r.stderr
        assert "chains = " in r.stdout
    # error convention of Parameters::readFile: message + non-zero exit
    bad = QFT.format(**QFT_DEFAULTS).replace("Mx_lat = 16", "Mx_lat = -3")
    r = run(drivers["driver_qft"], bad, tmp_path)
    assert r.returncode == 1 and "Mx_lat" in r.stderr
    bad = QFT.format(**QFT_DEFAULTS).replace("  dt = 0.10\n", "")
    r = run(drivers["driver_qft"], bad, tmp_path)
    assert r.returncode == 1 and "'dt'" in r.stderr and "missing" in r.stderr
    bad = QFT.format(**dict(QFT_DEFAULTS, coarsening="diagonal"))
    r = run(drivers["driver_qft"], bad, tmp_path)
    assert r.returncode == 1 and "allowed values" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("over", [
    dict(),                                                       # hierarchical, HMC coarse sampler
    dict(sampler="heatbath", n_samples=60000),
    dict(sampler="cluster", n_samples=60000),
    dict(sampler="hierarchical", coarsesampler="heatbath", renorm="nonperturbative", beta=6.0),
    dict(action="gff", coarsening="rotate", sampler="heatbath", n_samples=60000),
    dict(method="multilevel", n_max_level=2, epsilon=0.05),
    dict(action="gff", coarsening="rotate", sampler="hierarchical", coarsesampler="exact", n_max_level=2,
         n_samples=60000),
], ids=["hier-hmc", "heatbath", "cluster", "hier-hb-nonpert", "gff", "mlmc", "gff-hier-exact"])
def test_driver_qft_matches_analytic(drivers, tmp_path, over):
    r = run(drivers["driver_qft"], QFT.format(**dict(QFT_DEFAULTS, **over)), tmp_path)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr
    assert sigma_ratio(r.stdout) < 5.0, r.stdout[-1500:]


@pytest.mark.gpu
def test_driver_qft_twolevel(drivers, tmp_path):
    r = run(drivers["driver_qft"], QFT.format(**dict(QFT_DEFAULTS, method="twolevel", n_samples=20000)), tmp_path)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr
    m = re.search(r"delta QoI: Avg \+/- Err = ([-0-9.]+) \+/- ([0-9.]+)", r.stdout)
    f = re.search(r"QoI\[fine\]: Var \+/- Err = ([-0-9.]+)", r.stdout)
    d = re.search(r"delta QoI: Var \+/- Err = ([-0-9.]+)", r.stdout)
    assert m and f and d, r.stdout
    assert "Two level sampler statistics" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("over", [
    dict(),                                                                 # config C1: HO, HMC
    dict(action="rotor", sampler="hierarchical", n_samples=400000),
    dict(action="rotor", sampler="cluster", n_samples=100000),
    dict(action="rotor", method="multilevel", epsilon=2.0e-3),
    dict(action="harmonicoscillator", sampler="multilevel", n_samples=100000),
    dict(action="harmonicoscillator", sampler="hierarchical", coarsesampler="exact", n_samples=100000),
], ids=["ho-hmc", "rotor-hier", "rotor-cluster", "rotor-mlmc", "ho-multilevelsampler", "ho-hier-exact"])
def test_driver_qm_matches_analytic(drivers, tmp_path, over):
    r = run(drivers["driver_qm"], QM.format(**dict(QM_DEFAULTS, **over)), tmp_path)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr
    assert sigma_ratio(r.stdout) < 5.0, r.stdout[-1500:]


@pytest.mark.gpu
def test_driver_two_processes_two_gpus(drivers, tmp_path):
    """one process per GPU (examples/run_multi_gpu.sh): chains sharded over the processes, packed
    Statistics moments all-reduced with NCCL; twice the chains, the same analytic value"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    p = tmp_path / "parameters.in"
    p.write_text(QFT.format(**dict(QFT_DEFAULTS, n_samples=200000)))
    r = subprocess.run([os.path.join(ROOT, "examples", "run_multi_gpu.sh"), "2", drivers["driver_qft"], str(p), "64"],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr
    assert r.stdout.count("Single level MC") == 1  # only the master prints
    assert "on 2 x 64 chains" in r.stdout
    m = re.search(r"Q: # samples   = ([0-9]+)", r.stdout)
    assert m and int(m.group(1)) >= 200000
    assert sigma_ratio(r.stdout) < 5.0, r.stdout[-1500:]


@pytest.mark.gpu
def test_driver_multilevel_two_processes_two_gpus(drivers, tmp_path):
    """the multilevel method over two processes / GPUs (the reference refuses more than one rank,
    driver_qft.cc:409-414): chains sharded, every Statistics query of the allocation loop all-reduced
    (mlmcpi_comm_attach), both processes walk through the same loop"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    p = tmp_path / "parameters.in"
    p.write_text(QFT.format(**dict(QFT_DEFAULTS, method="multilevel", n_max_level=2, epsilon=0.05)))
    r = subprocess.run([os.path.join(ROOT, "examples", "run_multi_gpu.sh"), "2", drivers["driver_qft"], str(p), "64"],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr
    assert r.stdout.count("tolerance epsilon") == 1  # only the master prints
    assert sigma_ratio(r.stdout) < 5.0, r.stdout[-1500:]
