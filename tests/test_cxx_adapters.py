"""The C++ adapter layer (include/mlmcpi/adapters.hh: the reference's class names over the
C-ABI) compiles and links against libmlmcpi.so; on a GPU box the example driver -- written the
way the reference's driver_qft.cc / MonteCarloSingleLevel use the reference classes -- runs."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "mlmcpathintegral_b200")


def build_example(tmp_path):
    exe = str(tmp_path / "driver_qft_schwinger")
    cmd = ["g++", "-std=c++17", "-O2", "-w", f"-I{ROOT}/include", f"{ROOT}/examples/driver_qft_schwinger.cc",
           f"-L{LIBDIR}", "-lmlmcpi", f"-Wl,-rpath,{LIBDIR}", "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


def test_adapters_compile_and_link(tmp_path):
    exe = build_example(tmp_path)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in r.stderr  # fails loudly, never falls back


@pytest.mark.gpu
def test_example_driver_runs(tmp_path):
    exe = build_example(tmp_path)
    r = subprocess.run([exe, "16", "4.0", "3000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    m = re.search(r"sum\(force\) = ([-+0-9.eE]+)", out)
    assert m and abs(float(m.group(1))) < 1e-9
    m = re.search(r"Avg \+/- Err = ([0-9.]+) \+/- ([0-9.]+)", out)
    assert m, out
    avg, err = float(m.group(1)), float(m.group(2))
    # quenchedschwinger_chit_analytical(beta = 4, P = 256) recorded from the reference? not in the
    # fixtures: check the physically required range and a sane error bar instead
    assert 0.0 < avg < 10.0 and 0.0 < err < 0.5 * max(avg, 0.05), out
    assert "level 0 [finest]" in out
