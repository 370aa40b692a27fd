"""CPU-side checks of the C-ABI library: it loads, exports every symbol that
include/mlmcpi.h declares, refuses to work without a GPU (no CPU fallback), and its
host-side integer geometry / renormalisation / statistics-finalisation agree with
the golden vectors recorded from the reference (bit-exact for integers)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.util import load, qm_model, scalar, unhex

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mp():
    import mlmcpathintegral_b200 as mp
    return mp


def test_library_exports_every_declared_symbol(mp):
    hdr = open(os.path.join(ROOT, "include", "mlmcpi.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mlmcpi_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) > 40
    lib = C.CDLL(mp._lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(mp._lib.SIGNATURES), declared ^ set(mp._lib.SIGNATURES)
    assert lib.mlmcpi_version() == 100


def test_python_constants_match_the_header(mp):
    """every enumerator of include/mlmcpi.h that the ctypes layer names (options, streams, samplers, QoIs, models,
    coarsenings) has the header's value"""
    hdr = open(os.path.join(ROOT, "include", "mlmcpi.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    enums = dict((k, int(v)) for k, v in re.findall(r"\b(MLMCPI_[A-Z0-9_]+)\s*=\s*(-?\d+)", hdr))
    assert len(enums) > 40
    checked = 0
    for name, value in vars(mp._lib).items():
        if not name.isupper() or not isinstance(value, int):
            continue
        for key in ("MLMCPI_" + name, "MLMCPI_MODEL_" + name, "MLMCPI_E" + name[2:] if name.startswith("E_") else None):
            if key and key in enums:
                assert enums[key] == value, (name, value, key, enums[key])
                checked += 1
                break
    assert checked >= 30, checked
    for opt in [k for k in enums if k.startswith("MLMCPI_OPT_")]:
        assert hasattr(mp._lib, opt[len("MLMCPI_"):]), opt


def test_every_entry_point_sets_its_device():
    """every C-ABI entry point that takes a context, sampler, statistics or multilevel object starts with a
    DeviceGuard: the library restores the caller's current device on return, so an entry point without one runs
    on the wrong device in a process whose context is not on device 0 (found by the two-process driver run:
    Statistics::hard_reset on rank 1 failed with 'invalid argument' and rank 0 waited in the all-reduce)"""
    src = open(os.path.join(ROOT, "mlmcpathintegral_b200", "csrc", "capi.cu")).read()
    missing = []
    pat = r"^(?:int|void|uint64_t|const char \*|double|void \*|size_t)\s+(mlmcpi_[a-z0-9_]+)\s*\(([^)]*)\)\s*\{"
    for m in re.finditer(pat, src, re.M):
        name, args = m.group(1), m.group(2)
        if not any(t in args for t in ("mlmcpi_ctx", "mlmcpi_stats", "mlmcpi_mlmc", "mlmcpi_sampler")):
            continue
        body = src[m.end():m.end() + 600].split("\n}\n")[0]
        if "DeviceGuard" not in body and name not in ("mlmcpi_last_error", "mlmcpi_device", "mlmcpi_stream",
                                                      "mlmcpi_world_size", "mlmcpi_launch_count"):
            missing.append(name)
    assert not missing, missing


def test_no_cpu_fallback(mp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert mp._lib.lib.mlmcpi_create(C.byref(h), 0, 1, None) == mp._lib.E_CUDA
    with pytest.raises(mp.MlmcpiError):
        mp.Context()


def test_oracle_is_not_reachable_from_the_product():
    """the product package must not import, link or dlopen anything under oracle/"""
    pkg = os.path.join(ROOT, "mlmcpathintegral_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".hh", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("no cpu fallback", ""), os.path.join(dirpath, f)


def test_lattice2d_maps_bit_exact(mp):
    L = mp._lib.lib
    for c in load("lattice")["lattice2d"]:
        Mt, Mx, rot, lo = c["Mt"], c["Mx"], c["rotated"], c["lo"]
        k = 0
        for i in range(lo, Mt + 3):
            for j in range(lo, Mx + 3):
                want = c["vertex_cart2lin"][k]
                k += 1
                if want != 0xFFFFFFFF:
                    assert L.mlmcpi_vertex_cart2lin(Mt, Mx, rot, i, j) == want
        ii, jj, mm = C.c_int(), C.c_int(), C.c_int()
        nb = (C.c_uint32 * 8)()
        for ell in range(c["n_vertices"]):
            L.mlmcpi_vertex_lin2cart(Mt, Mx, rot, ell, C.byref(ii), C.byref(jj))
            assert [ii.value, jj.value] == c["vertex_lin2cart"][2 * ell:2 * ell + 2]
            L.mlmcpi_neighbours(Mt, Mx, rot, ell, nb)
            assert list(nb) == c["neighbours"][8 * ell:8 * ell + 8]
        if "link_cart2lin" in c:
            k = 0
            for i in range(lo, Mt + 3):
                for j in range(lo, Mx + 3):
                    for mu in range(2):
                        assert L.mlmcpi_link_cart2lin(Mt, Mx, i, j, mu) == c["link_cart2lin"][k]
                        k += 1
            for ell in range(c["n_edges"]):
                L.mlmcpi_link_lin2cart(Mt, Mx, ell, C.byref(ii), C.byref(jj), C.byref(mm))
                assert [ii.value, jj.value, mm.value] == c["link_lin2cart"][3 * ell:3 * ell + 3]
        a, b, r = C.c_int(), C.c_int(), C.c_int()
        ok = L.mlmcpi_coarse_shape(Mt, Mx, c["ctype"], c["level"], C.byref(a), C.byref(b), C.byref(r))
        assert bool(ok) == bool(c["has_coarse"])
        if c["has_coarse"]:
            nv = c["n_vertices"]
            co, fo, mv = (C.c_uint32 * nv)(), (C.c_uint32 * nv)(), (C.c_uint32 * nv)()
            cnt = (C.c_int * 2)()
            assert L.mlmcpi_coarsening_lists(Mt, Mx, c["ctype"], c["level"], co, fo, mv, cnt) == 0
            assert list(co)[:cnt[0]] == c["coarse"]
            assert list(fo)[:cnt[1]] == c["fineonly"]
            assert list(mv)[:cnt[0]] == c["map_vals"]


def test_coarse_models(mp):
    from oracle import pyoracle as po
    for c in load("qm"):
        o = qm_model(po, c)
        m = mp.Model(model=o.model, M_lat=o.M_lat, a_lat=o.a_lat, T_final=o.T_final, m0=o.m0,
                     mu2=o.mu2, lambda_=o.lambda_, x0=o.x0)
        mc = mp.coarse_model(m, renorm=c["ip"][1])
        assert mc.m0 == scalar(c["coarse_m0"]) and mc.M_lat == o.M_lat // 2
        assert mc.a_lat == o.T_final / mc.M_lat
    for c in load("schwinger"):
        m = mp.schwinger(c["Mt"], c["Mx"], c["beta"], c["ctype"], 0)
        mc = mp.coarse_model(m, renorm=c["renorm"], level=0, ctype=c["ctype"])
        assert mc.beta == scalar(c["coarse_beta"])
        assert mp.sample_size(mc) == len(c["restrict"])
    for c in load("gff"):
        m = mp.gff(c["Mt"], c["Mx"], c["mass"], c["ctype"], 0)
        assert m.gff_mu2 == scalar(c["mu2"])
        mc = mp.coarse_model(m, level=0, ctype=c["ctype"])
        assert abs(mc.gff_mu2 - scalar(c["coarse_mu2"])) <= 4e-16 * mc.gff_mu2
        assert mp.sample_size(mc) == len(c["restrict"])


def test_error_codes(mp):
    L = mp._lib.lib
    c = mp.Model()
    odd = mp.rotor(7)
    assert L.mlmcpi_coarse_model(C.byref(odd), 0, 0, 0, 4.0, C.byref(c)) == mp._lib.E_INVAL
    sw = mp.schwinger(5, 7, 1.0)
    assert L.mlmcpi_coarse_model(C.byref(sw), 0, 0, 0, 0.0, C.byref(c)) == mp._lib.E_INVAL
    r = mp.rotor(8)
    assert L.mlmcpi_coarse_model(C.byref(r), 2, 0, 0, 4.0, C.byref(c)) == mp._lib.E_UNSUPPORTED


def test_stats_finalize_matches_reference_statistics(mp):
    """one chain: finalize(packed moments) == Statistics of the reference"""
    g = load("scalars")["statistics"]
    q, want, k_max = unhex(g["q"]), unhex(g["out"]), g["k_max"]
    # build the packed vector on the host exactly as statistics.cc:4-27 accumulates it
    S = np.zeros(k_max)
    avg = np.zeros(4)
    hist = []
    for n, Q in enumerate(q, start=1):
        hist.insert(0, Q)
        hist = hist[:k_max]
        for p in range(4):
            avg[p] = ((n - 1.0) * avg[p] + Q ** (p + 1)) / n
        for k in range(len(hist)):
            Nk = n - k
            S[k] = ((Nk - 1.0) * S[k] + hist[0] * hist[k]) / Nk
    packed = np.concatenate([[1.0, float(len(q)), float(len(q)), avg[0]], avg, S])
    out = mp.Statistics.finalize(packed, k_max)
    got = np.array([out["average"], out["variance"], out["variance_error"], out["tau_int"],
                    out["error"], out["samples"]])
    assert np.allclose(got, want, rtol=1e-13, atol=0)


def test_analytic_results_match_reference(mp):
    """host quadrature (composite Gauss-Legendre) + bisection against the values the reference
    computes with GSL-style QAWO/QAG and its bisection solver (golden, recorded from oracle/_ref)"""
    L = mp._lib.lib
    an = load("scalars")["analytic"]

    def near(got, want, what):
        # chi_t values far below 1e-12 are rounding noise in both implementations
        assert abs(got - want) <= 1e-9 * abs(want) + 1e-14, (what, got, want)

    for g in an["schwinger_chit_grid"]:
        b, P = g["beta"], g["n_plaq"]
        near(L.mlmcpi_schwinger_chit_analytical(b, P), float.fromhex(g["exact"]), ("exact", b, P))
        near(L.mlmcpi_schwinger_chit_perturbative(b, P), float.fromhex(g["perturbative"]), ("pert", b, P))
        near(L.mlmcpi_schwinger_var_chit_continuum(b, P), float.fromhex(g["var_continuum"]), ("var", b, P))
    assert np.isnan(L.mlmcpi_schwinger_chit_analytical(2500.0, 64))
    for g in an["schwinger_betacoarse_nonperturbative"]:
        m = mp.schwinger(g["Mt"], g["Mt"], g["beta"], g["ctype"], 0)
        mc = mp.coarse_model(m, renorm=mp.RENORM_NONPERTURBATIVE, ctype=g["ctype"])
        want = float.fromhex(g["beta_coarse"])
        assert abs(mc.beta - want) <= 1e-10 * want, (g, mc.beta, want)
    for w, key in enumerate(["rotor_chit_exact_32", "rotor_chit_perturbative_32", "rotor_chit_continuum_32"]):
        near(L.mlmcpi_rotor_chit(0.25, 4.0 / 32, 4.0, w), float.fromhex(an[key]), key)
    for w in range(3):
        near(L.mlmcpi_rotor_chit(0.25, 4.0 / 256, 4.0, w), float.fromhex(an["rotor_chit_256"][w]), ("rotor256", w))
    near(L.mlmcpi_gff_phi_squared_analytical(10.0, 16, 16), float.fromhex(an["gff_phi_squared_10_16"]), "gff")
    near(L.mlmcpi_ho_xsquared_analytical(1.0, 1.0, 4.0 / 32, 32, 0), float.fromhex(an["ho_x2_32"]), "ho")
    near(L.mlmcpi_ho_xsquared_analytical(1.0, 1.0, 4.0 / 32, 32, 1), float.fromhex(an["ho_x2_continuum"]), "ho cont")
    sh = load("scalars")["Sigma_hat"]
    for (xi, p), y in zip(sh["args"], sh["y"]):
        assert L.mlmcpi_sigma_hat(xi, p) == float.fromhex(y)


def test_gff_coarse_model_carries_the_gibbs_smoothed_action(mp):
    """GFFAction::coarse_action (gffaction.hh:201-208): n_gibbs_smooth = 2, omega = 1 on EVERY coarse
    level (there is no silent fall-back to the 5-point action; level 1 of BASELINE config C3 has 32768
    vertices = MLMCPI_GFF_DENSE_MAX)"""
    mc = mp.coarse_model(mp.gff(32, 32, 10.0), ctype=mp.COARSEN_ROTATE)
    assert (mc.gff_n_gibbs, mc.gff_omega, mp.sample_size(mc)) == (2, 1.0, 512)
    big = mp.coarse_model(mp.gff(256, 256, 10.0), ctype=mp.COARSEN_ROTATE)
    assert big.gff_n_gibbs == 2 and mp.sample_size(big) == 32768
    hdr = open(os.path.join(ROOT, "include", "mlmcpi.h")).read()
    assert "#define MLMCPI_GFF_DENSE_MAX 32768" in hdr


def test_comm_library_exports_every_declared_symbol(mp):
    """libmlmcpi_comm.so (NCCL all-reduce of the statistics moments) loads and exports what
    include/mlmcpi_comm.h declares; a single-process communicator needs neither NCCL traffic nor a GPU"""
    hdr = open(os.path.join(ROOT, "include", "mlmcpi_comm.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mlmcpi_comm_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) == 9, declared
    path = os.path.join(os.path.dirname(mp._lib.LIB_PATH), "libmlmcpi_comm.so")
    if not os.path.exists(path):
        pytest.skip("NCCL not installed: libmlmcpi_comm.so not built")
    lib = C.CDLL(path)
    assert not [s for s in sorted(declared) if not hasattr(lib, s)]
    lib.mlmcpi_comm_world_size.argtypes = [C.c_void_p]
    assert lib.mlmcpi_comm_world_size(None) == 1 and lib.mlmcpi_comm_rank(None) == 0
