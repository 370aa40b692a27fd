"""Pin the C restatement (oracle/mlmcpi_oracle.c) against the golden vectors that
tools/make_golden.py recorded from the REFERENCE's own translation units
(oracle/_ref), and -- where oracle/_ref is present -- against the reference
library directly on fresh random inputs.

Tolerances: integer maps exact; arithmetic that involves no Bessel function or erf
is expected bit-exact (EXACT) -- the restatement follows the reference's order of
operations; Bessel/erf paths 1e-12 relative (the reference build here uses shimmed
GSL special functions)."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import load, qm_model, rel_err, scalar, unhex

TOL = 1e-12


def assert_exact(a, b, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, what
    assert np.array_equal(a, b), f"{what}: max abs diff {np.max(np.abs(a - b))}"


# ------------------------------------------------------------------ lattice


def test_lattice2d_maps_bit_exact(orc):
    L = orc.lib
    for c in load("lattice")["lattice2d"]:
        Mt, Mx, rot, lo = c["Mt"], c["Mx"], c["rotated"], c["lo"]
        assert L.orc_n_vertices(Mt, Mx, rot) == c["n_vertices"]
        k = 0
        for i in range(lo, Mt + 3):
            for j in range(lo, Mx + 3):
                want = c["vertex_cart2lin"][k]
                k += 1
                if want == 0xFFFFFFFF:
                    continue
                assert L.orc_vertex_cart2lin(Mt, Mx, rot, i, j) == want, (c["ctype"], c["level"], i, j)
        ii, jj = C.c_int(), C.c_int()
        nb = (C.c_uint32 * 8)()
        for ell in range(c["n_vertices"]):
            L.orc_vertex_lin2cart(Mt, Mx, rot, ell, C.byref(ii), C.byref(jj))
            assert [ii.value, jj.value] == c["vertex_lin2cart"][2 * ell:2 * ell + 2]
            L.orc_neighbours(Mt, Mx, rot, ell, nb)
            assert list(nb) == c["neighbours"][8 * ell:8 * ell + 8]
        if "link_cart2lin" in c:
            k = 0
            for i in range(lo, Mt + 3):
                for j in range(lo, Mx + 3):
                    for mu in range(2):
                        assert L.orc_link_cart2lin(Mt, Mx, i, j, mu) == c["link_cart2lin"][k]
                        k += 1
            mm = C.c_int()
            for ell in range(c["n_edges"]):
                L.orc_link_lin2cart(Mt, Mx, ell, C.byref(ii), C.byref(jj), C.byref(mm))
                assert [ii.value, jj.value, mm.value] == c["link_lin2cart"][3 * ell:3 * ell + 3]
        a, b, r = C.c_int(), C.c_int(), C.c_int()
        ok = L.orc_coarse_shape(Mt, Mx, c["ctype"], c["level"], C.byref(a), C.byref(b), C.byref(r))
        assert bool(ok) == bool(c["has_coarse"])
        if c["has_coarse"]:
            nv = c["n_vertices"]
            co, fo, mv = (C.c_uint32 * nv)(), (C.c_uint32 * nv)(), (C.c_uint32 * nv)()
            cnt = (C.c_int * 2)()
            assert L.orc_coarsening_lists(Mt, Mx, c["ctype"], c["level"], co, fo, mv, cnt) == 0
            assert list(co)[:cnt[0]] == c["coarse"] == c["map_keys"]
            assert list(fo)[:cnt[1]] == c["fineonly"]
            assert list(mv)[:cnt[0]] == c["map_vals"]


def test_lattice1d(orc):
    for c in load("lattice")["lattice1d"]:
        M = c["M"]
        assert c["T"] / M == scalar(c["a_lat"])
        nb = [((l - 1 + M) % M, (l + 1) % M) for l in range(M)]
        assert [v for p in nb for v in p] == c["neighbours"]


# ------------------------------------------------------------------ scalars


def test_scalars(orc):
    L = orc.lib
    g = load("scalars")
    xs, ys = unhex(g["mod_2pi"]["x"]), unhex(g["mod_2pi"]["y"])
    assert_exact([L.orc_mod_2pi(x) for x in xs], ys, "mod_2pi")
    zs, ys = unhex(g["fast_bessel_I0_scaled"]["z"]), unhex(g["fast_bessel_I0_scaled"]["y"])
    for z, y in zip(zs, ys):
        got = L.orc_fast_bessel_I0_scaled(z)
        if z > 100:
            assert got == y  # series branches: no Bessel library involved
        else:
            assert abs(got - y) <= TOL * abs(y)


def test_bessel_against_scipy(orc):
    sp = pytest.importorskip("scipy.special")
    x = np.concatenate([np.linspace(0, 30, 301), np.linspace(30, 700, 200)])
    got = np.array([orc.lib.orc_bessel_I0(v) for v in x])
    assert np.max(np.abs(got / sp.i0(x) - 1)) < 1e-14
    x = np.concatenate([x, np.linspace(700, 5000, 50)])
    got = np.array([orc.lib.orc_bessel_I0_scaled(v) for v in x])
    assert np.max(np.abs(got / sp.i0e(x) - 1)) < 1e-14


def test_distribution_pdfs(orc):
    L = orc.lib
    g = load("scalars")
    for p in g["dist_pdf"]:
        want = float.fromhex(p["y"])
        d, prm, x, xp, xm = p["dist"], p["param"], p["x"], p["x_p"], p["x_m"]
        got = [lambda: L.orc_expsin2_pdf(x, prm), lambda: L.orc_expcos_pdf(prm, x, xp, xm),
               lambda: L.orc_besselproduct_pdf(prm, x, xp, xm),
               lambda: L.orc_approxbessel_pdf(prm, x, xp, xm)][d]()
        assert abs(got - want) <= TOL * abs(want), p
    for z in g["Znorm_inv"]:
        alpha = (C.c_double * 17)()
        L.orc_besselproduct_alpha(z["beta"], alpha)
        got = L.orc_besselproduct_Znorm_inv(alpha, z["phi"], z["rescaled"])
        want = float.fromhex(z["y"])
        assert abs(got - want) <= TOL * abs(want)


def test_pdfs_are_normalised(orc):
    """independent check of the restated pdfs: they integrate to one"""
    L = orc.lib
    x = np.linspace(-np.pi, np.pi, 1001)
    for f in (lambda v: L.orc_expsin2_pdf(v, 7.0), lambda v: L.orc_expcos_pdf(3.0, v, 0.9, 0.2),
              lambda v: L.orc_besselproduct_pdf(4.0, v, 0.9, 0.2),
              lambda v: L.orc_approxbessel_pdf(20.0, v, 0.9, 0.2)):
        y = np.array([f(v) for v in x])
        assert abs(np.trapezoid(y, x) - 1.0) < 2e-3


def test_statistics(orc):
    g = load("scalars")["statistics"]
    got = orc.statistics(g["k_max"], unhex(g["q"]))
    assert_exact(got, unhex(g["out"]), "statistics")


# --------------------------------------------------------------- QM actions


@pytest.mark.parametrize("c", load("qm"), ids=lambda c: c["name"])
def test_qm_golden(orc, c):
    m = qm_model(po, c)
    x, p0 = unhex(c["x"]), unhex(c["p0"])
    rotor = c["kind"] == po.ROTOR
    bessel = rotor
    assert orc.action(m, x) == scalar(c["S"])
    assert_exact(orc.force(m, x), unhex(c["force"]), "force")
    for (xm, xp), w in zip([(0.3, -0.2), (-1.1, 2.5), (3.0, -3.0)], c["W"]):
        assert_exact(orc.W(m, xm, xp), unhex(w), "W")
    got = orc.cond_action(m, x)
    want = scalar(c["cond_S"])
    assert abs(got - want) <= (TOL if bessel else 0.0) * abs(want)
    assert orc.qoi(m, po.QOI_X2, x)[0] == scalar(c["qoi_x2"])
    if rotor:
        assert orc.qoi(m, po.QOI_ROTOR_CHI, x)[0] == scalar(c["qoi_chi"])
        assert_exact(orc.overrelax_sweep(m, x, coloured=False), unhex(c["overrelax_lex"]), "OR lex")
        assert_exact(orc.overrelax_sweep(m, x, coloured=True), unhex(c["overrelax_coloured"]), "OR col")
    lf = c["leapfrog"]
    xl, pl = orc.leapfrog(m, lf["nt"], lf["dt"], x, p0)
    assert_exact(xl, unhex(lf["x"]), "leapfrog x")
    assert_exact(pl, unhex(lf["p"]), "leapfrog p")
    mc = orc.coarse_model(m, renorm=c["ip"][1], T_final=c["dp"][0])
    assert mc.m0 == scalar(c["coarse_m0"])
    xc = orc.restrict(m, mc, x)
    assert_exact(xc, unhex(c["restrict"]), "restrict")
    n = len(x)
    from tools.make_golden import noncompact
    assert_exact(orc.prolong(m, noncompact(n // 2, 0.3), x), unhex(c["prolong"]), "prolong")
    assert orc.action(mc, xc) == scalar(c["coarse_S"])
    tl = c["twolevel"]
    tp, pc = unhex(tl["theta_prime"]), unhex(tl["phi_coarse"])
    d = [orc.action(m, tp) - orc.action(m, x),
         orc.action(mc, orc.restrict(m, mc, x)) - orc.action(mc, pc),
         orc.cond_action(m, x) - orc.cond_action(m, tp)]
    want = unhex(tl["deltas"])
    assert d[0] == want[0] and d[1] == want[1]
    assert abs(d[2] - want[2]) <= (TOL * max(abs(orc.cond_action(m, x)), 1.0) if bessel else 0.0)


# ---------------------------------------------------------------- Schwinger


@pytest.mark.parametrize("c", load("schwinger"), ids=lambda c: c["name"])
def test_schwinger_golden(orc, c):
    from tools.make_golden import angles
    m = po.schwinger(c["Mt"], c["Mx"], c["beta"], c["ctype"], 0)
    x, p0 = unhex(c["x"]), unhex(c["p0"])
    assert orc.action(m, x) == scalar(c["S"])
    assert_exact(orc.force(m, x), unhex(c["force"]), "force")
    chi, Q = orc.qoi(m, po.QOI_SCHWINGER_CHI, x)
    assert chi == scalar(c["qoi_chi"])
    assert Q == int(round(np.sqrt(scalar(c["qoi_chi"]) * 4 * np.pi ** 2) / (2 * np.pi))) or Q < 0
    assert orc.qoi(m, po.QOI_AVG_PLAQUETTE, x)[0] == scalar(c["qoi_plaq"])
    assert_exact(orc.overrelax_sweep(m, x, coloured=False), unhex(c["overrelax_lex"]), "OR lex")
    assert_exact(orc.overrelax_sweep(m, x, coloured=True), unhex(c["overrelax_coloured"]), "OR col")
    got, want = orc.cond_action(m, x), scalar(c["cond_S"])
    assert abs(got - want) <= TOL * abs(want)
    lf = c["leapfrog"]
    xl, pl = orc.leapfrog(m, lf["nt"], lf["dt"], x, p0)
    assert_exact(xl, unhex(lf["x"]), "leapfrog x")
    assert_exact(pl, unhex(lf["p"]), "leapfrog p")
    mc = orc.coarse_model(m, renorm=c["renorm"], level=0, ctype=c["ctype"])
    assert mc.beta == scalar(c["coarse_beta"])
    xc = orc.restrict(m, mc, x)
    assert_exact(xc, unhex(c["restrict"]), "restrict")
    assert orc.action(mc, xc) == scalar(c["coarse_S"])
    assert_exact(orc.prolong(m, angles(len(xc), 0.3), x), unhex(c["prolong"]), "prolong")
    tl = c["twolevel"]
    tp, pc = unhex(tl["theta_prime"]), unhex(tl["phi_coarse"])
    want = unhex(tl["deltas"])
    assert orc.action(m, tp) - orc.action(m, x) == want[0]
    assert orc.action(mc, xc) - orc.action(mc, pc) == want[1]
    d2 = orc.cond_action(m, x) - orc.cond_action(m, tp)
    assert abs(d2 - want[2]) <= TOL * max(abs(orc.cond_action(m, x)), 1.0)


# ---------------------------------------------------------------------- GFF


@pytest.mark.parametrize("c", load("gff"), ids=lambda c: c["name"])
def test_gff_golden(orc, c):
    from tools.make_golden import noncompact
    m = po.gff(c["Mt"], c["Mx"], c["mass"], c["ctype"], 0)
    assert m.gff_mu2 == scalar(c["mu2"])
    x, p0 = unhex(c["x"]), unhex(c["p0"])
    assert orc.action(m, x) == scalar(c["S"])
    assert_exact(orc.force(m, x), unhex(c["force"]), "force")
    assert orc.qoi(m, po.QOI_PHI2, x)[0] == scalar(c["qoi_phi2"])
    assert_exact(orc.overrelax_sweep(m, x, coloured=False), unhex(c["overrelax_lex"]), "OR lex")
    assert orc.cond_action(m, x) == scalar(c["cond_S"])
    lf = c["leapfrog"]
    xl, pl = orc.leapfrog(m, lf["nt"], lf["dt"], x, p0)
    assert_exact(xl, unhex(lf["x"]), "leapfrog x")
    assert_exact(pl, unhex(lf["p"]), "leapfrog p")
    mc = orc.coarse_model(m, level=0, ctype=c["ctype"])
    assert abs(mc.gff_mu2 - scalar(c["coarse_mu2"])) <= 4e-16 * mc.gff_mu2
    xc = orc.restrict(m, mc, x)
    assert_exact(xc, unhex(c["restrict"]), "restrict")
    assert_exact(orc.prolong(m, noncompact(len(xc), 0.3), x), unhex(c["prolong"]), "prolong")
    if "level1" in c:
        l1 = c["level1"]
        mc.gff_mu2 = scalar(c["coarse_mu2"])
        x1 = unhex(l1["x"])
        assert_exact(orc.force(mc, x1), unhex(l1["force"]), "level1 force")
        assert orc.cond_action(mc, x1) == scalar(l1["cond_S"])
        assert_exact(orc.overrelax_sweep(mc, x1, coloured=False), unhex(l1["overrelax_lex"]), "l1 OR")
        mcc = orc.coarse_model(mc, level=1, ctype=c["ctype"])
        xcc = orc.restrict(mc, mcc, x1)
        assert_exact(xcc, unhex(l1["restrict"]), "level1 restrict")
        assert_exact(orc.prolong(mc, noncompact(len(xcc), 0.3), x1), unhex(l1["prolong"]), "l1 prolong")


# -------------------------------------------- direct comparison with oracle/_ref

needs_ref = pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built here")


@needs_ref
def test_random_states_against_reference(orc):
    rng = np.random.default_rng(1234)
    R = po.ref()
    for Mt, Mx, ctype, beta in [(16, 16, po.BOTH, 3.3), (12, 16, po.BOTH, 40.0), (8, 16, po.TEMPORAL, 5.0),
                                (16, 8, po.SPATIAL, 11.0)]:
        a = R.action(po.SCHWINGER, [Mt, Mx, ctype, 0], [beta])
        m = po.schwinger(Mt, Mx, beta, ctype)
        x = rng.uniform(-np.pi, np.pi, a.n)
        assert orc.action(m, x) == a.evaluate(x)
        assert_exact(orc.force(m, x), a.force(x), "force")
        assert orc.qoi(m, po.QOI_SCHWINGER_CHI, x)[0] == a.qoi(po.QOI_SCHWINGER_CHI, x)
        assert abs(orc.cond_action(m, x) / a.cond_evaluate(x) - 1) < TOL
    for M, m0 in [(64, 0.25), (256, 0.25), (32, 30.0)]:
        a = R.action(po.ROTOR, [M, 0], [4.0, m0])
        m = po.rotor(M, 4.0, m0)
        w = np.pi if m0 < 1 else 0.3  # large m0/a: rough states underflow the pdf to 0
        x = rng.uniform(-w, w, M)
        assert orc.action(m, x) == a.evaluate(x)
        assert_exact(orc.force(m, x), a.force(x), "force")
        assert abs(orc.cond_action(m, x) / a.cond_evaluate(x) - 1) < TOL


@needs_ref
def test_draws_follow_reference_distributions(orc):
    """the Philox restatements of the four draw() algorithms sample the same
    distributions as the reference's own draw() (two-sample KS test)"""
    st = pytest.importorskip("scipy.stats")
    R = po.ref()
    n = 20000
    cases = [(0, 6.0, 0, 0, lambda r: orc.lib.orc_expsin2_draw(C.byref(r), 6.0)),
             (1, 4.0, 0.9, 0.2, lambda r: orc.lib.orc_expcos_draw(C.byref(r), 4.0, 0.9, 0.2)),
             (1, 4.0, 2.9, -2.8, lambda r: orc.lib.orc_expcos_draw(C.byref(r), 4.0, 2.9, -2.8)),
             (2, 4.0, 0.9, 0.2, lambda r: orc.lib.orc_besselproduct_draw(C.byref(r), 4.0, 0.9, 0.2)),
             (2, 1.5, -2.0, 2.5, lambda r: orc.lib.orc_besselproduct_draw(C.byref(r), 1.5, -2.0, 2.5)),
             (3, 16.0, 0.9, 0.2, lambda r: orc.lib.orc_approxbessel_draw(C.byref(r), 16.0, 0.9, 0.2)),
             (3, 16.0, -2.0, 2.5, lambda r: orc.lib.orc_approxbessel_draw(C.byref(r), 16.0, -2.0, 2.5))]
    def tight(beta, xp, xm, variant=1):
        def f(r):
            orc.lib.orc_set_expcos_envelope(variant)
            v = orc.lib.orc_expcos_draw(C.byref(r), beta, xp, xm)
            orc.lib.orc_set_expcos_envelope(0)
            return v
        return f
    # the product's tighter ExpCos envelopes (incl. the uniform-proposal branch tau < 1/2)
    cases += [(1, b, xp, xm, tight(b, xp, xm))
              for (b, xp, xm) in [(4.0, 0.9, 0.2), (4.0, 2.9, -2.8), (0.1, 0.3, -0.4), (0.6, 1.0, 1.2),
                                  (300.0, -1.0, -0.9), (2.0, 3.0, -0.1)]]
    # variant 2: Taylor-bound envelope for tau = 2 beta |cos(dx/2)| >= 64 (tau = 64.3, 599, 2047, 142
    # with the pi shift), chord / uniform branches below
    cases += [(1, b, xp, xm, tight(b, xp, xm, 2))
              for (b, xp, xm) in [(33.0, 0.5, 0.05), (300.0, -1.0, -0.9), (1024.0, 0.31, 0.25),
                                  (500.0, 3.0, -0.1), (4.0, 0.9, 0.2), (0.1, 0.3, -0.4)]]
    for dist, prm, xp, xm, draw in cases:
        want = np.zeros(n)
        R.lib.ref_dist_draw(dist, prm, xp, xm, 4711, n, want.ctypes.data_as(po.c_double_p))
        got = np.array([draw(orc.rng(99, po.STREAM_FILL2, 0, 0, k)) for k in range(n)])
        assert st.ks_2samp(got, want).pvalue > 1e-3, (dist, prm, xp, xm)


def test_ho_exact_sampler_factor(orc):
    """orc_ho_exact_factor restates HarmonicOscillatorAction::build_covariance
    (qm/harmonicoscillatoraction.cc:38-56): L L^T P = 1 for the cyclic tridiagonal precision matrix P"""
    for M, T, m0, mu2 in ((32, 4.0, 1.0, 1.0), (12, 3.0, 0.7, 2.5), (64, 4.0, 0.25, 0.3)):
        m = po.ho(M, T, m0, mu2)
        L = orc.ho_exact_factor(m)
        a = T / M
        P = np.zeros((M, M))
        for i in range(M):
            P[i, i] = a * m0 * mu2 + 2 * m0 / a
            P[i, (i + 1) % M] += -m0 / a
            P[i, (i - 1) % M] += -m0 / a
        assert np.allclose(L, np.tril(L))
        assert np.max(np.abs(L @ L.T @ P - np.eye(M))) < 1e-10


@needs_ref
def test_ho_exact_sampler_matches_reference_draws(orc):
    """exact draws of the restatement (Philox) and of the reference's own
    HarmonicOscillatorAction::draw (mt19937_64) follow the same Gaussian: second moments within
    statistical error, single-site marginals by a two-sample KS test"""
    st = pytest.importorskip("scipy.stats")
    R = po.ref()
    M, n = 32, 20000
    a = R.action(po.HO, [M, 0], [4.0, 1.0, 1.0])
    want = np.zeros((n, M))
    R.lib.ref_ho_exact_draws(a.h, n, want.ctypes.data_as(po.c_double_p))
    m = po.ho(M, 4.0, 1.0, 1.0)
    got = np.array([orc.ho_exact_draw(m, 77, 3, k) for k in range(n)])
    for site in (0, 5, 31):
        assert st.ks_2samp(got[:, site], want[:, site]).pvalue > 1e-3
    C_got, C_want = got.T @ got / n, want.T @ want / n
    # entries of a sample covariance have standard error ~ sqrt((C_ii C_jj + C_ij^2) / n)
    tol = 6.0 * np.sqrt(2.0 / n) * np.max(np.diag(C_want))
    assert np.max(np.abs(C_got - C_want)) < 2 * tol
    C_exact = orc.ho_exact_factor(m) @ orc.ho_exact_factor(m).T
    assert np.max(np.abs(C_got - C_exact)) < tol


@needs_ref
@pytest.mark.parametrize("M,ctype", [(8, po.ROTATE), (16, po.ROTATE), (8, po.BOTH)])
def test_gff_dense_coarse_action_matches_reference(orc, M, ctype):
    """numpy restatement of GFFAction::buildMatrices against the reference's own coarse GFF action
    (n_gibbs_smooth = 2, omega = 1: gffaction.hh:201-208), S = phi^T Q_hat phi / 2, and its exact
    sampler (second moments of gffaction.cc:200-213 draws = Sigma_hat)"""
    R = po.ref()
    fine = R.action(po.GFF, [M, M, ctype], [3.0])
    coarse = fine.coarse()
    m = po.gff(M, M, 3.0, ctype)
    mc = orc.coarse_model(m, 0, 0, ctype)
    assert abs(coarse.param(0) - mc.gff_mu2) < 1e-15
    mats = po.gff_dense_matrices(orc, mc, 2, 1.0)
    rng = np.random.default_rng(M)
    for _ in range(3):
        x = rng.normal(size=coarse.n)
        want = coarse.evaluate(x)
        got = 0.5 * x @ mats["Q_hat"] @ x
        assert abs(got - want) <= 1e-9 * abs(want), (got, want)
    n = 20000
    draws = np.zeros((n, coarse.n))
    R.lib.ref_gff_exact_draws(coarse.h, n, draws.ctypes.data_as(po.c_double_p))
    C = draws.T @ draws / n
    tol = 6.0 * np.sqrt(2.0 / n) * np.max(np.diag(mats["Sigma_hat"]))
    assert np.max(np.abs(C - mats["Sigma_hat"])) < tol
