"""GPU parity tests: the CUDA path, called through the C-ABI, against
(a) the golden vectors recorded from the reference's own translation units,
(b) the C oracle on seeded random inputs (several chains per call), and
(c) size-independent properties at the full BASELINE.json sizes.

Tolerances (north_star): indexing / coarsening maps / topological charge bit-exact;
action, force, QoI values <= 1e-12 relative (norm-wise for vectors); stochastic
kernels are compared draw by draw with the oracle's restatement of the reference's
sampling algorithms on the same Philox stream (angles compared modulo 2 pi)."""
import zlib

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import load, qm_model, scalar, unhex

pytestmark = pytest.mark.gpu

TOL = 1e-12
SEED = 0x5EED0001


@pytest.fixture(scope="module")
def mp():
    import mlmcpathintegral_b200 as mp
    return mp


@pytest.fixture(scope="module")
def ctx(mp):
    c = mp.Context(0, seed=SEED)
    yield c
    c.close()


def to_mp(mp, o):
    """oracle model -> product model (same fields)"""
    return mp.Model(model=o.model, M_lat=o.M_lat, Mt_lat=o.Mt_lat, Mx_lat=o.Mx_lat,
                    rotated=o.rotated, coarsening=o.coarsening, a_lat=o.a_lat, T_final=o.T_final,
                    m0=o.m0, mu2=o.mu2, lambda_=o.lambda_, x0=o.x0, beta=o.beta, gff_mu2=o.gff_mu2)


def dev(ctx, a):
    a = np.atleast_2d(np.asarray(a, dtype=np.float64))
    return ctx.to_device(a)


def host(t):
    return t.detach().cpu().numpy()


def close(got, want, tol=TOL, what=""):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = max(float(np.max(np.abs(want))), 1e-300)
    err = float(np.max(np.abs(got - want))) / scale
    assert err <= tol, f"{what}: relative error {err:.3e} > {tol}"


def ang_close(got, want, tol=1e-10, what="", max_bad=0):
    d = np.asarray(got) - np.asarray(want)
    d = d - 2 * np.pi * np.round(d / (2 * np.pi))
    bad = int(np.sum(np.abs(d) > tol))
    assert bad <= max_bad, f"{what}: {bad} entries differ by more than {tol} (max {np.max(np.abs(d)):.3e})"


# ----------------------------------------------------------- golden fixtures


@pytest.mark.parametrize("c", load("schwinger"), ids=lambda c: c["name"])
def test_schwinger_golden(mp, ctx, c):
    from tools.make_golden import angles
    m = mp.schwinger(c["Mt"], c["Mx"], c["beta"], c["ctype"], 0)
    x = unhex(c["x"])
    xd = dev(ctx, x)
    close(host(ctx.action(m, xd))[0], scalar(c["S"]), what="S")
    close(host(ctx.force(m, xd))[0], unhex(c["force"]), what="force")
    chi, Q = ctx.qoi(m, mp.QOI_SCHWINGER_CHI, xd, with_charge=True)
    want_chi = scalar(c["qoi_chi"])
    assert abs(host(chi)[0] - want_chi) <= TOL * max(want_chi, 1.0)
    assert int(host(Q)[0]) ** 2 == int(round(want_chi * 4 * np.pi ** 2 / (2 * np.pi) ** 2))
    close(host(ctx.qoi(m, mp.QOI_AVG_PLAQUETTE, xd))[0], scalar(c["qoi_plaq"]), what="plaq")
    y = xd.clone()
    ctx.overrelax_sweep(m, y)
    ang_close(host(y)[0], unhex(c["overrelax_coloured"]), what="overrelax")
    close(host(ctx.cond_action(m, xd))[0], scalar(c["cond_S"]), what="cond_S")
    lf = c["leapfrog"]
    y, p = xd.clone(), dev(ctx, unhex(c["p0"]))
    ctx.leapfrog(m, lf["nt"], lf["dt"], y, p)
    close(host(y)[0], unhex(lf["x"]), what="leapfrog x")
    close(host(p)[0], unhex(lf["p"]), what="leapfrog p")
    mc = mp.coarse_model(m, renorm=c["renorm"], level=0, ctype=c["ctype"])
    xc = ctx.state(mc, 1)
    ctx.restrict(m, xd, xc)
    ang_close(host(xc)[0], unhex(c["restrict"]), tol=1e-14, what="restrict")
    close(host(ctx.action(mc, xc))[0], scalar(c["coarse_S"]), what="coarse S")
    y = xd.clone()
    ctx.prolong(m, dev(ctx, angles(mp.sample_size(mc), 0.3)), y)
    assert np.array_equal(host(y)[0], unhex(c["prolong"]))  # pure copies / halvings: exact
    tl = c["twolevel"]
    tp, pc = dev(ctx, unhex(tl["theta_prime"])), dev(ctx, unhex(tl["phi_coarse"]))
    want = unhex(tl["deltas"])
    got = [host(ctx.action(m, tp) - ctx.action(m, xd))[0],
           host(ctx.action(mc, xc) - ctx.action(mc, pc))[0],
           host(ctx.cond_action(m, xd) - ctx.cond_action(m, tp))[0]]
    scale = max(abs(scalar(c["S"])), abs(scalar(c["cond_S"])), 1.0)
    assert np.max(np.abs(np.array(got) - want)) <= TOL * scale


@pytest.mark.parametrize("c", load("qm"), ids=lambda c: c["name"])
def test_qm_golden(mp, ctx, c):
    from tools.make_golden import noncompact
    m = to_mp(mp, qm_model(po, c))
    x = unhex(c["x"])
    xd = dev(ctx, x)
    rotor = c["kind"] == po.ROTOR
    close(host(ctx.action(m, xd))[0], scalar(c["S"]), what="S")
    close(host(ctx.force(m, xd))[0], unhex(c["force"]), what="force")
    close(host(ctx.cond_action(m, xd))[0], scalar(c["cond_S"]), what="cond_S")
    close(host(ctx.qoi(m, mp.QOI_X2, xd))[0], scalar(c["qoi_x2"]), what="x2")
    if rotor:
        chi, Q = ctx.qoi(m, mp.QOI_ROTOR_CHI, xd, with_charge=True)
        want = scalar(c["qoi_chi"])
        assert abs(host(chi)[0] - want) <= TOL * max(want, 1.0)
        assert int(host(Q)[0]) ** 2 == int(round(want * 4 * np.pi ** 2 * m.T_final / (2 * np.pi) ** 2))
        y = xd.clone()
        ctx.overrelax_sweep(m, y)
        ang_close(host(y)[0], unhex(c["overrelax_coloured"]), what="overrelax")
    lf = c["leapfrog"]
    y, p = xd.clone(), dev(ctx, unhex(c["p0"]))
    ctx.leapfrog(m, lf["nt"], lf["dt"], y, p)
    close(host(y)[0], unhex(lf["x"]), what="leapfrog x")
    close(host(p)[0], unhex(lf["p"]), what="leapfrog p")
    mc = mp.coarse_model(m, renorm=c["ip"][1])
    xc = ctx.state(mc, 1)
    ctx.restrict(m, xd, xc)
    assert np.array_equal(host(xc)[0], unhex(c["restrict"]))
    y = xd.clone()
    ctx.prolong(m, dev(ctx, noncompact(len(x) // 2, 0.3)), y)
    assert np.array_equal(host(y)[0], unhex(c["prolong"]))
    close(host(ctx.action(mc, xc))[0], scalar(c["coarse_S"]), what="coarse S")


@pytest.mark.parametrize("c", load("gff"), ids=lambda c: c["name"])
def test_gff_golden(mp, ctx, c):
    from tools.make_golden import noncompact
    m = mp.gff(c["Mt"], c["Mx"], c["mass"], c["ctype"], 0)
    xd = dev(ctx, unhex(c["x"]))
    close(host(ctx.action(m, xd))[0], scalar(c["S"]), what="S")
    close(host(ctx.force(m, xd))[0], unhex(c["force"]), what="force")
    close(host(ctx.qoi(m, mp.QOI_PHI2, xd))[0], scalar(c["qoi_phi2"]), what="phi2")
    lf = c["leapfrog"]
    y, p = xd.clone(), dev(ctx, unhex(c["p0"]))
    ctx.leapfrog(m, lf["nt"], lf["dt"], y, p)
    close(host(y)[0], unhex(lf["x"]), what="leapfrog x")
    close(host(p)[0], unhex(lf["p"]), what="leapfrog p")
    mc = mp.coarse_model(m, level=0, ctype=c["ctype"])
    xc = ctx.state(mc, 1)
    ctx.restrict(m, xd, xc)
    assert np.array_equal(host(xc)[0], unhex(c["restrict"]))
    # the coarse action of the reference: Gibbs-smoothed dense precision matrix (gffaction.cc:25-28,
    # 133-174; coarse_action() sets n_gibbs_smooth = 2, omega = 1)
    assert mc.gff_n_gibbs == 2 and mc.gff_omega == 1.0
    close(host(ctx.action(mc, xc))[0], scalar(c["coarse_S_gibbs"]), tol=1e-9, what="coarse S (Q_hat)")
    y = xd.clone()
    ctx.prolong(m, dev(ctx, noncompact(mp.sample_size(mc), 0.3)), y)
    assert np.array_equal(host(y)[0], unhex(c["prolong"]))
    if c["ctype"] == po.ROTATE:
        close(host(ctx.cond_action(m, xd))[0], scalar(c["cond_S"]), what="cond_S")
        l1 = c["level1"]
        mc.gff_mu2 = scalar(c["coarse_mu2"])
        x1 = dev(ctx, unhex(l1["x"]))
        close(host(ctx.force(mc, x1))[0], unhex(l1["force"]), what="level1 force")
        close(host(ctx.cond_action(mc, x1))[0], scalar(l1["cond_S"]), what="level1 cond_S")
        mcc = mp.coarse_model(mc, level=1, ctype=c["ctype"])
        xcc = ctx.state(mcc, 1)
        ctx.restrict(mc, x1, xcc)
        assert np.array_equal(host(xcc)[0], unhex(l1["restrict"]))
        y = x1.clone()
        ctx.prolong(mc, dev(ctx, noncompact(mp.sample_size(mcc), 0.3)), y)
        assert np.array_equal(host(y)[0], unhex(l1["prolong"]))


# ------------------------------------------- oracle, seeded random inputs, B > 1

MODELS = {
    "ho32": lambda: po.ho(32),
    "quartic64": lambda: po.quartic(64, 4.0, 1.0, 1.0, 1.0, 1.0),
    "rotor256": lambda: po.rotor(256, 4.0, 0.25),
    "rotor48": lambda: po.rotor(48, 4.0, 2.0),
    "schw_both_b4_32": lambda: po.schwinger(32, 32, 4.0, po.BOTH),
    "schw_both_b16_32x16": lambda: po.schwinger(32, 16, 16.0, po.BOTH),
    "schw_both_12x20": lambda: po.schwinger(12, 20, 2.0, po.BOTH),
    "schw_both_b512_16": lambda: po.schwinger(16, 16, 512.0, po.BOTH),
    "schw_temporal": lambda: po.schwinger(16, 12, 5.0, po.TEMPORAL),
    "schw_spatial": lambda: po.schwinger(12, 16, 5.0, po.SPATIAL),
    "gff_rot16": lambda: po.gff(16, 16, 10.0, po.ROTATE, 0),
    "gff_rot16_l1": lambda: po.gff(16, 16, 10.0, po.ROTATE, 1),
}


def random_state(o, rng, B, smooth=1.0):
    n = po.oracle().sample_size(o)
    if o.model in (po.ROTOR, po.SCHWINGER):
        return rng.uniform(-np.pi * smooth, np.pi * smooth, (B, n))
    return rng.normal(0.0, 0.7, (B, n))


def coarse_of(orc, o):
    if o.model == po.SCHWINGER:
        return orc.coarse_model(o, 0, 0, o.coarsening)
    if o.model == po.GFF:
        return orc.coarse_model(o, 0, o.rotated, po.ROTATE)
    return orc.coarse_model(o, 0, 0, 0, o.T_final)


@pytest.mark.parametrize("name", list(MODELS))
def test_deterministic_kernels_against_oracle(mp, ctx, orc, name):
    o = MODELS[name]()
    m = to_mp(mp, o)
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    B = 5
    x = random_state(o, rng, B)
    xd = dev(ctx, x)
    S = host(ctx.action(m, xd))
    close(S, [orc.action(o, x[b]) for b in range(B)], what="action")
    F = host(ctx.force(m, xd))
    close(F, [orc.force(o, x[b]) for b in range(B)], what="force")
    p = rng.normal(size=x.shape)
    y, pd = xd.clone(), dev(ctx, p)
    ctx.leapfrog(m, 10, 0.02, y, pd)
    want = [orc.leapfrog(o, 10, 0.02, x[b], p[b]) for b in range(B)]
    close(host(y), [w[0] for w in want], tol=1e-11, what="leapfrog x")
    close(host(pd), [w[1] for w in want], tol=1e-11, what="leapfrog p")
    if o.model in (po.ROTOR, po.SCHWINGER, po.GFF):
        y = xd.clone()
        ctx.overrelax_sweep(m, y)
        want = np.array([orc.overrelax_sweep(o, x[b], coloured=True) for b in range(B)])
        if o.model == po.GFF:
            close(host(y), want, what="overrelax")
        else:
            ang_close(host(y), want, what="overrelax")
    oc = coarse_of(orc, o)
    mc = to_mp(mp, oc)
    xc = ctx.state(mc, B)
    ctx.restrict(m, xd, xc)
    want = np.array([orc.restrict(o, oc, x[b]) for b in range(B)])
    ang_close(host(xc), want, tol=1e-14, what="restrict")
    xcr = random_state(oc, rng, B)
    y = xd.clone()
    ctx.prolong(m, dev(ctx, xcr), y)
    assert np.array_equal(host(y), np.array([orc.prolong(o, xcr[b], x[b]) for b in range(B)]))
    xs = random_state(o, rng, B, smooth=min(1.2 / np.sqrt(o.beta), 0.3) if o.beta > 8 else 1.0)
    close(host(ctx.cond_action(m, dev(ctx, xs))), [orc.cond_action(o, xs[b]) for b in range(B)],
          tol=1e-11, what="cond_action")
    qois = {po.HO: [po.QOI_X2], po.QUARTIC: [po.QOI_X2], po.ROTOR: [po.QOI_X2, po.QOI_ROTOR_CHI],
            po.SCHWINGER: [po.QOI_SCHWINGER_CHI, po.QOI_AVG_PLAQUETTE], po.GFF: [po.QOI_PHI2]}[o.model]
    for q in qois:
        got, Q = ctx.qoi(m, q, xd, with_charge=True)
        want = [orc.qoi(o, q, x[b]) for b in range(B)]
        wv = np.array([w[0] for w in want])
        assert np.max(np.abs(host(got) - wv)) <= 1e-11 * max(np.max(np.abs(wv)), 1.0)
        if q in (po.QOI_ROTOR_CHI, po.QOI_SCHWINGER_CHI):
            assert list(host(Q)) == [w[1] for w in want]  # integer topological charge: exact


STOCHASTIC_CASES = [(n, e) for n in MODELS for e in ((0, 1, 2) if n.startswith("schw") else (2,))]


@pytest.fixture
def envelope(request, ctx, orc):
    """ExpCos proposal: 0 = the reference's envelope, 1 = chord bound, 2 = the product's default
    (chord bound + Taylor bound for tau >= 64)"""
    ctx.set_expcos_envelope(request.param)
    orc.lib.orc_set_expcos_envelope(request.param)
    yield request.param
    ctx.set_expcos_envelope(2)
    orc.lib.orc_set_expcos_envelope(0)


@pytest.mark.parametrize("name,envelope", STOCHASTIC_CASES, indirect=["envelope"])
def test_stochastic_kernels_against_oracle(mp, ctx, orc, name, envelope):
    o = MODELS[name]()
    m = to_mp(mp, o)
    rng = np.random.default_rng(zlib.crc32(name.encode()) + 1)
    B, chain0, draw = 4, 7, (3 << 32) + 11
    compact = o.model in (po.ROTOR, po.SCHWINGER)
    cmp = ang_close if compact else (lambda a, b, tol=1e-10, what="", max_bad=0: close(a, b, tol, what))
    # initial state and momenta
    want = np.array([orc.init_state(o, SEED, draw, chain0 + b) for b in range(B)])
    close(host(ctx.init_state(m, B, chain0, draw)), want, tol=1e-14, what="init_state") if np.any(want) \
        else None
    want = np.array([orc.hmc_momentum(o, SEED, draw, chain0 + b) for b in range(B)])
    close(host(ctx.hmc_momentum(m, B, chain0, draw)), want, tol=1e-13, what="momentum")
    # one HMC step (short trajectory so that rounding differences stay ~1e-13)
    x = random_state(o, rng, B)
    xd = dev(ctx, x)
    acc, diag = ctx.hmc_step(m, 6, 0.03, xd, chain0, draw)
    res = [orc.hmc_step(o, 6, 0.03, SEED, draw, chain0 + b, x[b]) for b in range(B)]
    assert list(host(acc)) == [r[0] for r in res]
    close(host(xd), [r[1] for r in res], tol=1e-11, what="hmc state")
    want_diag = np.array([r[2] for r in res])
    assert np.max(np.abs(host(diag) - want_diag)) <= 1e-10 * max(np.max(np.abs(want_diag[:, 1:])), 1.0)
    # heat bath sweep
    if o.model in (po.ROTOR, po.SCHWINGER, po.GFF):
        x = random_state(o, rng, B)
        xd = dev(ctx, x)
        ctx.heatbath_sweep(m, xd, chain0, draw)
        want = np.array([orc.heatbath_sweep(o, SEED, draw, chain0 + b, x[b]) for b in range(B)])
        cmp(host(xd), want, tol=1e-9, what="heatbath")
    # fill-in: in place, and fused with the prolongation
    oc = coarse_of(orc, o)
    mc = to_mp(mp, oc)
    xc = random_state(oc, rng, B)
    x0 = random_state(o, rng, B)
    pro = np.array([orc.prolong(o, xc[b], x0[b]) for b in range(B)])
    want = np.array([orc.fill(o, SEED, draw, chain0 + b, pro[b]) for b in range(B)])
    xd = dev(ctx, pro)
    ctx.fill(m, xd, chain0, draw)
    cmp(host(xd), want, tol=1e-9, what="fill")
    xd2 = dev(ctx, x0)
    ctx.prolong_fill(m, dev(ctx, xc), xd2, chain0, draw)
    cmp(host(xd2), want, tol=1e-9, what="prolong_fill")
    # ... and fused with the two reductions of the trial state: same theta', and S_f / S_cond equal
    # to the stand-alone evaluations of that state
    xd3 = dev(ctx, x0)
    Sf_f, Sc_f = ctx.prolong_fill_eval(m, dev(ctx, xc), xd3, chain0, draw)
    assert np.array_equal(host(xd3), host(xd2))
    close(host(Sf_f), host(ctx.action(m, xd3)), tol=1e-11, what="fused S_f")
    close(host(Sc_f), host(ctx.cond_action(m, xd3)), tol=1e-10, what="fused S_cond")
    # two-level Metropolis-Hastings step
    xf = random_state(o, rng, B, smooth=min(1.2 / np.sqrt(o.beta), 0.3) if o.beta > 8 else 1.0)
    xfd, xcd = dev(ctx, xf), dev(ctx, xc)
    Sf, Sc = ctx.action(m, xfd), ctx.cond_action(m, xfd)
    Sf0, Sc0 = host(Sf).copy(), host(Sc).copy()
    acc, deltas = ctx.twolevel_step(m, mc, xcd, xfd, Sf, Sc, chain0, draw)
    res = [orc.twolevel_step(o, oc, SEED, draw, chain0 + b, xc[b], xf[b], Sf0[b], Sc0[b])
           for b in range(B)]
    want_d = np.array([r[4] for r in res])
    assert np.max(np.abs(host(deltas) - want_d)) <= 1e-9 * max(np.max(np.abs(want_d)), 1.0)
    assert list(host(acc)) == [r[0] for r in res]
    cmp(host(xfd), np.array([r[1] for r in res]), tol=1e-9, what="twolevel state")
    close(host(Sf), [r[2] for r in res], tol=1e-10, what="cached S_f")


@pytest.mark.parametrize("beta", [6.0, 96.0])
def test_hierarchical_draw_equals_explicit_cascade(mp, ctx, beta):
    """HierarchicalSampler::draw (hierarchicalsampler.cc:55-81) inside the library -- fused
    fill-in + reductions, per-level action caches -- against the same cascade composed from the
    public single-purpose entry points (restrict, hmc_step, action, cond_action, twolevel_step)"""
    import torch
    L, B, chain0, nt, dt = 3, 6, 3, 5, 0.05
    m = mp.schwinger(16, 16, beta)
    smp = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_HMC, n_levels=L, nt=nt, dt=dt,
                     renorm=mp.RENORM_PERTURBATIVE, chain0=chain0)
    models = [smp.level_model(l) for l in range(L)]
    x0 = ctx.init_state(m, B, chain0, 5)
    for k in range(3):
        ctx.heatbath_sweep(m, x0, chain0, k)
    smp.set_state(x0)
    out = x0.clone()
    x = [x0.clone()] + [ctx.state(models[l], B) for l in range(1, L)]
    level_draw = lambda d, l: (d << 12) | (l << 8)
    n_acc = 0
    for d in range(4):
        smp.draw(out)
        for l in range(1, L):
            ctx.restrict(models[l - 1], x[l - 1], x[l])
        mask, _ = ctx.hmc_step(models[L - 1], nt, dt, x[L - 1], chain0, level_draw(d, L - 1))
        mask = mask.bool()
        for l in range(L - 2, -1, -1):
            keep = x[l].clone()
            Sf, Sc = ctx.action(models[l], x[l]), ctx.cond_action(models[l], x[l])
            acc, _ = ctx.twolevel_step(models[l], models[l + 1], x[l + 1], x[l], Sf, Sc, chain0,
                                       level_draw(d, l))
            x[l][~mask] = keep[~mask]  # `if (not accept) break`, hierarchicalsampler.cc:73-74
            mask = mask & acc.bool()
        n_acc += int(mask.sum())
        ang_close(host(out), host(x[0]), tol=1e-11, what=f"draw {d}")
    assert 0 < n_acc < 4 * B  # both branches of the cascade were exercised


def test_cached_cascade_equals_literal_cascade(mp, ctx):
    """MLMCPI_OPT_CASCADE_CACHE: the hierarchical Schwinger draw that keeps the coarse levels tentative (no
    restriction chain, cached level actions) against the literal sequence of hierarchicalsampler.cc:55-81:
    identical states and acceptance counters over many draws, 2 to 4 levels, with set_state and autotune
    (which advance coarse states behind the cache's back) in between"""
    for M, beta, L, B in [(32, 9.0, 2, 16), (64, 64.0, 3, 12), (64, 16.0, 4, 8), (32, 4.0, 3, 10)]:
        m = mp.schwinger(M, M, beta)
        res = []
        for cache in (0, 1):
            ctx.set_option(mp._lib.OPT_CASCADE_CACHE, cache)
            smp = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_HMC, n_levels=L, nt=6, dt=0.05,
                             renorm=mp.RENORM_PERTURBATIVE, chain0=7)
            x = ctx.init_state(m, B, 7, 3)
            for k in range(3):
                ctx.heatbath_sweep(m, x, 7, k)
            smp.set_state(x)
            states = []
            for d in range(12):
                smp.draw(x)
                states.append(host(x).copy())
                if d == 4:
                    smp.autotune(0.8, 2, 2 * B)
                    smp.set_dt(0.05)
                if d == 8:
                    smp.set_state(x)
            res.append((states, smp.p_accept(), host(smp.get_state()).copy()))
            smp.close()
        ctx.set_option(mp._lib.OPT_CASCADE_CACHE, 1)
        # (identical up to the last bits of the coarse angles: theta'_l stands in for restrict(fill(prolong(
        # theta'_l))), which reproduces it up to rounding)
        for d, (a, b) in enumerate(zip(res[0][0], res[1][0])):
            ang_close(a, b, tol=1e-9, what=f"{M}^2 beta {beta} {L} levels, draw {d}")
        assert res[0][1] == res[1][1]
        ang_close(res[0][2], res[1][2], tol=1e-9, what="get_state")
        assert 0.0 < res[1][1][0] < 1.0 or L == 2


def test_cached_cascade_other_coarse_samplers(mp, ctx):
    """the cached cascade with a coarse sampler that advances a COPY of the coarsest state (heat bath, cluster; GFF
    with heat bath on the Gibbs-smoothed dense coarse actions and with the exact coarse sampler) against the literal
    sequence: same states (up to the rounding of the cached level actions), same acceptance counters"""
    cases = [("schwinger", 32, 9.0, 3, dict(kind=mp.SAMPLER_HEATBATH, n_sweep_overrelax=2, n_sweep_heatbath=1)),
             ("schwinger", 64, 64.0, 3, dict(kind=mp.SAMPLER_CLUSTER, n_updates=20)),
             ("schwinger", 32, 16.0, 2, dict(kind=mp.SAMPLER_CLUSTER, n_updates=10)),
             ("gff", 16, 3.0, 3, dict(kind=mp.SAMPLER_HEATBATH, n_sweep_overrelax=2, n_sweep_heatbath=1)),
             ("gff", 32, 10.0, 4, dict(kind=mp.SAMPLER_HEATBATH, n_sweep_overrelax=1, n_sweep_heatbath=1)),
             ("gff", 16, 3.0, 2, dict(kind=mp.SAMPLER_EXACT))]
    for name, M, par, L, kw in cases:
        B = 12
        if name == "schwinger":
            m = mp.schwinger(M, M, par)
            kw = dict(kw, renorm=mp.RENORM_PERTURBATIVE)
        else:
            m = mp.gff(M, M, par, mp.COARSEN_ROTATE)
            kw = dict(kw, ctype=mp.COARSEN_ROTATE)
            with pytest.raises(mp.MlmcpiError):  # any other coarsening of a GFF hierarchy is refused
                mp.Sampler(ctx, m, B, n_levels=L, **dict(kw, ctype=mp.COARSEN_BOTH))
        res = []
        for cache in (0, 1):
            ctx.set_option(mp._lib.OPT_CASCADE_CACHE, cache)
            smp = mp.Sampler(ctx, m, B, n_levels=L, chain0=3, **kw)
            x = smp.get_state()
            states = []
            for d in range(10):
                smp.draw(x)
                states.append(host(x).copy())
                if d == 5:
                    smp.set_state(x)
            res.append((states, smp.p_accept(), host(smp.get_state()).copy()))
            smp.close()
        ctx.set_option(mp._lib.OPT_CASCADE_CACHE, 1)
        cmp = ang_close if name == "schwinger" else (lambda a, b, tol, what: close(a, b, tol, what))
        for d, (a, b) in enumerate(zip(res[0][0], res[1][0])):
            cmp(a, b, tol=1e-9, what=f"{name} {M}^2 {L} levels {kw['kind']}, draw {d}")
        assert res[0][1] == res[1][1], (name, M, L, res[0][1], res[1][1])
        cmp(res[0][2], res[1][2], tol=1e-9, what="get_state")


def test_sampler_qoi_fused_charge(mp, ctx):
    """mlmcpi_sampler_qoi: the susceptibility QoI the cached cascade maintains from the charge sum of the fill-in
    kernel equals mlmcpi_qoi of the chains' states, draw by draw, for HMC / cluster / heat-bath coarse samplers, beta
    above and below 8 (both fill-in distributions), across set_state and autotune; other QoIs and samplers without
    the fused path fall back to the evaluation pass"""
    for M, beta, L, kw in [(32, 9.0, 2, dict(kind=mp.SAMPLER_HMC, nt=6, dt=0.05)),
                           (64, 64.0, 3, dict(kind=mp.SAMPLER_CLUSTER, n_updates=20)),
                           (32, 4.0, 3, dict(kind=mp.SAMPLER_HMC, nt=6, dt=0.05)),
                           (32, 16.0, 2, dict(kind=mp.SAMPLER_HEATBATH, n_sweep_overrelax=2, n_sweep_heatbath=1))]:
        m = mp.schwinger(M, M, beta)
        B = 16
        smp = mp.Sampler(ctx, m, B, n_levels=L, renorm=mp.RENORM_PERTURBATIVE, chain0=2, **kw)
        x = smp.get_state()
        changed = 0
        for d in range(12):
            before = host(smp.qoi(mp.QOI_SCHWINGER_CHI)).copy()
            smp.draw(x)
            got = host(smp.qoi(mp.QOI_SCHWINGER_CHI))
            want = host(ctx.qoi(m, mp.QOI_SCHWINGER_CHI, smp.get_state()))
            assert np.max(np.abs(got - want)) <= 1e-9 * max(1.0, np.max(np.abs(want))), (M, beta, L, d)
            changed += int(np.sum(got != before))
            close(host(smp.qoi(mp.QOI_AVG_PLAQUETTE)), host(ctx.qoi(m, mp.QOI_AVG_PLAQUETTE, smp.get_state())),
                  tol=1e-13, what="other QoI")
            if d == 4 and kw["kind"] == mp.SAMPLER_HMC:
                smp.autotune(0.8, 2, 2 * B)
            if d == 7:
                y = ctx.init_state(m, B, 9, 1)
                smp.set_state(y)
        smp.close()


def test_draw_host_async_hands_back_accepted_states(mp, ctx):
    """mlmcpi_sampler_draw_host_async: chains resident, QoI of every chain and the states of the ACCEPTED
    chains written to the host buffer on a second stream (pinned memory: by a masked copy kernel straight
    over the host link; pageable memory: full copy).  A buffer initialised with the chains' states
    therefore tracks them draw by draw, exactly like the d_x_out of mlmcpi_sampler_draw."""
    import torch
    m = mp.schwinger(32, 32, 9.0)
    B = 24
    # pinned + copy engine (default: accept flags to the host, one copy per run of accepted chains, two snapshot
    # buffers), pinned + masked copy kernel over the host link, pageable (full copy)
    for pinned, engine in ((True, 1), (True, 0), (False, 1)):
        ctx.set_option(mp._lib.OPT_HOST_COPY_ENGINE, engine)
        s = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_HMC, n_levels=2, nt=8, dt=0.05, renorm=mp.RENORM_PERTURBATIVE)
        x0 = ctx.init_state(m, B, 0, 1)
        for k in range(3):
            ctx.heatbath_sweep(m, x0, 0, k)
        s.set_state(x0)
        h_x = torch.empty(B, mp.sample_size(m), dtype=torch.float64, pin_memory=pinned)
        h_q = torch.empty(B, dtype=torch.float64, pin_memory=pinned)
        h_x.copy_(s.get_state())
        hx = h_x if pinned else h_x.numpy()
        hq = h_q if pinned else h_q.numpy()
        changed = 0
        for d in range(8):
            before = h_x.clone()
            s.draw_host_async(mp.QOI_SCHWINGER_CHI, hq, hx)
            s.wait_host()
            now = s.get_state()
            assert torch.equal(h_x, now.cpu()), (pinned, d)
            # (the QoI comes from mlmcpi_sampler_qoi: the charge sum of the fill-in kernel, equal up to rounding)
            q_ref = ctx.qoi(m, mp.QOI_SCHWINGER_CHI, now).cpu()
            assert float((h_q - q_ref).abs().max()) <= 1e-9 * max(1.0, float(q_ref.abs().max()))
            changed += int((h_x != before).any(dim=1).sum())
        assert 0 < changed < 8 * B  # accepted and rejected draws both occurred
        if pinned:  # back-to-back calls without waiting in between (two alternating host buffers, as bench.py does)
            h2 = [h_x.clone().pin_memory(), h_x.clone().pin_memory()]
            q2 = [h_q.clone().pin_memory(), h_q.clone().pin_memory()]
            for d in range(6):
                s.draw_host_async(mp.QOI_SCHWINGER_CHI, q2[d & 1], h2[d & 1])
            s.wait_host()
            now = s.get_state().cpu()
            # the buffer of the last call holds every chain whose LAST draw was accepted; all rows are states the
            # chain has been in (rows of rejected draws keep an older state of that buffer)
            last = h2[5 & 1]
            q_ref = ctx.qoi(m, mp.QOI_SCHWINGER_CHI, s.get_state()).cpu()
            assert float((q2[5 & 1] - q_ref).abs().max()) <= 1e-9 * max(1.0, float(q_ref.abs().max()))
            same = (last == now).all(dim=1)
            assert same.any()
        s.close()
    ctx.set_option(mp._lib.OPT_HOST_COPY_ENGINE, 1)


def test_ho_exact_sampler(mp, ctx, orc):
    """HarmonicOscillatorAction::draw (Cholesky sampler): draw by draw against the oracle, and as the
    coarse sampler of a hierarchy (sampler = 'exact' of hierarchicalsampler.hh)"""
    o = po.ho(32, 4.0, 1.0, 1.0)
    m = to_mp(mp, o)
    B, chain0, draw = 5, 11, (1 << 33) + 4
    got = host(ctx.exact_draw(m, B, chain0, draw))
    want = np.array([orc.ho_exact_draw(o, SEED, draw, chain0 + b) for b in range(B)])
    close(got, want, tol=1e-11, what="exact draw")
    with pytest.raises(mp.MlmcpiError):
        ctx.exact_draw(mp.rotor(32), 2)
    # <x^2> from independent exact samples: no autocorrelation, analytic mean
    want_x2 = float.fromhex(load("scalars")["analytic"]["ho_x2_32"])
    for levels in (1, 2):
        Bc = 4096
        smp = mp.Sampler(ctx, mp.ho(32 * levels), Bc, kind=mp.SAMPLER_EXACT, n_levels=levels,
                         renorm=mp.RENORM_PERTURBATIVE)
        st = mp.Statistics(ctx, 10, Bc)
        mm = mp.ho(32 * levels)
        x = ctx.state(mm, Bc)
        for k in range(40):
            smp.draw(x)
            if k >= 10:
                st.record(ctx.qoi(mm, mp.QOI_X2, x))
        out = mp.Statistics.finalize(st.pack(), 10)
        ref = want_x2 if levels == 1 else mp._lib.lib.mlmcpi_ho_xsquared_analytical(1.0, 1.0, 4.0 / 64, 64, 0)
        assert abs(out["average"] - ref) < 5 * out["error"], (levels, out, ref)
        if levels == 1:
            assert out["tau_int"] < 1.1


def test_gff_dense_coarse_level(mp, ctx, orc):
    """GFF coarse level as the reference builds it (dense Q_hat, exact Cholesky sampler + 2 Gibbs
    sweeps): action against the numpy restatement of buildMatrices, second moments of the exact
    draws against Sigma_hat, and a hierarchical sampler with sampler = 'exact' on the coarsest
    level against gff_phi_squared_analytical (with the un-smoothed 5-point coarse action the
    two-level acceptance collapses beyond 16 x 16)"""
    m = mp.gff(16, 16, 3.0)
    mc = mp.coarse_model(m, ctype=mp.COARSEN_ROTATE)
    o = po.gff(16, 16, 3.0)
    oc = orc.coarse_model(o, 0, 0, po.ROTATE)
    mats = po.gff_dense_matrices(orc, oc, 2, 1.0)
    rng = np.random.default_rng(3)
    x = rng.normal(size=(4, mp.sample_size(mc)))
    want = 0.5 * np.einsum("bi,ij,bj->b", x, mats["Q_hat"], x)
    close(host(ctx.action(mc, dev(ctx, x))), want, tol=1e-9, what="dense action")
    B = 8192
    d = host(ctx.exact_draw(mc, B, 0, 1))
    C = d.T @ d / B
    tol = 6.0 * np.sqrt(2.0 / B) * np.max(np.diag(mats["Sigma_hat"]))
    assert np.max(np.abs(C - mats["Sigma_hat"])) < tol
    # fine level: plain Cholesky sample of the 5-point action
    d = host(ctx.exact_draw(m, B, 0, 2))
    Sigma = np.linalg.inv(po.gff_dense_matrices(orc, o, 0, 1.0)["Q"])
    assert np.max(np.abs(d.T @ d / B - Sigma)) < 6.0 * np.sqrt(2.0 / B) * np.max(np.diag(Sigma))
    # (two levels: with three or more the intermediate steps pair the dense Q_hat action of a level
    # with the 5-point conditional fill-in, gffconditionedfineaction.cc:7-25, and accept rarely)
    for M, levels in ((16, 2), (32, 2)):
        mm = mp.gff(M, M, 10.0)
        Bc = 1024
        smp = mp.Sampler(ctx, mm, Bc, kind=mp.SAMPLER_EXACT, n_levels=levels, ctype=mp.COARSEN_ROTATE)
        assert smp.level_model(1).gff_n_gibbs == 2
        st = mp.Statistics(ctx, 20, Bc)
        xx = ctx.init_state(mm, Bc, 0, 0)
        smp.set_state(xx)
        for k in range(120):
            smp.draw(xx)
            if k >= 40:
                st.record(ctx.qoi(mm, mp.QOI_PHI2, xx))
        out = mp.Statistics.finalize(st.pack(), 20)
        ref = mp._lib.lib.mlmcpi_gff_phi_squared_analytical(10.0, M, M)
        p = smp.p_accept()
        assert p[0] > 0.9, p  # (0.09 at 16^2 and 0 at 32^2 with the 5-point coarse action)
        assert abs(out["average"] - ref) < 5 * out["error"], (M, out, ref, p)


def test_per_dof_updates_reproduce_the_reference_lexicographic_sweep(mp, ctx):
    """Action::overrelaxation_update(state, ell) through mlmcpi_dof_update: one call per degree of freedom
    in the reference's lexicographic order reproduces the sweep recorded from the reference's own
    OverrelaxedHeatBathSampler loop (golden `overrelax_lex`); the heat-bath variant equals the coloured
    sweep's update of the same link on the same input (same variate, same arithmetic)."""
    for c in load("schwinger")[:2]:
        m = mp.schwinger(c["Mt"], c["Mx"], c["beta"], c["ctype"], 0)
        x = dev(ctx, np.tile(unhex(c["x"]), (3, 1)))
        for ell in range(x.shape[1]):
            ctx.dof_update(m, x, ell)
        ang_close(host(x)[0], unhex(c["overrelax_lex"]), tol=1e-11, what="lexicographic overrelaxation (Schwinger)")
        assert np.array_equal(host(x)[0], host(x)[2])
    for c in load("gff")[:1]:
        m = mp.gff(c["Mt"], c["Mx"], c["mass"], c["ctype"], 0)
        x = dev(ctx, unhex(c["x"]))
        for ell in range(x.shape[1]):
            ctx.dof_update(m, x, ell)
        close(host(x)[0], unhex(c["overrelax_lex"]), tol=1e-12, what="lexicographic overrelaxation (GFF)")
    # heat bath: link 0 belongs to colour 0, which the coloured sweep updates first, from the same
    # neighbours, with the variate of (chain, draw, link)
    c = load("schwinger")[0]
    m = mp.schwinger(c["Mt"], c["Mx"], c["beta"], c["ctype"], 0)
    a, b = dev(ctx, np.tile(unhex(c["x"]), (4, 1))), dev(ctx, np.tile(unhex(c["x"]), (4, 1)))
    ctx.dof_update(m, a, 0, heatbath=True, chain0=3, draw=11)
    ctx.heatbath_sweep(m, b, 3, 11)
    assert np.array_equal(host(a)[:, 0], host(b)[:, 0])
    assert len(set(host(a)[:, 0])) == 4  # one variate stream per chain


def test_schwinger_force_elementwise(mp, ctx, orc):
    """the force component by component (the norm-wise bound of close() says nothing about a component whose two
    sines nearly cancel): |dS/dtheta (CUDA) - oracle| <= 8 eps beta for EVERY link -- the branch-free sine of the
    leapfrog kernels is accurate to 2.2e-16 absolute, a component is beta (sin P - sin P'), the rest is the rounding
    of the plaquette sums -- on random states, on nearly flat states (all sines small: cancellation between
    neighbouring plaquettes) and at plaquette angles near +-pi, for beta = 1 ... 4096; the same for the positions
    after a short leapfrog trajectory (errors scale with dt^2 beta)"""
    rng = np.random.default_rng(23)
    eps = np.finfo(np.float64).eps
    for Mt, Mx, beta in [(16, 16, 1.0), (32, 16, 64.0), (64, 64, 1024.0), (128, 128, 4096.0)]:
        o = po.schwinger(Mt, Mx, beta)
        m = mp.schwinger(Mt, Mx, beta)
        n = 2 * Mt * Mx
        flat = 1e-3 * rng.normal(size=n)
        pis = rng.choice([-np.pi, np.pi], size=n) * (1 + 1e-9 * rng.normal(size=n))
        x = np.stack([rng.uniform(-np.pi, np.pi, n), flat, 0.25 * pis, rng.uniform(-30, 30, n)])
        got = host(ctx.force(m, dev(ctx, x)))
        want = np.array([orc.force(o, x[b]) for b in range(x.shape[0])])
        err = np.abs(got - want)
        assert err.max() <= 8 * eps * beta * max(1.0, np.abs(x).max() / np.pi), (Mt, beta, err.max() / (eps * beta))
        p = rng.normal(size=x.shape)
        y, pd = dev(ctx, x), dev(ctx, p)
        ctx.leapfrog(m, 4, 0.01, y, pd)
        res = [orc.leapfrog(o, 4, 0.01, x[b], p[b]) for b in range(x.shape[0])]
        ex = np.abs(host(y) - np.array([r[0] for r in res])).max()
        ep = np.abs(host(pd) - np.array([r[1] for r in res])).max()
        assert ex <= 64 * eps * max(np.abs(x).max(), 1e-2 * beta), (Mt, beta, ex)
        assert ep <= 64 * eps * max(1.0, 0.04 * beta) * max(1.0, np.abs(x).max() / np.pi), (Mt, beta, ep)


def test_schwinger_heatbath_paired_variates(mp, ctx, orc):
    """heatbath_pair_kernel (two links of a colour per Philox block) against the oracle's restatement of the same
    variate map, on lattices where a row holds an ODD number of links of a colour (Mt / 2 odd: the last link has no
    partner), for small and large beta (retries on the links' own streams); and the single-link entry point
    mlmcpi_dof_update(heatbath) for both roles: every link of colour 0 -- which the sweep updates first, from the
    original neighbours -- gets the sweep's value"""
    rng = np.random.default_rng(5)
    orc.lib.orc_set_expcos_envelope(2)  # the product's default envelope (the oracle's default is the reference's)
    for Mt, Mx, beta in [(6, 4, 2.0), (10, 6, 40.0), (12, 8, 900.0), (34, 6, 3.0)]:
        o = po.schwinger(Mt, Mx, beta)
        m = mp.schwinger(Mt, Mx, beta)
        B, chain0, draw = 3, 5, (2 << 32) + 7
        x = rng.uniform(-np.pi, np.pi, (B, 2 * Mt * Mx))
        xd = dev(ctx, x)
        ctx.heatbath_sweep(m, xd, chain0, draw)
        want = np.array([orc.heatbath_sweep(o, SEED, draw, chain0 + b, x[b]) for b in range(B)])
        ang_close(host(xd), want, tol=1e-9, what=f"paired heat bath {Mt}x{Mx} beta {beta}")
        got = host(xd)
        for j in range(0, Mx, 2):          # colour 0: mu = 0, even rows
            for i in range(Mt):
                ell = 2 * (Mt * j + i)
                a = dev(ctx, x)
                ctx.dof_update(m, a, ell, heatbath=True, chain0=chain0, draw=draw)
                assert np.array_equal(host(a)[:, ell], got[:, ell]), (Mt, Mx, i, j)
    orc.lib.orc_set_expcos_envelope(0)


def test_gff_dense_action_2048_vertices_against_reference(mp, ctx):
    """the dense coarse action built ON THE DEVICE (cuSOLVER / cuBLAS, csrc/gff.cu) and evaluated with
    one DGEMM over all chains, at the largest size the reference's own buildMatrices finishes here in
    minutes: 64 x 64 fine lattice, rotated coarse level of 2048 vertices; values recorded from the
    reference's GFFAction::evaluate (tools/make_golden.py --gff-dense, tests/golden/gff_dense.json)"""
    from tools.make_golden import noncompact
    for c in load("gff_dense"):
        m = mp.gff(c["Mt"], c["Mt"], c["mass"])
        mc = mp.coarse_model(m, ctype=mp.COARSEN_ROTATE)
        assert mp.sample_size(mc) == c["n_coarse"] and mc.gff_n_gibbs == 2
        xf = np.array([noncompact(mp.sample_size(m), sh) for sh in c["shifts"]])
        xc = ctx.state(mc, len(c["shifts"]))
        ctx.restrict(m, dev(ctx, xf), xc)
        want = [scalar(v) for v in c["coarse_S_gibbs"]]
        close(host(ctx.action(mc, xc)), want, tol=1e-9, what="dense Q_hat action, 2048 vertices")


# --------------------------------------------------- statistics accumulators


def test_statistics_against_oracle(mp, ctx, orc):
    rng = np.random.default_rng(5)
    B, n, k_max = 6, 300, 10
    q = np.zeros((n, B))
    v = rng.normal(size=B)
    for k in range(n):
        v = 0.7 * v + rng.normal(size=B)
        q[k] = v
    st = mp.Statistics(ctx, k_max, B)
    for k in range(n):
        st.record(dev(ctx, q[k])[0])
    packed = st.pack()
    assert packed[0] == B and packed[1] == B * n and packed[2] == B * n
    # single-chain restriction: chain b alone reproduces Statistics exactly
    for b in range(B):
        st1 = mp.Statistics(ctx, k_max, 1)
        for k in range(n):
            st1.record(dev(ctx, q[k, b:b + 1])[0])
        out = mp.Statistics.finalize(st1.pack(), k_max)
        want = orc.statistics(k_max, q[:, b])
        got = [out["average"], out["variance"], out["variance_error"], out["tau_int"], out["error"],
               out["samples"]]
        assert np.allclose(got, want, rtol=1e-11)
    # all chains: the "average over ranks" of statistics.cc:30-35,64-79
    out = mp.Statistics.finalize(packed, k_max)
    per = np.array([orc.statistics(k_max, q[:, b]) for b in range(B)])
    assert abs(out["average"] - per[:, 0].mean()) < 1e-12
    assert out["samples"] == B * n


# ------------------------------------------- properties at the BASELINE sizes


def test_schwinger_512_properties(mp, ctx):
    m = mp.schwinger(512, 512, 4.0)
    B = 4
    x = ctx.init_state(m, B, 0, 1)
    S0 = host(ctx.action(m, x))
    # gauge invariance: the force sums to zero (SURVEY 8c pin 3)
    F = ctx.force(m, x)
    assert float(F.sum(dim=1).abs().max()) < 1e-8
    # overrelaxation leaves the action unchanged
    y = x.clone()
    ctx.overrelax_sweep(m, y)
    assert np.max(np.abs(host(ctx.action(m, y)) - S0)) <= 1e-11 * np.max(S0)
    assert float((y - x).abs().max()) > 1e-3
    # integer topological charge, consistent with the double-valued QoI
    chi, Q = ctx.qoi(m, mp.QOI_SCHWINGER_CHI, x, with_charge=True)
    assert np.allclose(host(chi), host(Q).astype(float) ** 2, rtol=0, atol=1e-6)
    # restrict o prolong = identity on the coarse links (mod 2 pi)
    mc = mp.coarse_model(m)
    xc = ctx.init_state(mc, B, 0, 2)
    ctx.prolong(m, xc, y)
    xc2 = ctx.state(mc, B)
    ctx.restrict(m, y, xc2)
    ang_close(host(xc2), host(xc), tol=1e-14, what="restrict o prolong")
    # the fill-in keeps the coarse links: restrict(fill(prolong(xc))) = xc
    ctx.prolong_fill(m, xc, y, 0, 3)
    ctx.restrict(m, y, xc2)
    ang_close(host(xc2), host(xc), tol=1e-12, what="restrict o fill o prolong")
    # leapfrog: reversible and (nearly) energy conserving
    p = ctx.hmc_momentum(m, B, 0, 4)
    H0 = host(ctx.action(m, x) + 0.5 * (p * p).sum(dim=1))
    y, q = x.clone(), p.clone()
    ctx.leapfrog(m, 20, 0.01, y, q)
    H1 = host(ctx.action(m, y) + 0.5 * (q * q).sum(dim=1))
    # second-order symplectic integrator: the (extensive) energy error is O(dt^2)
    assert 0 < np.max(np.abs(H1 - H0)) < 1e-4 * np.max(np.abs(H0))
    y2, q2 = x.clone(), p.clone()
    ctx.leapfrog(m, 40, 0.005, y2, q2)
    H2 = host(ctx.action(m, y2) + 0.5 * (q2 * q2).sum(dim=1))
    ratio = np.abs(H1 - H0) / np.abs(H2 - H0)
    assert np.all((ratio > 3.0) & (ratio < 5.0)), ratio
    q.neg_()
    ctx.leapfrog(m, 20, 0.01, y, q)
    assert float((y - x).abs().max()) < 1e-9


def test_schwinger_1024_and_gff_256_properties(mp, ctx):
    """size-independent properties at the remaining BASELINE sizes: Schwinger 1024 x 1024 (config 4)
    at the continuum-limit coupling, GFF 256 x 256 with coarsening rotate (config 2)"""
    m = mp.schwinger(1024, 1024, 4096.0)
    B = 2
    x = ctx.init_state(m, B, 0, 1)
    for k in range(2):
        ctx.heatbath_sweep(m, x, 0, k)
    S0 = host(ctx.action(m, x))
    assert float(ctx.force(m, x).sum(dim=1).abs().max()) < 1e-6
    y = x.clone()
    ctx.overrelax_sweeps(m, y, 3)  # one-pass kernel, 1024 threads per block
    assert np.max(np.abs(host(ctx.action(m, y)) - S0)) <= 1e-11 * np.max(S0)
    chi, Q = ctx.qoi(m, mp.QOI_SCHWINGER_CHI, x, with_charge=True)
    assert np.allclose(host(chi), host(Q).astype(float) ** 2, rtol=0, atol=1e-5)
    mc = mp.coarse_model(m, renorm=mp.RENORM_PERTURBATIVE)
    xc = ctx.state(mc, B)
    ctx.restrict(m, x, xc)
    xc2 = ctx.state(mc, B)
    Sf, Sc = ctx.prolong_fill_eval(m, xc, y, 0, 3)  # ApproximateBesselProduct path, tau ~ 8000
    ctx.restrict(m, y, xc2)
    ang_close(host(xc2), host(xc), tol=1e-12, what="restrict o fill o prolong")
    close(host(Sf), host(ctx.action(m, y)), tol=1e-11, what="fused S_f at 1024^2")
    close(host(Sc), host(ctx.cond_action(m, y)), tol=1e-10, what="fused S_cond at 1024^2")
    # GFF 256^2: overrelaxation conserves the action, restrict o prolong = id on both level types,
    # the heat bath reaches the analytic <phi^2>
    g = mp.gff(256, 256, 10.0)
    Bg = 64
    phi = ctx.init_state(g, Bg, 0, 0)
    S0 = host(ctx.action(g, phi))
    y = phi.clone()
    ctx.overrelax_sweep(g, y)
    assert np.max(np.abs(host(ctx.action(g, y)) - S0)) <= 1e-11 * np.max(S0)
    gl = g
    for level in range(3):
        gc = mp.coarse_model(gl, level=level, ctype=mp.COARSEN_ROTATE)
        xc = ctx.init_state(gc, 2, 0, 5 + level)
        fine = ctx.state(gl, 2)
        ctx.prolong(gl, xc, fine)
        back = ctx.state(gc, 2)
        ctx.restrict(gl, fine, back)
        assert bool((back == xc).all()), level
        gl = gc
        gl.gff_n_gibbs = 0


def test_rowmarch_equals_generic_leapfrog(mp, ctx):
    """the tiled hot kernel (Mt % 32 == 0) and the generic fallback agree"""
    rng = np.random.default_rng(3)
    orc = po.oracle()
    for Mt, Mx in [(32, 8), (64, 48), (96, 33)]:
        o = po.schwinger(Mt, Mx, 3.0)
        m = to_mp(mp, o)
        x, p = rng.uniform(-3, 3, (2, 2 * Mt * Mx)), rng.normal(size=(2, 2 * Mt * Mx))
        xd, pd = dev(ctx, x), dev(ctx, p)
        ctx.leapfrog(m, 4, 0.05, xd, pd)
        want = [orc.leapfrog(o, 4, 0.05, x[b], p[b]) for b in range(2)]
        close(host(xd), [w[0] for w in want], what=f"x {Mt}x{Mx}")
        close(host(pd), [w[1] for w in want], what=f"p {Mt}x{Mx}")


def test_edge_shapes(mp, ctx, orc):
    """smallest lattices, a single chain, ragged extents, widths beyond one thread block, the
    1024 x 1024 lattice of BASELINE config 4, and 1-D paths on both sides of the register-resident
    HMC kernel (M = 32 k) -- deterministic kernels against the oracle / against each other"""
    rng = np.random.default_rng(11)
    for Mt, Mx, B in [(2, 2, 1), (4, 2, 3), (2, 6, 2), (1056, 4, 1), (2048, 3, 1)]:
        o = po.schwinger(Mt, Mx, 2.5)
        m = to_mp(mp, o)
        x, p = rng.uniform(-3, 3, (B, 2 * Mt * Mx)), rng.normal(size=(B, 2 * Mt * Mx))
        xd, pd = dev(ctx, x), dev(ctx, p)
        close(host(ctx.action(m, xd)), [orc.action(o, x[b]) for b in range(B)], what=f"S {Mt}x{Mx}")
        close(host(ctx.force(m, xd)), [orc.force(o, x[b]) for b in range(B)], what=f"F {Mt}x{Mx}")
        ctx.leapfrog(m, 3, 0.05, xd, pd)
        want = [orc.leapfrog(o, 3, 0.05, x[b], p[b]) for b in range(B)]
        close(host(xd), [w[0] for w in want], what=f"x {Mt}x{Mx}")
        close(host(pd), [w[1] for w in want], what=f"p {Mt}x{Mx}")
        got, Q = ctx.qoi(m, mp.QOI_SCHWINGER_CHI, dev(ctx, x), with_charge=True)
        assert list(host(Q)) == [orc.qoi(o, po.QOI_SCHWINGER_CHI, x[b])[1] for b in range(B)]
    # 1024 x 1024: the three leapfrog kernels (two-step TMA pipeline, one-step pipeline, generic)
    m = mp.schwinger(1024, 1024, 4.0)
    x0 = ctx.init_state(m, 2, 0, 1)
    p0 = ctx.hmc_momentum(m, 2, 0, 1)
    res = []
    for variant, fuse in ((0, 1), (0, 0), (2, 0)):
        ctx.set_option(mp._lib.OPT_LEAPFROG_VARIANT, variant)
        ctx.set_option(mp._lib.OPT_LEAPFROG_FUSE, fuse)
        x, p = x0.clone(), p0.clone()
        ctx.leapfrog(m, 5, 0.02, x, p)
        res.append((host(x).copy(), host(p).copy()))
    ctx.set_option(mp._lib.OPT_LEAPFROG_VARIANT, 0)
    ctx.set_option(mp._lib.OPT_LEAPFROG_FUSE, 1)
    for x, p in res[1:]:
        assert np.array_equal(x, res[0][0]) and np.array_equal(p, res[0][1])
    # 1-D paths: M = 32 k uses the register-resident HMC kernel, M = 33 / 40 the generic one
    for M in (2, 3, 32, 33, 40, 64, 128, 256, 288):
        for o in (po.rotor(M, 4.0, 0.25), po.ho(M), po.quartic(M)):
            mm = to_mp(mp, o)
            x = rng.uniform(-2, 2, (3, M))
            xd = dev(ctx, x)
            acc, diag = ctx.hmc_step(mm, 7, 0.04, xd, 5, 21)
            want = [orc.hmc_step(o, 7, 0.04, SEED, 21, 5 + b, x[b]) for b in range(3)]
            assert list(host(acc)) == [w[0] for w in want], (M, o.model)
            close(host(xd), [w[1] for w in want], tol=1e-11, what=f"hmc {M}")
            wd = np.array([w[2] for w in want])
            assert np.max(np.abs(host(diag) - wd)) <= 1e-10 * max(np.max(np.abs(wd[:, 1:])), 1.0)


def test_leapfrog_k_stage_pipeline_is_bit_identical(mp, ctx, orc):
    """leapfrog_rowpipek_kernel (2 / 4 leapfrog steps per HBM pass, compile-time block size) against the
    one-step pipeline and the round-1 two-step kernel: identical bits, for every specialised Mt, ragged
    last chunks (Mx not a multiple of the 32 rows per block), chunks shorter than the pipeline depth,
    trajectories whose length leaves 1, 2 or 3 steps for the fall-back kernels, and against the oracle"""
    rng = np.random.default_rng(17)
    for Mt, Mx, B, nt in [(64, 64, 3, 9), (128, 128, 2, 12), (128, 40, 2, 7), (256, 34, 2, 6), (512, 12, 1, 5),
                          (128, 10, 2, 4), (64, 200, 2, 11), (128, 96, 2, 19)]:
        m = mp.schwinger(Mt, Mx, 5.0)
        x0 = dev(ctx, rng.uniform(-3, 3, (B, 2 * Mt * Mx)))
        p0 = dev(ctx, rng.normal(size=(B, 2 * Mt * Mx)))
        res = {}
        for fuse in (0, 2, 3, 4, 5, 1):     # 5 = eight steps per pass (Mt <= 128, else four)
            ctx.set_option(mp._lib.OPT_LEAPFROG_FUSE, fuse)
            x, p = x0.clone(), p0.clone()
            ctx.leapfrog(m, nt, 0.03, x, p)
            res[fuse] = (host(x).copy(), host(p).copy())
        ctx.set_option(mp._lib.OPT_LEAPFROG_FUSE, 1)
        for fuse in (2, 3, 4, 5, 1):
            assert np.array_equal(res[fuse][0], res[0][0]), (Mt, Mx, nt, fuse, "theta")
            assert np.array_equal(res[fuse][1], res[0][1]), (Mt, Mx, nt, fuse, "p")
        o = po.schwinger(Mt, Mx, 5.0)
        xo, po_ = orc.leapfrog(o, nt, 0.03, host(x0)[0], host(p0)[0])
        close(res[1][0][0], xo, tol=1e-11, what=f"theta vs oracle {Mt}x{Mx}")
        close(res[1][1][0], po_, tol=1e-11, what=f"p vs oracle {Mt}x{Mx}")
    # rows-per-block option: 8 rows (the halo is then as large as the chunk) and 64
    m = mp.schwinger(128, 128, 5.0)
    x0, p0 = ctx.init_state(m, 2, 0, 1), ctx.hmc_momentum(m, 2, 0, 1)
    out = []
    for rows in (0, 8, 64):
        ctx.set_option(mp._lib.OPT_LEAPFROG_ROWS, rows)
        x, p = x0.clone(), p0.clone()
        ctx.leapfrog(m, 8, 0.05, x, p)
        out.append((host(x).copy(), host(p).copy()))
    ctx.set_option(mp._lib.OPT_LEAPFROG_ROWS, 0)
    for x, p in out[1:]:
        assert np.array_equal(x, out[0][0]) and np.array_equal(p, out[0][1])


def test_overrelax_one_pass_equals_colour_passes(mp, ctx):
    """the one-pass row-pipelined overrelaxation sweep (all four colours, out of place) gives the
    bits of the four colour passes, for chunked and wrapped lattices and several sweeps in a row"""
    for Mt, Mx, B in [(2, 2, 2), (4, 2, 1), (6, 4, 3), (32, 8, 2), (64, 96, 2), (96, 34, 2), (512, 512, 2),
                      (1024, 64, 1)]:
        m = mp.schwinger(Mt, Mx, 2.0)
        x0 = ctx.init_state(m, B, 0, 3)
        ref = x0.clone()
        ctx.set_option(mp._lib.OPT_OVERRELAX_ONE_PASS, 0)
        for _ in range(3):
            ctx.overrelax_sweep(m, ref)
        ctx.set_option(mp._lib.OPT_OVERRELAX_ONE_PASS, 1)
        got = x0.clone()
        for _ in range(3):
            ctx.overrelax_sweep(m, got)
        assert bool((got == ref).all()), (Mt, Mx)
    # inside the sampler: 3 overrelaxation sweeps + 1 heat bath per draw, both settings
    m = mp.schwinger(64, 64, 3.0)
    outs = []
    for one_pass in (0, 1):
        ctx.set_option(mp._lib.OPT_OVERRELAX_ONE_PASS, one_pass)
        smp = mp.Sampler(ctx, m, 4, kind=mp.SAMPLER_HEATBATH, n_sweep_overrelax=3, n_sweep_heatbath=1)
        x = ctx.init_state(m, 4, 0, 9)
        smp.set_state(x)
        for _ in range(3):
            smp.draw(x)
        outs.append(host(x).copy())
    ctx.set_option(mp._lib.OPT_OVERRELAX_ONE_PASS, 1)
    ang_close(outs[0], outs[1], tol=1e-12, what="sampler with one-pass overrelaxation")


def test_gff_one_pass_sweeps_equal_colour_passes(mp, ctx):
    """the one-pass row-pipelined GFF sweep (both colours, two columns per thread, out of place) gives the bits
    of the two colour passes: overrelaxation (odd and even numbers of sweeps: with and without the copy back),
    and overrelaxation + heat-bath sequences inside the sampler, for chunked, wrapped and narrow lattices"""
    for Mt, Mx, B in [(64, 4, 3), (64, 6, 5), (128, 64, 4), (256, 256, 3), (96, 34, 2), (2048, 8, 1), (32, 32, 2)]:
        m = mp.gff(Mt, Mx, 10.0)
        x0 = ctx.init_state(m, B, 0, 3)
        for n in (3, 4):
            outs = []
            for one_pass in (0, 1):
                ctx.set_option(mp._lib.OPT_OVERRELAX_ONE_PASS, one_pass)
                x = x0.clone()
                ctx.overrelax_sweeps(m, x, n)
                outs.append(x)
            assert bool((outs[0] == outs[1]).all()), (Mt, Mx, n)
            assert not bool((outs[0] == x0).all())
        outs = []
        for one_pass in (0, 1):
            ctx.set_option(mp._lib.OPT_OVERRELAX_ONE_PASS, one_pass)
            smp = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_HEATBATH, n_sweep_overrelax=2, n_sweep_heatbath=2)
            x = x0.clone()
            smp.set_state(x)
            for _ in range(3):
                smp.draw(x)
            outs.append(x.clone())
            smp.close()
        ctx.set_option(mp._lib.OPT_OVERRELAX_ONE_PASS, 1)
        assert bool((outs[0] == outs[1]).all()), (Mt, Mx, "sampler")


def test_fused_qm_hierarchy_equals_kernel_sequence(mp, ctx):
    """HierarchicalSampler::draw for 1-D paths as ONE kernel (one warp per chain, all levels on chip)
    against the sequence of single-purpose kernels: same states, same acceptance counters"""
    cases = [(mp.rotor(64, 4.0, 0.25), 2), (mp.rotor(256, 4.0, 0.25), 3), (mp.rotor(128, 4.0, 0.25), 2),
             (mp.ho(128), 3), (mp.ho(64), 2), (mp.quartic(256), 4), (mp.quartic(64), 2)]
    for m, L in cases:
        B = 37
        res = []
        for fused in (0, 1):
            ctx.set_option(mp._lib.OPT_FUSED_QM_HIERARCHY, fused)
            smp = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_HMC, n_levels=L, nt=12, dt=0.08,
                             renorm=mp.RENORM_PERTURBATIVE, chain0=5)
            x = ctx.init_state(m, B, 5, 2) if m.model == mp.ROTOR else ctx.exact_draw(mp.ho(m.M_lat), B, 5, 2)
            smp.set_state(x)
            launches0 = ctx.launches
            for _ in range(6):
                smp.draw(x)
            res.append((host(x).copy(), smp.p_accept(), ctx.launches - launches0))
        ctx.set_option(mp._lib.OPT_FUSED_QM_HIERARCHY, 1)
        # (bit-identical for the rotor; the compiler contracts a few multiply-adds of the Gaussian
        # fill-in differently inside the fused kernel: last-bit differences for the oscillators)
        diff = np.max(np.abs(res[0][0] - res[1][0]))
        assert diff <= (0.0 if m.model == mp.ROTOR else 1e-12), (m.model, m.M_lat, L, diff)
        assert res[0][1] == res[1][1]
        assert res[1][2] < res[0][2] / 3  # one kernel (+ the masked copy) per draw
        assert 0.0 < res[1][1][0] <= 1.0 and min(res[1][1]) < 1.0  # (per-level rates are conditional ones)


def test_rotor_c2_properties(mp, ctx):
    """C2 shape: M_lat = 256, 8192 chains"""
    m = mp.rotor(256, 4.0, 0.25)
    B = 8192
    x = ctx.init_state(m, B, 0, 0)
    S0 = ctx.action(m, x)
    y = x.clone()
    ctx.overrelax_sweep(m, y)
    assert float((ctx.action(m, y) - S0).abs().max()) < 1e-10 * float(S0.abs().max())
    chi, Q = ctx.qoi(m, mp.QOI_ROTOR_CHI, x, with_charge=True)
    assert np.allclose(host(chi) * m.T_final, host(Q).astype(float) ** 2, atol=1e-8)
    mc = mp.coarse_model(m)
    xc = ctx.state(mc, B)
    ctx.restrict(m, x, xc)
    assert bool((xc == x[:, ::2]).all())


# ------------------------------------------ statistical parity with analytics


def _mean_err(v):
    v = np.asarray(v)
    return v.mean(), v.std(ddof=1) / np.sqrt(len(v))


def test_ho_hmc_matches_analytic_x2(mp, ctx):
    """C1: harmonic oscillator, HMC, template parameters; <x^2> against
    HarmonicOscillatorAction::Xsquared_analytical (recorded from the reference)"""
    m = mp.ho(32, 4.0, 1.0, 1.0)
    B = 4096
    # trajectory length 73 * 0.05: no lattice mode is close to a resonance cos(omega t) = +-1
    # (with the template's t = 10 one mode has |cos| = 0.996 and needs ~1000 burn-in draws)
    s = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_HMC, nt=73, dt=0.05)
    x = ctx.state(m, B)
    for _ in range(100):
        s.draw(x)
    vals = []
    for _ in range(10):
        s.draw(x)
        vals.append(host(ctx.qoi(m, mp.QOI_X2, x)))
    per_chain = np.mean(vals, axis=0)
    mean, err = _mean_err(per_chain)
    want = float.fromhex(load("scalars")["analytic"]["ho_x2_32"])
    assert abs(mean - want) < 5 * err + 1e-4, (mean, err, want)
    assert 0.5 < s.p_accept()[0] <= 1.0


def test_rotor_hierarchical_matches_exact_chit(mp, ctx):
    """rotor M=32, hierarchical sampler (3 levels, HMC on the coarsest level):
    susceptibility against RotorAction::chit_exact"""
    m = mp.rotor(32, 4.0, 0.25)
    B = 8192
    s = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_HMC, n_levels=3, nt=20, dt=0.1,
                   renorm=mp.RENORM_PERTURBATIVE)
    x = ctx.init_state(m, B, 0, 0)
    s.set_state(x)
    for _ in range(40):
        s.draw(x)
    vals = []
    for _ in range(20):
        s.draw(x)
        vals.append(host(ctx.qoi(m, mp.QOI_ROTOR_CHI, x)))
    mean, err = _mean_err(np.mean(vals, axis=0))
    want = float.fromhex(load("scalars")["analytic"]["rotor_chit_exact_32"])
    assert abs(mean - want) < 5 * err, (mean, err, want)
    pa = s.p_accept()
    assert all(0.05 < p <= 1.0 for p in pa), pa


def test_schwinger_samplers_match_analytic_chit(mp, ctx):
    """quenched Schwinger 8x8, beta = 4 (P = 64 plaquettes): heat bath + overrelaxation
    and the hierarchical sampler against quenchedschwinger_chit_analytical"""
    want = float.fromhex(load("scalars")["analytic"]["schwinger_chit_analytical_4_64"])
    m = mp.schwinger(8, 8, 4.0)
    B = 4096
    for kw in (dict(kind=mp.SAMPLER_HEATBATH, n_sweep_overrelax=2, n_sweep_heatbath=1),
               dict(kind=mp.SAMPLER_HMC, n_levels=2, nt=20, dt=0.1)):
        s = mp.Sampler(ctx, m, B, **kw)
        x = ctx.init_state(m, B, 0, 0)
        s.set_state(x)
        for _ in range(60):
            s.draw(x)
        vals = []
        for _ in range(30):
            s.draw(x)
            vals.append(host(ctx.qoi(m, mp.QOI_SCHWINGER_CHI, x)))
        mean, err = _mean_err(np.mean(vals, axis=0))
        assert abs(mean - want) < 5 * err, (kw, mean, err, want)
        s.close()


@pytest.mark.parametrize("beta", [9.0, 16.0])
def test_schwinger_hierarchical_approx_path_matches_analytic_chit(mp, ctx, beta):
    """beta > 8 selects the ApproximateBesselProduct fill-in (quenchedschwingerconditionedfineaction.cc:45-48,
    250-285) and, here, the fused by-product evaluation of S_f / S_cond: two-level cascade with a heat-bath
    coarse sampler on 8x8, started from a thermalised state, against quenchedschwinger_chit_analytical.
    (From a HOT start the reference's own cascade leaves a fraction of the chains frozen in |Q| > 0 sectors --
    scratch/ref_phys_hot.py -- and so does this library; the reference starts cold.)"""
    L, B = 8, 4096
    m = mp.schwinger(L, L, beta)
    want = mp._lib.lib.mlmcpi_schwinger_chit_analytical(beta, L * L)
    x = ctx.init_state(m, B, 0, 0)
    for k in range(50):
        ctx.heatbath_sweep(m, x, 0, k)
    s = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_HEATBATH, n_levels=2, renorm=mp.RENORM_PERTURBATIVE,
                   n_sweep_overrelax=2, n_sweep_heatbath=1)
    s.set_state(x)
    for _ in range(100):
        s.draw(x)
    vals = []
    for _ in range(300):
        s.draw(x)
        vals.append(host(ctx.qoi(m, mp.QOI_SCHWINGER_CHI, x)))
    mean, err = _mean_err(np.mean(vals, axis=0))
    assert abs(mean - want) < 5 * err, (beta, mean, err, want)
    pa = s.p_accept()
    # the reference's own cascade: 0.753 (beta = 9) and 0.892 (beta = 16), scratch/ref_phys.py
    assert abs(pa[0] - (0.753 if beta == 9.0 else 0.89)) < 0.02, pa
    s.close()


def test_gff_heatbath_matches_analytic_phi2(mp, ctx):
    want = float.fromhex(load("scalars")["analytic"]["gff_phi_squared_10_16"])
    m = mp.gff(16, 16, 10.0)
    B = 2048
    s = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_HEATBATH, n_sweep_overrelax=1, n_sweep_heatbath=1)
    x = ctx.init_state(m, B, 0, 0)
    s.set_state(x)
    for _ in range(50):
        s.draw(x)
    vals = []
    for _ in range(20):
        s.draw(x)
        vals.append(host(ctx.qoi(m, mp.QOI_PHI2, x)))
    mean, err = _mean_err(np.mean(vals, axis=0))
    assert abs(mean - want) < 5 * err, (mean, err, want)


def test_host_entry_point(mp, ctx):
    """mlmcpi_sampler_draw_host: host buffers in, QoI and state out"""
    m = mp.schwinger(32, 32, 4.0)
    B = 8
    s = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_HMC, n_levels=2, nt=10, dt=0.05)
    x_in = host(ctx.init_state(m, B, 0, 5))
    q = np.zeros(B)
    x_out = np.zeros_like(x_in)
    s.draw_host(x_in, mp.QOI_SCHWINGER_CHI, q, x_out)
    orc = po.oracle()
    o = po.schwinger(32, 32, 4.0)
    for b in range(B):
        assert abs(q[b] - orc.qoi(o, po.QOI_SCHWINGER_CHI, x_out[b])[0]) < 1e-9
    w = s.work()
    assert w["leapfrog_site_steps"] == B * 11 * 16 * 16 and w["filled_fine_sites"] == B * 32 * 32
    # the host entry point issues the draw range by range (upload overlapped with compute): same
    # draws as the device entry point, for the 2-D cascade and for the fused 1-D cascade
    for mm, Bc, kw in ((mp.schwinger(64, 64, 6.0), 48, dict(n_levels=2, nt=8, dt=0.05)),
                       (mp.rotor(256, 4.0, 0.25), 16384, dict(n_levels=3, nt=10, dt=0.1))):
        xs = ctx.init_state(mm, Bc, 0, 5)
        for k in range(3):
            ctx.heatbath_sweep(mm, xs, 0, k)
        a = mp.Sampler(ctx, mm, Bc, kind=mp.SAMPLER_HMC, renorm=mp.RENORM_PERTURBATIVE, **kw)
        b = mp.Sampler(ctx, mm, Bc, kind=mp.SAMPLER_HMC, renorm=mp.RENORM_PERTURBATIVE, **kw)
        xa = xs.clone()
        a.set_state(xa)
        h_in, h_out, hq = host(xs).copy(), np.zeros((Bc, mp.sample_size(mm))), np.zeros(Bc)
        for d in range(3):
            a.draw(xa)
            b.draw_host(h_in, mp.QOI_X2 if mm.model == mp.ROTOR else mp.QOI_AVG_PLAQUETTE, hq, h_out)
            h_in = h_out.copy()
            ang_close(h_out, host(xa), tol=1e-12, what=f"draw_host vs draw, step {d}")
        assert a.p_accept() == b.p_accept()


def test_error_paths(mp, ctx):
    with pytest.raises(mp.MlmcpiError):
        ctx.overrelax_sweep(mp.ho(16), ctx.state(mp.ho(16), 1))       # not defined for HO
    with pytest.raises(mp.MlmcpiError):
        ctx.fill(mp.schwinger(7, 8, 1.0), ctx.state(mp.schwinger(7, 8, 1.0), 1))  # odd extent
    with pytest.raises(mp.MlmcpiError):
        ctx.qoi(mp.rotor(16), mp.QOI_PHI2, ctx.state(mp.rotor(16), 1))
    with pytest.raises(mp.MlmcpiError):
        m = mp.gff(8, 8, 1.0, mp.COARSEN_BOTH)
        ctx.fill(m, ctx.state(m, 1))                                   # GFF fill needs rotate


# ------------------------------------ multilevel sampler and multilevel Monte Carlo


def test_statistics_reset_semantics(mp, ctx, orc):
    """Statistics::reset clears the short-term mean and count, not the long-term moments"""
    rng = np.random.default_rng(11)
    B, k_max = 4, 5
    q = rng.normal(size=(60, B))
    st = mp.Statistics(ctx, k_max, B)
    for k in range(20):
        st.record(dev(ctx, q[k])[0])
    st.reset()
    for k in range(20, 60):
        st.record(dev(ctx, q[k])[0])
    out = mp.Statistics.finalize(st.pack(), k_max)
    assert out["samples"] == 40 * B
    assert abs(out["average"] - q[20:].mean()) < 1e-12
    per = np.array([orc.statistics(k_max, q[:, b]) for b in range(B)])
    # variance / tau_int use all 60 samples (long-term), like the reference
    S0 = np.mean([np.mean(q[:, b] ** 2) for b in range(B)])
    assert abs(out["variance"] - (60 * B) / (60 * B - 1.0) * (S0 - q.mean() ** 2)) < 1e-12
    assert per.shape == (B, 6)


def test_rotor_multilevel_sampler_matches_exact_chit(mp, ctx):
    """MultilevelSampler level walk (sampler/multilevelsampler.cc:71-112), rotor M = 32"""
    m = mp.rotor(32, 4.0, 0.25)
    B = 4096
    s = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_HMC, n_levels=3, nt=20, dt=0.1,
                   renorm=mp.RENORM_PERTURBATIVE, multilevel=True, qoi=mp.QOI_ROTOR_CHI,
                   n_autocorr_window=10)
    x = ctx.state(m, B)
    for _ in range(30):
        s.draw(x)
    vals = []
    for _ in range(20):
        s.draw(x)
        vals.append(host(ctx.qoi(m, mp.QOI_ROTOR_CHI, x)))
    mean, err = _mean_err(np.mean(vals, axis=0))
    want = float.fromhex(load("scalars")["analytic"]["rotor_chit_exact_32"])
    assert abs(mean - want) < 5 * err, (mean, err, want)
    t_indep, n_indep = s.independence()
    assert all(t >= 1.0 for t in t_indep) and n_indep[0] == 50
    assert s.cost_per_sample(2) > 0


def test_rotor_multilevel_mc_matches_exact_chit(mp, ctx):
    """MonteCarloMultiLevel::evaluate: sum_l <Y_l> = chi_t within the estimated error"""
    m = mp.rotor(32, 4.0, 0.25)
    B = 2048
    mc = mp.MultilevelMC(ctx, m, B, n_level=3, epsilon=2e-3, qoi=mp.QOI_ROTOR_CHI, n_burnin=30,
                         n_autocorr_window=10, n_min_samples_qoi=4 * B, max_iterations=20,
                         kind=mp.SAMPLER_HMC, nt=20, dt=0.1, renorm=mp.RENORM_PERTURBATIVE)
    converged = mc.evaluate()
    value, error, levels = mc.result()
    want = float.fromhex(load("scalars")["analytic"]["rotor_chit_exact_32"])
    assert converged, levels
    assert error < 3e-3 and abs(value - want) < 5 * error + 1e-3, (value, error, want, levels)
    assert all(lv["variance"] > 0 and lv["tau_int"] >= 1.0 and lv["cost_eff_usec"] > 0 for lv in levels)
    assert all(lv["samples"] >= lv["n_target"] for lv in levels)


# ------------------------------------------------------------- cluster samplers


def test_rotor_cluster_update_against_oracle(mp, ctx, orc):
    rng = np.random.default_rng(21)
    for M, m0 in [(32, 0.25), (64, 2.0), (16, 40.0)]:  # the last one: clusters wrap the ring
        o = po.rotor(M, 4.0, m0)
        m = to_mp(mp, o)
        B, chain0 = 6, 3
        x = rng.uniform(-np.pi, np.pi, (B, M))
        xd = dev(ctx, x)
        ctx.cluster_update(m, xd, chain0, 17, 5)
        want = np.array([orc.cluster_update(o, SEED, 17, 5, chain0 + b, x[b]) for b in range(B)])
        ang_close(host(xd), want, tol=1e-10, what=f"cluster M={M}")
        assert np.abs(host(xd) - x).max() > 1e-3


def test_schwinger_from_cluster_against_oracle(mp, ctx, orc):
    rng = np.random.default_rng(22)
    o = po.schwinger(8, 12, 3.0)
    m = to_mp(mp, o)
    B = 3
    psi = rng.uniform(-np.pi, np.pi, (B, 8 * 12))
    x = ctx.state(m, B)
    ctx.schwinger_from_cluster(m, dev(ctx, psi), x, 5, 9)
    want = np.array([orc.schwinger_from_cluster(o, SEED, 9, 5 + b, psi[b]) for b in range(B)])
    ang_close(host(x), want, tol=1e-10, what="links from cluster")
    # every plaquette angle is fixed by the rotor chain (gauge invariant): check one chain
    xo = want[0]
    P = orc.qoi(o, po.QOI_AVG_PLAQUETTE, xo)[0]
    assert abs(host(ctx.qoi(m, mp.QOI_AVG_PLAQUETTE, x))[0] - P) < 1e-12


def test_cluster_samplers_match_analytic_chit(mp, ctx):
    want_rotor = float.fromhex(load("scalars")["analytic"]["rotor_chit_exact_32"])
    want_schw = float.fromhex(load("scalars")["analytic"]["schwinger_chit_analytical_4_64"])
    for m, qoi, want in [(mp.rotor(32, 4.0, 0.25), mp.QOI_ROTOR_CHI, want_rotor),
                         (mp.schwinger(8, 8, 4.0), mp.QOI_SCHWINGER_CHI, want_schw)]:
        B = 4096
        s = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_CLUSTER, n_updates=10)
        x = ctx.state(m, B)
        for _ in range(30):
            s.draw(x)
        vals = []
        for _ in range(20):
            s.draw(x)
            vals.append(host(ctx.qoi(m, qoi, x)))
        mean, err = _mean_err(np.mean(vals, axis=0))
        assert abs(mean - want) < 5 * err, (m.model, mean, err, want)
        s.close()


def test_gff_hierarchical_matches_analytic_phi2(mp, ctx):
    """GFF 16x16, rotate coarsening, 2 levels (unrotated 16^2 -> rotated), HMC on the coarse level (a
    reversible coarse kernel, as delayed acceptance requires); the two-level step uses the 5-point action
    on both levels, which the Metropolis-Hastings correction makes exact for the fine-level
    distribution.  Chains are thermalised with overrelaxed heat-bath sweeps first (the slow zero mode),
    so a biased two-level kernel would show up as a drift away from the analytic value."""
    want = float.fromhex(load("scalars")["analytic"]["gff_phi_squared_10_16"])
    m = mp.gff(16, 16, 10.0)
    B = 4096
    hb = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_HEATBATH, n_sweep_overrelax=2, n_sweep_heatbath=1)
    x = ctx.init_state(m, B, 0, 0)
    hb.set_state(x)
    for _ in range(150):
        hb.draw(x)
    s = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_HMC, n_levels=2, ctype=mp.COARSEN_ROTATE, nt=20, dt=0.1)
    assert s.level_model(1).rotated == 1 and s.level_model(1).Mt_lat == 16
    s.set_state(x)
    for _ in range(60):
        s.draw(x)
    vals = []
    for _ in range(40):
        s.draw(x)
        vals.append(host(ctx.qoi(m, mp.QOI_PHI2, x)))
    mean, err = _mean_err(np.mean(vals, axis=0))
    assert abs(mean - want) < 5 * err, (mean, err, want)
    assert s.p_accept()[0] > 0.05, s.p_accept()
