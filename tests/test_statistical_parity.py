"""Third parity leg of north_star: "sampled observables match the reference within combined
statistical error bars at the same integrated autocorrelation".

The fixtures in tests/golden/stats.json were printed by the REFERENCE'S OWN DRIVERS
(src/driver_qft.cc / src/driver_qm.cc compiled byte for byte into oracle/_ref/driver_q*, run on the
CPU by tools/make_golden_stats.py with parameter files derived from the reference's templates).
Every test below runs the same sampler configuration through the CUDA library (B chains side by
side) and compares

  * the estimator with the reference's, within the combined errors (and with the analytic value),
  * tau_int of the QoI series -- the reference's estimator (common/statistics.cc:82-90) with the
    reference's window -- with the reference's tau_int,
  * the acceptance rate of every level with the reference's.

tau_int tolerance: the relative error of the windowed estimator is ~ sqrt(2 (2 W + 1) / N) per
series (Madras-Sokal), 3-5 % for the fixtures; the test allows 25 % (+ 0.5 absolute).  Local
heat-bath sweeps are exempt from the tau_int comparison: the kernels sweep colours, the reference
sweeps lexicographically -- same stationary distribution, different autocorrelation.
"""
import numpy as np
import pytest

from tests.util import load

pytestmark = pytest.mark.gpu

SEED = 0x5EED0001
STATS = load("stats")


@pytest.fixture(scope="module")
def mp():
    import mlmcpathintegral_b200 as mp
    return mp


@pytest.fixture(scope="module")
def ctx(mp):
    c = mp.Context(0, seed=SEED)
    yield c
    c.close()


def _val(s):
    return float(str(s).strip("'"))


def _build(mp, ov):
    """(model, qoi, sampler kwargs) from the overrides of a reference parameter file"""
    lat = ov["lattice"]
    kw = {}
    if "quantumfieldtheory" in ov:
        action = ov["quantumfieldtheory"]["action"].strip("'")
        if action == "quenchedschwinger":
            m = mp.schwinger(int(lat["Mt_lat"]), int(lat["Mx_lat"]), _val(ov["schwinger"]["beta"]))
            qoi = mp.QOI_SCHWINGER_CHI
            rn = ov["schwinger"].get("renormalisation", "'none'").strip("'")
            kw["ctype"] = mp.COARSEN_BOTH
        else:
            m = mp.gff(int(lat["Mt_lat"]), int(lat["Mx_lat"]), _val(ov["gff"]["mass"]))
            qoi, rn = mp.QOI_PHI2, "none"
            kw["ctype"] = mp.COARSEN_ROTATE
    else:
        m = mp.rotor(int(lat["M_lat"]), _val(lat["T_final"]), _val(ov["rotor"]["m0"]))
        qoi = mp.QOI_ROTOR_CHI
        rn = ov["rotor"].get("renormalisation", "'none'").strip("'")
    kw["renorm"] = {"none": mp.RENORM_NONE, "perturbative": mp.RENORM_PERTURBATIVE,
                    "nonperturbative": mp.RENORM_NONPERTURBATIVE}[rn]
    sampler = ov["singlelevelmc"]["sampler"].strip("'")
    kinds = {"HMC": mp.SAMPLER_HMC, "heatbath": mp.SAMPLER_HEATBATH, "cluster": mp.SAMPLER_CLUSTER,
             "exact": mp.SAMPLER_EXACT}
    if sampler == "hierarchical":
        kw["n_levels"] = int(ov["hierarchical"]["n_max_level"])
        kw["kind"] = kinds[ov["hierarchical"]["coarsesampler"].strip("'")]
    else:
        kw["n_levels"] = 1
        kw["kind"] = kinds[sampler]
    kw["nt"] = int(ov.get("hmc", {}).get("nt", 100))
    kw["dt"] = float(ov.get("hmc", {}).get("dt", 0.1))
    kw["n_updates"] = int(ov.get("clusteralgorithm", {}).get("n_updates", 10))
    kw["n_sweep_overrelax"] = int(ov.get("heatbath", {}).get("n_sweep_overrelax", 10))
    kw["n_sweep_heatbath"] = int(ov.get("heatbath", {}).get("n_sweep_heatbath", 1))
    return m, qoi, kw


def run_case(mp, ctx, name, B, n_burnin, n_draws):
    ref = STATS[name]
    m, qoi, kw = _build(mp, ref["overrides"])
    s = mp.Sampler(ctx, m, B, **kw)
    x = s.get_state()  # (draw() only overwrites the output of chains that accepted)
    for _ in range(n_burnin):
        s.draw(x)
    s.reset_stats()  # acceptance over the measured draws only
    st = mp.Statistics(ctx, ref["window"], B)
    means = None
    for _ in range(n_draws):
        s.draw(x)
        q = ctx.qoi(m, qoi, x)
        st.record(q)
        means = q.clone() if means is None else means + q
    r = mp.Statistics.finalize(st.pack(), ref["window"])
    r["acceptance"] = s.p_accept()
    chain_means = (means / n_draws).cpu().numpy()
    r["error_chains"] = float(chain_means.std(ddof=1) / np.sqrt(B))  # independent chains: model-free error
    st.close()
    s.close()
    return ref, r


def check(name, ref, r, tau=True, acceptance=True, analytic=True, n_sigma=4.5):
    err = max(r["error"], r["error_chains"])
    comb = float(np.hypot(err, ref["error"]))
    msg = (f"{name}: CUDA {r['average']:.5f} +- {err:.5f} (tau_int {r['tau_int']:.2f}, acceptance "
           f"{np.round(r['acceptance'], 3)}), reference driver {ref['average']:.5f} +- {ref['error']:.5f} "
           f"(tau_int {ref['tau_int']:.2f}, acceptance {ref['acceptance']}), analytic {ref.get('analytical')}")
    print(msg)
    assert abs(r["average"] - ref["average"]) <= n_sigma * comb, msg
    # the analytic value only where the reference's own run is consistent with it: some reference
    # configurations are biased BY CONSTRUCTION and parity means reproducing that (hierarchical sampler
    # with a cluster coarse sampler and few updates per draw -- correlated "independent" proposals:
    # reference 2.0135 +- 0.0217 vs exact 1.9339 at 16^2, beta = 4; GFF with a heat-bath coarse sampler,
    # which samples the 5-point action while the two-level step evaluates Q_hat: 0.3020 vs 0.3380)
    if analytic and ref.get("analytical") is not None and \
            abs(ref["average"] - ref["analytical"]) <= 3.0 * ref["error"]:
        assert abs(r["average"] - ref["analytical"]) <= n_sigma * err, msg
    if tau:
        assert abs(r["tau_int"] - ref["tau_int"]) <= 0.25 * ref["tau_int"] + 0.5, msg
    if acceptance and ref["acceptance"]:
        got = np.array(r["acceptance"])
        want = np.array(ref["acceptance"])
        assert got.shape == want.shape, msg
        assert np.all(np.abs(got - want) <= 0.02 + 0.05 * want), msg


# (case, chains, burn-in draws, measured draws)
SCHWINGER = [
    ("schwinger16_b4_hier2_cluster", 2048, 300, 1500),
    ("schwinger32_b16_hier2_cluster", 2048, 200, 600),
    ("schwinger64_b64_hier2_cluster", 1024, 100, 400),
    ("schwinger16_b4_hier2_hmc", 2048, 300, 1000),
    ("schwinger16_b4_cluster", 2048, 100, 500),
]


@pytest.mark.parametrize("name,B,n_burnin,n_draws", SCHWINGER, ids=[c[0] for c in SCHWINGER])
def test_schwinger_matches_reference_driver(mp, ctx, name, B, n_burnin, n_draws):
    ref, r = run_case(mp, ctx, name, B, n_burnin, n_draws)
    check(name, ref, r)


def test_schwinger_heatbath_matches_reference_driver(mp, ctx):
    # colour sweeps here, lexicographic sweeps in the reference: tau_int is not comparable
    ref, r = run_case(mp, ctx, "schwinger16_b4_heatbath", 2048, 200, 1000)
    check("schwinger16_b4_heatbath", ref, r, tau=False)
    assert 0.4 * ref["tau_int"] <= r["tau_int"] <= 2.5 * ref["tau_int"]


ROTOR = [("rotor32_hier3_hmc", 4096, 200, 1000), ("rotor64_cluster", 4096, 200, 1000)]


@pytest.mark.parametrize("name,B,n_burnin,n_draws", ROTOR, ids=[c[0] for c in ROTOR])
def test_rotor_matches_reference_driver(mp, ctx, name, B, n_burnin, n_draws):
    ref, r = run_case(mp, ctx, name, B, n_burnin, n_draws)
    check(name, ref, r, analytic=False)
    lat = ref["overrides"]["lattice"]
    T, M = _val(lat["T_final"]), int(lat["M_lat"])
    exact = mp._lib.lib.mlmcpi_rotor_chit(_val(ref["overrides"]["rotor"]["m0"]), T / M, T, 0)
    err = max(r["error"], r["error_chains"])
    assert abs(r["average"] - exact) <= 4.5 * err, (r["average"], err, exact)


# ------------------------------------------------------------------ Gaussian free field
# The reference's GFF hierarchy carries the Gibbs-smoothed action Q_hat on every coarse level
# (gffaction.hh:201-208) whatever the coarse sampler is.
#  * coarsesampler = 'HMC': the trajectories integrate the 5-point force but are accepted with
#    Action::evaluate = Q_hat, so the coarse chain samples Q_hat: a consistent algorithm, unbiased, and the
#    estimator must agree with the reference's (and the analytic value) within errors.
#  * coarsesampler = 'heatbath': the coarse chain samples the 5-POINT action while the two-level steps
#    evaluate Q_hat.  The reference's own estimate is biased (0.3020 +- 0.0001 vs the analytic 0.3380 at
#    16^2, two levels) and -- the chain not being reversible w.r.t. anything the steps assume -- its
#    stationary distribution depends on the coarse kernel's details (lexicographic sweeps there, colour
#    sweeps run forwards or backwards here).  Two levels agree within errors; with three and four levels
#    the estimators agree to a few per cent and the acceptance pattern (high on the finest level, a few
#    per cent on the intermediate ones) is the reference's.
#  * coarsesampler = 'exact' yields avg = 0, p = nan in the reference (GFFAction::draw never sets
#    MCMCStep::accept, so hierarchicalsampler.cc:73 breaks out of every draw): no fixture.
# (the chains start from exact samples of the fine-level action, csrc/capi.cu:gff_equilibrium_start; the
# reference's single chain runs 10^4 constructor draws before the first sample is recorded)
GFF = [("gff16_hier2_heatbath", 2048, 100, 400), ("gff16_hier2_hmc", 2048, 300, 1500),
       ("gff16_hier3_hmc", 2048, 300, 3000)]


@pytest.mark.parametrize("name,B,n_burnin,n_draws", GFF, ids=[c[0] for c in GFF])
def test_gff_matches_reference_driver(mp, ctx, name, B, n_burnin, n_draws):
    ref, r = run_case(mp, ctx, name, B, n_burnin, n_draws)
    # tau_int: colour sweeps vs lexicographic sweeps on the coarsest level, and series that are
    # longer than the window where an intermediate level accepts a few per cent -- within a factor of two
    check(name, ref, r, tau=False)
    assert 0.5 * ref["tau_int"] <= r["tau_int"] <= 2.0 * ref["tau_int"], (r["tau_int"], ref["tau_int"])


GFF_INCONSISTENT = [("gff16_hier3_heatbath", 2048, 1000, 2000), ("gff32_hier4_heatbath", 1024, 1000, 2000)]


@pytest.mark.parametrize("name,B,n_burnin,n_draws", GFF_INCONSISTENT, ids=[c[0] for c in GFF_INCONSISTENT])
def test_gff_heatbath_hierarchy_follows_reference_driver(mp, ctx, name, B, n_burnin, n_draws):
    ref, r = run_case(mp, ctx, name, B, n_burnin, n_draws)
    print(f"{name}: CUDA {r['average']:.5f} (tau_int {r['tau_int']:.1f}, acceptance {np.round(r['acceptance'], 3)}), "
          f"reference driver {ref['average']:.5f} +- {ref['error']:.5f} (tau_int {ref['tau_int']:.1f}, "
          f"acceptance {ref['acceptance']}), analytic {ref['analytical']}")
    assert abs(r["average"] - ref["average"]) <= 0.08 * ref["average"]
    # both sit below the analytic value: the bias of the configuration, not of the implementation
    assert r["average"] < ref["analytical"] and ref["average"] < ref["analytical"]
    got, want = np.array(r["acceptance"]), np.array(ref["acceptance"])
    assert np.all(got > 0.4 * want) and np.all(got < 2.5 * want + 1e-12), (got, want)
    assert 0.5 * ref["tau_int"] <= r["tau_int"] <= 2.0 * ref["tau_int"]


def test_gff_256_four_levels_every_level_accepts(mp, ctx):
    """BASELINE config C3: GFF 256 x 256, coarsening rotate, 4 levels = 65536 / 32768 / 16384 / 8192
    vertices, Q_hat (dense, built on the device) on the three coarse levels.  (a) heat-bath coarse
    sampler as the config names it: every level accepts (the acceptance pattern of the reference at
    32^2, tests/golden/stats.json: high on the finest level, a few per cent on the intermediate ones);
    (b) exact coarse sampler, two levels: unbiased, <phi^2> within errors of gff_phi_squared_analytical."""
    m = mp.gff(256, 256, 10.0)
    B = 64
    s = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_HEATBATH, n_levels=4, ctype=mp.COARSEN_ROTATE,
                   n_sweep_overrelax=10, n_sweep_heatbath=1)
    assert [mp.sample_size(s.level_model(l)) for l in range(4)] == [65536, 32768, 16384, 8192]
    assert all(s.level_model(l).gff_n_gibbs == 2 for l in (1, 2, 3))
    x = s.get_state()
    for _ in range(150):
        s.draw(x)
    p = s.p_accept()
    print("gff256 4 levels, acceptance per level:", p)
    assert p[3] == 1.0 and all(pl > 0.0 for pl in p), p
    s.close()
    exact = mp._lib.lib.mlmcpi_gff_phi_squared_analytical(10.0, 256, 256)
    s = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_EXACT, n_levels=2, ctype=mp.COARSEN_ROTATE)
    s.get_state(x)
    means = None
    n_draws = 60
    for k in range(20 + n_draws):
        s.draw(x)
        if k >= 20:
            q = ctx.qoi(m, mp.QOI_PHI2, x)
            means = q.clone() if means is None else means + q
    cm = (means / n_draws).cpu().numpy()
    mean, err = float(cm.mean()), float(cm.std(ddof=1) / np.sqrt(B))
    p = s.p_accept()
    print(f"gff256 2 levels exact coarse sampler: <phi^2> = {mean:.6f} +- {err:.6f}, analytic {exact:.6f}, "
          f"acceptance {p}")
    assert p[0] > 0.5, p
    assert abs(mean - exact) <= 4.5 * err, (mean, err, exact)
    s.close()
