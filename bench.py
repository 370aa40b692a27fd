#!/usr/bin/env python
"""bench.py -- headline benchmark: quenched Schwinger model 512x512, batched
HierarchicalSampler draw = HMC on the coarsest level + conditioned fill-in on the two
finer levels + action differences / accept / topological susceptibility
(BASELINE.json configs[3]; metric: lattice site-updates/s, ESS/s reported beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one HierarchicalSampler::draw for every chain of the batch.  A
site-update = one lattice site advanced by one leapfrog step, or one
fine-level site filled in; both are counted from what the kernels actually executed.

Arms
  default            CUDA path through the C-ABI; `value` with states resident in HBM,
                     `e2e` through the host-buffer entry point (mlmcpi_sampler_draw_host:
                     pinned host state -> device every step, QoI back to the host).
  --impl reference   the reference's own CPU implementation (its translation units in
                     oracle/_ref, else the C port in oracle/), same cascade on a bounded
                     sample of the workload, all host cores (one chain per process -- the
                     reference's MPI mode is exactly independent chains).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "lattice site-updates/s (Schwinger HMC + hierarchical fill-in)"
UNIT = "site-updates/s"


# The five configurations of BASELINE.json.  `schwinger512` (configs[3]) is the one the metric
# is quoted on and what the driver measures; the others are run on request (--workload) so that
# every named shape has a measured throughput (profiles/r01_configs.md).
WORKLOADS = {
    "ho32": dict(config="configs[0]: driver_qm harmonic oscillator, single-level HMC, parameters_qm_template.in "
                        "defaults (M_lat=32, T=4, m0=mu2=1, nt=100, dt=0.1)",
                 model="ho", lattice=32, beta=None, levels=1, chains=65536, sampler="HMC", qoi="QOI_X2"),
    "rotor256": dict(config="configs[1]: driver_qm topological rotor, hierarchical sampler, M_lat=256, 3 levels, "
                            "HMC on the coarsest level",
                     model="rotor", lattice=256, beta=None, levels=3, chains=8192, sampler="HMC",
                     qoi="QOI_ROTOR_CHI"),
    "gff256": dict(config="configs[2]: driver_qft GFF 256x256, 4 levels (coarsening rotate), checkerboard "
                          "overrelaxed heat bath on the coarsest level, conditioned Gaussian fill-in above",
                   model="gff", lattice=256, beta=None, levels=4, chains=512, sampler="heatbath", qoi="QOI_PHI2"),
    "gff256_mlmc": dict(config="configs[2] as MonteCarloMultiLevel: GFF 256x256, 4 levels (coarsening rotate), every level's "
                               "sampler the overrelaxed heat bath, two-level steps with the Gibbs-smoothed coarse actions",
                        model="gff", lattice=256, beta=None, levels=4, chains=256, sampler="heatbath", qoi="QOI_PHI2"),
    "schwinger512_heatbath": dict(config="driver_qft quenched Schwinger 512x512, single-level overrelaxed heat bath "
                                         "sampler with the parameters_qft_template.in defaults (10 overrelaxation "
                                         "sweeps + 1 heat-bath sweep per draw)",
                                  model="schwinger", lattice=512, beta=1024.0, levels=1, chains=512,
                                  sampler="heatbath", qoi="QOI_SCHWINGER_CHI", n_sweep_overrelax=10),
    "schwinger512": dict(config="configs[3]", model="schwinger", lattice=512, beta=1024.0, levels=3, chains=512,
                         sampler="HMC", qoi="QOI_SCHWINGER_CHI"),
    "schwinger1024": dict(config="configs[4]: quenched Schwinger 1024x1024, hierarchical sampler (3 levels), chains "
                                 "sharded over the GPUs, NCCL allreduce of the QoI moments",
                          model="schwinger", lattice=1024, beta=1024.0, levels=3, chains=128, sampler="HMC",
                          qoi="QOI_SCHWINGER_CHI"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="schwinger512", choices=sorted(WORKLOADS),
                    help="BASELINE.json config to run; the default is the headline one (configs[3])")
    ap.add_argument("--lattice", type=int, default=None)
    ap.add_argument("--beta", type=float, default=None,
                    help="fine-level coupling; default = continuum-limit point beta/P = 2^-8 at 512^2 "
                         "(ApproximateBesselProduct fill-in); --beta 4 runs the BesselProduct regime")
    ap.add_argument("--levels", type=int, default=None)
    ap.add_argument("--chains", type=int, default=None, help="chains per GPU")
    ap.add_argument("--nt", type=int, default=100)
    ap.add_argument("--dt", type=float, default=0.1)
    ap.add_argument("--thermalise", type=int, default=20, help="overrelaxed heat-bath sweeps before timing")
    ap.add_argument("--autotune", type=int, default=1, help="tune the HMC step size as HMCSampler does")
    ap.add_argument("--ess-draws", type=int, default=40, help="draws of the ESS leg (cluster coarse sampler); 0 = skip")
    ap.add_argument("--ess-updates", type=int, default=100, help="cluster updates per draw in the ESS leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the beta = 4, sweep-roofline and MLMC legs")
    ap.add_argument("--mlmc-chains", type=int, default=64, help="chains per GPU and level of the MLMC leg (configs[4])")
    ap.add_argument("--mlmc-epsilon", type=float, default=0.25, help="tolerance of the MLMC leg on V chi_t (~ 6.5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    a = ap.parse_args()
    w = WORKLOADS[a.workload]
    for k in ("lattice", "beta", "levels", "chains"):
        if getattr(a, k) is None:
            setattr(a, k, w.get(k))
    return a


def workload_config(a):
    w = WORKLOADS[a.workload]
    if w["model"] == "schwinger" and w["sampler"] == "HMC":
        return {
            "workload": f"driver_qft quenched Schwinger {a.lattice}x{a.lattice}, hierarchical sampler "
                        f"({a.levels} levels, coarsening both, perturbative renormalisation), HMC coarse "
                        f"sampler nt={a.nt} dt={a.dt}, QoI topological susceptibility",
            "lattice": [a.lattice, a.lattice], "beta": a.beta, "levels": a.levels, "nt": a.nt, "dt": a.dt,
        }
    return {"workload": w["config"], "lattice": [a.lattice], "levels": a.levels, "nt": a.nt, "dt": a.dt}


def build_model(mp, a):
    w = WORKLOADS[a.workload]
    if w["model"] == "schwinger":
        return mp.schwinger(a.lattice, a.lattice, a.beta)
    if w["model"] == "gff":
        return mp.gff(a.lattice, a.lattice, 10.0)
    if w["model"] == "rotor":
        return mp.rotor(a.lattice)
    return mp.ho(a.lattice)


# --------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    """one process = one independent chain (the reference's MPI decomposition):
    n_draws full hierarchical cascades, timed after one untimed warm-up cascade"""
    (lattice, beta, levels, nt, dt, n_draws, seed, use_ref) = args
    import numpy as np

    from oracle import pyoracle as po
    rng = np.random.default_rng(seed)
    work = 0.0
    if use_ref:
        R = po.ref()
        acts = [R.action(po.SCHWINGER, [lattice, lattice, po.BOTH, 1], [beta])]
        for _ in range(levels - 1):
            acts.append(acts[-1].coarse())
        x = [rng.uniform(-np.pi, np.pi, a.n) for a in acts]

        def cascade():
            w = 0.0
            for l in range(1, levels):  # hierarchicalsampler.cc:57-60
                x[l] = acts[l].copy_from_fine(x[l - 1])
            _, x[levels - 1], _, _ = acts[-1].hmc_draws(nt, dt, 1, int(rng.integers(1 << 30)),
                                                       x[levels - 1])
            w += (nt + 1) * acts[-1].n / 2
            for l in range(levels - 2, -1, -1):  # twolevelmetropolisstep.cc:35-97
                a, ac = acts[l], acts[l + 1]
                Sf, Sc = a.evaluate(x[l]), a.cond_evaluate(x[l])
                tp = a.cond_fill(a.copy_from_coarse(x[l + 1], x[l]))
                dS = (a.evaluate(tp) - Sf) + (ac.evaluate(ac.copy_from_fine(x[l])) - ac.evaluate(x[l + 1])) \
                    + (Sc - a.cond_evaluate(tp))
                if dS < 0 or rng.uniform() < np.exp(-dS):
                    x[l] = tp
                w += a.n / 2
            return w
    else:
        orc = po.oracle()
        mods = [po.schwinger(lattice, lattice, beta)]
        for l in range(levels - 1):
            mods.append(orc.coarse_model(mods[-1], 1, l, po.BOTH))
        x = [rng.uniform(-np.pi, np.pi, orc.sample_size(m)) for m in mods]
        state = {"draw": 0}

        def cascade():
            w = 0.0
            state["draw"] += 1
            for l in range(1, levels):
                x[l] = orc.restrict(mods[l - 1], mods[l], x[l - 1])
            _, x[levels - 1], _ = orc.hmc_step(mods[-1], nt, dt, seed, state["draw"], 0, x[levels - 1])
            w += (nt + 1) * orc.sample_size(mods[-1]) / 2
            for l in range(levels - 2, -1, -1):
                Sf, Sc = orc.action(mods[l], x[l]), orc.cond_action(mods[l], x[l])
                _, x[l], _, _, _ = orc.twolevel_step(mods[l], mods[l + 1], seed, state["draw"] * 16 + l,
                                                     0, x[l + 1], x[l], Sf, Sc)
                w += orc.sample_size(mods[l]) / 2
            return w

    cascade()  # warm-up (page-in, Bessel tables)
    t0 = time.perf_counter()
    for _ in range(n_draws):
        work += cascade()
    return work, time.perf_counter() - t0


def cpu_arm(a, n_draws, cores=None):
    """returns (site-updates/s over all cores, cores, kind, sample description)"""
    import multiprocessing as mp

    from oracle import pyoracle as po
    if not os.path.exists(po.ORACLE_SO):
        po.build(ref=False)
    use_ref = po.have_ref()
    cores = cores or len(os.sched_getaffinity(0))
    jobs = [(a.lattice, a.beta, a.levels, a.nt, a.dt, n_draws, 1000 + c, use_ref) for c in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    work = sum(r[0] for r in res)
    t_max = max(r[1] for r in res)
    kind = "reference" if use_ref else "port"
    sample = (f"{cores} independent chains (one process per core), {n_draws} full hierarchical "
              f"cascade(s) each on the {a.lattice}x{a.lattice} workload after 1 warm-up cascade; "
              f"slowest process {t_max:.2f} s, wall incl. start-up {wall:.1f} s")
    return work / t_max, cores, kind, sample, work, t_max


def reference_main(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if WORKLOADS[a.workload]["model"] != "schwinger" or WORKLOADS[a.workload]["sampler"] != "HMC":
        emit({"impl": "reference", "unavailable": "the reference arm times the Schwinger HMC workloads only"})
        return
    draws_total = a.steps
    # bounded sample: one cascade at 512^2 is ~0.5 s per core; every process does one
    # untimed warm-up cascade, then `steps` timed ones
    value, cores, kind, sample, work, t = cpu_arm(a, max(1, draws_total))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t / max(1, a.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (U(-pi,pi) start states)", "config": workload_config(a),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# --------------------------------------------------------------------------- GPU arm
class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 20 ms from before the warm-up; only the
    samples inside the timed window [t0, t1] (host clock) are summarised"""
    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.QUERY}",
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self, t0, t1):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, sm_all = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [s.strip() for s in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clk, cmax = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            mx.append(cmax)
            sm_all.append(clk)
            if t0 - 0.02 <= ts <= t1 + 0.02:
                sm.append(clk)
                for n, v in zip(names, parts[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        self.f.close()
        os.unlink(self.f.name)
        use = sm if sm else sm_all[-3:]
        if use:
            use.sort()
            out = {"sm_mhz": use[len(use) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm), "window_s": t1 - t0}
        return out


def bind_to_gpu_cpus(index):
    """run this process on the CPUs next to its GPU (NVML's ideal affinity), so that the pinned host buffers
    it allocates afterwards are NUMA-local to the GPU's PCIe root"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and 64 * w + b < n_cpu]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"gpu": index, "cpus": "%d-%d (%d)" % (allowed[0], allowed[-1], len(allowed)) if allowed else None}
    except Exception as e:  # no NVML / no permission: stay where we are
        return {"gpu": index, "cpus": None, "note": "affinity not set: %s" % type(e).__name__}


def _timed(torch, fn, n):
    """CUDA-event time (ms) of n calls of fn on the current stream, after one untimed call"""
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def ess_leg(mp, ctx, torch, dist, a, m, B, rank, world):
    """ESS/s on the topological susceptibility from an ERGODIC sampler on the same lattice and beta: the
    hierarchical sampler with the cluster coarse sampler (hierarchical: coarsesampler = 'cluster',
    sampler/quenchedschwingerclustersampler.cc:40-86 -- the reference's topology-changing move).  With the
    HMC coarse sampler of the throughput leg every chain stays in the sector it starts in at this coupling
    (coarse beta ~ 64: no tunnelling, in the reference as here; profiles/r02_summary.md section 1)."""
    k_max = 20
    s = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_CLUSTER, n_levels=a.levels, renorm=mp.RENORM_PERTURBATIVE,
                   ctype=mp.COARSEN_BOTH, n_updates=a.ess_updates, chain0=(world + rank) * B)
    st = mp.Statistics(ctx, k_max, B)
    x = s.get_state()
    packed = ctx.empty(8 + k_max)
    n_burn = 60  # (10 left V chi_t 1 % low over the first 40 draws: 2.7 sigma with the statistics of 8 GPUs)
    for _ in range(n_burn):  # the start state is a draw of the cascade itself (cascade start): short burn-in
        s.draw(x)
    s.reset_stats()
    ctx.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.ess_draws):
        s.draw(x)
        st.record(s.qoi(mp.QOI_SCHWINGER_CHI))
        st.pack_device(packed)
        if world > 1:
            dist.all_reduce(packed)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    r = mp.Statistics.finalize(packed.cpu().numpy(), k_max)
    exact = mp._lib.lib.mlmcpi_schwinger_chit_analytical(a.beta, a.lattice * a.lattice)
    out = {
        "sampler": f"hierarchical ({a.levels} levels), coarse sampler 'cluster' with {a.ess_updates} updates per draw",
        "chains_per_gpu": B, "draws": a.ess_draws, "burnin_draws": n_burn, "ms_per_draw": ms / a.ess_draws,
        "acceptance_per_level": s.p_accept(),
        "qoi": {"name": "QOI_SCHWINGER_CHI", "average": r["average"], "error": r["error"], "tau_int": r["tau_int"],
                "window": k_max, "samples": r["samples"], "analytic": exact,
                "deviation_sigma": abs(r["average"] - exact) / r["error"] if r["error"] > 0 else None},
        "ess_per_s": r["samples"] / r["tau_int"] / (ms * 1e-3),
    }
    st.close()
    s.close()
    del x
    torch.cuda.empty_cache()
    return out


def beta4_leg(mp, ctx, torch, a, rank):
    """the BesselProduct regime (beta <= 8 on every level: quenchedschwingerconditionedfineaction.hh:39-45)
    of the same hierarchy, timed: HMC on the coarsest level + two exact-conditional fill-ins"""
    B = max(32, a.chains // 4)
    m = mp.schwinger(a.lattice, a.lattice, 4.0)
    s = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_HMC, n_levels=a.levels, nt=a.nt, dt=a.dt,
                   renorm=mp.RENORM_PERTURBATIVE, ctype=mp.COARSEN_BOTH, chain0=rank * B)
    x = ctx.init_state(m, B, rank * B, 0)
    for k in range(5):
        ctx.overrelax_sweep(m, x)
        ctx.heatbath_sweep(m, x, rank * B, 2000 + k)
    s.set_state(x)
    s.draw(x)
    ms = _timed(torch, lambda: s.draw(x), 3) / 3
    w = s.work()
    units = w["leapfrog_site_steps"] + w["filled_fine_sites"]
    out = {"beta": 4.0, "fill_in": "BesselProduct (exact conditional, beta <= 8)", "chains_per_gpu": B,
           "ms_per_step": ms, "site_updates_per_s_per_gpu": units / (ms * 1e-3),
           "filled_fine_sites_per_s_per_gpu": w["filled_fine_sites"] / (ms * 1e-3),
           "acceptance_per_level": s.p_accept(),
           "note": "two-level acceptance collapses with the volume at fixed beta (16^2: 0.2, 512^2: < 1e-3), in the "
                   "reference as here; this leg times the kernels of the regime"}
    s.close()
    del x
    torch.cuda.empty_cache()
    return out


def mlmc_leg(mp, torch, dist, a, local, rank, world):
    """BASELINE configs[4]: quenched Schwinger 1024 x 1024, MULTILEVEL MONTE CARLO (MonteCarloMultiLevel::evaluate,
    montecarlo/montecarlomultilevel.cc:71-204) with the chains of every level sharded over the GPUs; the library's
    own host-side decisions (sub-sampling on tau_int, sample allocation from the variances and costs of all levels)
    run over the chains of ALL ranks through mlmcpi_set_allreduce (NCCL), so the ranks stay in lockstep.  3 MLMC
    levels 1024^2 / 512^2 / 256^2; the sampler of every level is the hierarchical sampler down to 128^2 with the
    cluster coarse sampler (the ergodic choice, see ess_leg)."""
    import time as _time
    L_fine, n_level, B = 1024, 3, a.mlmc_chains
    beta = float(L_fine * L_fine) / 256.0
    ctx = mp.Context(local, seed=0x5EED0002)
    if world > 1:
        ctx.attach_process_group()
    m = mp.schwinger(L_fine, L_fine, beta)
    torch.cuda.synchronize()
    t0 = _time.perf_counter()
    mc = mp.MultilevelMC(ctx, m, B, n_level=n_level, epsilon=a.mlmc_epsilon, qoi=mp.QOI_SCHWINGER_CHI, n_burnin=20,
                         n_autocorr_window=20, n_min_samples_qoi=4 * B * world, max_iterations=8, chain0=rank * B,
                         kind=mp.SAMPLER_CLUSTER, n_levels=n_level + 1, renorm=mp.RENORM_PERTURBATIVE,
                         ctype=mp.COARSEN_BOTH, n_updates=100)
    ctx.sync()
    t1 = _time.perf_counter()
    launches0 = ctx.launches
    converged = mc.evaluate()
    ctx.sync()
    t2 = _time.perf_counter()
    value, error, levels = mc.result()
    tt = torch.tensor([t1 - t0, t2 - t1], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    setup_s, eval_s = (float(v) for v in tt.cpu())
    analytic = mp._lib.lib.mlmcpi_schwinger_chit_perturbative(beta, L_fine * L_fine)
    fine_sites = sum(lv["samples"] * (L_fine >> l) ** 2 for l, lv in enumerate(levels))
    out = {
        "workload": "configs[4]: quenched Schwinger %dx%d, multilevel Monte Carlo, %d levels, chains sharded over %d "
                    "GPU(s), NCCL all-reduce of the QoI moments and of every host-side decision" % (L_fine, L_fine, n_level, world),
        "beta": beta, "chains_per_gpu_and_level": B, "epsilon": a.mlmc_epsilon, "converged": bool(converged),
        "estimate": value, "error": error, "analytic_perturbative": analytic,
        "deviation_sigma": abs(value - analytic) / error if error > 0 else None,
        "levels": levels, "setup_s": setup_s, "evaluate_s": eval_s,
        "samples_per_s": sum(lv["samples"] for lv in levels) / eval_s,
        "sampled_lattice_sites_per_s": fine_sites / eval_s,
        "gpu_launches": ctx.launches - launches0,
    }
    mc.close()
    ctx.close()
    torch.cuda.empty_cache()
    return out


def sweeps_leg(mp, ctx, torch, a, m, x, peak):
    """roofline of the update sweeps (north_star: >= 90 % of HBM on the Schwinger / GFF sweeps), CUDA events
    in this very run: algorithmic bytes (32 B per Schwinger site, 16 B per GFF vertex) / time"""
    out = {}
    B, sites = x.shape[0], a.lattice * a.lattice
    n = 10
    ms = _timed(torch, lambda: ctx.overrelax_sweeps(m, x, 2), n) / (2 * n)
    out["schwinger_overrelaxation"] = {"kernel": "overrelax_rowpipe_kernel (4 colours in one pass)", "ms_per_sweep": ms,
                                       "bytes_per_sweep": 32.0 * sites * B}
    ms = _timed(torch, lambda: ctx.heatbath_sweep(m, x, 0, 77), 4) / 4
    out["schwinger_heatbath"] = {"kernel": "heat-bath sweep (ExpCos rejection sampler per link)", "ms_per_sweep": ms,
                                 "bytes_per_sweep": 32.0 * sites * B}
    g = mp.gff(256, 256, 10.0)
    Bg = 2048
    xg = ctx.init_state(g, Bg, 0, 0)
    ms = _timed(torch, lambda: ctx.overrelax_sweeps(g, xg, 2), n) / (2 * n)
    out["gff_overrelaxation"] = {"kernel": "gff::sweep_rowpipe_kernel<false> (2 colours in one pass)", "ms_per_sweep": ms,
                                 "bytes_per_sweep": 16.0 * 65536 * Bg}
    hb = mp.Sampler(ctx, g, Bg, kind=mp.SAMPLER_HEATBATH, n_sweep_overrelax=0, n_sweep_heatbath=2)
    hb.set_state(xg)
    ms = _timed(torch, lambda: hb.draw(None), n) / (2 * n)
    out["gff_heatbath"] = {"kernel": "gff::sweep_rowpipe_kernel<true> (2 colours in one pass, 1 normal per vertex)",
                           "ms_per_sweep": ms, "bytes_per_sweep": 16.0 * 65536 * Bg}
    hb.close()
    for v in out.values():
        v["achieved_gbs"] = v["bytes_per_sweep"] / (v["ms_per_sweep"] * 1e-3) / 1e9
        v["peak_gbs"] = peak
        v["frac"] = v["achieved_gbs"] / peak
    del xg
    torch.cuda.empty_cache()
    return out


def gff_mlmc_main(a):
    """GFF 256^2 through mlmcpi_mlmc_* (MonteCarloMultiLevel::evaluate): one JSON line"""
    import torch

    import mlmcpathintegral_b200 as mp
    ctx = mp.Context(0, seed=0x5EED0003)
    m = mp.gff(a.lattice, a.lattice, 10.0)
    B = a.chains
    t0 = time.perf_counter()
    mc = mp.MultilevelMC(ctx, m, B, n_level=a.levels, epsilon=a.mlmc_epsilon if a.mlmc_epsilon < 0.1 else 0.004,
                         qoi=mp.QOI_PHI2, n_burnin=20, n_autocorr_window=20, n_min_samples_qoi=4 * B,
                         max_iterations=6, kind=mp.SAMPLER_HEATBATH, n_levels=1, renorm=mp.RENORM_NONE,
                         ctype=mp.COARSEN_ROTATE, n_sweep_overrelax=10, n_sweep_heatbath=1)
    ctx.sync()
    t1 = time.perf_counter()
    converged = mc.evaluate()
    ctx.sync()
    t2 = time.perf_counter()
    value, error, levels = mc.result()
    exact = mp._lib.lib.mlmcpi_gff_phi_squared_analytical(10.0, a.lattice, a.lattice)
    sites = sum(lv["samples"] * (a.lattice ** 2 >> l) for l, lv in enumerate(levels))
    emit({"metric": "lattice site-updates/s (gff256_mlmc)", "value": sites / (t2 - t1), "unit": "sampled lattice sites/s",
          "n_gpus": 1, "steps": 1, "warmup": 0, "ms_per_step": 1e3 * (t2 - t1), "higher_is_better": True,
          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "gpu_launches": ctx.launches,
          "config": {"workload": WORKLOADS[a.workload]["config"], "chains_per_level": B, "converged": bool(converged),
                     "setup_s": t1 - t0, "evaluate_s": t2 - t1, "estimate": value, "error": error, "analytic": exact,
                     "levels": levels,
                     "note": "the reference's configuration: heat-bath samplers draw from the 5-point action of a level "
                             "while the two-level steps evaluate the Gibbs-smoothed Q_hat, which biases the estimate "
                             "(reference driver at 16^2, hierarchical sampler: 0.3020 for the analytic 0.3380)"}})


def gpu_main(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import mlmcpathintegral_b200 as mp

    if a.workload == "gff256_mlmc":
        return gff_mlmc_main(a)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    ctx = mp.Context(local, seed=0x5EED0001)
    B = a.chains
    w = WORKLOADS[a.workload]
    is_schwinger = w["model"] == "schwinger"
    QOI = getattr(mp, w["qoi"])
    m = build_model(mp, a)
    kind = mp.SAMPLER_HMC if w["sampler"] == "HMC" else mp.SAMPLER_HEATBATH
    sampler = mp.Sampler(ctx, m, B, kind=kind, n_levels=a.levels, nt=a.nt, dt=a.dt,
                         renorm=mp.RENORM_PERTURBATIVE if w["model"] != "gff" else mp.RENORM_NONE,
                         ctype=mp.COARSEN_ROTATE if w["model"] == "gff" else mp.COARSEN_BOTH,
                         n_sweep_overrelax=w.get("n_sweep_overrelax", 1), n_sweep_heatbath=1, chain0=rank * B)
    k_max = 10
    stats = mp.Statistics(ctx, k_max, B)
    # start state: hot (U(-pi,pi), Action::initialise_state) at small beta, cold at large beta,
    # then thermalised by overrelaxed heat-bath sweeps on the fine level (untimed)
    if is_schwinger:
        x = ctx.init_state(m, B, rank * B, 0) if a.beta <= 8 else ctx.state(m, B)
    else:
        x = ctx.init_state(m, B, rank * B, 0)
    if w["model"] == "gff" and a.levels > 1:
        # the GFF hierarchy keeps the start state the library gives it (csrc/capi.cu: cascade start for the
        # heat-bath coarse sampler -- a thermalised fine-level state is a metastable start for that chain)
        x = sampler.get_state()
        for _ in range(20):
            sampler.draw(x)
    else:
        if w["model"] != "ho":  # (the harmonic oscillator has no heat bath; HMC burn-in below)
            for k in range(a.thermalise):
                ctx.overrelax_sweep(m, x)
                ctx.heatbath_sweep(m, x, rank * B, 1000 + k)
        sampler.set_state(x)
    tuned = None
    if a.autotune and kind == mp.SAMPLER_HMC:  # HMCSampler::autotune_stepsize (hmcsampler.cc:72-113)
        dt0 = a.dt
        for _ in range(12):  # bring dt into the bisection bracket [dt/2, 2 dt] of the reference
            sampler.set_dt(dt0)
            dt_t, p_t, ok = sampler.autotune(0.8, 8, 2 * B)
            if ok or p_t > 0.8:
                break
            dt0 *= 0.5
        tuned = {"dt": dt_t, "p_accept": p_t, "converged": ok}
    packed = ctx.empty(8 + k_max)

    def step():
        sampler.draw(x)
        # QoI of the chains' states (x holds them: draw() overwrote the accepted chains).  mlmcpi_sampler_qoi: for the
        # susceptibility of a hierarchical Schwinger sampler the charge of the trial state came out of the fill-in kernel
        stats.record(sampler.qoi(QOI))
        stats.pack_device(packed)
        if world > 1:  # QoI moments + autocorrelation sums: the only inter-GPU traffic
            dist.all_reduce(packed)

    clocks = ClockSampler(local) if rank == 0 else None
    for _ in range(a.warmup):
        step()
    stats.hard_reset()
    ctx.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ctx.profile(True)
    ctx.profile_read()
    launches0 = ctx.launches
    sampler.reset_stats()  # acceptance rates of the timed region (they decide how many chains a level fills in)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record()
    for _ in range(a.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    lf_ms, lf_launches, lf_bytes = ctx.profile_read()
    ctx.profile(False)
    launches = ctx.launches - launches0
    work = sampler.work()
    p_timed = sampler.p_accept()
    if is_schwinger and len(p_timed) > 1:
        # a level is only filled in for the chains whose cascade is still alive (hierarchicalsampler.cc:73-74; the
        # fill-in kernel skips the others): count the fine sites that WERE filled, from the measured acceptance rates
        filled, alive = 0.0, 1.0
        for l in range(len(p_timed) - 2, -1, -1):
            alive *= p_timed[l + 1]
            lm = sampler.level_model(l)
            filled += B * lm.Mt_lat * lm.Mx_lat * alive
        work["filled_fine_sites_nominal"] = work["filled_fine_sites"]
        work["filled_fine_sites"] = filled
    units_per_step = work["leapfrog_site_steps"] + work["filled_fine_sites"] + work["sweep_site_updates"]
    t = torch.tensor([ms], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    clk = clocks.stop(t_wall0, t_wall1) if clocks else None
    value = units_per_step * world * a.steps / (ms_max * 1e-3)
    st = mp.Statistics.finalize(packed.cpu().numpy(), k_max)
    p_acc = sampler.p_accept()

    # ---- e2e: the public host API a driver calls.  As Sampler::draw(state) does, every step hands the NEW
    #      STATES of all chains (and their QoI) back to the host; the chains themselves stay resident on the
    #      device, as the reference's samplers keep phi_state_cur.  The device-to-host copy of step k runs on a
    #      second stream while step k+1 computes (mlmcpi_sampler_draw_host_async, two pinned buffers that are
    #      allocated after the process has been bound to the CPUs next to its GPU).
    e2e = None
    if not a.no_e2e:
        n = mp.sample_size(m)
        affinity0 = os.sched_getaffinity(0)
        numa = bind_to_gpu_cpus(local)
        h_x = [torch.empty(B, n, dtype=torch.float64, pin_memory=True) for _ in range(2)]
        h_q = [torch.empty(B, dtype=torch.float64, pin_memory=True) for _ in range(2)]
        x_now = sampler.get_state()
        for k in range(2):  # the host buffers start as the chains' states: draw() only overwrites accepted chains
            h_x[k].copy_(x_now)
        del x_now
        for k in range(2):  # warm-up
            sampler.draw_host_async(QOI, h_q[k], h_x[k])
        sampler.wait_host()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        k_e2e = max(2, min(a.steps, 6))
        t_e0 = time.perf_counter()
        for k in range(k_e2e):
            sampler.draw_host_async(QOI, h_q[k & 1], h_x[k & 1])
        sampler.wait_host()  # the last copy has landed
        t_full = time.perf_counter() - t_e0
        # the same API handing back the QoI only (what a driver that evaluates on the device needs)
        t_e0 = time.perf_counter()
        for k in range(k_e2e):
            sampler.draw_host_async(QOI, h_q[k & 1], None)
        sampler.wait_host()
        t_qoi = time.perf_counter() - t_e0
        te = torch.tensor([t_full, t_qoi], dtype=torch.float64, device=ctx.device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        t_full, t_qoi = (float(v) for v in te.cpu())
        # the ceiling: what the copy engine moves device -> pinned host when every rank copies at the same time
        link = None
        try:
            probe = torch.empty(B * n // 2, dtype=torch.float64, device=ctx.device)  # half the states: 1 GiB
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t_l0 = time.perf_counter()
            h_x[0].view(-1)[:probe.numel()].copy_(probe, non_blocking=True)
            torch.cuda.synchronize()
            link = probe.numel() * 8 / (time.perf_counter() - t_l0) / 1e9
            del probe
        except Exception:  # pragma: no cover
            link = None
        chk = float(h_q[(k_e2e - 1) & 1].mean())  # the host really holds the result
        os.sched_setaffinity(0, affinity0)  # (the cpu_baseline leg uses every core)
        acc_all = 1.0
        for pl in sampler.p_accept():
            acc_all *= pl
        d2h = int(B * n * 8 * acc_all) + B * 8
        e2e = {"value": units_per_step * world * k_e2e / t_full, "unit": UNIT,
               "h2d_bytes_per_step": 0, "d2h_bytes_per_step": d2h, "steps": k_e2e,
               "ms_per_step": 1e3 * t_full / k_e2e,
               "d2h_gbs_per_gpu": d2h * k_e2e / t_full / 1e9,
               "d2h_note": "states of the chains whose draw was accepted (%.0f %% of %d chains x %.1f MiB; a rejected draw "
                           "leaves the caller's state untouched, as Sampler::draw does) + the QoI of every chain"
                           % (100 * acc_all, B, n * 8 / 2 ** 20),
               "api": "mlmcpi_sampler_draw_host_async + mlmcpi_sampler_wait_host: chains resident on the device, the new "
                      "states of the accepted chains and the QoI of all chains handed back to pinned host buffers every "
                      "step by the copy engine (one copy per run of accepted chains, issued when the step's accept flags "
                      "have reached the host), the copies of step k overlapped with the draw of step k+1",
               "inputs": "none per step: a sampler's only input is its own previous state (resident) and the Philox "
                         "counters; the per-step host traffic is the OUTPUT, as in Sampler::draw(state)",
               "limited_by": "the host link: %.2f GiB of states per step and GPU" % (B * n * 8 * acc_all / 2 ** 30),
               "host_link_gbs_per_gpu": link,
               "host_link_note": "copy-engine D2H of 1 GiB into the same pinned buffer, all ranks at the same time "
                                 "(rank 0's figure): the ceiling of d2h_gbs_per_gpu",
               "qoi_only": {"value": units_per_step * world * k_e2e / t_qoi, "ms_per_step": 1e3 * t_qoi / k_e2e,
                            "d2h_bytes_per_step": B * 8,
                            "note": "same API with h_x_out = NULL: the QoI of every chain to the host each step"},
               "numa": numa, "mean_qoi_on_host": chk}

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, which = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, which = 6650.0, "fallback (B200_PROFILING.md)"
    headline = is_schwinger and kind == mp.SAMPLER_HMC
    # ---- further measured legs (outside the timed region of `value`): every rank runs the ESS leg (its
    #      moments are all-reduced), rank 0 alone the single-GPU kernel legs
    ess = sweeps = beta4 = None
    if headline and a.ess_draws > 0:
        ess = ess_leg(mp, ctx, torch, dist, a, m, B, rank, world)
    if headline and not a.no_extra and rank == 0:
        sweeps = sweeps_leg(mp, ctx, torch, a, m, x, peak)
        beta4 = beta4_leg(mp, ctx, torch, a, rank)
    if world > 1:
        dist.barrier()
    mlmc = None
    if headline and not a.no_extra:  # every rank: the MLMC chains are sharded
        del x
        torch.cuda.empty_cache()
        mlmc = mlmc_leg(mp, torch, dist, a, local, rank, world)

    if rank == 0:
        achieved = lf_bytes / (lf_ms * 1e-3) / 1e9 if lf_ms > 0 else None
        if not is_schwinger:
            achieved = None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("leapfrog_dram_bytes_per_launch")
        steps_per_launch = (a.nt + 1) * a.steps / lf_launches if lf_launches else None
        roofline = {
            "bound": "hbm",
            "kernel": "leapfrog_rowpipek_kernel<4, 128, 8> (four leapfrog steps per launch on the 128^2 coarsest level; "
                      "the last, kick-only step of a trajectory runs in leapfrog_rowpipe_kernel)",
            # algorithmic bytes (64 B per site and leapfrog step) / CUDA-event time of the
            # leapfrog launches; above 1.0 because the kernel blocks four steps per pass over HBM
            "achieved": achieved, "peak": peak, "peak_source": which, "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None,
            "steps_per_hbm_pass": steps_per_launch,
            "traffic": traffic, "traffic_source": "static: ncu --set full capture of this kernel at this shape, "
                                                  "profiles/roofline_traffic.json (not re-measured in this run)",
            "launches": lf_launches,
            "avg_launch_ms": lf_ms / lf_launches if lf_launches else None,
            "algorithmic_bytes_per_launch": lf_bytes / lf_launches if lf_launches else None,
            "algorithmic_bytes_per_site_step": 64,
            "share_of_step": lf_ms / ms if ms > 0 else None,
        }
        cfg = workload_config(a)
        cfg.update({
            "chains_per_gpu": B, "parallelism": f"independent chains x{world} GPUs, NCCL allreduce of QoI moments",
            "l2": "inputs larger than L2 (per-GPU fine states %.1f GiB)" % (B * mp.sample_size(m) * 8 / 2 ** 30),
            "site_updates_per_step_per_gpu": {k: v for k, v in work.items()},
            "value_per_gpu": value / world,
            "acceptance_per_level": p_acc, "hmc_autotune": tuned,
        })
        timed_qoi = {"name": w["qoi"], "average": st["average"], "error": st["error"], "tau_int": st["tau_int"],
                     "samples": st["samples"]}
        if ess is not None:
            # the QoI / ESS figures of the line come from the ergodic sampler (ess_leg); the series of the
            # throughput leg is reported beside them and labelled for what it is
            cfg["qoi"] = ess["qoi"]
            cfg["ess_per_s"] = ess["ess_per_s"]
            cfg["ess"] = ess
            cfg["qoi_of_timed_steps"] = dict(timed_qoi, note=(
                "HMC coarse sampler at coarse beta ~ %.0f: no tunnelling between topological sectors (in the "
                "reference as here), every chain keeps the charge of its thermalised start state; the QoI and "
                "ESS/s of this line are measured with the cluster coarse sampler instead (config.ess)"
                % (a.beta / 4 ** (a.levels - 1))))
        else:
            cfg["qoi"] = timed_qoi
            cfg["ess_per_s"] = st["samples"] / st["tau_int"] / (ms_max * 1e-3)
        if sweeps is not None:
            cfg["extra"] = {"schwinger%d_beta4" % a.lattice: beta4}
        if mlmc is not None:
            cfg.setdefault("extra", {})["schwinger1024_mlmc"] = mlmc
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_max / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (U(-pi,pi) start states, Philox4x32-10)", "config": cfg,
            "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
        }
        if sweeps is not None:
            line["roofline_sweeps"] = sweeps
        if not is_schwinger or kind != mp.SAMPLER_HMC:
            line["metric"] = "lattice site-updates/s (%s)" % a.workload
            line["roofline"] = None  # 1-D paths / Gaussian fields: see profiles/r01_summary.md section 4
        if world == 1 and not a.no_cpu_baseline and is_schwinger and kind == mp.SAMPLER_HMC:
            # bounded sample: ~0.19 s per cascade and core -> about 0.75 x cpu_seconds of CPU work per core
            v, cores, kind, sample, _, _ = cpu_arm(a, max(1, int(a.cpu_seconds / 0.25)))
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """the one JSON line of the contract, on the real stdout"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def main():
    global _REAL_STDOUT
    # libraries (NCCL prints its version banner to stdout) must not pollute the one-line
    # contract: everything else that writes to fd 1 goes to stderr
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        reference_main(a)
    else:
        gpu_main(a)


if __name__ == "__main__":
    main()
