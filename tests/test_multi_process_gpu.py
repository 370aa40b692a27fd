"""Multilevel Monte Carlo and the MultilevelSampler over MORE THAN ONE process (BASELINE config
C5): chains sharded over the processes, every Statistics query of a host-side decision taken over the
chains of all of them (mlmcpi_set_allreduce), so all processes walk through the same allocation loop.
Two processes share cuda:0 here and exchange through gloo (NCCL refuses two ranks on one device); on a
multi-GPU box the transport is ncclAllReduce (libmlmcpi_comm.so: mlmcpi_comm_attach) -- the library
code under test is the same."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import load  # noqa: E402

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import mlmcpathintegral_b200 as mp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = mp.Context(0)
    ctx.attach_process_group()
    assert ctx.world_size == world
    m = mp.rotor(32, 4.0, 0.25)
    B = 1024
    # rank r owns the global chains [rB, (r+1)B)
    mc = mp.MultilevelMC(ctx, m, B, n_level=3, epsilon=2e-3, qoi=mp.QOI_ROTOR_CHI, n_burnin=30,
                         n_autocorr_window=10, n_min_samples_qoi=4 * B * world, max_iterations=20,
                         chain0=rank * B, kind=mp.SAMPLER_HMC, nt=20, dt=0.1, renorm=mp.RENORM_PERTURBATIVE)
    converged = mc.evaluate()
    value, error, levels = mc.result()
    # MultilevelSampler: the level walk follows tau_int over all processes
    s = mp.Sampler(ctx, m, B, kind=mp.SAMPLER_HMC, n_levels=3, nt=20, dt=0.1, renorm=mp.RENORM_PERTURBATIVE,
                   multilevel=1, qoi=mp.QOI_ROTOR_CHI, n_autocorr_window=10, chain0=rank * B)
    x = ctx.init_state(m, B, rank * B, 0)
    s.set_state(x)
    st = mp.Statistics(ctx, 10, B)
    for _ in range(40):
        s.draw(x)
    per_chain = torch.zeros(B, dtype=torch.float64, device=x.device)
    for _ in range(60):
        s.draw(x)
        q = ctx.qoi(m, mp.QOI_ROTOR_CHI, x)
        st.record(q)
        per_chain += q / 60
    # error from the scatter of the chain means over ALL chains (no autocorrelation estimate needed)
    mom = torch.stack([per_chain.sum(), (per_chain ** 2).sum()]).cpu()
    dist.all_reduce(mom)
    n_all = B * world
    mean_all = mom[0].item() / n_all
    err_all = ((mom[1].item() / n_all - mean_all ** 2) / (n_all - 1)) ** 0.5
    packed = torch.from_numpy(st.pack().copy())
    dist.all_reduce(packed)
    ms = mp.Statistics.finalize(packed.numpy(), 10)
    ret[rank] = dict(converged=converged, value=value, error=error, levels=levels, launches=ctx.launches,
                     indep=s.independence(), ms=ms, mean_all=mean_all, err_all=err_all)
    dist.barrier()
    dist.destroy_process_group()


def test_multilevel_mc_two_processes_lockstep():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    world = 2
    mgr = tmp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    tmp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    a, b = ret[0], ret[1]
    # identical decisions and identical global estimators on both processes
    assert a["converged"] and b["converged"], (a["levels"], b["levels"])
    assert a["value"] == b["value"] and a["error"] == b["error"]
    for la, lb in zip(a["levels"], b["levels"]):
        assert la["samples"] == lb["samples"] and la["n_target"] == lb["n_target"]
        assert la["mean"] == lb["mean"] and la["variance"] == lb["variance"] and la["tau_int"] == lb["tau_int"]
        assert la["samples"] >= la["n_target"]
        assert la["samples"] % (2 * 1024) == 0  # every batched draw adds B samples on each process
    assert a["launches"] == b["launches"]  # the same sequence of kernels
    want = float.fromhex(load("scalars")["analytic"]["rotor_chit_exact_32"])
    assert a["error"] < 3e-3 and abs(a["value"] - want) < 5 * a["error"] + 1e-3, (a["value"], a["error"], want)
    # MultilevelSampler: same level walk, consistent estimate
    assert a["indep"] == b["indep"]
    assert a["ms"]["samples"] == 2 * 1024 * 60
    assert abs(a["ms"]["average"] - a["mean_all"]) < 1e-12
    # The level walk proposes states that are only ~tau_int apart, which leaves chi_t about 2 % low -- in the
    # reference's algorithm itself: its own classes walked by multilevelsampler.cc:71-112 with the thresholds
    # ceil(tau_int) = (2, 3, 2) observed here give 0.15178 +/- 0.00033 over 380 000 draws (scratch/ref_mls.py)
    # against the exact 0.15485.  Parity is with that value.
    ref_walk, ref_walk_err = 0.15178, 0.00033
    assert abs(a["mean_all"] - ref_walk) < 5 * (a["err_all"] ** 2 + ref_walk_err ** 2) ** 0.5, (a["mean_all"], a["err_all"])
    assert abs(a["mean_all"] - want) < 0.05 * want
