"""The N > 1 path on CPU (gloo, world_size 2): chains are sharded over ranks with no data-path
collective; the only exchange is the allreduce(sum) of the packed moment vector that
mlmcpi_stats_pack produces and mlmcpi_stats_finalize consumes (the quantities Statistics
averages over MPI ranks in the reference, common/statistics.cc:30-35,64-79)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K_MAX = 8


def packed_from_chains(q):
    """host restatement of what the device accumulators hold: q[n_samples, n_chains] ->
    {n_chains, n_longterm, n_shortterm, sum avg, sum avg1..4, sum S_k} (statistics.cc:4-27 per chain)"""
    n, B = q.shape
    S = np.zeros((K_MAX, B))
    avg = np.zeros((4, B))
    hist = []
    for s in range(1, n + 1):
        Q = q[s - 1]
        hist.insert(0, Q)
        hist = hist[:K_MAX]
        for p in range(4):
            avg[p] = ((s - 1.0) * avg[p] + Q ** (p + 1)) / s
        for k in range(len(hist)):
            Nk = s - k
            S[k] = ((Nk - 1.0) * S[k] + hist[0] * hist[k]) / Nk
    return np.concatenate([[B, n * B, n * B, avg[0].sum()], avg.sum(axis=1), S.sum(axis=1)])


def chains(n, B, chain0):
    """deterministic AR(1) chains keyed by the GLOBAL chain index (sharding-invariant)"""
    out = np.zeros((n, B))
    for b in range(B):
        rng = np.random.default_rng(1000 + chain0 + b)
        v = 0.0
        for s in range(n):
            v = 0.6 * v + rng.normal()
            out[s, b] = v
    return out


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import mlmcpathintegral_b200 as mp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, n = 3, 120
    local = packed_from_chains(chains(n, B, rank * B))  # rank r owns chains [rB, (r+1)B)
    t = torch.from_numpy(local.copy())
    dist.all_reduce(t)  # the one collective of the path
    out = mp.Statistics.finalize(t.numpy(), K_MAX)
    ret[rank] = (out, local)
    dist.destroy_process_group()


def test_sharded_statistics_allreduce_gloo():
    import mlmcpathintegral_b200 as mp
    world, B, n = 2, 3, 120
    mgr = tmp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    tmp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    # both ranks hold the same global estimators
    for k in ret[0][0]:
        assert ret[0][0][k] == ret[1][0][k]
    # ... equal to the single-process result over all 6 chains (sharding invariance)
    allq = np.concatenate([chains(n, B, r * B) for r in range(world)], axis=1)
    want = mp.Statistics.finalize(packed_from_chains(allq), K_MAX)
    for k, v in want.items():
        assert abs(ret[0][0][k] - v) <= 1e-12 * max(abs(v), 1.0), k
    assert want["samples"] == world * B * n
    # ... and the global mean is the reference's unweighted mean over ranks of per-rank means
    per_rank = [mp.Statistics.finalize(ret[r][1], K_MAX)["average"] for r in range(world)]
    assert abs(np.mean(per_rank) - want["average"]) < 1e-13


def test_chain_sharding_offsets():
    """bench.py's sharding: rank r uses chain0 = r * B, so Philox streams never overlap"""
    B, world = 512, 8
    owned = [set(range(r * B, (r + 1) * B)) for r in range(world)]
    assert len(set().union(*owned)) == B * world
