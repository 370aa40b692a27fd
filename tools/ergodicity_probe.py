#!/usr/bin/env python
"""Acceptance / chi_t / tau_int of the quenched Schwinger hierarchical sampler along the
continuum-limit line beta / P = const (qoi/qft/qoi2dsusceptibility.hh:27-29), for the coarse
samplers the reference offers (hierarchical: coarsesampler = 'HMC' | 'cluster' | 'heatbath').

    python tools/ergodicity_probe.py [--sizes 64,128,256,512] [--ratio 256] [--samplers cluster,HMC]

Prints one JSON line per (lattice, sampler); profiles/r02_ergodicity.md is built from them.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch

    import mlmcpathintegral_b200 as mp
    from mlmcpathintegral_b200 import _lib

    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="64,128,256,512")
    ap.add_argument("--ratio", type=float, default=256.0, help="P / beta")
    ap.add_argument("--beta", type=float, default=None, help="fixed beta instead of the fixed ratio")
    ap.add_argument("--samplers", default="cluster,HMC,heatbath")
    ap.add_argument("--levels", type=int, default=3)
    ap.add_argument("--chains", type=int, default=512)
    ap.add_argument("--burnin", type=int, default=50)
    ap.add_argument("--draws", type=int, default=200)
    ap.add_argument("--n-updates", type=int, default=100)
    ap.add_argument("--window", type=int, default=20)
    a = ap.parse_args()
    ctx = mp.Context(0, seed=0x5EED0001)
    kinds = {"cluster": mp.SAMPLER_CLUSTER, "HMC": mp.SAMPLER_HMC, "heatbath": mp.SAMPLER_HEATBATH}
    for L in [int(s) for s in a.sizes.split(",")]:
        beta = a.beta if a.beta else L * L / a.ratio
        exact = _lib.lib.mlmcpi_schwinger_chit_analytical(beta, L * L)
        for name in a.samplers.split(","):
            B = a.chains if L <= 256 else max(64, a.chains // 4)
            m = mp.schwinger(L, L, beta)
            s = mp.Sampler(ctx, m, B, kind=kinds[name], n_levels=a.levels, nt=100, dt=0.1,
                           renorm=mp.RENORM_PERTURBATIVE, n_sweep_overrelax=10, n_sweep_heatbath=1,
                           n_updates=a.n_updates)
            tuned = None
            if name == "HMC":
                dt0 = 0.1
                for _ in range(12):
                    s.set_dt(dt0)
                    dt_t, p_t, ok = s.autotune(0.8, 8, 2 * B)
                    if ok or p_t > 0.8:
                        break
                    dt0 *= 0.5
                tuned = dt_t
            x = s.get_state()
            for _ in range(a.burnin):
                s.draw(x)
            st = mp.Statistics(ctx, a.window, B)
            ctx.sync()
            t0 = time.perf_counter()
            qs = []
            for _ in range(a.draws):
                s.draw(x)
                q = ctx.qoi(m, mp.QOI_SCHWINGER_CHI, x)
                st.record(q)
                qs.append(q.clone())
            ctx.sync()
            dt = time.perf_counter() - t0
            r = mp.Statistics.finalize(st.pack(), a.window)
            qs = torch.stack(qs).cpu().numpy()  # [draw][chain]
            # ensemble estimate (independent chains): mean over chains of the per-chain means
            cm = qs.mean(axis=0)
            out = dict(lattice=L, beta=beta, coarse_sampler=name, levels=a.levels, chains=B, draws=a.draws,
                       n_updates=a.n_updates if name == "cluster" else None, hmc_dt=tuned,
                       acceptance=s.p_accept(), chit=r["average"], chit_err_tau=r["error"],
                       chit_err_chains=float(cm.std(ddof=1) / np.sqrt(B)), tau_int=r["tau_int"],
                       exact=exact, ms_per_draw=1e3 * dt / a.draws,
                       frac_chains_moving=float((qs.std(axis=0) > 0).mean()))
            print(json.dumps(out), flush=True)
            st.close()
            s.close()
            del x
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
