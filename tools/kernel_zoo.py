#!/usr/bin/env python
"""Launch every hot-path kernel once at the BASELINE.json shapes (after one warm-up call each)
so that one ncu pass can tabulate duration, DRAM bytes and pipe utilisation per kernel:

  ncu --metrics <list> --clock-control none --csv --log-file zoo.csv python tools/kernel_zoo.py
  python tools/kernel_zoo.py --summarise zoo.csv
"""
import collections
import csv
import re
import sys

METRICS = ("gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,"
           "smsp__issue_active.avg.pct_of_peak_sustained_active,"
           "sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,"
           "smsp__thread_inst_executed_per_inst_executed.ratio")


def run():
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import mlmcpathintegral_b200 as mp
    ctx = mp.Context(0)
    for rep in range(2):  # second repetition is the measured one (warm)
        # --- Schwinger 512^2 (C4), 64 chains; coarse level 256^2
        for beta in (1024.0, 4.0):
            m = mp.schwinger(512, 512, beta)
            mc = mp.coarse_model(m, renorm=mp.RENORM_PERTURBATIVE)
            B = 64
            x = ctx.init_state(m, B, 0, 1) if beta < 8 else ctx.state(m, B)
            ctx.heatbath_sweep(m, x, 0, 1)
            ctx.overrelax_sweep(m, x)
            p = ctx.hmc_momentum(m, B, 0, 1)
            ctx.action(m, x)
            ctx.force(m, x)
            ctx.leapfrog(m, 2, 0.01, x, p)
            ctx.qoi(m, mp.QOI_SCHWINGER_CHI, x)
            ctx.qoi(m, mp.QOI_AVG_PLAQUETTE, x)
            xc = ctx.state(mc, B)
            ctx.restrict(m, x, xc)
            y = ctx.state(m, B)
            ctx.prolong(m, xc, y)
            ctx.fill(m, y, 0, 2)
            ctx.prolong_fill(m, xc, y, 0, 3)
            ctx.prolong_fill_eval(m, xc, y, 0, 5)
            ctx.cond_action(m, y)
            Sf, Sc = ctx.action(m, x), ctx.cond_action(m, x)
            ctx.twolevel_step(m, mc, xc, x, Sf, Sc, 0, 4)
        # --- rotor M = 256, 8192 chains (C2)
        m = mp.rotor(256, 4.0, 0.25)
        mc = mp.coarse_model(m, renorm=mp.RENORM_PERTURBATIVE)
        B = 8192
        x = ctx.init_state(m, B, 0, 1)
        ctx.action(m, x)
        ctx.force(m, x)
        ctx.hmc_step(m, 100, 0.1, x, 0, 1)
        ctx.overrelax_sweep(m, x)
        ctx.heatbath_sweep(m, x, 0, 1)
        xc = ctx.state(mc, B)
        ctx.restrict(m, x, xc)
        ctx.prolong_fill(m, xc, x, 0, 2)
        ctx.cond_action(m, x)
        ctx.qoi(m, mp.QOI_ROTOR_CHI, x)
        ctx.cluster_update(m, x, 0, 0, 10)
        # --- harmonic oscillator M = 32 (C1), 65536 chains
        m = mp.ho(32)
        x = ctx.state(m, 65536)
        ctx.hmc_step(m, 100, 0.1, x, 0, 1)
        ctx.exact_draw(m, 65536, 0, 2)
        # --- GFF 32^2: the dense coarse level (512 vertices), 512 chains
        mc = mp.coarse_model(mp.gff(32, 32, 10.0), ctype=mp.COARSEN_ROTATE)
        x = ctx.exact_draw(mc, 512, 0, 1)
        ctx.action(mc, x)
        # --- GFF 256^2 (C3), 64 chains
        m = mp.gff(256, 256, 10.0)
        B = 64
        x = ctx.init_state(m, B, 0, 1)
        p = ctx.hmc_momentum(m, B, 0, 1)
        ctx.action(m, x)
        ctx.leapfrog(m, 2, 0.01, x, p)
        ctx.overrelax_sweep(m, x)
        ctx.heatbath_sweep(m, x, 0, 1)
        mc = mp.coarse_model(m, ctype=mp.COARSEN_ROTATE)
        xc = ctx.state(mc, B)
        ctx.restrict(m, x, xc)
        ctx.prolong_fill(m, xc, x, 0, 2)
        ctx.cond_action(m, x)
        ctx.qoi(m, mp.QOI_PHI2, x)
        ctx.sync()


def summarise(path, peak=6537.0):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = collections.OrderedDict()
    for r in rows:
        key = (r["ID"], r["Kernel Name"], r["Grid Size"], r["Block Size"])
        per.setdefault(key, {})[r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
    items = list(per.items())
    items = items[len(items) // 2:]  # second (warm) repetition

    def val(d, k):
        v, u = d.get(k, ("0", ""))
        v = float(v.replace(",", ""))
        if k.endswith("duration.sum"):
            return v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)  # us
        if "bytes" in k:
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return v
    print("| kernel | grid x block | us | DRAM GB/s | % of 6537 | fp64 pipe % | issue % | warps % | regs | thr/inst |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for (kid, name, grid, block), d in items:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short).replace("<unnamed>::", "")
        us = val(d, "gpu__time_duration.sum")
        byts = val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum")
        gbs = byts / us / 1e3 if us > 0 else 0
        if us < 4.0:
            continue
        print(f"| `{short[:70]}` | {grid} x {block} | {us:.1f} | {gbs:.0f} | {100 * gbs / peak:.0f} | "
              f"{val(d, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{val(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{val(d, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{val(d, 'launch__registers_per_thread'):.0f} | "
              f"{val(d, 'smsp__thread_inst_executed_per_inst_executed.ratio'):.1f} |")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
        summarise(sys.argv[2])
    elif len(sys.argv) > 1 and sys.argv[1] == "--metrics":
        print(METRICS)
    else:
        run()
