import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import mlmcpathintegral_b200 as mp
ctx = mp.Context(0)
def run(name, M, par, L, kw, B=12, n=6):
    m = mp.schwinger(M, M, par) if name == "schwinger" else mp.gff(M, M, par, mp.COARSEN_ROTATE)
    if name == "schwinger": kw = dict(kw, renorm=mp.RENORM_PERTURBATIVE)
    res = []
    for cache in (0, 1):
        ctx.set_option(mp._lib.OPT_CASCADE_CACHE, cache)
        s = mp.Sampler(ctx, m, B, n_levels=L, chain0=3, **(kw if name == "schwinger" else dict(kw, ctype=mp.COARSEN_ROTATE)))
        x = s.get_state()
        st = [x.cpu().numpy().copy()]
        for d in range(n):
            s.draw(x); st.append(x.cpu().numpy().copy())
        res.append((st, s.p_accept())); s.close()
    ctx.set_option(mp._lib.OPT_CASCADE_CACHE, 1)
    print(name, M, L, kw.get("kind"), "p_accept", res[0][1], res[1][1])
    for d, (a, b) in enumerate(zip(res[0][0], res[1][0])):
        diff = np.abs(a - b).max(axis=1)
        print("  draw", d - 1, "chains differing:", int((diff > 1e-9).sum()), "max", diff.max())
HB = dict(kind=mp.SAMPLER_HEATBATH, n_sweep_overrelax=2, n_sweep_heatbath=1)
run("gff", 16, 3.0, 2, HB)
run("gff", 16, 3.0, 3, HB)
run("gff", 16, 3.0, 2, dict(kind=mp.SAMPLER_EXACT))
run("gff", 32, 10.0, 4, HB)
run("gff", 16, 3.0, 3, dict(kind=mp.SAMPLER_HMC, nt=10, dt=0.1))
