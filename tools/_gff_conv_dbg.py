import sys, numpy as np
sys.path.insert(0, '/root/repo')
import mlmcpathintegral_b200 as mp
ctx = mp.Context(0, seed=0x5EED0001)
def probe(M, kind, levels, B=2048, n=300, **kw):
    m = mp.gff(M, M, 10.0)
    exact = mp._lib.lib.mlmcpi_gff_phi_squared_analytical(10.0, M, M)
    s = mp.Sampler(ctx, m, B, kind=kind, n_levels=levels, ctype=mp.COARSEN_ROTATE, nt=20, dt=0.2, **kw)
    x = s.get_state()
    out = []
    for k in range(n):
        s.draw(x)
        if k in (0, 1, 10, 100, n - 1):
            out.append("%d: %.4f" % (k, float(ctx.qoi(m, mp.QOI_PHI2, x).mean())))
    print(M, kind, levels, "exact %.4f" % exact, out, "acc", np.round(s.p_accept(), 3), flush=True)
    s.close()
probe(16, mp.SAMPLER_HMC, 2)
probe(16, mp.SAMPLER_HMC, 3)
probe(16, mp.SAMPLER_HMC, 3)
probe(16, mp.SAMPLER_HEATBATH, 3)
probe(32, mp.SAMPLER_HEATBATH, 4, B=1024)
probe(256, mp.SAMPLER_HEATBATH, 4, B=64, n=100)
probe(256, mp.SAMPLER_EXACT, 2, B=64, n=60)
