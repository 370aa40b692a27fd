import sys, numpy as np
sys.path.insert(0, '/root/repo')
import mlmcpathintegral_b200 as mp
from oracle import pyoracle as po
orc = po.oracle()
ctx = mp.Context(0, seed=0x5EED0001)
M = 16
m0 = mp.gff(M, M, 10.0)
s = mp.Sampler(ctx, m0, 4, kind=mp.SAMPLER_HEATBATH, n_levels=3, ctype=mp.COARSEN_ROTATE)
m1, m2 = s.level_model(1), s.level_model(2)
print("level1", m1.Mt_lat, m1.rotated, m1.gff_mu2, m1.gff_n_gibbs, "level2", m2.Mt_lat, m2.rotated, m2.gff_mu2, m2.gff_n_gibbs)
o0 = po.gff(M, M, 10.0)
o1 = orc.coarse_model(o0, 0, 0, po.ROTATE)
o2 = orc.coarse_model(o1, 0, 1, po.ROTATE)
Q1 = po.gff_dense_matrices(orc, o1, 2, 1.0)["Q_hat"]
Q2 = po.gff_dense_matrices(orc, o2, 2, 1.0)["Q_hat"]
rng = np.random.default_rng(5)
x1 = rng.normal(size=(4, mp.sample_size(m1)))
x2 = rng.normal(size=(4, mp.sample_size(m2)))
print("S1 dev", ctx.action(m1, ctx.to_device(x1)).cpu().numpy(), "np", 0.5 * np.einsum("bi,ij,bj->b", x1, Q1, x1))
print("S2 dev", ctx.action(m2, ctx.to_device(x2)).cpu().numpy(), "np", 0.5 * np.einsum("bi,ij,bj->b", x2, Q2, x2))
# restrict level1 -> level2, prolong, cond action on level 1
xr = ctx.state(m2, 4)
ctx.restrict(m1, ctx.to_device(x1), xr)
print("restrict err", np.max(np.abs(xr.cpu().numpy() - np.array([orc.restrict(o1, o2, x1[b]) for b in range(4)]))))
print("cond1 dev", ctx.cond_action(m1, ctx.to_device(x1)).cpu().numpy(), "orc", [orc.cond_action(o1, x1[b]) for b in range(4)])
xf = ctx.state(m1, 4)
ctx.prolong_fill(m1, ctx.to_device(x2), xf, 0, 7)
xo = np.array([orc.fill(o1, 0x5EED0001, 7, b, orc.prolong(o1, x2[b])) for b in range(4)]) if hasattr(orc, "fill") else None
if xo is not None:
    print("fill err", np.max(np.abs(xf.cpu().numpy() - xo)))
# two-level step at level 1 through the C-ABI
Sf = ctx.action(m1, ctx.to_device(x1)); Sc = ctx.cond_action(m1, ctx.to_device(x1))
x1d = ctx.to_device(x1)
acc, deltas = ctx.twolevel_step(m1, m2, ctx.to_device(x2), x1d, Sf, Sc, 0, 9)
print("deltas dev", deltas.cpu().numpy())
# by hand
tp = ctx.state(m1, 4); ctx.prolong_fill(m1, ctx.to_device(x2), tp, 0, 9)
tpn = tp.cpu().numpy()
d0 = 0.5 * np.einsum("bi,ij,bj->b", tpn, Q1, tpn) - 0.5 * np.einsum("bi,ij,bj->b", x1, Q1, x1)
thC = np.array([orc.restrict(o1, o2, x1[b]) for b in range(4)])
d1 = 0.5 * np.einsum("bi,ij,bj->b", thC, Q2, thC) - 0.5 * np.einsum("bi,ij,bj->b", x2, Q2, x2)
d2 = np.array([orc.cond_action(o1, x1[b]) - orc.cond_action(o1, tpn[b]) for b in range(4)])
print("deltas np ", np.stack([d0, d1, d2], axis=1))
# equilibrium check: run sampler and print acceptance
x = ctx.state(m0, 4)
s.close()
s = mp.Sampler(ctx, m0, 1024, kind=mp.SAMPLER_HEATBATH, n_levels=3, ctype=mp.COARSEN_ROTATE)
x = ctx.state(m0, 1024)
for k in range(300):
    s.draw(x)
    if k % 100 == 99:
        print(k, s.p_accept(), float(ctx.qoi(m0, mp.QOI_PHI2, x).mean()))
