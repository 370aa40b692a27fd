import sys, numpy as np
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
from tests.util import load
ctx=mp.Context(0)
want=float.fromhex(load("scalars")["analytic"]["gff_phi_squared_10_16"])
m=mp.gff(16,16,10.0); B=2048
def run(**kw):
    s=mp.Sampler(ctx,m,B,ctype=mp.COARSEN_ROTATE,**kw)
    x=ctx.init_state(m,B,0,0); s.set_state(x)
    for _ in range(60): s.draw(x)
    vals=[]
    for _ in range(30):
        s.draw(x); vals.append(ctx.qoi(m,mp.QOI_PHI2,x).cpu().numpy())
    v=np.mean(vals,axis=0)
    print(kw, v.mean(), v.std(ddof=1)/np.sqrt(B), want, s.p_accept(), flush=True)
run(kind=mp.SAMPLER_HMC,n_levels=1,nt=20,dt=0.1)
run(kind=mp.SAMPLER_HMC,n_levels=2,nt=20,dt=0.1)
run(kind=mp.SAMPLER_HMC,n_levels=3,nt=20,dt=0.1)
run(kind=mp.SAMPLER_HEATBATH,n_levels=2,n_sweep_overrelax=0,n_sweep_heatbath=1)
run(kind=mp.SAMPLER_HEATBATH,n_levels=2,n_sweep_overrelax=1,n_sweep_heatbath=1)
