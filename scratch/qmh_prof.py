import sys, torch
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
ctx=mp.Context(0)
m=mp.rotor(256,4.0,0.25); B=8192
s=mp.Sampler(ctx,m,B,kind=mp.SAMPLER_HMC,n_levels=3,nt=100,dt=0.19,renorm=mp.RENORM_PERTURBATIVE)
x=ctx.init_state(m,B,0,1)
for k in range(5):
    ctx.overrelax_sweep(m,x); ctx.heatbath_sweep(m,x,0,k)
s.set_state(x)
for k in range(4): s.draw(x)
torch.cuda.synchronize()
