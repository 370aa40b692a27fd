import sys, torch
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
ctx=mp.Context(0)
beta=float(sys.argv[1])
m=mp.schwinger(512,512,beta); mc=mp.coarse_model(m, renorm=mp.RENORM_PERTURBATIVE)
B=128
x=ctx.state(m,B) if beta>8 else ctx.init_state(m,B,0,1)
for k in range(3):
    ctx.overrelax_sweep(m,x); ctx.heatbath_sweep(m,x,0,k)
xc=ctx.state(mc,B); ctx.restrict(m,x,xc)
y=ctx.state(m,B)
ctx.prolong_fill(m,xc,y,0,1)
ctx.prolong_fill_eval(m,xc,y,0,2)
ctx.prolong_fill_eval(m,xc,y,0,3)
torch.cuda.synchronize()
