import sys, torch
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
ctx=mp.Context(0)
m=mp.schwinger(512,512,1024.0); B=64
x=ctx.state(m,B)
for k in range(3):
    ctx.overrelax_sweep(m,x); ctx.heatbath_sweep(m,x,0,k)
torch.cuda.synchronize()
