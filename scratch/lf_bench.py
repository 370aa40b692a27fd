import sys, torch, numpy as np
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
from mlmcpathintegral_b200 import _lib
ctx=mp.Context(0)
def run(M,B,variant,R,nt=100,check=None):
    m=mp.schwinger(M,M,64.0)
    ctx.set_option(_lib.OPT_LEAPFROG_VARIANT,variant); ctx.set_option(_lib.OPT_LEAPFROG_ROWS,R)
    x=ctx.init_state(m,B,0,1); p=ctx.hmc_momentum(m,B,0,1)
    x0,p0=x.clone(),p.clone()
    ctx.leapfrog(m,3,0.01,x,p)   # warm
    res=(x.clone(),p.clone())
    x.copy_(x0); p.copy_(p0)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.leapfrog(m,nt,0.001,x,p); e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)
    byts=16.0*M*M*B*(4*nt+3)
    return res, byts/ms/1e6
for (M,B) in [(128,512),(128,256),(512,64),(256,128)]:
    ref,_=run(M,B,2,0,nt=5)
    for variant in (1,0):
        for R in (4,8,16,32):
            res,gbs=run(M,B,variant,R)
            err=max(float((res[0]-ref[0]).abs().max()), float((res[1]-ref[1]).abs().max()))
            print(f"M={M} B={B} variant={variant} R={R}: {gbs:8.1f} GB/s  frac={gbs/6537:.3f}  maxdiff_vs_generic={err:.2e}", flush=True)
