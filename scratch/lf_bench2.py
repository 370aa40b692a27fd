import sys, torch, numpy as np
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
from mlmcpathintegral_b200 import _lib
ctx=mp.Context(0)
def run(M,Mx,B,variant,fuse,R,nt=100,dt=0.001):
    m=mp.schwinger(M,Mx,64.0)
    ctx.set_option(_lib.OPT_LEAPFROG_VARIANT,variant); ctx.set_option(_lib.OPT_LEAPFROG_ROWS,R); ctx.set_option(_lib.OPT_LEAPFROG_FUSE,fuse)
    x=ctx.init_state(m,B,0,1); p=ctx.hmc_momentum(m,B,0,1)
    x0,p0=x.clone(),p.clone()
    ctx.leapfrog(m,7,0.01,x,p)
    res=(x.clone(),p.clone())
    x.copy_(x0); p.copy_(p0)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.leapfrog(m,nt,dt,x,p); e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)
    byts=16.0*M*Mx*B*(4*nt+3)
    return res, byts/ms/1e6, ms
for (M,Mx,B) in [(32,8,3),(64,40,2),(96,33,2),(128,128,512),(256,256,128),(512,512,64)]:
    ref,_,_=run(M,Mx,B,2,0,0,nt=5)
    for fuse,Rs in ((0,(0,)),(1,(8,16,32,64))):
        for R in Rs:
            if R>Mx and R!=0: continue
            res,gbs,ms=run(M,Mx,B,0,fuse,R)
            err=max(float((res[0]-ref[0]).abs().max()), float((res[1]-ref[1]).abs().max()))
            print(f"M={M}x{Mx} B={B} fuse={fuse} R={R}: {ms:8.2f} ms {gbs:8.1f} GB/s-equivalent  frac={gbs/6537:.3f}  maxdiff_vs_generic={err:.2e}", flush=True)
