import numpy as np, sys
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
from oracle import pyoracle as po
orc=po.oracle()
ctx=mp.Context(0)
rng=np.random.default_rng(0)
for M,m0 in [(64,0.25),(64,20.0),(32,20.0),(96,1.0)]:
    o=po.rotor(M,4.0,m0); m=mp.rotor(M,4.0,m0)
    x=rng.uniform(-3,3,(1,M)); p=rng.normal(size=(1,M))
    for nt in (0,1,2,7):
        xd,pd=ctx.to_device(x),ctx.to_device(p)
        ctx.leapfrog(m,nt,0.05,xd,pd)
        xo,po_=orc.leapfrog(o,nt,0.05,x[0],p[0])
        print(M,m0,nt,np.max(np.abs(xd.cpu().numpy()[0]-xo)),np.max(np.abs(pd.cpu().numpy()[0]-po_)))
