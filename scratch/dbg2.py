import numpy as np, sys
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
from oracle import pyoracle as po
from tests.util import load, unhex, qm_model
orc=po.oracle()
ctx=mp.Context(0)
c=[c for c in load("qm") if c["name"]=="rotor_large"][0]
o=qm_model(po,c)
m=mp.Model(model=o.model, M_lat=o.M_lat, a_lat=o.a_lat, T_final=o.T_final, m0=o.m0)
x=unhex(c["x"])[None,:]; p=unhex(c["p0"])[None,:]
for nt in (0,1,2,3,4,5,6,7):
    xd,pd=ctx.to_device(x),ctx.to_device(p)
    ctx.leapfrog(m,nt,0.05,xd,pd)
    xo,po_=orc.leapfrog(o,nt,0.05,x[0],p[0])
    print(nt,np.max(np.abs(xd.cpu().numpy()[0]-xo)),np.max(np.abs(pd.cpu().numpy()[0]-po_)), np.max(np.abs(po_)))
