import sys, torch, numpy as np
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
from oracle import pyoracle as po
ctx=mp.Context(0); orc=po.oracle()
m=mp.gff(4,4,10.0); B=70000
x=ctx.init_state(m,B,0,1)
S=ctx.action(m,x); q=ctx.qoi(m,mp.QOI_PHI2,x)
o=po.gff(4,4,10.0)
xs=x.cpu().numpy()
for b in (0,1,65534,65535,65536,69999):
    assert abs(S[b].item()-orc.action(o,xs[b]))<1e-12*max(1,abs(S[b].item())), b
    assert abs(q[b].item()-(xs[b]**2).mean())<1e-13
mc=mp.coarse_model(m,ctype=mp.COARSEN_ROTATE)
Sc=ctx.cond_action(m,x)
for b in (0,65535,65536,69999):
    assert abs(Sc[b].item()-orc.cond_action(o,xs[b]))<1e-12*max(1,abs(Sc[b].item())), b
print("GFF reductions with 70000 chains ok")
