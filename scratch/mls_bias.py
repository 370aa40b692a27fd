import sys, numpy as np, torch
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
from tests.util import load
ctx=mp.Context(0)
want=float.fromhex(load("scalars")["analytic"]["rotor_chit_exact_32"])
m=mp.rotor(32,4.0,0.25)
for B,warm,N,chain0 in ((4096,40,100,0),(4096,400,100,0),(1024,40,100,1024),(4096,40,100,1<<20)):
    s=mp.Sampler(ctx,m,B,kind=mp.SAMPLER_HMC,n_levels=3,nt=20,dt=0.1,renorm=mp.RENORM_PERTURBATIVE,multilevel=True,qoi=mp.QOI_ROTOR_CHI,n_autocorr_window=10,chain0=chain0)
    x=ctx.state(m,B)
    for _ in range(warm): s.draw(x)
    pc=torch.zeros(B,dtype=torch.float64,device='cuda'); blocks=[]
    for k in range(N):
        s.draw(x); q=ctx.qoi(m,mp.QOI_ROTOR_CHI,x); pc+=q/N
        blocks.append(q.mean().item())
    mean=pc.mean().item(); err=pc.std().item()/np.sqrt(B)
    b=np.array(blocks).reshape(5,-1).mean(1)
    print("B",B,"warm",warm,"chain0",chain0,"mean %.5f +/- %.5f want %.5f (%.1f sigma)"%(mean,err,want,(mean-want)/err),"blocks",np.round(b,4).tolist(),"indep",[round(t,2) for t in s.independence()[0]],"p_acc",[round(p,3) for p in s.p_accept()],flush=True)
    s.close()
