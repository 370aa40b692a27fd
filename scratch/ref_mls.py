# MultilevelSampler::draw (sampler/multilevelsampler.cc:71-112) walked with the REFERENCE's own classes
# (rotor M = 32, 3 levels, HMC coarse sampler nt = 20, dt = 0.1), fixed thresholds ceil(tau_int) = (2, 3, 2)
# as observed in the steady state: is the ~2 % low chi_t of the level walk the algorithm's?
import sys, numpy as np, time
sys.path.insert(0,'.')
from oracle import pyoracle as po
R=po.ref()
M=32
a0=R.action(po.ROTOR,[M,1],[4.0,0.25]); a1=a0.coarse(); a2=a1.coarse()
acts=[a0,a1,a2]
want=R.lib.ref_rotor_chit(a0.h,0)
thr=[int(v) for v in sys.argv[2:5]] if len(sys.argv)>4 else [2,3,2]
def run(ndraw,seed):
    rng=np.random.default_rng(seed)
    st=[np.zeros(a.n) for a in acts]
    st[2]=rng.uniform(-np.pi,np.pi,acts[2].n)
    Sf=[acts[l].evaluate(st[l]) for l in range(2)]; Sc=[acts[l].cond_evaluate(st[l]) for l in range(2)]
    t=[0,0,0]; q=[]; hs=seed*1000003
    for d in range(ndraw):
        level=2
        while level>=0:
            if level==2:
                hs+=1
                _,st[2],_,_=acts[2].hmc_draws(20,0.1,1,hs,st[2])
            else:
                f,c=acts[level],acts[level+1]
                tp=f.cond_fill(f.copy_from_coarse(st[level+1]))
                Sfp=f.evaluate(tp); Scp=f.cond_evaluate(tp)
                dS=(Sfp-Sf[level])+(c.evaluate(c.copy_from_fine(st[level]))-c.evaluate(st[level+1]))+(Sc[level]-Scp)
                if dS<0 or rng.random()<np.exp(-dS):
                    st[level]=tp; Sf[level]=Sfp; Sc[level]=Scp
            t[level]+=1
            if t[level]>=thr[level]:
                t[level]=0; level-=1
            else:
                level=2
        q.append(a0.qoi(1,st[0]))
    return np.array(q)
n=int(sys.argv[1]); t0=time.time()
q=run(n,5)[n//20:]
nb=40; b=q[:len(q)//nb*nb].reshape(nb,-1).mean(1)
print("thresholds",thr,"chi %.5f +/- %.5f want %.5f (%.1f sigma)  %.0f s"%(q.mean(),b.std(ddof=1)/np.sqrt(nb),want,(q.mean()-want)/(b.std(ddof=1)/np.sqrt(nb)),time.time()-t0))
