# reference's own classes: hierarchical two-level cascade at small lattice, chi_t vs analytic
import sys, numpy as np, time
sys.path.insert(0,'.')
from oracle import pyoracle as po
R = po.ref()
SCHW=3
def run(L,beta,ndraw,nburn,seed=1,renorm=1):
    rng=np.random.default_rng(seed)
    kind=[k for k in dir(po) if 'SCHW' in k.upper()]
    fine=R.action(po.SCHWINGER,[L,L,po.BOTH,renorm],[beta])
    coarse=fine.coarse()
    x=np.zeros(fine.n)
    # cold start -> heatbath thermalise
    x=fine.heatbath_sweep(x,50)
    S_f=fine.evaluate(x); S_c=fine.cond_evaluate(x)
    q=[];nacc=0
    for d in range(nburn+ndraw):
        xc=coarse.copy_from_fine(x)
        Sc_old=coarse.evaluate(xc)
        pc=coarse.heatbath_sweep(coarse.overrelax_sweep(xc,2),1)
        pc=coarse.heatbath_sweep(coarse.overrelax_sweep(pc,2),1)
        tp=fine.cond_fill(fine.copy_from_coarse(pc))
        S_fp=fine.evaluate(tp); S_cp=fine.cond_evaluate(tp)
        dS=(S_fp-S_f)+(Sc_old-coarse.evaluate(pc))+(S_c-S_cp)
        if dS<0 or rng.random()<np.exp(-dS):
            x=tp;S_f=S_fp;S_c=S_cp;nacc+=1
        if d>=nburn: q.append(fine.qoi(2,x))
    q=np.array(q)
    # binned error
    nb=50; b=q[:len(q)//nb*nb].reshape(nb,-1).mean(1)
    ex=R.lib.ref_schwinger_chit_analytical(beta,L*L)
    print(L,beta,"beta_c %.4f"%coarse.param(0),"p_acc %.3f"%(nacc/(nburn+ndraw)),"chi %.4f +/- %.4f exact %.4f"%(q.mean(),b.std(ddof=1)/np.sqrt(nb),ex),flush=True)
print([k for k in dir(po) if k.isupper()])
for L,beta in ((8,7.9),(8,9.0),(8,16.0)):
    t=time.time(); run(L,beta,int(sys.argv[1]),500); print(time.time()-t)
