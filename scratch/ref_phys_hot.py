import sys, numpy as np, time
sys.path.insert(0,'.')
from oracle import pyoracle as po
R = po.ref()
def run(L,beta,nchain,ndraw,nburn):
    fine=R.action(po.SCHWINGER,[L,L,po.BOTH,1],[beta]); coarse=fine.coarse()
    rng=np.random.default_rng(3)
    qm=[];acc=0
    for c in range(nchain):
        x=rng.uniform(-np.pi,np.pi,fine.n)
        S_f=fine.evaluate(x); S_c=fine.cond_evaluate(x); q=[]
        for d in range(nburn+ndraw):
            xc=coarse.copy_from_fine(x); Sc_old=coarse.evaluate(xc)
            pc=coarse.heatbath_sweep(coarse.overrelax_sweep(xc,2),1)
            tp=fine.cond_fill(fine.copy_from_coarse(pc))
            S_fp=fine.evaluate(tp); S_cp=fine.cond_evaluate(tp)
            dS=(S_fp-S_f)+(Sc_old-coarse.evaluate(pc))+(S_c-S_cp)
            if dS<0 or rng.random()<np.exp(-dS):
                x=tp;S_f=S_fp;S_c=S_cp;acc+=1
            if d>=nburn: q.append(fine.qoi(2,x))
        qm.append(np.mean(q))
    qm=np.array(qm)
    print(L,beta,"hot start: p_acc %.3f chi %.4f +/- %.4f exact %.4f ; chains with mean chi>1: %d of %d"%(acc/(nchain*(nburn+ndraw)),qm.mean(),qm.std()/np.sqrt(nchain),R.lib.ref_schwinger_chit_analytical(beta,L*L),(qm>1).sum(),nchain),flush=True)
run(8,9.0,64,1000,1500)
run(8,16.0,64,1000,1500)
