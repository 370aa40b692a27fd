import sys, numpy as np, time
sys.path.insert(0,'.')
from oracle import pyoracle as po
O=po.oracle(); R=po.ref()
def run(L,beta,ndraw,nburn,seed=7,env=2):
    O.lib.orc_set_expcos_envelope(env)
    fine=po.schwinger(L,L,beta); coarse=O.coarse_model(fine,renorm=1)
    x=np.zeros(O.sample_size(fine))
    for k in range(50): x=O.heatbath_sweep(fine,seed,k,0,x)
    S_f=O.action(fine,x); S_c=O.cond_action(fine,x)
    q=[];nacc=0;dr=100
    for d in range(nburn+ndraw):
        xc=O.restrict(fine,coarse,x)
        for r in range(2):
            xc=O.overrelax_sweep(coarse,O.overrelax_sweep(coarse,xc)); dr+=1
            xc=O.heatbath_sweep(coarse,seed,dr,0,xc)
        dr+=1
        acc,x,S_f,S_c,out=O.twolevel_step(fine,coarse,seed,dr,0,xc,x,S_f,S_c)
        nacc+=acc
        if d>=nburn: q.append(O.qoi(fine,po.QOI_SCHWINGER_CHI,x)[0])
    q=np.array(q); nb=50; b=q[:len(q)//nb*nb].reshape(nb,-1).mean(1)
    ex=R.lib.ref_schwinger_chit_analytical(beta,L*L)
    print(L,beta,"env",env,"beta_c %.4f"%coarse.beta,"p_acc %.3f"%(nacc/(nburn+ndraw)),"chi %.4f +/- %.4f exact %.4f"%(q.mean(),b.std(ddof=1)/np.sqrt(nb),ex),flush=True)
for L,beta in ((8,7.9),(8,9.0),(8,16.0)):
    run(L,beta,int(sys.argv[1]),500)
