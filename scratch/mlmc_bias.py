import sys, numpy as np, torch
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
ctx=mp.Context(0)
L,beta=16,4.0
fine=mp.schwinger(L,L,beta); coarse=mp.coarse_model(fine,renorm=mp.RENORM_PERTURBATIVE)
ex=mp._lib.lib.mlmcpi_schwinger_chit_analytical(beta,L*L)
# (1) stationarity of the MLMC-type two-level step: fine chains start from an equilibrated ensemble
B=8192
hs=mp.Sampler(ctx,fine,B,kind=mp.SAMPLER_HMC,n_levels=2,renorm=mp.RENORM_PERTURBATIVE,nt=20,dt=0.1)
x=ctx.init_state(fine,B,0,0); hs.set_state(x)
for k in range(400): hs.draw(x)
q0=ctx.qoi(fine,mp.QOI_SCHWINGER_CHI,x); print("equilibrated ensemble: Q0 %.4f +/- %.4f exact %.4f"%(q0.mean().item(),q0.std().item()/np.sqrt(B),ex))
for n_sub in (1,3,10):
    cs=mp.Sampler(ctx,coarse,B,kind=mp.SAMPLER_HMC,nt=100,dt=0.1,chain0=1<<24)
    xc=ctx.init_state(coarse,B,1<<24,0); cs.set_state(xc)
    for k in range(200): cs.draw(xc)
    xf=x.clone(); Sf=ctx.action(fine,xf); Sc=ctx.cond_action(fine,xf)
    out=[]; accs=[]
    for d in range(1,601):
        for r in range(n_sub): cs.draw(xc)
        acc,_=ctx.twolevel_step(fine,coarse,xc,xf,Sf,Sc,0,d+1000*n_sub); accs.append(acc.double().mean().item())
        if d in (1,10,30,100,200,400,600):
            q=ctx.qoi(fine,mp.QOI_SCHWINGER_CHI,xf); qc=ctx.qoi(coarse,mp.QOI_SCHWINGER_CHI,xc)
            out.append("d=%d: Q0 %.3f+/-%.3f Q1 %.3f"%(d,q.mean().item(),q.std().item()/np.sqrt(B),qc.mean().item()))
    print("coarse draws between proposals:",n_sub,"acc %.3f"%np.mean(accs)); print("   "+" | ".join(out),flush=True)
# (2) the MLMC estimator against n_burnin
for nb in (100,1000):
    mc=mp.MultilevelMC(ctx,fine,128,n_level=2,epsilon=0.05,qoi=mp.QOI_SCHWINGER_CHI,n_burnin=nb,n_autocorr_window=20,
        n_min_samples_qoi=100,max_iterations=50,kind=mp.SAMPLER_HMC,nt=100,dt=0.1,renorm=mp.RENORM_PERTURBATIVE)
    conv=mc.evaluate(); v,e,lv=mc.result()
    print("MLMC n_burnin",nb,"value %.4f +/- %.4f exact %.4f (%.1f sigma)"%(v,e,ex,(v-ex)/e),[ (int(l['samples']),round(l['mean'],4),round(l['tau_int'],2)) for l in lv],flush=True)
