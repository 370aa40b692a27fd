# diagnose the beta > 8 (ApproximateBesselProduct) hierarchical path at 8^2
import sys, numpy as np, torch
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
from oracle import pyoracle as po
R=po.ref()
ctx=mp.Context(0)
def chi_stats(qs):
    q=torch.stack(qs).double()          # [n][B]
    m=q.mean(0); return m.mean().item(), (m.std()/np.sqrt(m.numel())).item()
for L,beta in ((8,9.0),(8,16.0)):
    B=4096
    fine=mp.schwinger(L,L,beta); coarse=mp.coarse_model(fine,renorm=mp.RENORM_PERTURBATIVE)
    ex=R.lib.ref_schwinger_chit_analytical(beta,L*L)
    rfine=R.action(po.SCHWINGER,[L,L,po.BOTH,1],[beta])
    x=ctx.init_state(fine,B,0,0)
    for k in range(50): ctx.heatbath_sweep(fine,x,0,k)
    # --- (A) explicit cascade, stand-alone evaluations, torch accept
    xa=x.clone(); Sf=ctx.action(fine,xa); Sc=ctx.cond_action(fine,xa)
    xc=ctx.state(coarse,B); tp=ctx.state(fine,B); tp2=ctx.state(fine,B)
    g=torch.Generator(device='cuda'); g.manual_seed(1)
    qs=[];nacc=0;ntot=0;maxdf=0;maxdc=0;nbad=0;worst=None
    for d in range(1,1501):
        ctx.restrict(fine,xa,xc); ScC=ctx.action(coarse,xc)
        for r in range(2):
            ctx.overrelax_sweeps(coarse,xc,2); ctx.heatbath_sweep(coarse,xc,0,1000+2*d+r)
        Scc=ctx.action(coarse,xc)
        ctx.prolong_fill(fine,xc,tp,0,d)
        Sfp=ctx.action(fine,tp); Scp=ctx.cond_action(fine,tp)
        # fused evaluation on the same draw
        Sfe,Sce=ctx.prolong_fill_eval(fine,xc,tp2,0,d)
        assert torch.equal(tp,tp2)
        df=(Sfe-Sfp).abs(); dc=(Sce-Scp).abs()
        maxdf=max(maxdf,df.max().item()); maxdc=max(maxdc,dc.max().item())
        bad=(df>1e-8)|(dc>1e-8); nbad+=int(bad.sum())
        if bad.any() and worst is None:
            i=int(torch.nonzero(bad)[0]); worst=(d,i,xc[i].cpu().numpy().copy(),tp[i].cpu().numpy().copy(),Sfe[i].item(),Sfp[i].item(),Sce[i].item(),Scp[i].item())
        dS=(Sfp-Sf)+(ScC-Scc)+(Sc-Scp)
        acc=(dS<0)|(torch.rand(B,generator=g,device='cuda',dtype=torch.float64)<torch.exp(-dS))
        xa[acc]=tp[acc]; Sf[acc]=Sfp[acc]; Sc[acc]=Scp[acc]
        nacc+=int(acc.sum()); ntot+=B
        if d>500: qs.append(ctx.qoi(fine,mp.QOI_SCHWINGER_CHI,xa).clone())
    m,e=chi_stats(qs)
    print(L,beta,"A explicit standalone: p_acc %.3f chi %.4f +/- %.4f exact %.4f"%(nacc/ntot,m,e,ex))
    print("   fused-vs-standalone max |dS_f| %.3e max |dS_cond| %.3e  n_bad %d of %d"%(maxdf,maxdc,nbad,ntot))
    if worst:
        d,i,xcw,tpw,a,b,c_,d_=worst
        print("   first bad: draw",d,"chain",i,"Sf fused %.12g standalone %.12g ref %.12g | Scond fused %.12g standalone %.12g ref %.12g"%(a,b,rfine.evaluate(tpw),c_,d_,rfine.cond_evaluate(tpw)))
        np.save("gpurun_out/bad_xc_%g.npy"%beta,xcw); np.save("gpurun_out/bad_tp_%g.npy"%beta,tpw)
    # reference check of stand-alone on last theta'
    tpn=tp.cpu().numpy(); e1=max(abs(rfine.evaluate(tpn[i])-Sfp[i].item()) for i in range(64)); e2=max(abs(rfine.cond_evaluate(tpn[i])-Scp[i].item()) for i in range(64))
    print("   standalone vs reference on theta': %.3e %.3e"%(e1,e2))
    # --- (B) ctx.twolevel_step (fused inside, no caches)
    xb=x.clone(); Sf=ctx.action(fine,xb); Sc=ctx.cond_action(fine,xb); qs=[];nacc=0
    for d in range(1,1501):
        ctx.restrict(fine,xb,xc)
        for r in range(2):
            ctx.overrelax_sweeps(coarse,xc,2); ctx.heatbath_sweep(coarse,xc,0,1000+2*d+r)
        acc,_=ctx.twolevel_step(fine,coarse,xc,xb,Sf,Sc,0,d); nacc+=int(acc.sum())
        if d>500: qs.append(ctx.qoi(fine,mp.QOI_SCHWINGER_CHI,xb).clone())
    m,e=chi_stats(qs)
    print(L,beta,"B twolevel_step: p_acc %.3f chi %.4f +/- %.4f exact %.4f"%(nacc/(1500*B),m,e,ex))
    print("   cache drift: |Sf - action| %.3e |Sc - cond| %.3e"%((Sf-ctx.action(fine,xb)).abs().max().item(),(Sc-ctx.cond_action(fine,xb)).abs().max().item()))
    # --- (C) Sampler
    for kind,name in ((mp.SAMPLER_HEATBATH,"heatbath"),):
        s=mp.Sampler(ctx,fine,B,kind=kind,n_levels=2,renorm=mp.RENORM_PERTURBATIVE,n_sweep_overrelax=2,n_sweep_heatbath=1)
        xs=x.clone(); s.set_state(xs); qs=[]
        for d in range(1500):
            s.draw(xs)
            if d>=500: qs.append(ctx.qoi(fine,mp.QOI_SCHWINGER_CHI,xs).clone())
        m,e=chi_stats(qs)
        print(L,beta,"C Sampler",name,"p_acc",[round(p,3) for p in s.p_accept()],"chi %.4f +/- %.4f exact %.4f"%(m,e,ex),flush=True)
