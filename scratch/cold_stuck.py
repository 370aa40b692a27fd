# does the two-level step leave the COLD fine state?  (MLMC level 0 starts from theta = 0)
import sys, numpy as np, torch
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
ctx=mp.Context(0)
for L,beta in ((16,4.0),(16,6.0),(32,4.0)):
    B=1024
    fine=mp.schwinger(L,L,beta); coarse=mp.coarse_model(fine,renorm=mp.RENORM_PERTURBATIVE)
    cs=mp.Sampler(ctx,coarse,B,kind=mp.SAMPLER_HEATBATH,n_sweep_overrelax=2,n_sweep_heatbath=1)
    xc=ctx.init_state(coarse,B,0,0); cs.set_state(xc)
    for k in range(100): cs.draw(xc)
    for start in ("cold","thermal"):
        x=ctx.state(fine,B)
        if start=="thermal":
            x=ctx.init_state(fine,B,0,1)
            for k in range(100): ctx.heatbath_sweep(fine,x,0,k)
        Sf=ctx.action(fine,x); Sc=ctx.cond_action(fine,x)
        never=torch.ones(B,dtype=torch.bool,device='cuda'); hist=[]; accs=[]; dd=[]
        for d in range(1,2001):
            for r in range(3): cs.draw(xc)
            acc,deltas=ctx.twolevel_step(fine,coarse,xc,x,Sf,Sc,0,d)
            never&=(acc==0); accs.append(acc.double().mean().item())
            if d==1: dd=deltas[:4].cpu().numpy()
            if d in (1,10,100,300,1000,2000): hist.append((d,int(never.sum())))
        print(L,beta,start,"chains that never accepted after d draws:",hist,"mean acc first 100 %.3f last 100 %.3f"%(np.mean(accs[:100]),np.mean(accs[-100:])))
        print("   first-draw deltas (fine, coarse, trial):",np.round(dd,2).tolist(),flush=True)
