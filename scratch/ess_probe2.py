import sys, time, torch
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
ctx=mp.Context(0)
for L,beta,levels in ((16,4.0,2),(32,4.0,2),(32,16.0,2),(64,16.0,2),(64,16.0,3),(128,64.0,2),(256,256.0,2)):
    B=512
    m=mp.schwinger(L,L,beta)
    ref=mp._lib.lib.mlmcpi_schwinger_chit_analytical(beta, L*L)
    for kind,name in ((mp.SAMPLER_CLUSTER,"cluster"),(mp.SAMPLER_HMC,"hmc")):
        s=mp.Sampler(ctx,m,B,kind=kind,n_levels=levels,nt=20,dt=0.1,renorm=mp.RENORM_PERTURBATIVE,n_updates=10)
        x=ctx.init_state(m,B,0,0)
        for k in range(20):
            ctx.overrelax_sweep(m,x); ctx.heatbath_sweep(m,x,0,1000+k)
        s.set_state(x)
        st=mp.Statistics(ctx,20,B)
        for k in range(50): s.draw(x)
        for k in range(100):
            s.draw(x); st.record(ctx.qoi(m,mp.QOI_SCHWINGER_CHI,x))
        out=mp.Statistics.finalize(st.pack(),20)
        print(L,beta,levels,name,"p_acc",[round(p,3) for p in s.p_accept()],"chi %.4f +/- %.4f (exact %.4f) tau %.2f"%(out['average'],out['error'],ref,out['tau_int']), flush=True)
