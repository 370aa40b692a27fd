import sys, time, torch
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
ctx=mp.Context(0)
for L,beta,kind,name in ((8,9.0,mp.SAMPLER_CLUSTER,"cluster"),(8,16.0,mp.SAMPLER_CLUSTER,"cluster"),(8,16.0,mp.SAMPLER_HEATBATH,"heatbath"),(16,16.0,mp.SAMPLER_CLUSTER,"cluster"),(16,32.0,mp.SAMPLER_CLUSTER,"cluster"),(8,7.9,mp.SAMPLER_CLUSTER,"cluster-besselpath")):
    B=4096
    m=mp.schwinger(L,L,beta)
    ref=mp._lib.lib.mlmcpi_schwinger_chit_analytical(beta, L*L)
    s=mp.Sampler(ctx,m,B,kind=kind,n_levels=2,renorm=mp.RENORM_PERTURBATIVE,n_updates=20,n_sweep_overrelax=2,n_sweep_heatbath=1)
    x=ctx.init_state(m,B,0,0); s.set_state(x)
    st=mp.Statistics(ctx,100,B)
    for k in range(1500): s.draw(x)
    for k in range(3000):
        s.draw(x); st.record(ctx.qoi(m,mp.QOI_SCHWINGER_CHI,x))
    out=mp.Statistics.finalize(st.pack(),100)
    print(L,beta,name,"p_acc",[round(p,3) for p in s.p_accept()],"chi %.4f +/- %.4f exact %.4f  dev %.1f sigma  tau %.1f"%(out['average'],out['error'],ref,(out['average']-ref)/out['error'],out['tau_int']), flush=True)
