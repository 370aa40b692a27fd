import sys, time, torch
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
ctx=mp.Context(0)
L=512; beta=1024.0; B=512
m=mp.schwinger(L,L,beta)
ref=mp._lib.lib.mlmcpi_schwinger_chit_analytical(beta, L*L)
print("analytic E[V chi_t] =", ref)
for kind,name,levels in ((mp.SAMPLER_CLUSTER,"cluster",3),(mp.SAMPLER_CLUSTER,"cluster",4),(mp.SAMPLER_HMC,"hmc6",6)):
    s=mp.Sampler(ctx,m,B,kind=kind,n_levels=levels,nt=100,dt=0.1,renorm=mp.RENORM_PERTURBATIVE,n_updates=10)
    if kind==mp.SAMPLER_HMC:
        dt0=0.1
        for _ in range(12):
            s.set_dt(dt0); dt_t,p_t,ok=s.autotune(0.8,8,2*B)
            if ok or p_t>0.8: break
            dt0*=0.5
        print("tuned",dt_t,p_t)
    x=ctx.state(m,B)
    for k in range(10):
        ctx.overrelax_sweep(m,x); ctx.heatbath_sweep(m,x,0,1000+k)
    s.set_state(x)
    st=mp.Statistics(ctx,20,B)
    for k in range(30): s.draw(x)
    torch.cuda.synchronize(); t0=time.time()
    n=60
    for k in range(n):
        s.draw(x); st.record(ctx.qoi(m,mp.QOI_SCHWINGER_CHI,x))
    torch.cuda.synchronize(); dt=time.time()-t0
    out=mp.Statistics.finalize(st.pack(),20)
    print(name,levels,"ms/draw",1e3*dt/n,"p_acc",s.p_accept(),"chi",out['average'],"+/-",out['error'],"tau",out['tau_int'],"ESS/s",out['samples']/out['tau_int']/dt, flush=True)
