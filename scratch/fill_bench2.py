import sys, torch
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
ctx=mp.Context(0)
def timeit(f, n=5):
    f(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
for beta in (1024.0, 4.0):
    m=mp.schwinger(512,512,beta); mc=mp.coarse_model(m, renorm=mp.RENORM_PERTURBATIVE)
    B=256
    x=ctx.state(m,B) if beta>8 else ctx.init_state(m,B,0,1)
    for k in range(3):
        ctx.overrelax_sweep(m,x); ctx.heatbath_sweep(m,x,0,k)
    xc=ctx.state(mc,B); ctx.restrict(m,x,xc)
    y=ctx.state(m,B)
    d=[0]
    def pf():
        d[0]+=1; ctx.prolong_fill(m,xc,y,0,d[0])
    def pfe():
        d[0]+=1; ctx.prolong_fill_eval(m,xc,y,0,d[0])
    t_pf=timeit(pf); t_pfe=timeit(pfe)
    t_cond=timeit(lambda: ctx.cond_action(m,y))
    t_hb=timeit(lambda: ctx.heatbath_sweep(m,x,0,7))
    print(f"beta={beta}: prolong_fill {t_pf:.3f} ms  fused eval {t_pfe:.3f} ms  cond {t_cond:.3f} ms  heatbath sweep {t_hb:.3f} ms", flush=True)
