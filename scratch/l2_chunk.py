import sys, torch
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
ctx=mp.Context(0)
def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
m=mp.schwinger(512,512,1024.0); B=256
x=ctx.state(m,B)
for k in range(2):
    ctx.overrelax_sweep(m,x); ctx.heatbath_sweep(m,x,0,k)
def run(chunk, n_or=10, n_hb=1):
    def f():
        for c0 in range(0,B,chunk):
            xs=x[c0:c0+chunk]
            ctx.overrelax_sweeps(m,xs,n_or)
            for k in range(n_hb): ctx.heatbath_sweep(m,xs,c0,5)
    return f
for chunk in (256,):
    t=timeit(run(chunk))
    t_or=timeit(run(chunk,10,0)); 
    print(f"chunk {chunk:4d} ({chunk*4} MiB): 10 OR + 1 HB = {t:.2f} ms ; 10 OR = {t_or:.2f} ms -> {B*512*512*10/t_or/1e6:.1f} G site-updates/s", flush=True)
mg=mp.gff(256,256,10.0); Bg=512
xg=ctx.init_state(mg,Bg,0,0)
def rung(chunk):
    def f():
        for c0 in range(0,Bg,chunk):
            xs=xg[c0:c0+chunk]
            for k in range(10): ctx.overrelax_sweep(mg,xs)
            ctx.heatbath_sweep(mg,xs,c0,5)
    return f
for chunk in (512,):
    print(f"gff chunk {chunk} ({chunk*0.5} MiB): {timeit(rung(chunk)):.2f} ms", flush=True)
