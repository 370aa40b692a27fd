import sys, torch
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
from mlmcpathintegral_b200 import _lib
ctx=mp.Context(0)
def timeit(f, n=5):
    f(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
for (Mt,Mx,B) in ((256,256,512),(256,256,666),(256,256,333),(256,256,1024),(512,512,128)):
    m=mp.gff(Mt,Mx,10.0)
    x0=ctx.init_state(m,B,0,1)
    res={}
    for v in (0,1):
        ctx.set_option(_lib.OPT_OVERRELAX_ONE_PASS,v)
        x=x0.clone(); ctx.overrelax_sweeps(m,x,3); res[v]=x.clone()
        xe=x0.clone(); ctx.overrelax_sweeps(m,xe,4); res[(v,'e')]=xe
        t=timeit(lambda: ctx.overrelax_sweeps(m,x,10))
        gb=B*Mt*Mx*16*10/t/1e6
        print(f"gff {Mt}x{Mx} B={B} one_pass={v}: 10 OR sweeps {t:.3f} ms = {gb:.0f} GB/s algorithmic ({gb/65.37:.0f}% of 6537)",flush=True)
    print("   bit-identical (3 sweeps, 4 sweeps):",torch.equal(res[1],res[0]),torch.equal(res[(1,'e')],res[(0,'e')]), "max diff", (res[1]-res[0]).abs().max().item())
    # sampler: n_or + n_hb sequence
    outs={}
    for v in (0,1):
        ctx.set_option(_lib.OPT_OVERRELAX_ONE_PASS,v)
        s=mp.Sampler(ctx,m,B,kind=mp.SAMPLER_HEATBATH,n_sweep_overrelax=3,n_sweep_heatbath=2)
        x=x0.clone(); s.set_state(x)
        for k in range(3): s.draw(x)
        outs[v]=x.clone()
        s10=mp.Sampler(ctx,m,B,kind=mp.SAMPLER_HEATBATH,n_sweep_overrelax=10,n_sweep_heatbath=1)
        s10.set_state(x)
        t=timeit(lambda: s10.draw(x))
        print(f"   sampler draw (10 OR + 1 HB) one_pass={v}: {t:.3f} ms",flush=True)
        s.close(); s10.close()
    print("   sampler sequence bit-identical:",torch.equal(outs[0],outs[1]),(outs[0]-outs[1]).abs().max().item())
ctx.set_option(_lib.OPT_OVERRELAX_ONE_PASS,1)
