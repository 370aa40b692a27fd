import sys, torch
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
ctx=mp.Context(0)
def timeit(f, n=10):
    f(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
for name,m,B in (("rotor64",mp.rotor(64,4.0,0.25),8192),("rotor256",mp.rotor(256,4.0,0.25),8192),("ho32",mp.ho(32),65536),("quartic128",mp.quartic(128),8192),("rotor96",mp.rotor(96,4.0,0.25),8192)):
    x=ctx.init_state(m,B,0,1)
    d=[0]
    def f():
        d[0]+=1; ctx.hmc_step(m,100,0.05,x,0,d[0])
    t=timeit(f)
    print(f"{name}: hmc_step {t*1e3:.1f} us  {B*m.M_lat*101/t/1e6:.1f} G site-steps/s", flush=True)
