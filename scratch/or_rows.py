import sys, os, torch, subprocess
if len(sys.argv)>1:
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
    for (Mt,Mx,B) in ((512,512,256),(1024,1024,64)):
        m=mp.schwinger(Mt,Mx,4.0); x=ctx.init_state(m,B,0,1)
        t=timeit(lambda: ctx.overrelax_sweeps(m,x,10))
        print(f"R={os.environ.get('MLMCPI_OR_ROWS')} {Mt}x{Mx} B={B}: {t:.3f} ms = {B*Mt*Mx*320/t/1e6/65.37:.0f}%",flush=True)
else:
    for R in (16,32,64,128,256):
        subprocess.run([sys.executable,__file__,"x"],env=dict(os.environ,MLMCPI_OR_ROWS=str(R)))
