import sys, torch
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
ctx=mp.Context(0)
m=mp.schwinger(512,512,1024.0); B=128
x=ctx.init_state(m,B,0,1)
ctx.overrelax_sweeps(m,x,4)
torch.cuda.synchronize()
