import sys, torch
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
ctx=mp.Context(0)
m=mp.gff(256,256,10.0); B=512
x=ctx.init_state(m,B,0,1)
ctx.overrelax_sweeps(m,x,4)
torch.cuda.synchronize()
