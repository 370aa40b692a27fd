import os, sys, torch
sys.path.insert(0, "/root/repo")
import mlmcpathintegral_b200 as mp
ctx = mp.Context(0)
L = int(os.environ.get("MLMCPI_L", "512")); B = int(os.environ.get("MLMCPI_B", "512"))
beta = float(os.environ.get("MLMCPI_BETA", "1024"))
m = mp.schwinger(L, L, beta)
x = ctx.init_state(m, B, 0, 0)
for k in range(3):
    ctx.heatbath_sweep(m, x, 0, k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for k in range(n): ctx.heatbath_sweep(m, x, 0, 10 + k)
e1.record(); torch.cuda.synchronize()
print("lib", os.environ.get("MLMCPI_LIB", "default").split("/")[-1], "heat-bath sweep L", L, "B", B, "%.3f ms" % (e0.elapsed_time(e1) / n))
