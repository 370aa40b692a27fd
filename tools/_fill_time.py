import os, sys, torch
sys.path.insert(0, "/root/repo")
import mlmcpathintegral_b200 as mp
ctx = mp.Context(0)
L = int(os.environ.get("MLMCPI_L", "512")); B = int(os.environ.get("MLMCPI_B", "512"))
beta = float(os.environ.get("MLMCPI_BETA", "1024"))
m = mp.schwinger(L, L, beta)
mc = mp.coarse_model(m, renorm=mp.RENORM_PERTURBATIVE)
xc = ctx.init_state(mc, B, 0, 0)
for k in range(3):
    ctx.heatbath_sweep(mc, xc, 0, k)
x = ctx.state(m, B)
def run():
    if os.environ.get("MLMCPI_NOEVAL"):
        return ctx.prolong_fill(m, xc, x, 0, 5)
    return ctx.prolong_fill_eval(m, xc, x, 0, 5)
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = int(os.environ.get("MLMCPI_N", "5"))
e0.record()
for _ in range(n): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print("lib", os.environ.get("MLMCPI_LIB", "default").split("/")[-1], "fill+eval L", L, "B", B, "beta", beta, "%.3f ms, %.1f G fine sites/s" % (ms, L * L * B / ms / 1e6))
