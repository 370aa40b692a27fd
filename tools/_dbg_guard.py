"""guard-band probe: every GFF entry point of the cascade on buffers with sentinel bands on both sides"""
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, "/root/repo")
import mlmcpathintegral_b200 as mp
from mlmcpathintegral_b200 import _lib
L = _lib.lib
ctx = mp.Context(0)
B, G = 12, 4096
SENT = 7777.25
class Guarded:
    def __init__(self, n):
        self.n = n
        self.t = torch.full((n + 2 * G,), SENT, dtype=torch.float64, device="cuda")
    @property
    def ptr(self): return C.c_void_p(self.t.data_ptr() + 8 * G)
    def body(self): return self.t[G:G + self.n]
    def check(self, what):
        lo, hi = self.t[:G], self.t[G + self.n:]
        bad = int((lo != SENT).sum()) + int((hi != SENT).sum())
        if bad:
            idx_hi = torch.nonzero(hi != SENT).flatten()[:6].tolist()
            idx_lo = torch.nonzero(lo != SENT).flatten()[:6].tolist()
            print("   !!! OVERRUN in", what, "below:", idx_lo, "above:", idx_hi, "values", hi[hi != SENT][:4].tolist())
            self.t[:G] = SENT; self.t[G + self.n:] = SENT
        return bad
def ck(rc):
    if rc: raise RuntimeError(L.mlmcpi_last_error(ctx.h).decode())
for smoothing in (1, 0):
    ctx.set_option(_lib.OPT_GFF_COARSE_SMOOTHING, smoothing)
    for M, mass, Lv in ((16, 3.0, 3), (32, 10.0, 4)):
        m0 = mp.gff(M, M, mass, mp.COARSEN_ROTATE)
        s = mp.Sampler(ctx, m0, B, n_levels=Lv, chain0=3, kind=mp.SAMPLER_HEATBATH, n_sweep_overrelax=1, n_sweep_heatbath=1)
        models = [s.level_model(l) for l in range(Lv)]
        s.close()
        print("smoothing", smoothing, "M", M, [(mm.Mt_lat, mm.Mx_lat, mm.rotated, mm.gff_n_gibbs, mp.sample_size(mm)) for mm in models])
        for l, m in enumerate(models):
            N = mp.sample_size(m)
            x = Guarded(N * B); S = Guarded(2 * B)
            bufs = {"x": x, "S": S}
            def run(name, f):
                ck(f()); torch.cuda.synchronize()
                for k, b in bufs.items(): b.check(f"level {l} {name} [{k}]")
                v = float(x.body().abs().max())
                if not np.isfinite(v) or v > 50: print("   ??? level", l, name, "max|x|", v)
            run("init_state", lambda: L.mlmcpi_init_state(ctx.h, C.byref(m), x.ptr, B, 3, 0))
            run("heatbath", lambda: L.mlmcpi_heatbath_sweep(ctx.h, C.byref(m), x.ptr, B, 3, 5))
            run("overrelax", lambda: L.mlmcpi_overrelax_sweeps(ctx.h, C.byref(m), x.ptr, B, 3))
            run("action", lambda: L.mlmcpi_action(ctx.h, C.byref(m), x.ptr, B, S.ptr))
            run("exact_draw", lambda: L.mlmcpi_exact_draw(ctx.h, C.byref(m), x.ptr, B, 3, 7))
            run("hmc_step", lambda: L.mlmcpi_hmc_step(ctx.h, C.byref(m), 5, 0.1, x.ptr, B, 3, 9, None, None))
            if l + 1 < Lv:
                mc = models[l + 1]
                xc = Guarded(mp.sample_size(mc) * B); bufs["xc"] = xc
                run("init_state(coarse)", lambda: L.mlmcpi_init_state(ctx.h, C.byref(mc), xc.ptr, B, 3, 0))
                run("prolong_fill", lambda: L.mlmcpi_prolong_fill(ctx.h, C.byref(m), xc.ptr, x.ptr, B, 3, 11))
                run("prolong_fill_eval", lambda: L.mlmcpi_prolong_fill_eval(ctx.h, C.byref(m), xc.ptr, x.ptr, B, 3, 11, S.ptr))
                run("cond_action", lambda: L.mlmcpi_cond_action(ctx.h, C.byref(m), x.ptr, B, S.ptr))
                run("restrict", lambda: L.mlmcpi_restrict(ctx.h, C.byref(m), x.ptr, xc.ptr, B))
                run("prolong", lambda: L.mlmcpi_prolong(ctx.h, C.byref(m), xc.ptr, x.ptr, B))
                run("fill", lambda: L.mlmcpi_fill(ctx.h, C.byref(m), x.ptr, B, 3, 13))
print("done")
