import sys, torch
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
ctx=mp.Context(0)
for M,mass in ((16,10.0),(32,10.0),(64,10.0),(128,10.0),(256,10.0),(256,160.0)):
    m=mp.gff(M,M,mass)
    B=64
    s=mp.Sampler(ctx,m,B,kind=mp.SAMPLER_HEATBATH,n_levels=2,ctype=mp.COARSEN_ROTATE,n_sweep_overrelax=1,n_sweep_heatbath=1)
    x=ctx.init_state(m,B,0,0)
    for k in range(50):
        ctx.overrelax_sweep(m,x); ctx.heatbath_sweep(m,x,0,k)
    s.set_state(x)
    for k in range(20): s.draw(x)
    print(M,mass,s.p_accept(), float(ctx.qoi(m,mp.QOI_PHI2,x).mean()), mp._lib.lib.mlmcpi_gff_phi_squared_analytical(mass,M,M))
