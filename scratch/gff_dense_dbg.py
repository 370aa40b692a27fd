import sys, torch, numpy as np
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
ctx=mp.Context(0)
for M,levels,mass in ((16,2,10.0),(16,3,10.0),(32,3,10.0),(32,2,10.0)):
    mm=mp.gff(M,M,mass); Bc=1024
    smp=mp.Sampler(ctx,mm,Bc,kind=mp.SAMPLER_EXACT,n_levels=levels,ctype=mp.COARSEN_ROTATE)
    st=mp.Statistics(ctx,20,Bc)
    xx=ctx.init_state(mm,Bc,0,0); smp.set_state(xx)
    for k in range(200):
        smp.draw(xx)
        if k>=50: st.record(ctx.qoi(mm,mp.QOI_PHI2,xx))
    out=mp.Statistics.finalize(st.pack(),20)
    ref=mp._lib.lib.mlmcpi_gff_phi_squared_analytical(mass,M,M)
    print(M,levels,[ (smp.level_model(l).gff_n_gibbs, mp.sample_size(smp.level_model(l))) for l in range(levels)], smp.p_accept(), out['average'],out['error'],ref,(out['average']-ref)/out['error'], out['tau_int'])
    # coarsest-level exact sampler alone: phi^2 of the coarse distribution
