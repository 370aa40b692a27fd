import sys, torch, numpy as np
sys.path.insert(0,'.')
import mlmcpathintegral_b200 as mp
ctx=mp.Context(0)
# one-pass overrelaxation (shared-memory rings)
for Mt,Mx,B in ((6,4,2),(32,8,2),(64,34,2),(128,64,1)):
    m=mp.schwinger(Mt,Mx,2.0); x=ctx.init_state(m,B,0,1); ctx.overrelax_sweeps(m,x,3)
# fused fill + eval, leapfrog pipelines
m=mp.schwinger(64,64,100.0); mc=mp.coarse_model(m,renorm=mp.RENORM_PERTURBATIVE)
x=ctx.init_state(m,2,0,1); xc=ctx.state(mc,2); ctx.restrict(m,x,xc); y=ctx.state(m,2)
ctx.prolong_fill_eval(m,xc,y,0,3); p=ctx.hmc_momentum(m,2,0,1); ctx.leapfrog(m,5,0.01,x,p)
s=mp.Sampler(ctx,m,4,kind=mp.SAMPLER_HMC,n_levels=3,nt=5,dt=0.05,renorm=mp.RENORM_PERTURBATIVE)
xx=ctx.init_state(m,4,0,2); s.set_state(xx); s.draw(xx); s.draw(xx)
# fused 1-D hierarchy + register HMC
for mm,L in ((mp.rotor(128,4.0,0.25),3),(mp.ho(64),2)):
    s=mp.Sampler(ctx,mm,37,kind=mp.SAMPLER_HMC,n_levels=L,nt=6,dt=0.08,renorm=mp.RENORM_PERTURBATIVE)
    xx=ctx.state(mm,37); s.draw(xx); s.draw(xx)
# dense GFF, exact samplers
g=mp.coarse_model(mp.gff(16,16,3.0),ctype=mp.COARSEN_ROTATE); d=ctx.exact_draw(g,9,0,1); ctx.action(g,d)
ctx.exact_draw(mp.ho(32),7,0,1)
# GFF / Schwinger sweeps with the 3-D grids
gg=mp.gff(32,32,10.0); xg=ctx.init_state(gg,3,0,0); ctx.overrelax_sweep(gg,xg); ctx.heatbath_sweep(gg,xg,0,1)
g1=mp.coarse_model(gg,ctype=mp.COARSEN_ROTATE); g1.gff_n_gibbs=0; x1=ctx.init_state(g1,3,0,0); ctx.overrelax_sweep(g1,x1); ctx.heatbath_sweep(g1,x1,0,1)
m=mp.schwinger(12,20,2.0); x=ctx.init_state(m,3,0,1); ctx.heatbath_sweep(m,x,0,1)
ctx.sync(); print("sanitize workload done")
# end of round: one-pass GFF sweeps (OR + HB sequences, wrapped / chunked shapes), larger one-pass Schwinger OR,
# GFF hierarchical cascade on the cheaper index maps
for Mt,Mx,B in ((64,4,2),(64,6,3),(128,64,2),(96,34,1),(256,128,1)):
    gm=mp.gff(Mt,Mx,10.0); xg=ctx.init_state(gm,B,0,0); ctx.overrelax_sweeps(gm,xg,3)
    sg=mp.Sampler(ctx,gm,B,kind=mp.SAMPLER_HEATBATH,n_sweep_overrelax=2,n_sweep_heatbath=1); sg.set_state(xg); sg.draw(xg); sg.draw(xg); sg.close()
for Mt,Mx,B in ((512,128,1),(1024,64,1),(64,128,2)):
    m=mp.schwinger(Mt,Mx,2.0); x=ctx.init_state(m,B,0,1); ctx.overrelax_sweeps(m,x,2)
gh=mp.gff(32,32,10.0); sh=mp.Sampler(ctx,gh,5,kind=mp.SAMPLER_HEATBATH,n_levels=3,ctype=mp.COARSEN_ROTATE,n_sweep_overrelax=1,n_sweep_heatbath=1)
xh=ctx.init_state(gh,5,0,0); sh.set_state(xh); sh.draw(xh); sh.draw(xh)
ctx.sync(); print("sanitize workload 2 done")
