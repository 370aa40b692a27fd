#!/usr/bin/env python
"""Record known-answer vectors from the REFERENCE's own translation units.

Runs only where /root/reference exists (the build container): it drives
oracle/_ref/libmlmcpi_ref.so (reference .cc files compiled unmodified against the
Eigen/GSL header shims, see oracle/Makefile) and writes tests/golden/*.json.
Floats are stored as C99 hex strings so the fixtures are bit-exact.

Values that involve a Bessel function or erf come from the shim
(std::cyl_bessel_i / std::erf), not from GSL: tests treat those as 1e-12-relative,
everything else as exact reference arithmetic.

    python tools/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def hx(a):
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    return [float(v).hex() for v in a.ravel()]


def angles(n, shift=0.0):
    ell = np.arange(n, dtype=np.float64)
    x = 2.0 * np.sin(0.37 * ell + 0.11 + shift) + 0.5 * np.cos(1.3 * ell)
    return x - 2 * np.pi * np.floor(0.5 * (x + np.pi) / np.pi)


def noncompact(n, shift=0.0):
    ell = np.arange(n, dtype=np.float64)
    return 1.5 * np.sin(0.37 * ell + 0.11 + shift) + 0.25 * np.cos(1.3 * ell)


def lattice_cases(R):
    import ctypes as C
    cases = []
    for (Mt, Mx, ctype, level) in [(8, 8, po.BOTH, 0), (8, 8, po.BOTH, 1), (6, 4, po.TEMPORAL, 0),
                                   (4, 6, po.SPATIAL, 0), (8, 12, po.ALTERNATE, 0),
                                   (8, 12, po.ALTERNATE, 1), (8, 8, po.ROTATE, 0),
                                   (8, 8, po.ROTATE, 1), (8, 8, po.ROTATE, 2), (16, 8, po.ROTATE, 1),
                                   (4, 4, po.ROTATE, 1), (5, 7, po.BOTH, 0)]:
        info = (C.c_int * 6)()
        if R.lib.ref_lattice2d_info(Mt, Mx, ctype, level, info) != 0:
            continue
        mt, mx, rot, nv, ne, has_coarse = list(info)
        lo, hi_t, hi_x = -3, mt + 3, mx + 3
        n_cart = (hi_t - lo) * (hi_x - lo)
        c2l = (C.c_uint * n_cart)()
        l2c = (C.c_int * (2 * nv))()
        nb = (C.c_uint * (8 * nv))()
        R.lib.ref_lattice2d_vertex_maps(Mt, Mx, ctype, level, lo, hi_t, lo, hi_x, c2l, l2c, nb)
        case = dict(Mt0=Mt, Mx0=Mx, ctype=ctype, level=level, Mt=mt, Mx=mx, rotated=rot,
                    n_vertices=nv, n_edges=ne, has_coarse=has_coarse, lo=lo,
                    vertex_cart2lin=list(c2l), vertex_lin2cart=list(l2c), neighbours=list(nb))
        if not rot:
            lc2l = (C.c_uint * (2 * n_cart))()
            ll2c = (C.c_int * (3 * ne))()
            R.lib.ref_lattice2d_link_maps(Mt, Mx, ctype, level, lo, hi_t, lo, hi_x, lc2l, ll2c)
            case["link_cart2lin"] = list(lc2l)
            case["link_lin2cart"] = list(ll2c)
        if has_coarse:
            co, fo = (C.c_uint * nv)(), (C.c_uint * nv)()
            mk, mv = (C.c_uint * nv)(), (C.c_uint * nv)()
            cnt = (C.c_int * 3)()
            R.lib.ref_lattice2d_coarsening(Mt, Mx, ctype, level, co, fo, mk, mv, cnt)
            case["coarse"] = list(co)[:cnt[0]]
            case["fineonly"] = list(fo)[:cnt[1]]
            case["map_keys"] = list(mk)[:cnt[2]]
            case["map_vals"] = list(mv)[:cnt[2]]
        cases.append(case)
    one_d = []
    for M in (8, 32, 7):
        out = (C.c_double * 1)()
        nb = (C.c_uint * (2 * M))()
        Mc = R.lib.ref_lattice1d(M, 4.0, out, nb)
        one_d.append(dict(M=M, T=4.0, a_lat=float(out[0]).hex(), neighbours=list(nb), M_coarse=Mc))
    return dict(lattice2d=cases, lattice1d=one_d)


def qm_cases(R):
    cases = []
    specs = [
        ("ho", po.HO, [16, 0], [4.0, 1.0, 1.0]),
        ("ho_renorm1", po.HO, [32, 1], [4.0, 1.3, 0.7]),
        ("ho_renorm2", po.HO, [32, 2], [4.0, 1.3, 0.7]),
        ("quartic", po.QUARTIC, [16, 0], [4.0, 1.0, 1.0, 1.0, 1.0]),
        ("quartic2", po.QUARTIC, [32, 0], [3.0, 0.8, -1.0, 0.5, 0.3]),
        ("rotor", po.ROTOR, [16, 0], [4.0, 0.25]),
        ("rotor_renorm1", po.ROTOR, [32, 1], [4.0, 0.25]),
        ("rotor_large", po.ROTOR, [64, 0], [4.0, 20.0]),  # sigma/2 > 100: series branch
    ]
    for name, kind, ip, dp in specs:
        a = R.action(kind, ip, dp)
        n = a.n
        x = angles(n) if kind == po.ROTOR else noncompact(n)
        p0 = noncompact(n, 0.7)
        c = dict(name=name, kind=kind, ip=ip, dp=dp, x=hx(x), p0=hx(p0))
        c["S"] = hx(a.evaluate(x))
        c["force"] = hx(a.force(x))
        c["W"] = [hx(a.W(xm, xp)) for (xm, xp) in [(0.3, -0.2), (-1.1, 2.5), (3.0, -3.0)]]
        c["cond_S"] = hx(a.cond_evaluate(x))
        c["qoi_x2"] = hx(a.qoi(po.QOI_X2, x))
        if kind == po.ROTOR:
            c["qoi_chi"] = hx(a.qoi(po.QOI_ROTOR_CHI, x))
            c["overrelax_lex"] = hx(a.overrelax_sweep(x))
            even = np.arange(0, n, 2)
            odd = np.arange(1, n, 2)
            c["overrelax_coloured"] = hx(a.overrelax_sweep(x, idx=np.concatenate([even, odd])))
        xl, pl = a.leapfrog(7, 0.05, x, p0)
        c["leapfrog"] = dict(nt=7, dt=0.05, x=hx(xl), p=hx(pl))
        ac = a.coarse()
        c["coarse_m0"] = hx(ac.param(0))
        xc = ac.copy_from_fine(x)
        c["restrict"] = hx(xc)
        c["prolong"] = hx(a.copy_from_coarse(noncompact(n // 2, 0.3), x))
        c["coarse_S"] = hx(ac.evaluate(xc))
        tp = angles(n, 0.4) if kind == po.ROTOR else noncompact(n, 0.4)
        pc = angles(n // 2, 0.9) if kind == po.ROTOR else noncompact(n // 2, 0.9)
        c["twolevel"] = dict(theta_prime=hx(tp), phi_coarse=hx(pc),
                             deltas=hx(a.twolevel_deltas(ac, x, tp, pc)))
        cases.append(c)
    return cases


def schwinger_cases(R):
    cases = []
    specs = [("both_b4", 8, 8, po.BOTH, 0, 4.0), ("both_b16", 8, 8, po.BOTH, 0, 16.0),
             ("both_b1200", 8, 8, po.BOTH, 0, 1200.0),
             ("both_rect", 8, 12, po.BOTH, 0, 2.5), ("both_renorm1", 8, 8, po.BOTH, 1, 6.0),
             ("temporal", 8, 6, po.TEMPORAL, 0, 3.0), ("spatial", 6, 8, po.SPATIAL, 0, 3.0),
             ("alternate", 8, 8, po.ALTERNATE, 0, 9.0), ("alternate_renorm1", 8, 8, po.ALTERNATE, 1, 9.0)]
    for name, Mt, Mx, ctype, renorm, beta in specs:
        a = R.action(po.SCHWINGER, [Mt, Mx, ctype, renorm], [beta])
        n = a.n
        # large beta: a smooth state, otherwise every pdf underflows to zero
        scale = 0.02 if beta > 100 else 1.0
        x = scale * angles(n)
        p0 = noncompact(n, 0.7)
        c = dict(name=name, Mt=Mt, Mx=Mx, ctype=ctype, renorm=renorm, beta=beta, scale=scale,
                 x=hx(x), p0=hx(p0))
        c["S"] = hx(a.evaluate(x))
        c["force"] = hx(a.force(x))
        c["qoi_chi"] = hx(a.qoi(po.QOI_SCHWINGER_CHI, x))
        c["qoi_plaq"] = hx(a.qoi(po.QOI_AVG_PLAQUETTE, x))
        c["overrelax_lex"] = hx(a.overrelax_sweep(x))
        ell = np.arange(n)
        j = ell // (2 * Mt)
        i = (ell % (2 * Mt)) // 2
        mu = ell % 2
        col = np.where(mu == 0, j % 2, 2 + i % 2)
        order = np.concatenate([ell[col == k] for k in range(4)])
        c["overrelax_coloured"] = hx(a.overrelax_sweep(x, idx=order))
        c["cond_S"] = hx(a.cond_evaluate(x))
        xl, pl = a.leapfrog(5, 0.05, x, p0)
        c["leapfrog"] = dict(nt=5, dt=0.05, x=hx(xl), p=hx(pl))
        ac = a.coarse()
        c["coarse_beta"] = hx(ac.param(0))
        xc = ac.copy_from_fine(x)
        c["restrict"] = hx(xc)
        c["coarse_S"] = hx(ac.evaluate(xc))
        c["prolong"] = hx(a.copy_from_coarse(angles(ac.n, 0.3), x))
        tp = scale * angles(n, 0.4)
        pc = scale * angles(ac.n, 0.9)
        c["twolevel"] = dict(theta_prime=hx(tp), phi_coarse=hx(pc),
                             deltas=hx(a.twolevel_deltas(ac, x, tp, pc)))
        cases.append(c)
    return cases


def gff_cases(R):
    """fine-level (n_gibbs_smooth = 0) GFF pieces; the dense coarse-level action
    of gffaction.cc:25-28 is recorded as coarse_S for small lattices only"""
    cases = []
    for name, Mt, ctype, mass in [("rotate8", 8, po.ROTATE, 10.0), ("rotate16", 16, po.ROTATE, 3.0),
                                  ("both8", 8, po.BOTH, 10.0)]:
        a = R.action(po.GFF, [Mt, Mt, ctype], [mass])
        n = a.n
        x = noncompact(n)
        p0 = noncompact(n, 0.7)
        c = dict(name=name, Mt=Mt, Mx=Mt, ctype=ctype, mass=mass, x=hx(x), p0=hx(p0))
        c["mu2"] = hx(a.param(0))
        c["S"] = hx(a.evaluate(x))
        c["force"] = hx(a.force(x))
        c["qoi_phi2"] = hx(a.qoi(po.QOI_PHI2, x))
        c["overrelax_lex"] = hx(a.overrelax_sweep(x))
        c["cond_S"] = hx(a.cond_evaluate(x))
        xl, pl = a.leapfrog(5, 0.05, x, p0)
        c["leapfrog"] = dict(nt=5, dt=0.05, x=hx(xl), p=hx(pl))
        ac = a.coarse()
        c["coarse_mu2"] = hx(ac.param(0))
        xc = ac.copy_from_fine(x)
        c["restrict"] = hx(xc)
        c["prolong"] = hx(a.copy_from_coarse(noncompact(ac.n, 0.3), x))
        c["coarse_S_gibbs"] = hx(ac.evaluate(xc))
        if ctype == po.ROTATE:
            # second (rotated) level as a FINE level: build it via level-1 lattice
            x1 = noncompact(ac.n, 0.2)
            c["level1"] = dict(x=hx(x1), force=hx(ac.force(x1)), cond_S=hx(ac.cond_evaluate(x1)),
                               overrelax_lex=hx(ac.overrelax_sweep(x1)),
                               restrict=hx(ac.coarse().copy_from_fine(x1)),
                               prolong=hx(ac.copy_from_coarse(noncompact(ac.coarse().n, 0.3), x1)))
        cases.append(c)
    return cases


def gff_dense_cases(R, sizes=(64,)):
    """the reference's dense coarse-level action (gffaction.cc:25-28, 133-174) at the largest sizes its
    own buildMatrices finishes here in minutes (the Eigen shim inverts with plain O(N^3) loops):
    fine lattice Mt x Mt, coarsening rotate -> coarse level of Mt^2 / 2 vertices.  Only the inputs'
    generating formula and the action values are stored."""
    cases = []
    for Mt in sizes:
        a = R.action(po.GFF, [Mt, Mt, po.ROTATE], [10.0])
        ac = a.coarse()
        vals = []
        for shift in (0.0, 0.3, 1.1):
            xc = ac.copy_from_fine(noncompact(a.n, shift))
            vals.append(hx(ac.evaluate(xc)))
        cases.append(dict(Mt=Mt, mass=10.0, n_coarse=ac.n, shifts=[0.0, 0.3, 1.1], coarse_S_gibbs=vals))
    return cases


def scalar_cases(R):
    L = R.lib
    xs = [-7.5, -np.pi, -3.0, -1e-3, 0.0, 0.5, 3.0, np.pi, 3.2, 6.5, 100.25, -100.25]
    out = dict(mod_2pi=dict(x=hx(xs), y=hx([L.ref_mod_2pi(v) for v in xs])))
    zs = [0.5, 3.0, 25.0, 99.0, 150.0, 300.0, 1000.0, 1500.0]
    out["fast_bessel_I0_scaled"] = dict(z=hx(zs), y=hx([L.ref_fast_bessel_I0_scaled(v) for v in zs]))
    out["Sigma_hat"] = dict(args=[[2.5, 2], [2.5, 4], [16.0, 2], [0.1, 4], [1.0, 3], [1.0, 0]],
                            y=hx([L.ref_Sigma_hat(a, b) for a, b in
                                  [[2.5, 2], [2.5, 4], [16.0, 2], [0.1, 4], [1.0, 3], [1.0, 0]]]))
    out["log_nCk"] = dict(args=[[10, 3], [64, 32], [5, 0]],
                          y=hx([L.ref_log_nCk(a, b) for a, b in [[10, 3], [64, 32], [5, 0]]]))
    pts = []
    for dist, params in [(0, [0.5, 16.0, 250.0]), (1, [0.8, 4.0, 700.0]), (2, [0.5, 4.0, 8.0]),
                         (3, [9.0, 16.0, 1024.0])]:
        for prm in params:
            for (x, xp, xm) in [(0.3, 0.9, 0.2), (-2.0, 2.5, -2.9), (1.0, -0.4, 3.0), (3.1, 0.05, 0.0),
                                (-0.7, -3.0, 3.0)]:
                pts.append(dict(dist=dist, param=prm, x=x, x_p=xp, x_m=xm,
                                y=float(L.ref_dist_evaluate(dist, prm, x, xp, xm)).hex()))
    out["dist_pdf"] = pts
    out["Znorm_inv"] = [dict(beta=b, phi=ph, rescaled=r,
                             y=float(L.ref_besselproduct_Znorm_inv(b, ph, r)).hex())
                        for b in (0.5, 4.0, 8.0) for ph in (0.0, 1.3, -2.9) for r in (0, 1)]
    out["analytic"] = dict(
        schwinger_chit_perturbative_4_64=float(L.ref_schwinger_chit_perturbative(4.0, 64)).hex(),
        schwinger_var_chit_continuum_4_64=float(L.ref_schwinger_var_chit_continuum(4.0, 64)).hex(),
        schwinger_chit_analytical_4_64=float(L.ref_schwinger_chit_analytical(4.0, 64)).hex(),
        schwinger_chit_analytical_1_256=float(L.ref_schwinger_chit_analytical(1.0, 256)).hex(),
        gff_phi_squared_10_16=float(L.ref_gff_phi_squared_analytical(10.0, 16, 16)).hex(),
    )
    # grid of exact / perturbative chi_t values and non-perturbatively matched coarse couplings
    # (qft/quenchedschwingerrenormalisation.cc:7-64; ip = {Mt, Mx, coarsening, renormalisation})
    out["analytic"]["schwinger_chit_grid"] = [
        dict(beta=b, n_plaq=P, exact=float(L.ref_schwinger_chit_analytical(b, P)).hex(),
             perturbative=float(L.ref_schwinger_chit_perturbative(b, P)).hex(),
             var_continuum=float(L.ref_schwinger_var_chit_continuum(b, P)).hex())
        for b, P in [(0.5, 64), (1.0, 256), (4.0, 256), (4.0, 4096), (16.0, 1024), (31.9, 4096),
                     (32.1, 4096), (64.0, 16384), (256.0, 65536), (1024.0, 262144), (1999.0, 1048576)]]
    out["analytic"]["schwinger_betacoarse_nonperturbative"] = [
        dict(beta=b, Mt=M, ctype=ct,
             beta_coarse=float(R.action(po.SCHWINGER, [M, M, ct, 2], [b]).coarse().param(0)).hex())
        for b, M, ct in [(6.0, 16, po.BOTH), (16.0, 32, po.BOTH), (64.0, 64, po.BOTH),
                         (256.0, 128, po.BOTH), (9.0, 16, po.TEMPORAL), (4.5, 8, po.BOTH),
                         (3.0, 16, po.BOTH), (900.0, 512, po.BOTH)]]
    a = R.action(po.ROTOR, [256, 0], [4.0, 0.25])
    out["analytic"]["rotor_chit_256"] = [float(L.ref_rotor_chit(a.h, w)).hex() for w in range(3)]
    a = R.action(po.ROTOR, [32, 0], [4.0, 0.25])
    out["analytic"]["rotor_chit_exact_32"] = float(L.ref_rotor_chit(a.h, 0)).hex()
    out["analytic"]["rotor_chit_perturbative_32"] = float(L.ref_rotor_chit(a.h, 1)).hex()
    out["analytic"]["rotor_chit_continuum_32"] = float(L.ref_rotor_chit(a.h, 2)).hex()
    a = R.action(po.HO, [32, 0], [4.0, 1.0, 1.0])
    out["analytic"]["ho_x2_32"] = float(L.ref_ho_xsquared_analytical(a.h, 0)).hex()
    out["analytic"]["ho_x2_continuum"] = float(L.ref_ho_xsquared_analytical(a.h, 1)).hex()
    # statistics on a deterministic AR(1)-like sequence
    q = np.zeros(400)
    v = 0.3
    for k in range(400):
        v = 0.8 * v + np.sin(1.7 * k) + 0.2 * np.cos(0.31 * k * k)
        q[k] = v
    st = np.zeros(6)
    L.ref_statistics(10, len(q), q.ctypes.data_as(po.c_double_p), st.ctypes.data_as(po.c_double_p))
    out["statistics"] = dict(k_max=10, q=hx(q), out=hx(st))
    return out


def main():
    po.build(ref=True)
    R = po.ref()
    os.makedirs(OUT, exist_ok=True)
    if "--gff-dense" in sys.argv:  # slow (minutes): recorded separately
        path = os.path.join(OUT, "gff_dense.json")
        with open(path, "w") as f:
            json.dump(gff_dense_cases(R), f, separators=(",", ":"))
        print("wrote", path)
        return
    for name, data in [("lattice", lattice_cases(R)), ("qm", qm_cases(R)),
                       ("schwinger", schwinger_cases(R)), ("gff", gff_cases(R)),
                       ("scalars", scalar_cases(R))]:
        path = os.path.join(OUT, name + ".json")
        with open(path, "w") as f:
            json.dump(data, f, separators=(",", ":"))
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
