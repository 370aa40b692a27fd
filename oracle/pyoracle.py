"""TEST INFRASTRUCTURE ONLY (oracle/): ctypes access to

* ``libmlmcpi_oracle.so`` -- the plain-C restatement (``mlmcpi_oracle.c``), and
* ``_ref/libmlmcpi_ref.so`` -- the REFERENCE's own translation units behind the
  extern-C harness ``ref_harness.cc`` (exists only where /root/reference was
  available at build time; it is shipped to the GPU box as a prebuilt file).

Only tests/, tools/make_golden.py, __graft_entry__.smoke() and the cpu_baseline /
``--impl reference`` legs of bench.py import this module.  The product package
never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libmlmcpi_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmlmcpi_ref.so")

HO, QUARTIC, ROTOR, SCHWINGER, GFF = 0, 1, 2, 3, 4
BOTH, TEMPORAL, SPATIAL, ALTERNATE, ROTATE = 0, 1, 2, 3, 4
QOI_X2, QOI_ROTOR_CHI, QOI_SCHWINGER_CHI, QOI_AVG_PLAQUETTE, QOI_PHI2 = range(5)
(STREAM_INIT, STREAM_HMC_MOMENTUM, STREAM_HMC_ACCEPT, STREAM_HEATBATH, STREAM_FILL1,
 STREAM_FILL2, STREAM_FILL3, STREAM_TWOLEVEL_ACCEPT) = range(1, 9)

c_double_p = C.POINTER(C.c_double)
c_u32_p = C.POINTER(C.c_uint32)
c_int_p = C.POINTER(C.c_int)


def build(ref=True):
    """(Re)build the oracle library and, when the reference sources exist, _ref."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-j8", "-C", HERE, "ref"], check=True)


class Model(C.Structure):
    """mirror of orc_model (mlmcpi_oracle.h)"""
    _fields_ = [("model", C.c_int), ("M_lat", C.c_int), ("Mt_lat", C.c_int),
                ("Mx_lat", C.c_int), ("rotated", C.c_int), ("coarsening", C.c_int),
                ("a_lat", C.c_double), ("T_final", C.c_double), ("m0", C.c_double),
                ("mu2", C.c_double), ("lambda_", C.c_double), ("x0", C.c_double),
                ("beta", C.c_double), ("gff_mu2", C.c_double)]

    def copy(self):
        m = Model()
        C.memmove(C.byref(m), C.byref(self), C.sizeof(Model))
        return m


class Rng(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("c0", "c1", "c2", "a", "k0", "k1")]


def ho(M, T=4.0, m0=1.0, mu2=1.0):
    return Model(model=HO, M_lat=M, a_lat=T / M, T_final=T, m0=m0, mu2=mu2)


def quartic(M, T=4.0, m0=1.0, mu2=1.0, lam=1.0, x0=1.0):
    return Model(model=QUARTIC, M_lat=M, a_lat=T / M, T_final=T, m0=m0, mu2=mu2,
                 lambda_=lam, x0=x0)


def rotor(M, T=4.0, m0=0.25):
    return Model(model=ROTOR, M_lat=M, a_lat=T / M, T_final=T, m0=m0)


def level_coarsening(ctype, level):
    """how a level of a lattice with coarsening type `ctype` is coarsened"""
    if ctype == ALTERNATE:
        return TEMPORAL if level % 2 == 0 else SPATIAL
    return ctype


def schwinger(Mt, Mx, beta, ctype=BOTH, level=0):
    return Model(model=SCHWINGER, Mt_lat=Mt, Mx_lat=Mx, beta=beta,
                 coarsening=level_coarsening(ctype, level))


def gff(Mt, Mx, mass, ctype=ROTATE, level=0):
    rotated = int(ctype == ROTATE and level % 2 == 1)
    a = (np.sqrt(2.0) if rotated else 1.0) / Mt
    return Model(model=GFF, Mt_lat=Mt, Mx_lat=Mx, rotated=rotated, coarsening=ctype,
                 gff_mu2=a * a * mass * mass)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _arr(x):
    return np.ascontiguousarray(x, dtype=np.float64)


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        _oracle = Oracle(C.CDLL(ORACLE_SO))
    return _oracle


class Oracle:
    def __init__(self, lib):
        self.lib = L = lib
        MP = C.POINTER(Model)
        d, i, u32, u64, i64 = C.c_double, C.c_int, C.c_uint32, C.c_uint64, C.c_int64
        sig = {
            "orc_philox4x32_10": (None, [c_u32_p, c_u32_p, c_u32_p]),
            "orc_rng_init": (None, [C.POINTER(Rng), u64, i, u64, u32, u32]),
            "orc_rng_uniform2": (None, [C.POINTER(Rng), c_double_p, c_double_p]),
            "orc_rng_normal2": (None, [C.POINTER(Rng), c_double_p, c_double_p]),
            "orc_mod_2pi": (d, [d]),
            "orc_bessel_I0": (d, [d]),
            "orc_bessel_I0_scaled": (d, [d]),
            "orc_fast_bessel_I0_scaled": (d, [d]),
            "orc_vertex_cart2lin": (u32, [i, i, i, i, i]),
            "orc_vertex_lin2cart": (None, [i, i, i, u32, c_int_p, c_int_p]),
            "orc_link_cart2lin": (u32, [i, i, i, i, i]),
            "orc_link_lin2cart": (None, [i, i, u32, c_int_p, c_int_p, c_int_p]),
            "orc_n_vertices": (i, [i, i, i]),
            "orc_neighbours": (None, [i, i, i, u32, c_u32_p]),
            "orc_coarse_shape": (i, [i, i, i, i, c_int_p, c_int_p, c_int_p]),
            "orc_coarsening_lists": (i, [i, i, i, i, c_u32_p, c_u32_p, c_u32_p, c_int_p]),
            "orc_sample_size": (i, [MP]),
            "orc_action": (d, [MP, c_double_p]),
            "orc_force": (None, [MP, c_double_p, c_double_p]),
            "orc_W": (None, [MP, d, d, c_double_p, c_double_p]),
            "orc_overrelax_update": (None, [MP, c_double_p, u32]),
            "orc_overrelax_sweep_lex": (None, [MP, c_double_p]),
            "orc_n_colours": (i, [MP]),
            "orc_colour_of": (i, [MP, u32]),
            "orc_overrelax_sweep_coloured": (None, [MP, c_double_p]),
            "orc_prolong": (None, [MP, c_double_p, c_double_p]),
            "orc_restrict": (None, [MP, c_double_p, c_double_p]),
            "orc_cond_action": (d, [MP, c_double_p]),
            "orc_qoi": (d, [MP, i, c_double_p, C.POINTER(i64)]),
            "orc_leapfrog": (None, [MP, i, d, c_double_p, c_double_p]),
            "orc_coarse_model": (i, [MP, i, i, i, d, MP]),
            "orc_expsin2_pdf": (d, [d, d]),
            "orc_expcos_pdf": (d, [d, d, d, d]),
            "orc_besselproduct_alpha": (None, [d, c_double_p]),
            "orc_besselproduct_Znorm_inv": (d, [c_double_p, d, i]),
            "orc_besselproduct_pdf": (d, [d, d, d, d]),
            "orc_approxbessel_pdf": (d, [d, d, d, d]),
            "orc_expsin2_draw": (d, [C.POINTER(Rng), d]),
            "orc_set_expcos_envelope": (None, [i]),
            "orc_expcos_draw": (d, [C.POINTER(Rng), d, d, d]),
            "orc_besselproduct_draw": (d, [C.POINTER(Rng), d, d, d]),
            "orc_approxbessel_draw": (d, [C.POINTER(Rng), d, d, d]),
            "orc_init_state": (None, [MP, u64, u64, u32, c_double_p]),
            "orc_hmc_momentum": (None, [MP, u64, u64, u32, c_double_p]),
            "orc_ho_exact_factor": (i, [MP, c_double_p]),
            "orc_ho_exact_draw": (i, [MP, u64, u64, u32, c_double_p]),
            "orc_hmc_step": (i, [MP, i, d, u64, u64, u32, c_double_p, c_double_p]),
            "orc_heatbath_sweep_coloured": (None, [MP, u64, u64, u32, c_double_p]),
            "orc_fill": (None, [MP, u64, u64, u32, c_double_p]),
            "orc_twolevel_step": (i, [MP, MP, u64, u64, u32, c_double_p, c_double_p,
                                      c_double_p, c_double_p, c_double_p]),
            "orc_cluster_update": (None, [MP, u64, u64, i, u32, c_double_p]),
            "orc_schwinger_from_cluster": (None, [MP, u64, u64, u32, c_double_p, c_double_p]),
            "orc_statistics": (None, [i, i, c_double_p, c_double_p]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args

    # ---- convenience wrappers (numpy in / numpy out) ----
    def sample_size(self, m):
        return self.lib.orc_sample_size(C.byref(m))

    def action(self, m, x):
        x = _arr(x)
        return self.lib.orc_action(C.byref(m), _dp(x))

    def force(self, m, x):
        x = _arr(x)
        p = np.empty_like(x)
        self.lib.orc_force(C.byref(m), _dp(x), _dp(p))
        return p

    def W(self, m, x_m, x_p):
        a, b = C.c_double(), C.c_double()
        self.lib.orc_W(C.byref(m), x_m, x_p, C.byref(a), C.byref(b))
        return a.value, b.value

    def overrelax_sweep(self, m, x, coloured=True):
        x = _arr(x).copy()
        f = self.lib.orc_overrelax_sweep_coloured if coloured else self.lib.orc_overrelax_sweep_lex
        f(C.byref(m), _dp(x))
        return x

    def colours(self, m):
        n = self.sample_size(m)
        return np.array([self.lib.orc_colour_of(C.byref(m), l) for l in range(n)])

    def prolong(self, fine, xc, x=None):
        n = self.sample_size(fine)
        x = np.zeros(n) if x is None else _arr(x).copy()
        xc = _arr(xc)
        self.lib.orc_prolong(C.byref(fine), _dp(xc), _dp(x))
        return x

    def restrict(self, fine, coarse, xf):
        xf = _arr(xf)
        xc = np.zeros(self.sample_size(coarse))
        self.lib.orc_restrict(C.byref(fine), _dp(xf), _dp(xc))
        return xc

    def cond_action(self, fine, x):
        x = _arr(x)
        return self.lib.orc_cond_action(C.byref(fine), _dp(x))

    def qoi(self, m, qoi, x):
        x = _arr(x)
        Q = C.c_int64()
        v = self.lib.orc_qoi(C.byref(m), qoi, _dp(x), C.byref(Q))
        return v, Q.value

    def leapfrog(self, m, nt, dt, x, p):
        x, p = _arr(x).copy(), _arr(p).copy()
        self.lib.orc_leapfrog(C.byref(m), nt, dt, _dp(x), _dp(p))
        return x, p

    def coarse_model(self, fine, renorm=0, level=0, ctype=BOTH, T_final=0.0):
        c = Model()
        rc = self.lib.orc_coarse_model(C.byref(fine), renorm, level, ctype, T_final,
                                       C.byref(c))
        if rc != 0:
            raise ValueError("cannot coarsen")
        return c

    def init_state(self, m, seed, draw, chain):
        x = np.zeros(self.sample_size(m))
        self.lib.orc_init_state(C.byref(m), seed, draw, chain, _dp(x))
        return x

    def hmc_momentum(self, m, seed, draw, chain):
        p = np.zeros(self.sample_size(m))
        self.lib.orc_hmc_momentum(C.byref(m), seed, draw, chain, _dp(p))
        return p

    def ho_exact_factor(self, m):
        L = np.zeros((m.M_lat, m.M_lat))
        assert self.lib.orc_ho_exact_factor(C.byref(m), _dp(L)) == 0
        return L

    def ho_exact_draw(self, m, seed, draw, chain):
        x = np.zeros(m.M_lat)
        assert self.lib.orc_ho_exact_draw(C.byref(m), seed, draw, chain, _dp(x)) == 0
        return x

    def hmc_step(self, m, nt, dt, seed, draw, chain, x):
        x = _arr(x).copy()
        out = np.zeros(5)
        acc = self.lib.orc_hmc_step(C.byref(m), nt, dt, seed, draw, chain, _dp(x), _dp(out))
        return acc, x, out

    def heatbath_sweep(self, m, seed, draw, chain, x):
        x = _arr(x).copy()
        self.lib.orc_heatbath_sweep_coloured(C.byref(m), seed, draw, chain, _dp(x))
        return x

    def fill(self, fine, seed, draw, chain, x):
        x = _arr(x).copy()
        self.lib.orc_fill(C.byref(fine), seed, draw, chain, _dp(x))
        return x

    def twolevel_step(self, fine, coarse, seed, draw, chain, x_coarse, x_fine, S_fine, S_cond):
        x_fine = _arr(x_fine).copy()
        x_coarse = _arr(x_coarse)
        sf, sc = C.c_double(S_fine), C.c_double(S_cond)
        out = np.zeros(3)
        acc = self.lib.orc_twolevel_step(C.byref(fine), C.byref(coarse), seed, draw, chain,
                                         _dp(x_coarse), _dp(x_fine), C.byref(sf),
                                         C.byref(sc), _dp(out))
        return acc, x_fine, sf.value, sc.value, out

    def cluster_update(self, rotor, seed, update0, n_updates, chain, x):
        x = _arr(x).copy()
        self.lib.orc_cluster_update(C.byref(rotor), seed, update0, n_updates, chain, _dp(x))
        return x

    def schwinger_from_cluster(self, m, seed, draw, chain, psi):
        psi = _arr(psi)
        x = np.zeros(self.sample_size(m))
        self.lib.orc_schwinger_from_cluster(C.byref(m), seed, draw, chain, _dp(psi), _dp(x))
        return x

    def statistics(self, k_max, q):
        q = _arr(q)
        out = np.zeros(6)
        self.lib.orc_statistics(k_max, len(q), _dp(q), _dp(out))
        return out

    def rng(self, seed, stream, draw, chain, index):
        r = Rng()
        self.lib.orc_rng_init(C.byref(r), seed, stream, draw, chain, index)
        return r

    def uniform2(self, r):
        a, b = C.c_double(), C.c_double()
        self.lib.orc_rng_uniform2(C.byref(r), C.byref(a), C.byref(b))
        return a.value, b.value

    def normal2(self, r):
        a, b = C.c_double(), C.c_double()
        self.lib.orc_rng_normal2(C.byref(r), C.byref(a), C.byref(b))
        return a.value, b.value


# --------------------------------------------------------------------------- #
#  the reference's own translation units
# --------------------------------------------------------------------------- #

def gff_dense_matrices(orc, m, n_gibbs=2, omega=1.0):
    """numpy restatement of GFFAction::buildMatrices (qft/gffaction.cc:133-174) for the level
    described by the GFF model `m`: returns dict(Q=5-point precision, Q_eff=9-point effective
    precision, G=Gibbs iteration matrix ^ n_gibbs, Sigma_hat, Q_hat, U=upper Cholesky factor of Q).
    The coarse-level action of the reference is S = phi^T Q_hat phi / 2 (gffaction.cc:25-28) and its
    exact sampler draws phi = U^{-1} psi followed by n_gibbs lexicographic sweeps (gffaction.cc:200-213)."""
    try:  # single-threaded BLAS: a BLAS thread pool can deadlock next to torch's OpenMP runtime
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1):
            return _gff_dense_matrices(orc, m, n_gibbs, omega)
    except ImportError:
        return _gff_dense_matrices(orc, m, n_gibbs, omega)


def _gff_dense_matrices(orc, m, n_gibbs, omega):
    N = orc.sample_size(m)
    mu2 = m.gff_mu2
    nb = np.zeros((N, 8), dtype=np.uint32)
    for ell in range(N):
        orc.lib.orc_neighbours(m.Mt_lat, m.Mx_lat, m.rotated, ell, nb[ell].ctypes.data_as(c_u32_p))

    def precision(stencil):  # GFFAction::buildPrecisionMatrix, gffaction.cc:177-197
        Q = np.zeros((N, N))
        for ell in range(N):
            Q[ell, ell] += stencil[0]
            for j in range(len(stencil) - 1):
                for k in range(4):
                    Q[ell, nb[ell, 4 * j + k]] += stencil[j + 1]
        return Q

    Q = precision([4.0 + mu2, -1.0])
    d = 4.0 + 0.5 * mu2
    Q_eff = precision([d - 4.0 / d, -2.0 / d, -1.0 / d])
    Sigma, Sigma_eff = np.linalg.inv(Q), np.linalg.inv(Q_eff)
    M = np.tril(Q_eff)
    if abs(omega - 1.0) > 1e-14:
        M = M + np.diag((1.0 / omega - 1.0) * np.diag(Q_eff))
    G1 = np.eye(N) - np.linalg.solve(M, Q_eff)
    G = np.linalg.matrix_power(G1, n_gibbs) if n_gibbs > 0 else np.eye(N)
    Sigma_hat = Sigma_eff + G @ (Sigma - Sigma_eff) @ G.T
    return dict(Q=Q, Q_eff=Q_eff, G=G, Sigma=Sigma, Sigma_hat=Sigma_hat, Q_hat=np.linalg.inv(Sigma_hat),
                U=np.linalg.cholesky(Q).T, nb=nb)


def have_ref():
    return os.path.exists(REF_SO)


_ref = None


def ref():
    global _ref
    if _ref is None:
        _ref = Ref(C.CDLL(REF_SO))
    return _ref


class Ref:
    """thin ctypes layer over oracle/ref_harness.cc"""

    def __init__(self, lib):
        self.lib = L = lib
        d, i, u, vp = C.c_double, C.c_int, C.c_uint, C.c_void_p
        c_uint_p = C.POINTER(C.c_uint)
        sig = {
            "ref_lattice2d_info": (i, [u, u, i, i, c_int_p]),
            "ref_lattice2d_vertex_maps": (i, [u, u, i, i, i, i, i, i, c_uint_p, c_int_p, c_uint_p]),
            "ref_lattice2d_link_maps": (i, [u, u, i, i, i, i, i, i, c_uint_p, c_int_p]),
            "ref_lattice2d_coarsening": (i, [u, u, i, i, c_uint_p, c_uint_p, c_uint_p, c_uint_p, c_int_p]),
            "ref_lattice1d": (i, [u, d, c_double_p, c_uint_p]),
            "ref_action_create": (vp, [i, c_int_p, c_double_p]),
            "ref_action_destroy": (None, [vp]),
            "ref_action_coarse": (vp, [vp]),
            "ref_action_sample_size": (i, [vp]),
            "ref_action_param": (d, [vp, i]),
            "ref_action_evaluate": (d, [vp, c_double_p]),
            "ref_action_force": (None, [vp, c_double_p, c_double_p]),
            "ref_action_overrelax_sweep": (None, [vp, c_double_p, i, c_uint_p, i]),
            "ref_action_heatbath_sweep": (None, [vp, c_double_p, i]),
            "ref_action_copy_from_coarse": (None, [vp, c_double_p, i, c_double_p]),
            "ref_action_copy_from_fine": (None, [vp, c_double_p, i, c_double_p]),
            "ref_action_W": (None, [vp, d, d, c_double_p]),
            "ref_action_initialise_state": (None, [vp, c_double_p]),
            "ref_cond_create": (vp, [vp]),
            "ref_cond_destroy": (None, [vp]),
            "ref_cond_evaluate": (d, [vp, c_double_p, i]),
            "ref_cond_fill": (None, [vp, c_double_p, i]),
            "ref_qoi_evaluate": (d, [i, vp, c_double_p]),
            "ref_hmc_leapfrog": (None, [vp, u, d, c_double_p, c_double_p]),
            "ref_hmc_draws": (i, [vp, u, d, i, C.c_uint64, c_double_p, i, c_double_p, c_double_p]),
            "ref_twolevel_deltas": (None, [vp, vp, vp, c_double_p, c_double_p, c_double_p, c_double_p]),
            "ref_twolevel_draws": (d, [vp, vp, vp, i, c_double_p, c_double_p, c_double_p]),
            "ref_dist_evaluate": (d, [i, d, d, d, d]),
            "ref_dist_draw": (None, [i, d, d, d, C.c_uint64, i, c_double_p]),
            "ref_besselproduct_Znorm_inv": (d, [d, d, i]),
            "ref_mod_2pi": (d, [d]),
            "ref_mod_pi": (d, [d]),
            "ref_fast_bessel_I0_scaled": (d, [d]),
            "ref_Sigma_hat": (d, [d, u]),
            "ref_log_nCk": (d, [u, u]),
            "ref_schwinger_chit_analytical": (d, [d, u]),
            "ref_schwinger_chit_perturbative": (d, [d, u]),
            "ref_schwinger_var_chit_continuum": (d, [d, u]),
            "ref_gff_phi_squared_analytical": (d, [d, d, d]),
            "ref_rotor_chit": (d, [vp, i]),
            "ref_ho_xsquared_analytical": (d, [vp, i]),
            "ref_ho_exact_draws": (None, [vp, i, c_double_p]),
            "ref_gff_exact_draws": (None, [vp, i, c_double_p]),
            "ref_statistics": (None, [u, i, c_double_p, c_double_p]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args

    def action(self, kind, ip, dp):
        ipa = (C.c_int * len(ip))(*ip)
        dpa = (C.c_double * len(dp))(*dp)
        h = self.lib.ref_action_create(kind, ipa, dpa)
        if not h:
            raise ValueError("reference refused to construct the action")
        return RefAction(self, h, kind)


class RefAction:
    def __init__(self, ref_, handle, kind):
        self.ref, self.h, self.kind = ref_, handle, kind
        self.L = ref_.lib
        self.n = self.L.ref_action_sample_size(handle)
        self._cond = None

    def coarse(self):
        h = self.L.ref_action_coarse(self.h)
        if not h:
            raise ValueError("reference cannot coarsen this action")
        return RefAction(self.ref, h, self.kind)

    def param(self, which=0):
        return self.L.ref_action_param(self.h, which)

    def evaluate(self, x):
        x = _arr(x)
        return self.L.ref_action_evaluate(self.h, _dp(x))

    def force(self, x):
        x = _arr(x)
        p = np.empty_like(x)
        self.L.ref_action_force(self.h, _dp(x), _dp(p))
        return p

    def overrelax_sweep(self, x, n_sweeps=1, idx=None):
        x = _arr(x).copy()
        if idx is None:
            self.L.ref_action_overrelax_sweep(self.h, _dp(x), n_sweeps, None, 0)
        else:
            idx = np.ascontiguousarray(idx, dtype=np.uint32)
            self.L.ref_action_overrelax_sweep(
                self.h, _dp(x), n_sweeps, idx.ctypes.data_as(C.POINTER(C.c_uint)), len(idx))
        return x

    def heatbath_sweep(self, x, n_sweeps=1):
        x = _arr(x).copy()
        self.L.ref_action_heatbath_sweep(self.h, _dp(x), n_sweeps)
        return x

    def copy_from_coarse(self, xc, x=None):
        xc = _arr(xc)
        x = np.zeros(self.n) if x is None else _arr(x).copy()
        self.L.ref_action_copy_from_coarse(self.h, _dp(xc), len(xc), _dp(x))
        return x

    def copy_from_fine(self, xf):
        """self is the COARSE action"""
        xf = _arr(xf)
        x = np.zeros(self.n)
        self.L.ref_action_copy_from_fine(self.h, _dp(xf), len(xf), _dp(x))
        return x

    def W(self, x_m, x_p):
        out = np.zeros(2)
        self.L.ref_action_W(self.h, x_m, x_p, _dp(out))
        return out[0], out[1]

    def cond(self):
        if self._cond is None:
            self._cond = self.L.ref_cond_create(self.h)
            if not self._cond:
                raise ValueError("no conditioned fine action")
        return self._cond

    def cond_evaluate(self, x):
        x = _arr(x)
        return self.L.ref_cond_evaluate(self.cond(), _dp(x), len(x))

    def cond_fill(self, x):
        x = _arr(x).copy()
        self.L.ref_cond_fill(self.cond(), _dp(x), len(x))
        return x

    def qoi(self, qoi, x):
        x = _arr(x)
        return self.L.ref_qoi_evaluate(qoi, self.h, _dp(x))

    def leapfrog(self, nt, dt, x, p):
        x, p = _arr(x).copy(), _arr(p).copy()
        self.L.ref_hmc_leapfrog(self.h, nt, dt, _dp(x), _dp(p))
        return x, p

    def hmc_draws(self, nt, dt, n_draws, seed, x, qoi=-1):
        x = _arr(x).copy()
        q = np.zeros(max(n_draws, 1))
        sec = C.c_double()
        nacc = self.L.ref_hmc_draws(self.h, nt, dt, n_draws, seed, _dp(x), qoi,
                                    _dp(q) if qoi >= 0 else None, C.byref(sec))
        return nacc, x, q, sec.value

    def twolevel_deltas(self, coarse, theta_fine, theta_prime, phi_coarse):
        out = np.zeros(3)
        tf, tp, pc = _arr(theta_fine), _arr(theta_prime), _arr(phi_coarse)
        self.L.ref_twolevel_deltas(coarse.h, self.h, self.cond(), _dp(tf), _dp(tp), _dp(pc), _dp(out))
        return out
