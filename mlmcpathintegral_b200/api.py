"""Host-side handles over the C-ABI: Context (kernel entry points on torch CUDA
tensors), Sampler (HMCSampler / OverrelaxedHeatBathSampler / HierarchicalSampler of
the reference as one batched object) and Statistics.

torch is used for device memory and streams only; every computation is a call into
libmlmcpi.so.  Batched states are float64 tensors of shape [B, n_dof] in the
reference's dof order (one row == one reference SampleState::data)."""
import ctypes as C
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import (COARSEN_ALTERNATE, COARSEN_BOTH, COARSEN_ROTATE, COARSEN_SPATIAL,  # noqa: F401
                   COARSEN_TEMPORAL, GFF, HO, QOI_AVG_PLAQUETTE, QOI_PHI2, QOI_ROTOR_CHI,
                   QOI_SCHWINGER_CHI, QOI_X2, QUARTIC, RENORM_NONE, RENORM_NONPERTURBATIVE,
                   RENORM_PERTURBATIVE, ROTOR,
                   SAMPLER_CLUSTER, SAMPLER_EXACT, SAMPLER_HEATBATH, SAMPLER_HMC, SCHWINGER, MlmcParams, Model,
                   SamplerParams)

L = _lib.lib


class MlmcpiError(RuntimeError):
    pass


# ----------------------------------------------------------------- model makers
def ho(M, T=4.0, m0=1.0, mu2=1.0):
    """HarmonicOscillatorAction (action/qm/harmonicoscillatoraction.hh)"""
    return Model(model=HO, M_lat=M, a_lat=T / M, T_final=T, m0=m0, mu2=mu2)


def quartic(M, T=4.0, m0=1.0, mu2=1.0, lam=1.0, x0=1.0):
    """QuarticOscillatorAction (action/qm/quarticoscillatoraction.hh)"""
    return Model(model=QUARTIC, M_lat=M, a_lat=T / M, T_final=T, m0=m0, mu2=mu2, lambda_=lam, x0=x0)


def rotor(M, T=4.0, m0=0.25):
    """RotorAction (action/qm/rotoraction.hh)"""
    return Model(model=ROTOR, M_lat=M, a_lat=T / M, T_final=T, m0=m0)


def level_coarsening(ctype, level):
    if ctype == COARSEN_ALTERNATE:  # lattice/lattice2d.cc:39-48
        return COARSEN_TEMPORAL if level % 2 == 0 else COARSEN_SPATIAL
    return ctype


def schwinger(Mt, Mx, beta, ctype=COARSEN_BOTH, level=0):
    """QuenchedSchwingerAction (action/qft/quenchedschwingeraction.hh)"""
    return Model(model=SCHWINGER, Mt_lat=Mt, Mx_lat=Mx, beta=beta,
                 coarsening=level_coarsening(ctype, level))


def gff(Mt, Mx, mass, ctype=COARSEN_ROTATE, level=0):
    """GFFAction, fine-level 5-point form (action/qft/gffaction.hh:174-181)"""
    rotated = int(ctype == COARSEN_ROTATE and level % 2 == 1)
    a = (np.sqrt(2.0) if rotated else 1.0) / Mt
    return Model(model=GFF, Mt_lat=Mt, Mx_lat=Mx, rotated=rotated, coarsening=ctype,
                 gff_mu2=a * a * mass * mass)


def sample_size(m):
    return L.mlmcpi_sample_size(C.byref(m))


def coarse_model(fine, renorm=RENORM_NONE, level=0, ctype=COARSEN_BOTH, T_final=None):
    c = Model()
    rc = L.mlmcpi_coarse_model(C.byref(fine), renorm, level, ctype,
                               fine.T_final if T_final is None else T_final, C.byref(c))
    if rc:
        raise MlmcpiError(f"mlmcpi_coarse_model failed ({rc})")
    return c


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


class Context:
    """mlmcpi_ctx bound to a CUDA device and (by default) torch's current stream"""

    def __init__(self, device=0, seed=0x5EED0001, use_torch_stream=True):
        if not torch.cuda.is_available():
            raise MlmcpiError("no CUDA device: mlmcpathintegral_b200 has no CPU fallback")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        # torch's default stream is CUDA's legacy default stream (handle 0 == NULL)
        stream = (torch.cuda.current_stream(self.device).cuda_stream if use_torch_stream
                  else C.c_void_p(-1))  # MLMCPI_OWN_STREAM
        h = C.c_void_p()
        rc = L.mlmcpi_create(C.byref(h), device, seed, stream)
        if rc:
            raise MlmcpiError(f"mlmcpi_create failed ({rc})")
        self.h = h
        self._children = weakref.WeakSet()  # samplers / statistics living on this context

    def close(self):
        if self.h:
            for child in list(self._children):
                child.close()
            L.mlmcpi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise MlmcpiError(f"{L.mlmcpi_last_error(self.h).decode()} ({rc})")

    def sync(self):
        self._ck(L.mlmcpi_sync(self.h))

    def set_seed(self, seed):
        L.mlmcpi_set_seed(self.h, seed)

    def set_option(self, option, value):
        self._ck(L.mlmcpi_set_option(self.h, option, value))

    def set_expcos_envelope(self, variant=2):
        """ExpCos proposal: 0 reference envelope, 1 chord bound, 2 chord + Taylor bound (default)"""
        self._ck(L.mlmcpi_set_option(self.h, _lib.OPT_EXPCOS_ENVELOPE, int(variant)))

    def attach_process_group(self, group=None):
        """mlmcpi_set_allreduce with torch.distributed as the transport: from now on the library's
        own Statistics queries (MultilevelSampler, MonteCarloMultiLevel, HMC autotune) run over the
        chains of ALL processes of the group, so every process takes the same host-side decisions
        (common/statistics.cc:30-35,64-79 averages the same moments over MPI ranks).  NCCL groups
        reduce a device staging tensor, other backends (gloo) a host one."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world == 1:
            self._ck(L.mlmcpi_set_allreduce(self.h, None, None, 1, 0))
            self._allreduce_cb = None
            return
        on_device = dist.get_backend(group) == "nccl"
        stage = {}

        def hook(_user, d_buf, n):
            try:
                t = stage.get(n)
                if t is None:
                    t = stage[n] = torch.zeros(n, dtype=torch.float64,
                                               device=self.device if on_device else "cpu")
                tp = C.c_void_p(t.data_ptr())
                if on_device:
                    # a host-decision path (a few calls per allocation round): plain synchronisation
                    # orders the library's stream and the stream NCCL is issued on
                    self._ck(L.mlmcpi_copy(self.h, tp, d_buf, n))
                    self.sync()
                    dist.all_reduce(t, group=group)
                    torch.cuda.current_stream(self.device).synchronize()
                    self._ck(L.mlmcpi_copy(self.h, d_buf, tp, n))
                else:
                    self._ck(L.mlmcpi_download(self.h, tp, d_buf, n))  # synchronises
                    dist.all_reduce(t, group=group)
                    self._ck(L.mlmcpi_upload(self.h, d_buf, tp, n))
                    self.sync()
                return 0
            except Exception:  # never unwind through the C frames
                import traceback
                traceback.print_exc()
                return _lib.E_CUDA

        self._allreduce_cb = _lib.ALLREDUCE_FN(hook)  # keep the thunk alive
        self._ck(L.mlmcpi_set_allreduce(self.h, C.cast(self._allreduce_cb, C.c_void_p), None, world, rank))

    @property
    def world_size(self):
        return int(L.mlmcpi_world_size(self.h))

    @property
    def launches(self):
        return int(L.mlmcpi_launch_count(self.h))

    def profile(self, enable=True):
        L.mlmcpi_profile(self.h, int(enable))

    def profile_read(self):
        """(milliseconds, launches, algorithmic bytes) of the leapfrog kernel since the last read"""
        out = (C.c_double * 3)()
        self._ck(L.mlmcpi_profile_read(self.h, out))
        return out[0], int(out[1]), out[2]

    # ---- tensors
    def empty(self, *shape, dtype=torch.float64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def state(self, m, B):
        return torch.zeros(B, sample_size(m), dtype=torch.float64, device=self.device)

    def to_device(self, a):
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)

    # ---- group 1
    def init_state(self, m, B, chain0=0, draw=0):
        x = self.state(m, B)
        self._ck(L.mlmcpi_init_state(self.h, C.byref(m), _ptr(x), B, chain0, draw))
        return x

    def action(self, m, x):
        S = self.empty(x.shape[0])
        self._ck(L.mlmcpi_action(self.h, C.byref(m), _ptr(x), x.shape[0], _ptr(S)))
        return S

    def force(self, m, x):
        f = torch.empty_like(x)
        self._ck(L.mlmcpi_force(self.h, C.byref(m), _ptr(x), _ptr(f), x.shape[0]))
        return f

    def leapfrog(self, m, nt, dt, x, p):
        """in place on x, p"""
        self._ck(L.mlmcpi_leapfrog(self.h, C.byref(m), nt, dt, _ptr(x), _ptr(p), x.shape[0]))

    def hmc_momentum(self, m, B, chain0=0, draw=0):
        p = self.state(m, B)
        self._ck(L.mlmcpi_hmc_momentum(self.h, C.byref(m), _ptr(p), B, chain0, draw))
        return p

    def hmc_step(self, m, nt, dt, x, chain0=0, draw=0):
        """in place on x; returns (accept[B] int32, diag[B,5])"""
        B = x.shape[0]
        acc = self.empty(B, dtype=torch.int32)
        diag = self.empty(B, 5)
        self._ck(L.mlmcpi_hmc_step(self.h, C.byref(m), nt, dt, _ptr(x), B, chain0, draw, _ptr(acc),
                                   _ptr(diag)))
        return acc, diag

    # ---- group 2
    def overrelax_sweep(self, m, x):
        self._ck(L.mlmcpi_overrelax_sweep(self.h, C.byref(m), _ptr(x), x.shape[0]))

    def overrelax_sweeps(self, m, x, n_sweeps):
        self._ck(L.mlmcpi_overrelax_sweeps(self.h, C.byref(m), _ptr(x), x.shape[0], n_sweeps))

    def heatbath_sweep(self, m, x, chain0=0, draw=0):
        self._ck(L.mlmcpi_heatbath_sweep(self.h, C.byref(m), _ptr(x), x.shape[0], chain0, draw))

    def dof_update(self, m, x, ell, heatbath=False, chain0=0, draw=0):
        """Action::heatbath_update / overrelaxation_update of one degree of freedom on all chains"""
        self._ck(L.mlmcpi_dof_update(self.h, C.byref(m), _ptr(x), x.shape[0], ell, int(heatbath), chain0, draw))

    def prolong(self, fine, xc, x):
        self._ck(L.mlmcpi_prolong(self.h, C.byref(fine), _ptr(xc), _ptr(x), x.shape[0]))

    def restrict(self, fine, xf, xc):
        self._ck(L.mlmcpi_restrict(self.h, C.byref(fine), _ptr(xf), _ptr(xc), xf.shape[0]))

    def fill(self, fine, x, chain0=0, draw=0):
        self._ck(L.mlmcpi_fill(self.h, C.byref(fine), _ptr(x), x.shape[0], chain0, draw))

    def prolong_fill(self, fine, xc, x, chain0=0, draw=0):
        self._ck(L.mlmcpi_prolong_fill(self.h, C.byref(fine), _ptr(xc), _ptr(x), x.shape[0], chain0,
                                       draw))

    def prolong_fill_eval(self, fine, xc, x, chain0=0, draw=0):
        """prolong_fill plus (S_f(theta'), S_cond(theta')) of the new state, one pass"""
        S = self.empty(2 * x.shape[0])
        self._ck(L.mlmcpi_prolong_fill_eval(self.h, C.byref(fine), _ptr(xc), _ptr(x), x.shape[0], chain0,
                                            draw, _ptr(S)))
        return S[:x.shape[0]], S[x.shape[0]:]

    def exact_draw(self, m, B, chain0=0, draw=0):
        """HarmonicOscillatorAction::draw: independent exact samples for B chains"""
        x = self.state(m, B)
        self._ck(L.mlmcpi_exact_draw(self.h, C.byref(m), _ptr(x), B, chain0, draw))
        return x

    def cluster_update(self, rotor_model, x, chain0=0, update0=0, n_updates=1):
        self._ck(L.mlmcpi_cluster_update(self.h, C.byref(rotor_model), _ptr(x), x.shape[0], chain0,
                                         update0, n_updates))

    def schwinger_from_cluster(self, m, psi, x, chain0=0, draw=0):
        self._ck(L.mlmcpi_schwinger_from_cluster(self.h, C.byref(m), _ptr(psi), _ptr(x), x.shape[0],
                                                 chain0, draw))

    # ---- group 3
    def cond_action(self, fine, x):
        S = self.empty(x.shape[0])
        self._ck(L.mlmcpi_cond_action(self.h, C.byref(fine), _ptr(x), x.shape[0], _ptr(S)))
        return S

    def qoi(self, m, which, x, with_charge=False):
        B = x.shape[0]
        q = self.empty(B)
        Q = self.empty(B, dtype=torch.int64) if with_charge else None
        self._ck(L.mlmcpi_qoi(self.h, C.byref(m), which, _ptr(x), B, _ptr(q), _ptr(Q)))
        return (q, Q) if with_charge else q

    def twolevel_step(self, fine, coarse, xc, xf, Sf, Scond, chain0=0, draw=0):
        """in place on xf, Sf, Scond; returns (accept[B], deltas[B,3])"""
        B = xf.shape[0]
        acc = self.empty(B, dtype=torch.int32)
        deltas = self.empty(B, 3)
        self._ck(L.mlmcpi_twolevel_step(self.h, C.byref(fine), C.byref(coarse), _ptr(xc), _ptr(xf),
                                        _ptr(Sf), _ptr(Scond), B, chain0, draw, _ptr(acc),
                                        _ptr(deltas)))
        return acc, deltas


class Sampler:
    """Batched Sampler (sampler/sampler.hh:20-43): HMC or overrelaxed heat bath on one
    level, or the HierarchicalSampler cascade (sampler/hierarchicalsampler.cc) when
    n_levels > 1."""

    def __init__(self, ctx, fine, B, kind=SAMPLER_HMC, n_levels=1, renorm=RENORM_NONE,
                 ctype=COARSEN_BOTH, nt=100, dt=0.1, n_rep=1, n_sweep_overrelax=10,
                 n_sweep_heatbath=1, chain0=0, multilevel=False, qoi=QOI_X2, n_autocorr_window=20,
                 n_updates=10):
        self.ctx, self.fine, self.B, self.n_levels = ctx, fine, B, n_levels
        prm = SamplerParams(kind=kind, n_levels=n_levels, renorm=renorm, ctype=ctype, nt=nt, dt=dt,
                            n_rep=n_rep, n_sweep_overrelax=n_sweep_overrelax,
                            n_sweep_heatbath=n_sweep_heatbath, multilevel=int(multilevel), qoi=qoi,
                            n_autocorr_window=n_autocorr_window, n_updates=n_updates)
        h = C.c_void_p()
        ctx._ck(L.mlmcpi_sampler_create(ctx.h, C.byref(fine), C.byref(prm), B, chain0, C.byref(h)))
        self.h = h
        self.n = sample_size(fine)
        ctx._children.add(self)

    def close(self):
        if self.h:
            L.mlmcpi_sampler_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, x):
        self.ctx._ck(L.mlmcpi_sampler_set_state(self.h, _ptr(x)))

    def get_state(self, x=None):
        """the current state of every chain (draw() only overwrites its output where it accepted)"""
        if x is None:
            x = self.ctx.state(self.fine, self.B)
        self.ctx._ck(L.mlmcpi_sampler_get_state(self.h, _ptr(x)))
        return x

    def qoi(self, which):
        """QoI of the chains' current states (mlmcpi_sampler_qoi: maintained by the draws where the library can)"""
        q = self.ctx.empty(self.B)
        self.ctx._ck(L.mlmcpi_sampler_qoi(self.h, which, _ptr(q)))
        return q

    def draw(self, x_out=None, accept=None):
        self.ctx._ck(L.mlmcpi_sampler_draw(self.h, _ptr(x_out), _ptr(accept)))

    def draw_host(self, x_in=None, qoi=QOI_SCHWINGER_CHI, q_out=None, x_out=None):
        """host-buffer entry point: numpy/pinned arrays in and out, synchronous"""
        def hp(a):
            return None if a is None else C.c_void_p(a.ctypes.data if isinstance(a, np.ndarray)
                                                     else a.data_ptr())
        self.ctx._ck(L.mlmcpi_sampler_draw_host(self.h, hp(x_in), qoi, hp(q_out), hp(x_out)))

    def draw_host_async(self, qoi=QOI_SCHWINGER_CHI, q_out=None, x_out=None):
        """pipelined host hand-over (mlmcpi_sampler_draw_host_async): pinned host tensors / numpy arrays;
        the buffers of this call are complete after the next call or after wait_host()"""
        def hp(a):
            return None if a is None else C.c_void_p(a.ctypes.data if isinstance(a, np.ndarray)
                                                     else a.data_ptr())
        self.ctx._ck(L.mlmcpi_sampler_draw_host_async(self.h, qoi, hp(q_out), hp(x_out)))

    def wait_host(self):
        self.ctx._ck(L.mlmcpi_sampler_wait_host(self.h))

    def level_model(self, level):
        m = Model()
        self.ctx._ck(L.mlmcpi_sampler_level_model(self.h, level, C.byref(m)))
        return m

    def p_accept(self):
        out = (C.c_double * self.n_levels)()
        self.ctx._ck(L.mlmcpi_sampler_stats(self.h, out))
        return list(out)

    def reset_stats(self):
        self.ctx._ck(L.mlmcpi_sampler_reset_stats(self.h))

    def autotune(self, p_accept_target=0.8, n_rounds=100, n_samples=1000):
        """HMCSampler::autotune_stepsize; returns (dt, last acceptance, converged)"""
        dt, pa = C.c_double(), C.c_double()
        rc = L.mlmcpi_sampler_autotune(self.h, p_accept_target, n_rounds, n_samples, C.byref(dt),
                                       C.byref(pa))
        if rc not in (0, 1):
            self.ctx._ck(rc)
        return dt.value, pa.value, rc == 0

    def set_dt(self, dt):
        L.mlmcpi_sampler_set_dt(self.h, dt)

    def cost_per_sample(self, n_meas=10):
        """microseconds per chain-sample (CUDA events over n_meas batched draws)"""
        v = C.c_double()
        self.ctx._ck(L.mlmcpi_sampler_cost(self.h, n_meas, C.byref(v)))
        return v.value

    def independence(self):
        """multilevel sampler: (t_indep[l], n_indep[l]) per level"""
        out = (C.c_double * (2 * self.n_levels))()
        self.ctx._ck(L.mlmcpi_sampler_indep(self.h, out))
        return list(out[:self.n_levels]), [int(v) for v in out[self.n_levels:]]

    def work(self):
        out = (C.c_double * 3)()
        L.mlmcpi_sampler_work(self.h, out)
        return dict(leapfrog_site_steps=out[0], sweep_site_updates=out[1], filled_fine_sites=out[2])


class MultilevelMC:
    """MonteCarloMultiLevel (montecarlo/montecarlomultilevel.cc), batched over B chains"""

    def __init__(self, ctx, fine, B, n_level, epsilon, qoi, n_burnin=100, n_autocorr_window=20,
                 n_min_samples_qoi=100, max_iterations=0, chain0=0, **sampler_kw):
        self.ctx, self.n_level = ctx, n_level
        sp = dict(kind=SAMPLER_HMC, n_levels=n_level, renorm=RENORM_NONE, ctype=COARSEN_BOTH, nt=100,
                  dt=0.1, n_rep=1, n_sweep_overrelax=10, n_sweep_heatbath=1, multilevel=0, qoi=qoi,
                  n_autocorr_window=n_autocorr_window, n_updates=10)
        sp.update(sampler_kw)
        prm = MlmcParams(n_level=n_level, n_burnin=n_burnin, epsilon=epsilon,
                         n_autocorr_window=n_autocorr_window, n_min_samples_qoi=n_min_samples_qoi,
                         qoi=qoi, max_iterations=max_iterations, sampler=SamplerParams(**sp))
        h = C.c_void_p()
        ctx._ck(L.mlmcpi_mlmc_create(ctx.h, C.byref(fine), C.byref(prm), B, chain0, C.byref(h)))
        self.h = h
        ctx._children.add(self)

    def close(self):
        if self.h:
            L.mlmcpi_mlmc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def evaluate(self):
        """returns True when the sample allocation converged"""
        rc = L.mlmcpi_mlmc_evaluate(self.h)
        if rc not in (0, 1):
            self.ctx._ck(rc)
        return rc == 0

    def result(self):
        v, e = C.c_double(), C.c_double()
        lv = (C.c_double * (6 * self.n_level))()
        self.ctx._ck(L.mlmcpi_mlmc_result(self.h, C.byref(v), C.byref(e), lv))
        keys = ("samples", "mean", "variance", "tau_int", "cost_eff_usec", "n_target")
        levels = [dict(zip(keys, lv[6 * l:6 * l + 6])) for l in range(self.n_level)]
        return v.value, e.value, levels


class Statistics:
    """per-chain Statistics accumulators (common/statistics.hh) on the device"""

    def __init__(self, ctx, k_max, B):
        self.ctx, self.k_max, self.B = ctx, k_max, B
        h = C.c_void_p()
        ctx._ck(L.mlmcpi_stats_create(ctx.h, k_max, B, C.byref(h)))
        self.h = h
        ctx._children.add(self)

    def close(self):
        if self.h:
            L.mlmcpi_stats_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        """Statistics::reset: clears the short-term mean and sample count only"""
        self.ctx._ck(L.mlmcpi_stats_reset(self.h))

    def hard_reset(self):
        self.ctx._ck(L.mlmcpi_stats_hard_reset(self.h))

    def record(self, q):
        self.ctx._ck(L.mlmcpi_stats_record(self.h, _ptr(q)))

    def pack_device(self, out=None):
        if out is None:
            out = self.ctx.empty(L.mlmcpi_stats_packed_size(self.k_max))
        self.ctx._ck(L.mlmcpi_stats_pack_device(self.h, _ptr(out)))
        return out

    def pack(self):
        out = np.zeros(L.mlmcpi_stats_packed_size(self.k_max))
        self.ctx._ck(L.mlmcpi_stats_pack(self.h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    @staticmethod
    def finalize(packed, k_max):
        """{average, variance, variance_error, tau_int, error, samples} from a packed
        (and possibly all-reduced) moment vector"""
        packed = np.ascontiguousarray(packed, dtype=np.float64)
        out = np.zeros(6)
        rc = L.mlmcpi_stats_finalize(packed.ctypes.data_as(C.POINTER(C.c_double)), k_max,
                                     out.ctypes.data_as(C.POINTER(C.c_double)))
        if rc:
            raise MlmcpiError(f"mlmcpi_stats_finalize failed ({rc})")
        return dict(average=out[0], variance=out[1], variance_error=out[2], tau_int=out[3],
                    error=out[4], samples=out[5])
