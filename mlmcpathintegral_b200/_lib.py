"""ctypes binding of libmlmcpi.so (the C-ABI declared in include/mlmcpi.h).

The library is the product: there is no Python or CPU fallback.  If it has not been
built (``python -c 'import __graft_entry__ as g; g.build()'`` or ``make -C
mlmcpathintegral_b200/csrc``) importing this module raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MLMCPI_LIB: alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("MLMCPI_LIB") or os.path.join(HERE, "libmlmcpi.so")

HO, QUARTIC, ROTOR, SCHWINGER, GFF = 0, 1, 2, 3, 4
COARSEN_BOTH, COARSEN_TEMPORAL, COARSEN_SPATIAL, COARSEN_ALTERNATE, COARSEN_ROTATE = range(5)
RENORM_NONE, RENORM_PERTURBATIVE, RENORM_NONPERTURBATIVE = range(3)
QOI_X2, QOI_ROTOR_CHI, QOI_SCHWINGER_CHI, QOI_AVG_PLAQUETTE, QOI_PHI2 = range(5)
SAMPLER_HMC, SAMPLER_HEATBATH, SAMPLER_CLUSTER, SAMPLER_EXACT = 0, 1, 2, 3
E_INVAL, E_CUDA, E_NOMEM, E_UNSUPPORTED = -1, -2, -3, -4
OPT_EXPCOS_ENVELOPE, OPT_LEAPFROG_VARIANT, OPT_LEAPFROG_ROWS, OPT_LEAPFROG_FUSE = 1, 2, 3, 4
OPT_SWEEP_REVERSE, OPT_OVERRELAX_ONE_PASS, OPT_FUSED_QM_HIERARCHY, OPT_GFF_COARSE_SMOOTHING, OPT_CASCADE_CACHE, OPT_TAU_REFRESH = 5, 6, 7, 8, 9, 10
OPT_HOST_COPY_ENGINE = 11


class Model(C.Structure):
    """mlmcpi_model"""
    _fields_ = [("model", C.c_int), ("M_lat", C.c_int), ("Mt_lat", C.c_int), ("Mx_lat", C.c_int),
                ("rotated", C.c_int), ("coarsening", C.c_int), ("a_lat", C.c_double),
                ("T_final", C.c_double), ("m0", C.c_double), ("mu2", C.c_double),
                ("lambda_", C.c_double), ("x0", C.c_double), ("beta", C.c_double),
                ("gff_mu2", C.c_double), ("gff_n_gibbs", C.c_int), ("gff_omega", C.c_double)]


class SamplerParams(C.Structure):
    """mlmcpi_sampler_params"""
    _fields_ = [("kind", C.c_int), ("n_levels", C.c_int), ("renorm", C.c_int), ("ctype", C.c_int),
                ("nt", C.c_int), ("dt", C.c_double), ("n_rep", C.c_int),
                ("n_sweep_overrelax", C.c_int), ("n_sweep_heatbath", C.c_int),
                ("multilevel", C.c_int), ("qoi", C.c_int), ("n_autocorr_window", C.c_int),
                ("n_updates", C.c_int)]


class MlmcParams(C.Structure):
    """mlmcpi_mlmc_params"""
    _fields_ = [("n_level", C.c_int), ("n_burnin", C.c_int), ("epsilon", C.c_double),
                ("n_autocorr_window", C.c_int), ("n_min_samples_qoi", C.c_int), ("qoi", C.c_int),
                ("max_iterations", C.c_int), ("sampler", SamplerParams)]


# every symbol include/mlmcpi.h declares: name -> (restype, argtypes)
_vp, _i, _u32, _u64, _d, _sz = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_double, C.c_size_t
_MP = C.POINTER(Model)
# mlmcpi_allreduce_fn: int (*)(void *user, double *d_buf, size_t n)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t)
_ip = C.POINTER(C.c_int)
_u32p = C.POINTER(C.c_uint32)
_dp = C.POINTER(C.c_double)
SIGNATURES = {
    "mlmcpi_version": (_i, []),
    "mlmcpi_create": (_i, [C.POINTER(_vp), _i, _u64, _vp]),
    "mlmcpi_destroy": (None, [_vp]),
    "mlmcpi_last_error": (C.c_char_p, [_vp]),
    "mlmcpi_sync": (_i, [_vp]),
    "mlmcpi_set_allreduce": (_i, [_vp, _vp, _vp, _i, _i]),
    "mlmcpi_world_size": (_i, [_vp]),
    "mlmcpi_rank": (_i, [_vp]),
    "mlmcpi_device": (_i, [_vp]),
    "mlmcpi_stream": (_vp, [_vp]),
    "mlmcpi_set_seed": (_i, [_vp, _u64]),
    "mlmcpi_set_option": (_i, [_vp, _i, _i]),
    "mlmcpi_launch_count": (_u64, [_vp]),
    "mlmcpi_profile": (_i, [_vp, _i]),
    "mlmcpi_profile_read": (_i, [_vp, _dp]),
    "mlmcpi_alloc": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "mlmcpi_free": (_i, [_vp, _vp]),
    "mlmcpi_upload": (_i, [_vp, _vp, _vp, _sz]),
    "mlmcpi_download": (_i, [_vp, _vp, _vp, _sz]),
    "mlmcpi_copy": (_i, [_vp, _vp, _vp, _sz]),
    "mlmcpi_axpy": (_i, [_vp, _vp, _vp, _d, _vp, _sz]),
    "mlmcpi_sample_size": (_i, [_MP]),
    "mlmcpi_vertex_cart2lin": (_u32, [_i, _i, _i, _i, _i]),
    "mlmcpi_vertex_lin2cart": (None, [_i, _i, _i, _u32, _ip, _ip]),
    "mlmcpi_link_cart2lin": (_u32, [_i, _i, _i, _i, _i]),
    "mlmcpi_link_lin2cart": (None, [_i, _i, _u32, _ip, _ip, _ip]),
    "mlmcpi_neighbours": (None, [_i, _i, _i, _u32, _u32p]),
    "mlmcpi_coarse_shape": (_i, [_i, _i, _i, _i, _ip, _ip, _ip]),
    "mlmcpi_coarsening_lists": (_i, [_i, _i, _i, _i, _u32p, _u32p, _u32p, _ip]),
    "mlmcpi_coarse_model": (_i, [_MP, _i, _i, _i, _d, _MP]),
    "mlmcpi_init_state": (_i, [_vp, _MP, _vp, _i, _u32, _u64]),
    "mlmcpi_action": (_i, [_vp, _MP, _vp, _i, _vp]),
    "mlmcpi_force": (_i, [_vp, _MP, _vp, _vp, _i]),
    "mlmcpi_leapfrog": (_i, [_vp, _MP, _i, _d, _vp, _vp, _i]),
    "mlmcpi_hmc_momentum": (_i, [_vp, _MP, _vp, _i, _u32, _u64]),
    "mlmcpi_hmc_step": (_i, [_vp, _MP, _i, _d, _vp, _i, _u32, _u64, _vp, _vp]),
    "mlmcpi_overrelax_sweep": (_i, [_vp, _MP, _vp, _i]),
    "mlmcpi_overrelax_sweeps": (_i, [_vp, _MP, _vp, _i, _i]),
    "mlmcpi_heatbath_sweep": (_i, [_vp, _MP, _vp, _i, _u32, _u64]),
    "mlmcpi_prolong": (_i, [_vp, _MP, _vp, _vp, _i]),
    "mlmcpi_restrict": (_i, [_vp, _MP, _vp, _vp, _i]),
    "mlmcpi_fill": (_i, [_vp, _MP, _vp, _i, _u32, _u64]),
    "mlmcpi_sigma_hat": (C.c_double, [C.c_double, C.c_uint]),
    "mlmcpi_schwinger_chit_analytical": (C.c_double, [C.c_double, C.c_uint]),
    "mlmcpi_schwinger_chit_perturbative": (C.c_double, [C.c_double, C.c_uint]),
    "mlmcpi_schwinger_var_chit_continuum": (C.c_double, [C.c_double, C.c_uint]),
    "mlmcpi_rotor_chit": (C.c_double, [C.c_double, C.c_double, C.c_double, _i]),
    "mlmcpi_gff_phi_squared_analytical": (C.c_double, [C.c_double, _i, _i]),
    "mlmcpi_ho_xsquared_analytical": (C.c_double, [C.c_double, C.c_double, C.c_double, _i, _i]),
    "mlmcpi_schwinger_betacoarse_nonperturbative": (C.c_double, [C.c_double, C.c_uint, _i]),
    "mlmcpi_exact_draw": (_i, [_vp, _MP, _vp, _i, _u32, _u64]),
    "mlmcpi_prolong_fill": (_i, [_vp, _MP, _vp, _vp, _i, _u32, _u64]),
    "mlmcpi_prolong_fill_eval": (_i, [_vp, _MP, _vp, _vp, _i, _u32, _u64, _vp]),
    "mlmcpi_cluster_update": (_i, [_vp, _MP, _vp, _i, _u32, _u64, _i]),
    "mlmcpi_schwinger_from_cluster": (_i, [_vp, _MP, _vp, _vp, _i, _u32, _u64]),
    "mlmcpi_cond_action": (_i, [_vp, _MP, _vp, _i, _vp]),
    "mlmcpi_qoi": (_i, [_vp, _MP, _i, _vp, _i, _vp, _vp]),
    "mlmcpi_thermal_state": (_i, [_vp, _MP, _vp, _i, _u32]),
    "mlmcpi_twolevel_step": (_i, [_vp, _MP, _MP, _vp, _vp, _vp, _vp, _i, _u32, _u64, _vp, _vp]),
    "mlmcpi_sampler_create": (_i, [_vp, _MP, C.POINTER(SamplerParams), _i, _u32, C.POINTER(_vp)]),
    "mlmcpi_sampler_destroy": (None, [_vp]),
    "mlmcpi_sampler_set_state": (_i, [_vp, _vp]),
    "mlmcpi_sampler_draw": (_i, [_vp, _vp, _vp]),
    "mlmcpi_sampler_draw_host": (_i, [_vp, _vp, _i, _vp, _vp]),
    "mlmcpi_sampler_level_model": (_i, [_vp, _i, _MP]),
    "mlmcpi_dof_update": (_i, [_vp, _MP, _vp, _i, _i, _i, _u32, _u64]),
    "mlmcpi_sampler_get_state": (_i, [_vp, _vp]),
    "mlmcpi_sampler_qoi": (_i, [_vp, _i, _vp]),
    "mlmcpi_sampler_draw_host_async": (_i, [_vp, _i, _vp, _vp]),
    "mlmcpi_sampler_wait_host": (_i, [_vp]),
    "mlmcpi_sampler_stats": (_i, [_vp, _dp]),
    "mlmcpi_sampler_reset_stats": (_i, [_vp]),
    "mlmcpi_sampler_work": (_i, [_vp, _dp]),
    "mlmcpi_sampler_autotune": (_i, [_vp, _d, _i, _i, _dp, _dp]),
    "mlmcpi_sampler_set_dt": (_i, [_vp, _d]),
    "mlmcpi_sampler_cost": (_i, [_vp, _i, _dp]),
    "mlmcpi_sampler_indep": (_i, [_vp, _dp]),
    "mlmcpi_mlmc_create": (_i, [_vp, _MP, C.POINTER(MlmcParams), _i, _u32, C.POINTER(_vp)]),
    "mlmcpi_mlmc_destroy": (None, [_vp]),
    "mlmcpi_mlmc_evaluate": (_i, [_vp]),
    "mlmcpi_mlmc_result": (_i, [_vp, _dp, _dp, _dp]),
    "mlmcpi_stats_create": (_i, [_vp, _i, _i, C.POINTER(_vp)]),
    "mlmcpi_stats_destroy": (None, [_vp]),
    "mlmcpi_stats_reset": (_i, [_vp]),
    "mlmcpi_stats_hard_reset": (_i, [_vp]),
    "mlmcpi_stats_record": (_i, [_vp, _vp]),
    "mlmcpi_stats_pack": (_i, [_vp, _dp]),
    "mlmcpi_stats_pack_device": (_i, [_vp, _vp]),
    "mlmcpi_stats_packed_size": (_i, [_i]),
    "mlmcpi_stats_finalize": (_i, [_dp, _i, _dp]),
}


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: mlmcpathintegral_b200 has no CPU fallback; build the CUDA "
            "library first (python -c 'import __graft_entry__ as g; g.build()')")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(lib, name)  # AttributeError if the library does not export it
        f.restype = res
        f.argtypes = args
    return lib


lib = load()
