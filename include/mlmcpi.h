/* mlmcpi.h -- C-ABI of the B200-native sampler inner loop (libmlmcpi.so).
 *
 * The reference (eikehmueller/mlmcpathintegral) has no FFI layer: its seam is the
 * set of C++ abstract classes Action / ConditionedFineAction / Sampler / QoI
 * (SURVEY.md 8b).  Every entry point below replaces the loop body behind one of
 * those virtuals and cites it.  The C++ classes in include/mlmcpi/ that carry the
 * reference's names are thin adapters over this file.
 *
 * Conventions
 *   - every function returns 0 on success, a negative MLMCPI_E* code otherwise,
 *     never throws; mlmcpi_last_error(ctx) gives the message.
 *   - the hot-path functions take DEVICE pointers to fp64 and an integer batch
 *     of B independent chains.  A batched state is laid out [chain][dof] with the
 *     dof index exactly the reference's (lattice/lattice2d.hh:230-245 vertices,
 *     :348-353 links ell = 2*Mt*j + 2*i + mu; 1-D paths contiguous), so upload and
 *     download are plain copies and one chain is bit-identical to the reference's
 *     SampleState::data (common/samplestate.hh:48).
 *   - kernels are launched asynchronously on the context's stream; *_host entry
 *     points take HOST pointers, copy in, run, copy out and synchronise.
 *   - there is no CPU fallback: without a CUDA device mlmcpi_create fails.
 *
 * Random streams (counter-based Philox4x32-10, replaces std::mt19937_64,
 * SURVEY 7.3-5): counter = (index, chain, draw_lo, stream<<24 | call number),
 * key = (seed_lo, seed_hi ^ draw_hi); two 53-bit uniforms per call; normals by
 * Box-Muller.  `chain0` is the global index of the first chain of the batch (the
 * multi-GPU sharding offsets it), `draw` a caller-maintained draw counter.
 */
#ifndef MLMCPI_H
#define MLMCPI_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLMCPI_VERSION 100

enum { MLMCPI_OK = 0, MLMCPI_EINVAL = -1, MLMCPI_ECUDA = -2, MLMCPI_ENOMEM = -3,
       MLMCPI_EUNSUPPORTED = -4 };

/* models: action/qm/{harmonicoscillator,quarticoscillator,rotor}action.hh,
 * action/qft/{quenchedschwinger,gff}action.hh */
enum { MLMCPI_HO = 0, MLMCPI_QUARTIC = 1, MLMCPI_ROTOR = 2, MLMCPI_SCHWINGER = 3,
       MLMCPI_GFF = 4 };
/* lattice/lattice2d.hh:18-26 */
enum { MLMCPI_COARSEN_BOTH = 0, MLMCPI_COARSEN_TEMPORAL = 1, MLMCPI_COARSEN_SPATIAL = 2,
       MLMCPI_COARSEN_ALTERNATE = 3, MLMCPI_COARSEN_ROTATE = 4 };
/* action/renormalisation.hh:17-21 */
enum { MLMCPI_RENORM_NONE = 0, MLMCPI_RENORM_PERTURBATIVE = 1,
       MLMCPI_RENORM_NONPERTURBATIVE = 2 };
/* qoi/qm/qoixsquared.hh, qoi/qm/qoisusceptibility.hh, qoi/qft/qoi2dsusceptibility.hh,
 * qoi/qft/qoiavgplaquette.hh, qoi/qft/qoi2dphisquared.hh */
enum { MLMCPI_QOI_X2 = 0, MLMCPI_QOI_ROTOR_CHI = 1, MLMCPI_QOI_SCHWINGER_CHI = 2,
       MLMCPI_QOI_AVG_PLAQUETTE = 3, MLMCPI_QOI_PHI2 = 4 };
/* Philox4x32-10 streams: counter = (index, global chain, draw, stream << 24 | call), key = (seed, draw_hi).
 * Variate-to-degree-of-freedom maps that are not one stream per dof:
 *   FILL3, coarsening both: the two horizontal interior links (2i, 2j+1, 0), (2i+1, 2j+1, 0) of a coarse cell take
 *     the first attempts of their ExpCos draws from ONE normal pair and ONE uniform pair -- calls 0, 1 of the
 *     stream with index Mt j + 2i: (z0, u0) and (z1, u1); further attempts continue on the link's own stream
 *     (index Mt j + 2i from call 2, index Mt j + 2i + 1 from call 0);
 *   HEATBATH, quenched Schwinger model: the links of a colour in a lattice row, numbered n = i (mu = 0) or i / 2
 *     (mu = 1), are updated in pairs (2p, 2p + 1): calls 0, 1 of the stream of the EVEN link give (z0, u0), (z1, u1),
 *     the first ExpCos attempt of the even and of the odd link; further attempts continue on the link's own stream
 *     (even link from call 2, odd link from call 0).  mlmcpi_dof_update follows the same map;
 *   HEATBATH, Gaussian free field: the vertices 2q and 2q + 1 take the two Box-Muller normals of ONE block,
 *     index 2q: z0 the even, z1 the odd vertex (sweeps and mlmcpi_dof_update alike);
 *   CLUSTER: index 0 = (reflection angle, start site) of an update, index 1 + k = (forward, backward) uniform of
 *     the link between the sites k and k + 1. */
enum { MLMCPI_STREAM_INIT = 1, MLMCPI_STREAM_HMC_MOMENTUM = 2, MLMCPI_STREAM_HMC_ACCEPT = 3,
       MLMCPI_STREAM_HEATBATH = 4, MLMCPI_STREAM_FILL1 = 5, MLMCPI_STREAM_FILL2 = 6,
       MLMCPI_STREAM_FILL3 = 7, MLMCPI_STREAM_TWOLEVEL_ACCEPT = 8, MLMCPI_STREAM_CLUSTER = 9,
       MLMCPI_STREAM_GAUGE = 10, MLMCPI_STREAM_EXACT = 11 };
/* coarse-level samplers of sampler/hierarchicalsampler.hh */
/* MLMCPI_SAMPLER_CLUSTER: ClusterSampler (rotor, sampler/clustersampler.cc) or
 * QuenchedSchwingerClusterSampler (sampler/quenchedschwingerclustersampler.cc) */
/* MLMCPI_SAMPLER_EXACT: independent exact draws (harmonic oscillator: the Cholesky sampler of
 * qm/harmonicoscillatoraction.cc:38-66) */
enum { MLMCPI_SAMPLER_HMC = 0, MLMCPI_SAMPLER_HEATBATH = 1, MLMCPI_SAMPLER_CLUSTER = 2,
       MLMCPI_SAMPLER_EXACT = 3 };

/* One level of one model: the data members of the reference's action classes. */
typedef struct mlmcpi_model {
  int model;
  int M_lat;          /* 1-D: number of sites            (lattice/lattice1d.hh) */
  int Mt_lat, Mx_lat; /* 2-D                             (lattice/lattice2d.hh) */
  int rotated;        /* 2-D vertex lattice: rotated level of CoarsenRotate     */
  int coarsening;     /* how THIS level coarsens to the next: BOTH / TEMPORAL /
                         SPATIAL / ROTATE (ALTERNATE resolved per level)        */
  double a_lat;       /* 1-D lattice spacing T/M                                */
  double T_final;     /* 1-D total time                                         */
  double m0, mu2, lambda, x0; /* QM couplings                                   */
  double beta;        /* Schwinger coupling                                     */
  double gff_mu2;     /* GFF: a^2 m^2 (qft/gffaction.hh:174-181)                */
  int gff_n_gibbs;    /* GFF: n_gibbs_smooth (qft/gffaction.hh:161-168).  0: the 5-point action;
                         > 0: the Gibbs-smoothed coarse-level action S = phi^T Q_hat phi / 2 with
                         the dense precision matrix of gffaction.cc:133-174 (<= MLMCPI_GFF_DENSE_MAX
                         vertices), whatever sampler runs on the level -- as in the reference     */
  double gff_omega;   /* GFF: overrelaxation factor of the Gibbs smoother                      */
} mlmcpi_model;

/* Largest number of vertices for which the dense matrices of GFFAction::buildMatrices are formed
 * (gffaction.cc:133-174; on the device with cuSOLVER / cuBLAS: four N x N fp64 work matrices, about
 * 10 N^3 flops, once per level and context).  32768 = level 1 of BASELINE config C3 (256 x 256,
 * coarsening rotate: 65536 / 32768 / 16384 / 8192 vertices): 8.6 GB per matrix.  mlmcpi_coarse_model
 * gives EVERY coarse GFF level the reference's n_gibbs_smooth = 2, omega = 1 (gffaction.hh:201-208);
 * a level above this size makes mlmcpi_action / mlmcpi_exact_draw fail with MLMCPI_EUNSUPPORTED
 * (there is no silent fall-back to the 5-point action). */
#define MLMCPI_GFF_DENSE_MAX 32768

typedef struct mlmcpi_ctx mlmcpi_ctx;

/* ---- context ------------------------------------------------------------ */
int mlmcpi_version(void);
/* stream: the cudaStream_t all kernels and copies are issued on (e.g. torch's current
 * stream); NULL is CUDA's legacy default stream; MLMCPI_OWN_STREAM creates a private
 * non-blocking stream (the caller then orders its own work with mlmcpi_sync) */
#define MLMCPI_OWN_STREAM ((void *)(intptr_t)-1)
int mlmcpi_create(mlmcpi_ctx **ctx, int device, uint64_t seed, void *stream);
void mlmcpi_destroy(mlmcpi_ctx *ctx);
const char *mlmcpi_last_error(const mlmcpi_ctx *ctx);
int mlmcpi_sync(mlmcpi_ctx *ctx);
/* the device index and the cudaStream_t (as void *) of the context, for code that issues its own
 * work in order with the library's (e.g. the NCCL all-reduce of libmlmcpi_comm.so) */
int mlmcpi_device(const mlmcpi_ctx *ctx);
void *mlmcpi_stream(const mlmcpi_ctx *ctx);
int mlmcpi_set_seed(mlmcpi_ctx *ctx, uint64_t seed);
/* The one exchange between the processes of a run (one process per GPU): an in-place
 * SUM over all processes of n doubles in DEVICE memory, issued in order on the stream of ctx
 * (ncclAllReduce: mlmcpi_comm_attach of libmlmcpi_comm.so installs it; any other transport may be
 * plugged in).  Once set, every Statistics query the library makes for a host-side decision --
 * tau_int in MultilevelSampler::draw (sampler/multilevelsampler.cc:84-99) and in
 * MonteCarloMultiLevel::draw_coarse_sample (montecarlo/montecarlomultilevel.cc:170-190), the
 * variances / sample counts of the allocation loop (:139-158), the measured costs and the HMC
 * autotune acceptance -- is taken over the chains of ALL processes, which is what
 * mpi_allreduce_avg does inside the reference's Statistics (common/statistics.cc:30-35,64-79).
 * All processes then take identical decisions and stay in lockstep.  world_size / rank describe
 * the calling process; fn == NULL removes the hook. */
typedef int (*mlmcpi_allreduce_fn)(void *user, double *d_buf, size_t n);
int mlmcpi_set_allreduce(mlmcpi_ctx *ctx, mlmcpi_allreduce_fn fn, void *user, int world_size, int rank);
int mlmcpi_world_size(const mlmcpi_ctx *ctx);
int mlmcpi_rank(const mlmcpi_ctx *ctx);
/* options.  MLMCPI_OPT_EXPCOS_ENVELOPE: proposal of the ExpCos rejection sampler
 * (distribution/expcosdistribution.hh:50-65): 0 = the reference's Gaussian envelope
 * (variance 2 pi^2/tau, ~22 % acceptance), 1 = chord-bound envelope (variance
 * pi^2/(4 tau), ~64 % acceptance), 2 = chord bound for tau < 64 and the Taylor bound
 * 1 - cos x >= x^2/2 (1 - x^2/12) for tau >= 64 (> 89 % acceptance, 99.7 % at tau = 2048;
 * default).  The sampled distribution is the same (variant 2 truncates a tail mass < 1e-33). */
/* MLMCPI_OPT_LEAPFROG_VARIANT (2-D Schwinger leapfrog kernel; all variants compute the same
 * step): 0 = TMA/mbarrier row pipeline (default), 1 = register row march, 2 = generic.
 * MLMCPI_OPT_LEAPFROG_ROWS: lattice rows per thread block (0 = default).
 * MLMCPI_OPT_LEAPFROG_FUSE: leapfrog steps per pass over HBM (temporal blocking, variant 0 only):
 * 0 = one; 1 (default) = four for Mt in {64, 128, 256}, two for Mt = 512 (K-stage register pipeline
 * with compile-time block size), two for every other Mt (round-1 kernel); 2 / 3 = two / four steps
 * through the K-stage kernel; 4 = the round-1 two-step kernel; 5 = eight steps (Mt <= 128; 236 registers, two
 * blocks per SM: measured slower than four, 29.0 vs 25.4 us per step).  Same trajectory, bit for bit.
 * MLMCPI_OPT_SWEEP_REVERSE: 1 = the coloured sweeps visit the colours in descending order (the
 * exact reverse of the default; used to make a sequence of sweeps a reversible kernel).
 * MLMCPI_OPT_OVERRELAX_ONE_PASS: 1 (default) = a Schwinger overrelaxation sweep updates all four
 * colours in one pass over HBM (row pipeline, out of place), 0 = four colour passes; same result.
 * MLMCPI_OPT_FUSED_QM_HIERARCHY: 1 (default) = HierarchicalSampler::draw for 1-D paths with an HMC coarse
 * sampler runs as ONE kernel (one warp per chain, every level on chip), 0 = the sequence of
 * single-purpose kernels; same draw.
 * MLMCPI_OPT_CASCADE_CACHE: 1 (default) = HierarchicalSampler::draw of the quenched Schwinger model with the
 *   HMC coarse sampler keeps the coarse level states tentative until the whole cascade has accepted, which
 *   makes the per-draw restriction chain and its action / conditioned-action reductions redundant (their
 *   results are cached); 0 = the literal sequence of hierarchicalsampler.cc:55-81.  Same draws, bit for bit.
 * MLMCPI_OPT_TAU_REFRESH: the level walks of MultilevelSampler::draw and MonteCarloMultiLevel::draw_coarse_sample
 *   decide on tau_int of a device Statistics object (a pack kernel, a device-to-host copy, a stream
 *   synchronisation and, with several processes, an all-reduce).  The reference asks at every sample; here an
 *   answer is reused for a number of calls that doubles from 1 up to this value (default 16; 1 = every call).
 * MLMCPI_OPT_HOST_COPY_ENGINE: how mlmcpi_sampler_draw_host_async hands the accepted chains' states to a PINNED host
 *   buffer.  1 (default) = the copy engine: the step's accept flags go to the host, the next call (or
 *   mlmcpi_sampler_wait_host) issues one copy per run of accepted chains -- it blocks until the previous step's draw has
 *   finished -- 54 GB/s on the host link; 0 = a kernel stores the accepted rows straight into the (device-addressable)
 *   buffer, no host round trip, 38 GB/s.  Same bytes in the buffer.
 * MLMCPI_OPT_GFF_COARSE_SMOOTHING: 1 (default) = the coarse levels a sampler / multilevel driver builds
 *   for a GFF carry the reference's Gibbs-smoothed action Q_hat (gffaction.hh:201-208), whatever sampler
 *   runs on them; 0 = the plain 5-point action on every level (consistent with a heat-bath / HMC coarse
 *   sampler, but the two-level acceptance is ~ 0 beyond 16 x 16). */
enum { MLMCPI_OPT_EXPCOS_ENVELOPE = 1, MLMCPI_OPT_LEAPFROG_VARIANT = 2, MLMCPI_OPT_LEAPFROG_ROWS = 3,
       MLMCPI_OPT_LEAPFROG_FUSE = 4, MLMCPI_OPT_SWEEP_REVERSE = 5, MLMCPI_OPT_OVERRELAX_ONE_PASS = 6,
       MLMCPI_OPT_FUSED_QM_HIERARCHY = 7, MLMCPI_OPT_GFF_COARSE_SMOOTHING = 8, MLMCPI_OPT_CASCADE_CACHE = 9, MLMCPI_OPT_TAU_REFRESH = 10,
       MLMCPI_OPT_HOST_COPY_ENGINE = 11 };
int mlmcpi_set_option(mlmcpi_ctx *ctx, int option, int value);
/* number of kernels this context has launched so far */
uint64_t mlmcpi_launch_count(const mlmcpi_ctx *ctx);
/* CUDA-event timing of the leapfrog launches (the dominant kernel) on the context's
 * stream.  _read synchronises, returns and clears out = {milliseconds, launches,
 * algorithmic bytes (R theta, R p, W theta, W p per site-step)} */
int mlmcpi_profile(mlmcpi_ctx *ctx, int enable);
int mlmcpi_profile_read(mlmcpi_ctx *ctx, double out[3]);

/* ---- memory: SampleState storage (common/samplestate.hh:19-53) ----------- */
int mlmcpi_alloc(mlmcpi_ctx *ctx, size_t n_doubles, double **d_ptr); /* zero-filled */
int mlmcpi_free(mlmcpi_ctx *ctx, double *d_ptr);
int mlmcpi_upload(mlmcpi_ctx *ctx, double *d_dst, const double *h_src, size_t n);
int mlmcpi_download(mlmcpi_ctx *ctx, double *h_dst, const double *d_src, size_t n);
int mlmcpi_copy(mlmcpi_ctx *ctx, double *d_dst, const double *d_src, size_t n);
/* d_out[k] = d_a[k] + alpha * d_b[k] (e.g. the MLMC differences Y = Q_fine - Q_coarse of
 * montecarlo/montecarlotwolevel.cc:59 without a host round trip) */
int mlmcpi_axpy(mlmcpi_ctx *ctx, double *d_out, const double *d_a, double alpha, const double *d_b,
                size_t n);

/* ---- geometry (host, integer, bit-exact with lattice/lattice2d.{hh,cc}) -- */
int mlmcpi_sample_size(const mlmcpi_model *m);
uint32_t mlmcpi_vertex_cart2lin(int Mt, int Mx, int rotated, int i, int j);       /* lattice2d.hh:230-245 */
void mlmcpi_vertex_lin2cart(int Mt, int Mx, int rotated, uint32_t ell, int *i, int *j); /* :255-268 */
uint32_t mlmcpi_link_cart2lin(int Mt, int Mx, int i, int j, int mu);              /* :348-353 */
void mlmcpi_link_lin2cart(int Mt, int Mx, uint32_t ell, int *i, int *j, int *mu); /* :367-375 */
void mlmcpi_neighbours(int Mt, int Mx, int rotated, uint32_t ell, uint32_t nb[8]); /* lattice2d.cc:135-155 */
int mlmcpi_coarse_shape(int Mt, int Mx, int ctype, int level, int *Mt_c, int *Mx_c, int *rot_c); /* lattice2d.cc:20-81 */
/* coarse / fine-only vertex lists and fine->coarse map (lattice2d.cc:82-130);
 * buffers hold n_vertices entries each; counts = {n_coarse, n_fineonly} */
int mlmcpi_coarsening_lists(int Mt, int Mx, int ctype, int level, uint32_t *coarse,
                            uint32_t *fineonly, uint32_t *map_vals, int *counts);
/* Action::coarse_action() parameters (qm/{harmonicoscillator,rotor}renormalisation.hh,
 * qft/quenchedschwingerrenormalisation.hh:45-105, qft/gffaction.hh:201-208) */
int mlmcpi_coarse_model(const mlmcpi_model *fine, int renorm, int level, int ctype,
                        double T_final, mlmcpi_model *coarse);

/* ---- analytic results and coupling matching (host, setup time only) ------------------
 * The reference evaluates these with GSL quadrature / root finding (common/auxilliary.cc:44-209,
 * qft/quenchedschwingerrenormalisation.cc:7-64); here: composite Gauss-Legendre + bisection. */
double mlmcpi_sigma_hat(double xi, unsigned int p);                               /* auxilliary.cc:7-29 */
/* E[V chi_t] of the quenched Schwinger model (qoi/qft/qoi2dsusceptibility.cc:30-50); the exact
 * expression is defined for beta <= 2000 (NaN beyond, where the drivers use the perturbative one) */
double mlmcpi_schwinger_chit_analytical(double beta, unsigned int n_plaq);
double mlmcpi_schwinger_chit_perturbative(double beta, unsigned int n_plaq);
double mlmcpi_schwinger_var_chit_continuum(double beta, unsigned int n_plaq);
/* RotorAction::chit_exact / chit_perturbative / chit_continuum (qm/rotoraction.cc:92-115):
 * which = 0 / 1 / 2 */
double mlmcpi_rotor_chit(double m0, double a_lat, double T_final, int which);
double mlmcpi_gff_phi_squared_analytical(double mass, int Mt_lat, int Mx_lat);   /* auxilliary.cc:197-209 */
/* HarmonicOscillatorAction::Xsquared_analytical(_continuum) (qm/harmonicoscillatoraction.cc:69-80) */
double mlmcpi_ho_xsquared_analytical(double m0, double mu2, double a_lat, int M_lat, int continuum);
/* RenormalisedQuenchedSchwingerParameters::betacoarse_nonperturbative: n_plaq plaquettes on the
 * fine lattice, rho_refine = 4 (coarsening both) or 2; used by mlmcpi_coarse_model for
 * MLMCPI_RENORM_NONPERTURBATIVE */
double mlmcpi_schwinger_betacoarse_nonperturbative(double beta, unsigned int n_plaq, int rho_refine);

/* ---- group 1: action, force, HMC ------------------------------------------ */
/* Action::initialise_state (rotoraction.cc:82-85, quenchedschwingeraction.cc:198-204) */
int mlmcpi_init_state(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *d_x, int B,
                      uint32_t chain0, uint64_t draw);
/* Action::evaluate: d_S[B] */
int mlmcpi_action(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *d_x, int B, double *d_S);
/* Action::force: d_f[B][n] */
int mlmcpi_force(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *d_x, double *d_f, int B);
/* leapfrog trajectory of HMCSampler::single_step (sampler/hmcsampler.cc:31-46), in place */
int mlmcpi_leapfrog(mlmcpi_ctx *ctx, const mlmcpi_model *m, int nt, double dt, double *d_x,
                    double *d_p, int B);
/* momentum refresh (sampler/hmcsampler.cc:24-26) */
int mlmcpi_hmc_momentum(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *d_p, int B,
                        uint32_t chain0, uint64_t draw);
/* HMCSampler::single_step (sampler/hmcsampler.cc:22-69) for B chains: d_x updated
 * where accepted; d_accept[B] (int32) and d_diag[B][5] = {deltaH, S_cur, S_trial,
 * T_cur, T_trial} may be NULL */
int mlmcpi_hmc_step(mlmcpi_ctx *ctx, const mlmcpi_model *m, int nt, double dt, double *d_x,
                    int B, uint32_t chain0, uint64_t draw, int32_t *d_accept, double *d_diag);

/* ---- group 2: sweeps, prolongation, fill-in ------------------------------- */
/* one coloured sweep of Action::overrelaxation_update over all dofs
 * (sampler/overrelaxedheatbathsampler.cc:10-18; colours: SURVEY 7.4) */
int mlmcpi_overrelax_sweep(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *d_x, int B);
/* n_sweeps of them back to back (the n_sweep_overrelax loop of overrelaxedheatbathsampler.cc:10-18);
 * for the Schwinger model the sweeps ping-pong between d_x and a work buffer, one pass over HBM each */
int mlmcpi_overrelax_sweeps(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *d_x, int B, int n_sweeps);
/* one coloured sweep of Action::heatbath_update (overrelaxedheatbathsampler.cc:20-27) */
int mlmcpi_heatbath_sweep(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *d_x, int B,
                          uint32_t chain0, uint64_t draw);
/* Action::heatbath_update / Action::overrelaxation_update of ONE degree of freedom ell (the per-dof
 * interface of action/action.hh:85-110; rotoraction.cc:21-56, gffaction.cc:32-42,68-79,
 * quenchedschwingeraction.cc:46-65) on all chains: the building block the reference's
 * OverrelaxedHeatBathSampler loops over (overrelaxedheatbathsampler.cc:8-31, any order incl.
 * random_order).  One launch per call -- present for interface parity; the sweeps above are the
 * fast path.  heatbath != 0: variates of (chain0 + chain, draw, ell), as in mlmcpi_heatbath_sweep. */
int mlmcpi_dof_update(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *d_x, int B, int ell, int heatbath,
                      uint32_t chain0, uint64_t draw);
/* Action::copy_from_coarse of the fine action / Action::copy_from_fine of the coarse one */
int mlmcpi_prolong(mlmcpi_ctx *ctx, const mlmcpi_model *fine, const double *d_xc, double *d_x, int B);
int mlmcpi_restrict(mlmcpi_ctx *ctx, const mlmcpi_model *fine, const double *d_xf, double *d_xc, int B);
/* ConditionedFineAction::fill_fine_points, in place on a prolonged state */
int mlmcpi_fill(mlmcpi_ctx *ctx, const mlmcpi_model *fine, double *d_x, int B, uint32_t chain0,
                uint64_t draw);
/* prolongation and fill-in fused: theta' written straight from the coarse state
 * (TwoLevelMetropolisStep::draw lines 40-42 in one pass) */
int mlmcpi_prolong_fill(mlmcpi_ctx *ctx, const mlmcpi_model *fine, const double *d_xc,
                        double *d_x, int B, uint32_t chain0, uint64_t draw);
/* the same, and in the same pass the two reductions TwoLevelMetropolisStep::draw needs of the
 * new trial state (lines 48 and 65-66): d_S[0..B) = Action::evaluate(theta'), d_S[B..2B) =
 * ConditionedFineAction::evaluate(theta').  For the Schwinger model with coarsening `both` they
 * are accumulated from the cell's links while these are still in registers */
int mlmcpi_prolong_fill_eval(mlmcpi_ctx *ctx, const mlmcpi_model *fine, const double *d_xc,
                             double *d_x, int B, uint32_t chain0, uint64_t draw, double *d_S);

/* n_updates Wolff single-cluster updates of the rotor (ClusterSampler::single_cluster_update1d,
 * sampler/clustersampler.cc:88-132 with RotorAction::S_ell / new_reflection / flip,
 * action/qm/rotoraction.hh:226-253); `update0` numbers the first update (Philox draw counter) */
int mlmcpi_cluster_update(mlmcpi_ctx *ctx, const mlmcpi_model *rotor, double *d_x, int B,
                          uint32_t chain0, uint64_t update0, int n_updates);
/* Exact samplers, independent samples for every chain:
 *  - harmonic oscillator: HarmonicOscillatorAction::draw (qm/harmonicoscillatoraction.cc:59-66),
 *    x = L_cov y with L_cov the Cholesky factor of the inverse of the cyclic tridiagonal precision
 *    matrix (build_covariance :38-56);
 *  - GFF: GFFAction::draw (qft/gffaction.cc:200-213), phi = U^{-1} psi with U the upper Cholesky
 *    factor of the 5-point precision matrix, followed by gff_n_gibbs lexicographic sweeps of
 *    global_heatbath_update_eff (:45-65); at most MLMCPI_GFF_DENSE_MAX vertices.
 * The dense factors are computed on the host once per model and cached in the context. */
int mlmcpi_exact_draw(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *d_x, int B, uint32_t chain0,
                      uint64_t draw);
/* QuenchedSchwingerClusterSampler::draw lines 52-82: links from the rotor chain psi
 * (length Mt*Mx) followed by a random gauge transformation */
int mlmcpi_schwinger_from_cluster(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *d_psi,
                                  double *d_x, int B, uint32_t chain0, uint64_t draw);

/* ---- group 3: reductions, acceptance, QoIs -------------------------------- */
/* ConditionedFineAction::evaluate: d_S[B] */
int mlmcpi_cond_action(mlmcpi_ctx *ctx, const mlmcpi_model *fine, const double *d_x, int B,
                       double *d_S);
/* QoI::evaluate: d_q[B]; d_Qint[B] (may be NULL) receives the integer topological
 * charge for the two susceptibility QoIs */
int mlmcpi_qoi(mlmcpi_ctx *ctx, const mlmcpi_model *m, int qoi, const double *d_x, int B,
               double *d_q, int64_t *d_Qint);
/* Start state for chains that are advanced by two-level Metropolis steps (MonteCarloTwoLevel,
 * MonteCarloMultiLevel, MultilevelSampler, HierarchicalSampler): the reference's zero state
 * (twolevelmetropolisstep.cc:11-22) followed by 50 local heat-bath sweeps where the action has a heat
 * bath.  With B chains side by side EVERY chain has to forget its start, and the exactly cold state is
 * metastable under the two-level step (profiles/r01_summary.md 10.2). */
int mlmcpi_thermal_state(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *d_x, int B, uint32_t chain0);
/* TwoLevelMetropolisStep::draw (montecarlo/twolevelmetropolisstep.cc:35-89) for B
 * chains.  d_xc: coarse states phi_c; d_xf: current fine states theta (updated where
 * accepted); d_Sf / d_Scond: cached S_f(theta), S_cond(theta) (updated where
 * accepted; fill them with mlmcpi_action / mlmcpi_cond_action as set_state does);
 * d_accept[B], d_deltas[B][3] = {dS_fine, dS_coarse, dS_trial} may be NULL. */
int mlmcpi_twolevel_step(mlmcpi_ctx *ctx, const mlmcpi_model *fine, const mlmcpi_model *coarse,
                         const double *d_xc, double *d_xf, double *d_Sf, double *d_Scond, int B,
                         uint32_t chain0, uint64_t draw, int32_t *d_accept, double *d_deltas);

/* ---- samplers (sampler/hmcsampler.hh, overrelaxedheatbathsampler.hh,
 *      hierarchicalsampler.hh) as batched objects ---------------------------- */
typedef struct mlmcpi_sampler mlmcpi_sampler;
typedef struct mlmcpi_sampler_params {
  int kind;              /* MLMCPI_SAMPLER_*: the (coarse-level) sampler            */
  int n_levels;          /* 1 = single-level sampler; >1 = HierarchicalSampler      */
  int renorm, ctype;     /* coarse_action() chain                                   */
  int nt;                /* hmc: nt, dt, n_rep  (sampler/hmcsampler.hh:21-65)       */
  double dt;
  int n_rep;
  int n_sweep_overrelax; /* heatbath (sampler/overrelaxedheatbathsampler.hh)        */
  int n_sweep_heatbath;
  int multilevel;        /* 0: HierarchicalSampler cascade; 1: MultilevelSampler level walk
                            (sampler/multilevelsampler.cc:71-112), needs n_levels > 1      */
  int qoi;               /* multilevel: QoI of the per-level statistics Q_sampler[l]        */
  int n_autocorr_window; /* multilevel: window of those statistics (default 20)             */
  int n_updates;         /* cluster: cluster updates per draw (clusteralgorithm: n_updates)  */
} mlmcpi_sampler_params;
int mlmcpi_sampler_create(mlmcpi_ctx *ctx, const mlmcpi_model *fine,
                          const mlmcpi_sampler_params *prm, int B, uint32_t chain0,
                          mlmcpi_sampler **s);
void mlmcpi_sampler_destroy(mlmcpi_sampler *s);
/* Sampler::set_state / Sampler::draw on device states [B][n] */
int mlmcpi_sampler_set_state(mlmcpi_sampler *s, const double *d_x);
int mlmcpi_sampler_draw(mlmcpi_sampler *s, double *d_x_out, int32_t *d_accept);
/* the current states [B][n] of the chains.  mlmcpi_sampler_draw only overwrites d_x_out where the draw
 * was accepted (MCMCStep::copy_if_rejected = false, hierarchicalsampler.cc:78-80), so an output buffer
 * that is to hold the chains' states from the first draw on is initialised with this call */
int mlmcpi_sampler_get_state(mlmcpi_sampler *s, double *d_x);
/* QoI::evaluate (qoi/quantityofinterest.hh) of the chains' current states, d_q[B].  Equal to mlmcpi_qoi of
 * mlmcpi_sampler_get_state; for MLMCPI_QOI_SCHWINGER_CHI (qoi/qft/qoi2dsusceptibility.cc:7-27) of a hierarchical
 * sampler the values are maintained by the draws -- the fill-in kernel of the finest level returns the topological
 * charge of the trial state next to S_f and S_cond -- and the call copies B doubles */
int mlmcpi_sampler_qoi(mlmcpi_sampler *s, int qoi, double *d_q);
/* the same with HOST buffers: h_x_in (may be NULL: keep the current state) is
 * uploaded, one draw is made, the QoI of the new state is evaluated, and
 * h_q[B] (and h_x_out[B][n] if not NULL) are copied back; synchronous */
int mlmcpi_sampler_draw_host(mlmcpi_sampler *s, const double *h_x_in, int qoi, double *h_q,
                             double *h_x_out);
/* the same with the chains RESIDENT on the device (as Sampler keeps its phi_state_cur) and the hand-over
 * to the host pipelined: one draw, QoI of the new states, snapshot of the states, and the device-to-host
 * copies of h_q[B] / h_x_out[B][n] (either may be NULL) run on a second stream while the next draw
 * computes.  As Sampler::draw(state) does, h_x_out is only overwritten for the chains whose draw was
 * accepted (hierarchicalsampler.cc:78-80) -- initialise it with mlmcpi_sampler_get_state; with pinned
 * (device-addressable) host memory only those rows cross the host link.  Returns without synchronising: the host buffers of this call are complete after the NEXT call
 * or after mlmcpi_sampler_wait_host -- alternate two pinned buffers.  This is the entry point bench.py's
 * `e2e` figure goes through. */
int mlmcpi_sampler_draw_host_async(mlmcpi_sampler *s, int qoi, double *h_q, double *h_x_out);
int mlmcpi_sampler_wait_host(mlmcpi_sampler *s);
int mlmcpi_sampler_level_model(const mlmcpi_sampler *s, int level, mlmcpi_model *m);
/* MCMCStep::p_accept of every level (montecarlo/mcmcstep.hh:21-72): accepted / attempted steps of
 * that level -- in the hierarchical cascade a level is only attempted by the chains all coarser
 * levels accepted (hierarchicalsampler.cc:73-74), which is how HierarchicalSampler::show_stats
 * reports it; mlmcpi_sampler_reset_stats = MCMCStep::reset_stats on every level */
int mlmcpi_sampler_stats(mlmcpi_sampler *s, double *h_p_accept /* [n_levels] */);
int mlmcpi_sampler_reset_stats(mlmcpi_sampler *s);
/* elementary-update counters of the last draw for throughput accounting:
 * out = {leapfrog site-steps, sweep site-updates, filled fine sites} summed over chains */
int mlmcpi_sampler_work(const mlmcpi_sampler *s, double out[3]);
/* HMCSampler::autotune_stepsize (sampler/hmcsampler.cc:72-113) on the coarsest level:
 * returns 0 if tuned, 1 if not converged (dt then reverts, as in the reference) */
int mlmcpi_sampler_autotune(mlmcpi_sampler *s, double p_accept_target, int n_rounds, int n_samples,
                            double *dt_out, double *p_accept_out);
int mlmcpi_sampler_set_dt(mlmcpi_sampler *s, double dt);

/* time n_meas batched draws with CUDA events: microseconds per chain-sample
 * (cost_per_sample of hierarchicalsampler.cc:45-52, multilevelsampler.cc:60-67) */
int mlmcpi_sampler_cost(mlmcpi_sampler *s, int n_meas, double *usec_per_sample);
/* multilevel sampler: out[l] = average number of draws between independent samples on
 * level l (t_indep), out[n_levels + l] = number of independent samples (n_indep) */
int mlmcpi_sampler_indep(const mlmcpi_sampler *s, double *out);

/* ---- MonteCarloMultiLevel (montecarlo/montecarlomultilevel.cc:7-204), batched ------
 * Level l < L-1 owns B chains of (coarse sampler on level l+1, two-level step, Y_l =
 * Q_l - Q_{l+1}); level L-1 owns B chains of the coarsest sampler (Y = Q).  The B chains
 * play the role the MPI ranks would have: Statistics are averaged over them. */
typedef struct mlmcpi_mlmc mlmcpi_mlmc;
typedef struct mlmcpi_mlmc_params {
  int n_level;           /* multilevelmc: n_level                                     */
  int n_burnin;          /* multilevelmc: n_burnin (batched draws per level)           */
  double epsilon;        /* multilevelmc: epsilon (tolerance on the root-mean-square error) */
  int n_autocorr_window; /* statistics: n_autocorr_window                             */
  int n_min_samples_qoi; /* statistics: n_min_samples_qoi                             */
  int qoi;               /* MLMCPI_QOI_*                                              */
  int max_iterations;    /* safety bound on the adaptive do-while loop (0: none)      */
  mlmcpi_sampler_params sampler; /* the sampler built on every coarse level; its n_levels
                            is the hierarchy depth n_max_level counted from the FINE action */
} mlmcpi_mlmc_params;
int mlmcpi_mlmc_create(mlmcpi_ctx *ctx, const mlmcpi_model *fine, const mlmcpi_mlmc_params *prm,
                       int B, uint32_t chain0, mlmcpi_mlmc **out);
void mlmcpi_mlmc_destroy(mlmcpi_mlmc *m);
/* MonteCarloMultiLevel::evaluate (synchronous: the sample allocation is decided on the host) */
int mlmcpi_mlmc_evaluate(mlmcpi_mlmc *m);
/* numerical_result() / statistical_error() and the per-level table of
 * show_detailed_statistics(): level_out[l][6] = {samples, mean of Y_l, variance, tau_int,
 * effective cost (usec), n_target} */
int mlmcpi_mlmc_result(mlmcpi_mlmc *m, double *value, double *error, double *level_out);

/* ---- statistics (common/statistics.cc:4-97), one accumulator per chain ---- */
typedef struct mlmcpi_stats mlmcpi_stats;
int mlmcpi_stats_create(mlmcpi_ctx *ctx, int k_max, int B, mlmcpi_stats **st);
void mlmcpi_stats_destroy(mlmcpi_stats *st);
/* Statistics::reset (short-term mean and sample count only) / Statistics::hard_reset */
int mlmcpi_stats_reset(mlmcpi_stats *st);
int mlmcpi_stats_hard_reset(mlmcpi_stats *st);
/* Statistics::record_sample for every chain from d_q[B] */
int mlmcpi_stats_record(mlmcpi_stats *st, const double *d_q);
/* packed moment vector summed over the chains of this batch:
 * {n_chains, long-term samples (all chains), short-term samples (all chains), sum avg,
 *  sum avg_longterm, sum avg2_longterm, sum avg3_longterm, sum avg4_longterm,
 *  sum S_0..S_{k_max-1}}
 * (what Statistics allreduces across MPI ranks, statistics.cc:30-35,64-79);
 * length 8 + k_max, every entry additive.  Sum it over GPUs (NCCL allreduce) and hand it
 * to _finalize. */
int mlmcpi_stats_pack(mlmcpi_stats *st, double *h_packed);
/* the same vector left in device memory (no synchronisation), ready for ncclAllReduce */
int mlmcpi_stats_pack_device(mlmcpi_stats *st, double *d_packed);
int mlmcpi_stats_packed_size(int k_max);
/* out = {average, variance, variance_error, tau_int, error, samples}: average over the
 * short-term window, variance / tau_int over the long-term one, as in the reference */
int mlmcpi_stats_finalize(const double *h_packed, int k_max, double out[6]);

#ifdef __cplusplus
}
#endif
#endif /* MLMCPI_H */
