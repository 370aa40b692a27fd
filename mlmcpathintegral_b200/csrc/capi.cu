// capi.cu -- the extern "C" boundary of libmlmcpi.so (include/mlmcpi.h): context,
// memory, host-side geometry and renormalisation, model dispatch, the two-level
// Metropolis-Hastings step, batched sampler objects and per-chain statistics.
//
// Reference citations relative to /root/reference/src.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <vector>

#include "common.cuh"

// ================================================================== context
int ctx_fail(mlmcpi_ctx *ctx, int code, const char *what, const char *detail) {
  if (ctx) {
    ctx->err = what ? what : "error";
    if (detail) {
      ctx->err += ": ";
      ctx->err += detail;
    }
  }
  return code;
}

int ctx_allreduce_host(mlmcpi_ctx *ctx, double *h, size_t n) {
  if (!ctx->allreduce || ctx->world <= 1 || n == 0)
    return 0;
  if (ctx->reduce_buf_n < n) {
    if (ctx->reduce_buf)
      cudaFree(ctx->reduce_buf);
    ctx->reduce_buf = nullptr;
    ctx->reduce_buf_n = 0;
    MLMCPI_CUDA(cudaMalloc((void **)&ctx->reduce_buf, sizeof(double) * std::max<size_t>(n, 64)));
    ctx->reduce_buf_n = std::max<size_t>(n, 64);
  }
  MLMCPI_CUDA(cudaMemcpyAsync(ctx->reduce_buf, h, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  // (pageable source: the copy is staged before the call returns)
  const int rc = ctx->allreduce(ctx->allreduce_user, ctx->reduce_buf, n);
  if (rc)
    return ctx_fail(ctx, rc < 0 ? rc : MLMCPI_ECUDA, "all-reduce hook failed");
  MLMCPI_CUDA(cudaMemcpyAsync(h, ctx->reduce_buf, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  MLMCPI_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int ctx_check_launch(mlmcpi_ctx *ctx, const char *what) {
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return ctx_fail(ctx, MLMCPI_ECUDA, what, cudaGetErrorString(e));
  return 0;
}

static double *grow(mlmcpi_ctx *ctx, double **buf, size_t *cap, size_t n) {
  if (n <= *cap && *buf)
    return *buf;
  if (*buf)
    cudaFree(*buf); // implicit device synchronisation: nobody still uses the old buffer
  *buf = nullptr;
  *cap = 0;
  const size_t want = n + n / 8 + 64;
  if (cudaMalloc((void **)buf, want * sizeof(double)) != cudaSuccess) {
    cudaGetLastError();
    ctx_fail(ctx, MLMCPI_ENOMEM, "cudaMalloc failed for a work buffer");
    *buf = nullptr;
    return nullptr;
  }
  *cap = want;
  return *buf;
}

double *ctx_scratch(mlmcpi_ctx *ctx, size_t n) { return grow(ctx, &ctx->scratch, &ctx->scratch_n, n); }
double *ctx_work(mlmcpi_ctx *ctx, int which, size_t n) {
  return grow(ctx, &ctx->work[which], &ctx->work_n[which], n);
}

void prof_begin(mlmcpi_ctx *ctx) {
  if (!ctx->profile)
    return;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, ctx->stream);
  ctx->prof_events.push_back(e0);
  ctx->prof_events.push_back(e1);
}
void prof_end(mlmcpi_ctx *ctx, uint64_t launches, double algorithmic_bytes) {
  if (!ctx->profile || ctx->prof_events.empty())
    return;
  cudaEventRecord(ctx->prof_events.back(), ctx->stream);
  ctx->prof_launches += launches;
  ctx->prof_bytes += algorithmic_bytes;
}

// Every entry point runs on the device of its context, whatever device is current in the calling thread,
// and leaves the caller's current device as it found it.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(const mlmcpi_ctx *ctx) {
    if (ctx && cudaGetDevice(&prev) == cudaSuccess && prev != ctx->device)
      cudaSetDevice(ctx->device);
    else
      prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0)
      cudaSetDevice(prev);
  }
};

extern "C" {

int mlmcpi_profile(mlmcpi_ctx *ctx, int enable) {
  DeviceGuard device_guard(ctx);
  ctx->profile = enable != 0;
  return 0;
}

int mlmcpi_profile_read(mlmcpi_ctx *ctx, double out[3]) {
  DeviceGuard device_guard(ctx);
  MLMCPI_CUDA(cudaStreamSynchronize(ctx->stream));
  for (size_t k = 0; k + 1 < ctx->prof_events.size(); k += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->prof_events[k], ctx->prof_events[k + 1]) == cudaSuccess)
      ctx->prof_ms += ms;
    cudaEventDestroy(ctx->prof_events[k]);
    cudaEventDestroy(ctx->prof_events[k + 1]);
  }
  cudaGetLastError();
  ctx->prof_events.clear();
  out[0] = ctx->prof_ms;
  out[1] = (double)ctx->prof_launches;
  out[2] = ctx->prof_bytes;
  ctx->prof_ms = ctx->prof_bytes = 0.0;
  ctx->prof_launches = 0;
  return 0;
}

int mlmcpi_version(void) { return MLMCPI_VERSION; }

int mlmcpi_create(mlmcpi_ctx **out, int device, uint64_t seed, void *stream) {
  if (!out)
    return MLMCPI_EINVAL;
  *out = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0 || device < 0 || device >= n_dev) {
    cudaGetLastError();
    return MLMCPI_ECUDA; // no CPU fallback
  }
  mlmcpi_ctx *ctx = new (std::nothrow) mlmcpi_ctx;
  if (!ctx)
    return MLMCPI_ENOMEM;
  ctx->device = device;
  ctx->seed = seed;
  DeviceGuard device_guard(ctx); // (the caller's current device is restored on return)
  if (cudaSetDevice(device) != cudaSuccess) {
    delete ctx;
    return MLMCPI_ECUDA;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess)
    ctx->n_sm = prop.multiProcessorCount;
  if (stream != MLMCPI_OWN_STREAM) {
    ctx->stream = (cudaStream_t)stream; // NULL = the legacy default stream
  } else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete ctx;
      return MLMCPI_ECUDA;
    }
    ctx->own_stream = true;
  }
  *out = ctx;
  return 0;
}

void mlmcpi_destroy(mlmcpi_ctx *ctx) {
  DeviceGuard device_guard(ctx);
  if (!ctx)
    return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->scratch)
    cudaFree(ctx->scratch);
  if (ctx->reduce_buf)
    cudaFree(ctx->reduce_buf);
  for (int k = 0; k < MLMCPI_N_WORK; ++k)
    if (ctx->work[k])
      cudaFree(ctx->work[k]);
  for (auto &kv : ctx->ho_exact_factor)
    cudaFree(kv.second);
  gff::release_dense(ctx);
  if (ctx->own_stream)
    cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char *mlmcpi_last_error(const mlmcpi_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }

int mlmcpi_device(const mlmcpi_ctx *ctx) { return ctx ? ctx->device : -1; }
void *mlmcpi_stream(const mlmcpi_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int mlmcpi_sync(mlmcpi_ctx *ctx) {
  DeviceGuard device_guard(ctx);
  MLMCPI_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int mlmcpi_set_option(mlmcpi_ctx *ctx, int option, int value) {
  DeviceGuard device_guard(ctx);
  if (option == MLMCPI_OPT_EXPCOS_ENVELOPE && value >= 0 && value <= 2) {
    ctx->expcos_envelope = value;
    return 0;
  }
  if (option == MLMCPI_OPT_LEAPFROG_VARIANT && value >= 0 && value <= 2) {
    ctx->leapfrog_variant = value;
    return 0;
  }
  if (option == MLMCPI_OPT_SWEEP_REVERSE && (value == 0 || value == 1)) {
    ctx->sweep_reverse = value;
    return 0;
  }
  if (option == MLMCPI_OPT_LEAPFROG_FUSE && value >= 0 && value <= 5) {
    ctx->leapfrog_fuse = value;
    return 0;
  }
  if (option == MLMCPI_OPT_OVERRELAX_ONE_PASS && (value == 0 || value == 1)) {
    ctx->overrelax_one_pass = value;
    return 0;
  }
  if (option == MLMCPI_OPT_FUSED_QM_HIERARCHY && (value == 0 || value == 1)) {
    ctx->fused_qm_hierarchy = value;
    return 0;
  }
  if (option == MLMCPI_OPT_LEAPFROG_ROWS && value >= 0) {
    ctx->leapfrog_rows = value;
    return 0;
  }
  if (option == MLMCPI_OPT_HOST_COPY_ENGINE && (value == 0 || value == 1)) {
    ctx->host_copy_engine = value;
    return 0;
  }
  if (option == MLMCPI_OPT_TAU_REFRESH && value >= 1 && value <= 1024) {
    ctx->tau_refresh = value;
    return 0;
  }
  if (option == MLMCPI_OPT_CASCADE_CACHE && (value == 0 || value == 1)) {
    ctx->cascade_cache = value;
    return 0;
  }
  if (option == MLMCPI_OPT_GFF_COARSE_SMOOTHING && (value == 0 || value == 1)) {
    ctx->gff_coarse_smoothing = value;
    return 0;
  }
  return ctx_fail(ctx, MLMCPI_EINVAL, "unknown option or value");
}
int mlmcpi_set_seed(mlmcpi_ctx *ctx, uint64_t seed) {
  DeviceGuard device_guard(ctx);
  ctx->seed = seed;
  return 0;
}
int mlmcpi_set_allreduce(mlmcpi_ctx *ctx, mlmcpi_allreduce_fn fn, void *user, int world_size, int rank) {
  DeviceGuard device_guard(ctx);
  if (!ctx || world_size < 1 || rank < 0 || rank >= world_size)
    return ctx ? ctx_fail(ctx, MLMCPI_EINVAL, "bad world size / rank") : MLMCPI_EINVAL;
  ctx->allreduce = fn;
  ctx->allreduce_user = user;
  ctx->world = fn ? world_size : 1;
  ctx->rank = fn ? rank : 0;
  return 0;
}
int mlmcpi_world_size(const mlmcpi_ctx *ctx) { return ctx ? ctx->world : 1; }
int mlmcpi_rank(const mlmcpi_ctx *ctx) { return ctx ? ctx->rank : 0; }
uint64_t mlmcpi_launch_count(const mlmcpi_ctx *ctx) { return ctx->launches; }

// =================================================================== memory
int mlmcpi_alloc(mlmcpi_ctx *ctx, size_t n, double **d_ptr) {
  DeviceGuard device_guard(ctx);
  if (!d_ptr)
    return ctx_fail(ctx, MLMCPI_EINVAL, "null output pointer");
  *d_ptr = nullptr;
  if (n == 0)
    return 0;
  if (cudaMalloc((void **)d_ptr, n * sizeof(double)) != cudaSuccess) {
    cudaGetLastError();
    return ctx_fail(ctx, MLMCPI_ENOMEM, "cudaMalloc failed");
  }
  MLMCPI_CUDA(cudaMemsetAsync(*d_ptr, 0, n * sizeof(double), ctx->stream)); // samplestate.hh:30-33
  return 0;
}
int mlmcpi_free(mlmcpi_ctx *ctx, double *d_ptr) {
  DeviceGuard device_guard(ctx);
  if (d_ptr)
    MLMCPI_CUDA(cudaFree(d_ptr));
  return 0;
}
int mlmcpi_upload(mlmcpi_ctx *ctx, double *d_dst, const double *h_src, size_t n) {
  DeviceGuard device_guard(ctx);
  MLMCPI_CUDA(cudaMemcpyAsync(d_dst, h_src, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
int mlmcpi_download(mlmcpi_ctx *ctx, double *h_dst, const double *d_src, size_t n) {
  DeviceGuard device_guard(ctx);
  MLMCPI_CUDA(cudaMemcpyAsync(h_dst, d_src, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  MLMCPI_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int mlmcpi_copy(mlmcpi_ctx *ctx, double *d_dst, const double *d_src, size_t n) {
  DeviceGuard device_guard(ctx);
  MLMCPI_CUDA(cudaMemcpyAsync(d_dst, d_src, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}

__global__ void axpy_kernel(double *out, const double *a, double alpha, const double *b, size_t n) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n)
    out[k] = a[k] + alpha * b[k];
}
int mlmcpi_axpy(mlmcpi_ctx *ctx, double *d_out, const double *d_a, double alpha, const double *d_b, size_t n) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !d_out || !d_a || !d_b)
    return MLMCPI_EINVAL;
  if (n == 0)
    return 0;
  axpy_kernel<<<cdiv((long long)n, 256), 256, 0, ctx->stream>>>(d_out, d_a, alpha, d_b, n);
  MLMCPI_LAUNCHED("axpy");
  return 0;
}

// ================================================================= geometry
int mlmcpi_sample_size(const mlmcpi_model *m) {
  switch (m->model) {
  case MLMCPI_HO:
  case MLMCPI_QUARTIC:
  case MLMCPI_ROTOR:
    return m->M_lat;
  case MLMCPI_SCHWINGER:
    return 2 * m->Mt_lat * m->Mx_lat;
  case MLMCPI_GFF:
    return m->rotated ? m->Mt_lat * m->Mx_lat / 2 : m->Mt_lat * m->Mx_lat;
  }
  return -1;
}

uint32_t mlmcpi_vertex_cart2lin(int Mt, int Mx, int rotated, int i, int j) {
  if (!rotated)
    return (uint32_t)(Mt * ((j + Mx) % Mx) + ((i + Mt) % Mt));
  const int odd = i & 1; // on a rotated lattice i and j have the same parity
  const int ih = (((i + Mt) - odd) / 2) % (Mt / 2);
  const int jh = (((j + Mx) - (j & 1)) / 2) % (Mx / 2);
  return (uint32_t)((Mt / 2) * jh + ih + odd * (Mt * Mx / 4));
}

void mlmcpi_vertex_lin2cart(int Mt, int Mx, int rotated, uint32_t ell, int *i, int *j) {
  if (!rotated) {
    *j = (int)(ell / Mt);
    *i = (int)(ell % Mt);
    return;
  }
  const uint32_t quarter = (uint32_t)(Mt * Mx / 4);
  const int odd = ell >= quarter ? 1 : 0;
  const uint32_t r = ell - odd * quarter;
  *j = 2 * (int)(r / (Mt / 2)) + odd;
  *i = 2 * (int)(r % (Mt / 2)) + odd;
}

uint32_t mlmcpi_link_cart2lin(int Mt, int Mx, int i, int j, int mu) {
  return (uint32_t)(2 * (Mt * ((j + Mx) % Mx) + ((i + Mt) % Mt)) + mu);
}

void mlmcpi_link_lin2cart(int Mt, int Mx, uint32_t ell, int *i, int *j, int *mu) {
  (void)Mx;
  const uint32_t site = ell >> 1;
  *mu = (int)(ell & 1u);
  *j = (int)(site / Mt);
  *i = (int)(site % Mt);
}

void mlmcpi_neighbours(int Mt, int Mx, int rotated, uint32_t ell, uint32_t nb[8]) {
  // nearest neighbours first, then the four next-nearest ones
  static const int d_plain[8][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
  static const int d_rot[8][2] = {{1, 1}, {1, -1}, {-1, 1}, {-1, -1}, {2, 0}, {-2, 0}, {0, 2}, {0, -2}};
  int i, j;
  mlmcpi_vertex_lin2cart(Mt, Mx, rotated, ell, &i, &j);
  const int(*d)[2] = rotated ? d_rot : d_plain;
  for (int k = 0; k < 8; ++k)
    nb[k] = mlmcpi_vertex_cart2lin(Mt, Mx, rotated, i + d[k][0], j + d[k][1]);
}

static bool level_factors(int Mt, int Mx, int ctype, int level, int *rt, int *rx) {
  const bool rotated = (ctype == MLMCPI_COARSEN_ROTATE) && (level & 1);
  *rt = *rx = 1;
  if (ctype == MLMCPI_COARSEN_BOTH || (ctype == MLMCPI_COARSEN_ROTATE && rotated))
    *rt = *rx = 2;
  else if (ctype == MLMCPI_COARSEN_TEMPORAL || (ctype == MLMCPI_COARSEN_ALTERNATE && !(level & 1)))
    *rt = 2;
  else if (ctype == MLMCPI_COARSEN_SPATIAL || (ctype == MLMCPI_COARSEN_ALTERNATE && (level & 1)))
    *rx = 2;
  else if (ctype != MLMCPI_COARSEN_ROTATE)
    return false;
  return (Mt % *rt == 0) && (Mx % *rx == 0);
}

int mlmcpi_coarse_shape(int Mt, int Mx, int ctype, int level, int *Mt_c, int *Mx_c, int *rot_c) {
  int rt, rx;
  const bool ok = level_factors(Mt, Mx, ctype, level, &rt, &rx);
  *Mt_c = (Mt % rt == 0) ? Mt / rt : Mt;
  *Mx_c = (Mx % rx == 0) ? Mx / rx : Mx;
  *rot_c = (ctype == MLMCPI_COARSEN_ROTATE) && !(level & 1);
  return ok && *Mt_c > 1 && *Mx_c > 1;
}

int mlmcpi_coarsening_lists(int Mt, int Mx, int ctype, int level, uint32_t *coarse,
                            uint32_t *fineonly, uint32_t *map_vals, int *counts) {
  int Mtc, Mxc, rotc, rt, rx;
  if (!mlmcpi_coarse_shape(Mt, Mx, ctype, level, &Mtc, &Mxc, &rotc))
    return MLMCPI_EINVAL;
  level_factors(Mt, Mx, ctype, level, &rt, &rx);
  const int rotated = (ctype == MLMCPI_COARSEN_ROTATE) && (level & 1);
  const int nv = rotated ? Mt * Mx / 2 : Mt * Mx;
  int nc = 0, nf = 0;
  // walking the linear index in ascending order yields both lists already sorted
  for (int ell = 0; ell < nv; ++ell) {
    int i, j;
    mlmcpi_vertex_lin2cart(Mt, Mx, rotated, ell, &i, &j);
    bool is_c;
    if (ctype == MLMCPI_COARSEN_ROTATE && !rotated)
      is_c = ((i + j) % 2 == 0);
    else
      is_c = (i % rt == 0) && (j % rx == 0);
    if (is_c) {
      map_vals[nc] = mlmcpi_vertex_cart2lin(Mtc, Mxc, rotc, i / rt, j / rx);
      coarse[nc++] = ell;
    } else {
      fineonly[nf++] = ell;
    }
  }
  counts[0] = nc;
  counts[1] = nf;
  return 0;
}

// common/auxilliary.cc:7-29
static double sigma_hat(double xi, unsigned p) {
  if (p == 0)
    return 1.0;
  if (p & 1u)
    return 0.0;
  double num = 0.0, den = 1.0;
  for (unsigned m = 1; m < 100; ++m) {
    const double e = std::exp(-0.5 * xi * m * m);
    num += 2. * std::pow((double)m, (double)p) * e;
    den += 2. * e;
  }
  return num / den;
}

int mlmcpi_coarse_model(const mlmcpi_model *fine, int renorm, int level, int ctype, double T_final,
                        mlmcpi_model *coarse) {
  if (!fine || !coarse)
    return MLMCPI_EINVAL;
  *coarse = *fine;
  const double a = fine->a_lat;
  switch (fine->model) {
  case MLMCPI_HO:
  case MLMCPI_QUARTIC:
  case MLMCPI_ROTOR:
    if (fine->M_lat % 2 || fine->M_lat < 2)
      return MLMCPI_EINVAL;
    coarse->M_lat = fine->M_lat / 2;
    coarse->a_lat = T_final / coarse->M_lat;
    coarse->T_final = T_final;
    if (fine->model == MLMCPI_HO) { // qm/harmonicoscillatorrenormalisation.hh:46-79
      if (renorm == MLMCPI_RENORM_PERTURBATIVE)
        coarse->m0 = fine->m0 * (1. - 0.5 * a * a * fine->mu2);
      else if (renorm == MLMCPI_RENORM_NONPERTURBATIVE)
        coarse->m0 = fine->m0 / (1. + 0.5 * a * a * fine->mu2);
      if (renorm != MLMCPI_RENORM_NONE)
        coarse->mu2 = fine->mu2 * (1. + 0.25 * a * a * fine->mu2);
    } else if (fine->model == MLMCPI_ROTOR) { // qm/rotorrenormalisation.hh:38-57, .cc:8-14
      if (renorm == MLMCPI_RENORM_PERTURBATIVE) {
        const double xi = T_final / fine->m0;
        const double s2 = sigma_hat(xi, 2), s4 = sigma_hat(xi, 4);
        const double deltaI = 0.5 * (1. - 2. * xi * s2 + 0.5 * xi * xi * (s4 - s2 * s2)) /
                              (1. - 2. * xi * s2 + xi * xi * (s4 - s2 * s2));
        coarse->m0 = (1. + deltaI * a / fine->m0) * fine->m0;
      } else if (renorm == MLMCPI_RENORM_NONPERTURBATIVE) {
        return MLMCPI_EUNSUPPORTED; // the reference errors out as well
      }
    } // quartic: qm/quarticoscillatoraction.hh:105-110, parameters unchanged
    return 0;
  case MLMCPI_SCHWINGER: { // qft/quenchedschwingerrenormalisation.hh:45-105
    int Mtc, Mxc, rotc;
    if (ctype == MLMCPI_COARSEN_ROTATE ||
        !mlmcpi_coarse_shape(fine->Mt_lat, fine->Mx_lat, ctype, level, &Mtc, &Mxc, &rotc))
      return MLMCPI_EINVAL;
    coarse->Mt_lat = Mtc;
    coarse->Mx_lat = Mxc;
    const bool both = (ctype == MLMCPI_COARSEN_BOTH);
    const double beta = fine->beta;
    coarse->beta = (both ? 0.25 : 0.5) * beta;
    if (beta > 4.0 && renorm == MLMCPI_RENORM_PERTURBATIVE)
      coarse->beta = (both ? 0.25 : 0.5) * (1. + (both ? 1.5 : 0.5) / beta) * beta;
    else if (beta > 4.0 && renorm == MLMCPI_RENORM_NONPERTURBATIVE) {
      // qft/quenchedschwingerrenormalisation.cc:7-64 (chi_t matching; host quadrature + bisection)
      if (beta > 2000.0)
        return MLMCPI_EUNSUPPORTED; // Phi_chit is unstable there (auxilliary.cc:45-51)
      coarse->beta = mlmcpi_schwinger_betacoarse_nonperturbative(
          beta, (unsigned)fine->Mt_lat * (unsigned)fine->Mx_lat, both ? 4 : 2);
    }
    if (ctype == MLMCPI_COARSEN_ALTERNATE)
      coarse->coarsening = ((level + 1) % 2 == 0) ? MLMCPI_COARSEN_TEMPORAL : MLMCPI_COARSEN_SPATIAL;
    return 0;
  }
  case MLMCPI_GFF: { // qft/gffaction.hh:174-181, 201-208
    int Mtc, Mxc, rotc;
    if (!mlmcpi_coarse_shape(fine->Mt_lat, fine->Mx_lat, ctype, level, &Mtc, &Mxc, &rotc))
      return MLMCPI_EINVAL;
    const double af = (fine->rotated ? std::sqrt(2.) : 1.) / fine->Mt_lat;
    const double ac = (rotc ? std::sqrt(2.) : 1.) / Mtc;
    coarse->Mt_lat = Mtc;
    coarse->Mx_lat = Mxc;
    coarse->rotated = rotc;
    coarse->gff_mu2 = ac * ac * (fine->gff_mu2 / (af * af));
    // GFFAction::coarse_action (gffaction.hh:201-208): n_gibbs_smooth = 2, omega = 1
    coarse->gff_n_gibbs = 2;
    coarse->gff_omega = 1.0;
    return 0;
  }
  }
  return MLMCPI_EINVAL;
}

} // extern "C"

// constants of BesselProductDistribution (distribution/besselproductdistribution.hh:52-80)
void besselproduct_setup(double beta, BesselProductConst *bp) {
  static std::mutex mtx;
  static std::map<double, BesselProductConst> cache;
  std::lock_guard<std::mutex> lock(mtx);
  auto it = cache.find(beta);
  if (it != cache.end()) {
    *bp = it->second;
    return;
  }
  BesselProductConst c;
  c.beta = beta;
  c.I0_twobeta = std::cyl_bessel_i(0.0, 2 * beta);
  c.log_I0_twobeta = std::log(c.I0_twobeta);
  c.sigma_beta = M_PI / std::sqrt(2 * c.log_I0_twobeta);
  const unsigned kmax = 16, nmax = 32;
  std::vector<double> lf(2 * nmax + 2, 0.0); // log n! as a running sum of logs
  for (unsigned n = 2; n < lf.size(); ++n)
    lf[n] = lf[n - 1] + std::log((double)n);
  auto log_nCk = [&](unsigned n, unsigned k) { return lf[n] - lf[k] - lf[n - k]; };
  double alpha0 = 1.0;
  for (unsigned k = 0; k <= kmax; ++k) {
    double s = 0.0;
    for (unsigned n = k; n <= nmax; ++n)
      for (unsigned m = k; m <= nmax; ++m)
        s += std::pow(0.5 * beta, 2.0 * (n + m)) *
             std::exp(log_nCk(2 * n, n - k) + log_nCk(2 * m, m - k) - 2 * (lf[n] + lf[m]));
    double alpha = ((k == 0) ? 2 : 4) * M_PI * s;
    if (k == 0)
      alpha0 = alpha;
    else
      alpha /= alpha0;
    c.alphaZ[k] = alpha;
  }
  cache[beta] = c;
  *bp = c;
}

// ===================================================== small generic kernels
// one warp per chain: lanes stride over the per-block partials, fixed-order shuffle tree
__global__ void reduce_finish_kernel(const double *partial, int nblk, int B, int nout, int epi,
                                     double scale0, double scale1, double *out, int64_t *Qint, const int32_t *mask) {
  const int chain = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (chain >= B || (mask && !mask[chain]))
    return; // (masked chains: their partial sums were not written)
  double s[3] = {0.0, 0.0, 0.0};
  for (int k = 0; k < nout; ++k) {
    const double *p = partial + ((size_t)k * B + chain) * nblk;
    double acc = 0.0;
    for (int b = lane; b < nblk; b += 32)
      acc += p[b];
    s[k] = warp_sum(acc);
  }
  if (lane != 0)
    return;
  if (epi == EPI_CHI) {
    out[chain] = scale0 * s[0] * s[0];
    if (Qint)
      Qint[chain] = -(int64_t)llrint(s[1]);
  } else {
    out[chain] = scale0 * s[0];
    if (nout > 1)
      out[B + chain] = scale1 * s[1];
    if (nout > 2)
      out[2 * (size_t)B + chain] = s[2];
  }
}

int launch_reduce_finish(mlmcpi_ctx *ctx, const double *partial, int nblk, int B, int nout, int epi,
                         double scale0, double scale1, double *out, int64_t *Qint, const int32_t *mask) {
  reduce_finish_kernel<<<cdiv(B, 4), 128, 0, ctx->stream>>>(partial, nblk, B, nout, epi, scale0,
                                                             scale1, out, Qint, mask);
  MLMCPI_LAUNCHED("reduce_finish");
  return 0;
}

namespace {

struct SqNormF {
  const double *p;
  size_t n;
  __device__ void operator()(int chain, long long s, double acc[1]) const {
    const double v = p[(size_t)chain * n + s];
    acc[0] += v * v;
  }
};

// sampler/hmcsampler.cc:48-67
__global__ void hmc_accept_kernel(int B, uint32_t chain0, uint64_t seed, uint64_t draw,
                                  const double *S_cur, const double *S_trial, const double *T_cur,
                                  const double *T_trial, int32_t *accept, double *diag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= B)
    return;
  const double deltaH = (S_trial[c] - S_cur[c]) + (T_trial[c] - T_cur[c]);
  bool acc = deltaH < 0.0;
  if (!acc) {
    Rng r = rng_init(seed, MLMCPI_STREAM_HMC_ACCEPT, draw, chain0 + c, 0);
    double u0, u1;
    rng_uniform2(r, u0, u1);
    acc = u0 < exp(-deltaH);
  }
  accept[c] = acc ? 1 : 0;
  if (diag) {
    double *d = diag + 5 * (size_t)c;
    d[0] = deltaH;
    d[1] = S_cur[c];
    d[2] = S_trial[c];
    d[3] = T_cur[c];
    d[4] = T_trial[c];
  }
}

// dst[chain] = src[chain] where accept[chain]: blocks of a rejected chain leave at once (the chain index is
// block-uniform: grid.y folds the chains, no per-thread division).  WRAP: angles are reduced to [-pi, pi)
// on the way (what Action::copy_from_fine's mod_2pi would make of them).
// dst2: a second destination (the caller's output state of a draw next to the sampler's own), or nullptr
template <bool WRAP>
__global__ void masked_copy_kernel(double2 *dst, const double2 *src, size_t n2, int B, const int32_t *accept,
                                   double2 *dst2 = nullptr) {
  for (int chain = blockIdx.y; chain < B; chain += gridDim.y) {
    if (!accept[chain])
      continue;
    const double2 *s = src + (size_t)chain * n2;
    double2 *d = dst + (size_t)chain * n2;
    double2 *d2 = dst2 ? dst2 + (size_t)chain * n2 : nullptr;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (size_t)gridDim.x * blockDim.x) {
      double2 v = s[t];
      if (WRAP) {
        v.x = mod_2pi(v.x);
        v.y = mod_2pi(v.y);
      }
      d[t] = v;
      if (d2)
        d2[t] = v;
    }
  }
}
__global__ void masked_copy1_kernel(double *dst, const double *src, size_t n, int B,
                                    const int32_t *accept) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * B)
    return;
  if (accept[t / n])
    dst[t] = src[t];
}

// montecarlo/twolevelmetropolisstep.cc:48-81.  red = {S_f', S_cond'}; ScC = S_c(theta_C),
// Scc = S_c(phi_c).  Sf_out[c] = S_f of the state after the step (may alias Sf_in; nullptr: not
// wanted), Scond is updated in place where accepted unless update_scond == 0.
__global__ void twolevel_accept_kernel(int B, uint32_t chain0, uint64_t seed, uint64_t draw,
                                       const double *red, const double *ScC_all, const double *Scc_all,
                                       const double *Sf_in, double *Sf_out, double *Scond,
                                       int update_scond, const int32_t *mask, int32_t *accept,
                                       double *deltas) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= B)
    return;
  const double Sf_prime = red[c], Scond_prime = red[B + c], ScC = ScC_all[c], Scc = Scc_all[c];
  const double Sf_cur = Sf_in[c];
  const double dS_fine = Sf_prime - Sf_cur;
  const double dS_coarse = ScC - Scc;
  const double dS_trial = Scond[c] - Scond_prime;
  const double dS = dS_fine + dS_coarse + dS_trial;
  bool acc = dS < 0.0;
  if (!acc) {
    Rng r = rng_init(seed, MLMCPI_STREAM_TWOLEVEL_ACCEPT, draw, chain0 + c, 0);
    double u0, u1;
    rng_uniform2(r, u0, u1);
    acc = u0 < exp(-dS);
  }
  if (mask && !mask[c])
    acc = false; // the cascade already stopped for this chain (hierarchicalsampler.cc:73-74)
  if (Sf_out)
    Sf_out[c] = acc ? Sf_prime : Sf_cur;
  if (acc && update_scond)
    Scond[c] = Scond_prime;
  accept[c] = acc ? 1 : 0;
  if (deltas) { // (no proposal was made for a masked chain: its sums are not defined)
    const bool live = !(mask && !mask[c]);
    deltas[3 * (size_t)c] = live ? dS_fine : 0.0;
    deltas[3 * (size_t)c + 1] = live ? dS_coarse : 0.0;
    deltas[3 * (size_t)c + 2] = live ? dS_trial : 0.0;
  }
}

__global__ void or_accept_kernel(int B, int32_t *acc, const int32_t *step) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < B)
    acc[c] = (acc[c] | step[c]) ? 1 : 0;
}
__global__ void set_i32_kernel(int B, int32_t *a, int32_t v) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < B)
    a[c] = v;
}
__global__ void count_accept_kernel(int B, const int32_t *acc, unsigned long long *counter) {
  int v = 0;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < B; c += gridDim.x * blockDim.x)
    v += acc[c];
  const double s = block_sum((double)v);
  if (threadIdx.x == 0 && s > 0)
    atomicAdd(counter, (unsigned long long)(s + 0.5));
}

} // namespace

int launch_half_sqnorm(mlmcpi_ctx *ctx, const double *d_p, size_t n, int B, double *d_T) {
  return site_reduce<1>(ctx, "half_sqnorm", SqNormF{d_p, n}, (long long)n, B, EPI_SCALE, 0.5, 0.0,
                        d_T, nullptr);
}

int launch_hmc_accept(mlmcpi_ctx *ctx, int B, uint32_t chain0, uint64_t draw, const double *S_cur,
                      const double *S_trial, const double *T_cur, const double *T_trial,
                      int32_t *accept, double *diag) {
  hmc_accept_kernel<<<cdiv(B, 128), 128, 0, ctx->stream>>>(B, chain0, ctx->seed, draw, S_cur, S_trial,
                                                          T_cur, T_trial, accept, diag);
  MLMCPI_LAUNCHED("hmc_accept");
  return 0;
}

int launch_masked_copy(mlmcpi_ctx *ctx, double *dst, const double *src, size_t n, int B,
                       const int32_t *accept, bool wrap_angles, double *dst2) {
  if (dst2 && (n % 2 != 0 || (uintptr_t)dst2 % 16 != 0 || (uintptr_t)dst % 16 != 0 || (uintptr_t)src % 16 != 0)) {
    int rc = launch_masked_copy(ctx, dst, src, n, B, accept, wrap_angles, nullptr);
    return rc ? rc : launch_masked_copy(ctx, dst2, src, n, B, accept, wrap_angles, nullptr);
  }
  if (n % 2 == 0 && ((uintptr_t)dst % 16 == 0) && ((uintptr_t)src % 16 == 0)) {
    const size_t n2 = n / 2;
    const dim3 grid((unsigned)std::min<size_t>(64, (n2 + 1023) / 1024), (unsigned)std::min(B, 65535));
    if (wrap_angles)
      masked_copy_kernel<true><<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<double2 *>(dst),
                                                             reinterpret_cast<const double2 *>(src), n2, B, accept,
                                                             reinterpret_cast<double2 *>(dst2));
    else
      masked_copy_kernel<false><<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<double2 *>(dst),
                                                              reinterpret_cast<const double2 *>(src), n2, B, accept,
                                                              reinterpret_cast<double2 *>(dst2));
  } else {
    if (wrap_angles)
      return ctx_fail(ctx, MLMCPI_EINVAL, "masked copy with angle reduction needs an even, aligned state");
    masked_copy1_kernel<<<cdiv((long long)n * B, 256), 256, 0, ctx->stream>>>(dst, src, n, B, accept);
  }
  MLMCPI_LAUNCHED("masked_copy");
  return 0;
}

// =========================================================== model dispatch
#define DISPATCH(m, fn, ...)                                                                       \
  do {                                                                                             \
    if (!ctx || !(m))                                                                              \
      return MLMCPI_EINVAL;                                                                        \
    if (B <= 0)                                                                                    \
      return ctx_fail(ctx, MLMCPI_EINVAL, "batch size must be positive");                          \
    switch ((m)->model) {                                                                          \
    case MLMCPI_HO:                                                                                \
    case MLMCPI_QUARTIC:                                                                           \
    case MLMCPI_ROTOR:                                                                             \
      if ((m)->M_lat < 2)                                                                          \
        return ctx_fail(ctx, MLMCPI_EINVAL, "M_lat must be at least 2");                           \
      return qm::fn(__VA_ARGS__);                                                                  \
    case MLMCPI_SCHWINGER:                                                                         \
      if ((m)->Mt_lat < 2 || (m)->Mx_lat < 2)                                                      \
        return ctx_fail(ctx, MLMCPI_EINVAL, "lattice extents must be at least 2");                 \
      return schwinger::fn(__VA_ARGS__);                                                           \
    case MLMCPI_GFF:                                                                               \
      if ((m)->Mt_lat < 2 || (m)->Mx_lat < 2)                                                      \
        return ctx_fail(ctx, MLMCPI_EINVAL, "lattice extents must be at least 2");                 \
      return gff::fn(__VA_ARGS__);                                                                 \
    }                                                                                              \
    return ctx_fail(ctx, MLMCPI_EINVAL, "unknown model");                                          \
  } while (0)

// d_ScC / d_Scc: S_c(theta_C) and S_c(phi_c) when the caller already knows them (nullptr: computed
// here); d_Sf_out: S_f of the state after the step (nullptr: d_Sf is updated in place);
// update_scond: keep d_Scond current
struct TwoLevelKnown {
  const double *d_ScC = nullptr, *d_Scc = nullptr;
  double *d_Sf_out = nullptr;
  bool sf_in_place = true, update_scond = true;
};
static int twolevel_step_impl(mlmcpi_ctx *ctx, const mlmcpi_model *fine, const mlmcpi_model *coarse,
                              const double *d_xc, double *d_xf, double *d_Sf, double *d_Scond, int B,
                              uint32_t chain0, uint64_t draw, const int32_t *d_mask,
                              int32_t *d_accept, double *d_deltas,
                              const TwoLevelKnown &known = TwoLevelKnown());

extern "C" {

int mlmcpi_init_state(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *d_x, int B, uint32_t chain0,
                      uint64_t draw) {
  DeviceGuard device_guard(ctx);
  DISPATCH(m, init_state, ctx, m, d_x, B, chain0, draw);
}
int mlmcpi_action(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *d_x, int B, double *d_S) {
  DeviceGuard device_guard(ctx);
  DISPATCH(m, action, ctx, m, d_x, B, d_S);
}
int mlmcpi_force(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *d_x, double *d_f, int B) {
  DeviceGuard device_guard(ctx);
  DISPATCH(m, force, ctx, m, d_x, d_f, B);
}
int mlmcpi_leapfrog(mlmcpi_ctx *ctx, const mlmcpi_model *m, int nt, double dt, double *d_x,
                    double *d_p, int B) {
  DeviceGuard device_guard(ctx);
  if (nt < 0)
    return ctx_fail(ctx, MLMCPI_EINVAL, "nt must be non-negative");
  DISPATCH(m, leapfrog, ctx, m, nt, dt, d_x, d_p, B);
}
int mlmcpi_hmc_momentum(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *d_p, int B, uint32_t chain0,
                        uint64_t draw) {
  DeviceGuard device_guard(ctx);
  DISPATCH(m, hmc_momentum, ctx, m, d_p, B, chain0, draw);
}
int mlmcpi_hmc_step(mlmcpi_ctx *ctx, const mlmcpi_model *m, int nt, double dt, double *d_x, int B,
                    uint32_t chain0, uint64_t draw, int32_t *d_accept, double *d_diag) {
  DeviceGuard device_guard(ctx);
  if (nt < 0)
    return ctx_fail(ctx, MLMCPI_EINVAL, "nt must be non-negative");
  DISPATCH(m, hmc_step, ctx, m, nt, dt, d_x, B, chain0, draw, d_accept, d_diag);
}
int mlmcpi_overrelax_sweep(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *d_x, int B) {
  DeviceGuard device_guard(ctx);
  DISPATCH(m, overrelax_sweep, ctx, m, d_x, B);
}
int mlmcpi_overrelax_sweeps(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *d_x, int B, int n_sweeps) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !m || B <= 0 || n_sweeps < 0)
    return MLMCPI_EINVAL;
  if (m->model == MLMCPI_SCHWINGER && m->Mt_lat >= 2 && m->Mx_lat >= 2 && m->Mt_lat % 2 == 0 &&
      m->Mx_lat % 2 == 0)
    return schwinger::overrelax_sweeps(ctx, m, d_x, B, n_sweeps);
  if (m->model == MLMCPI_GFF)
    return gff::overrelax_sweeps(ctx, m, d_x, B, n_sweeps);
  for (int k = 0; k < n_sweeps; ++k) {
    const int rc = mlmcpi_overrelax_sweep(ctx, m, d_x, B);
    if (rc)
      return rc;
  }
  return 0;
}
int mlmcpi_heatbath_sweep(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *d_x, int B,
                          uint32_t chain0, uint64_t draw) {
  DeviceGuard device_guard(ctx);
  DISPATCH(m, heatbath_sweep, ctx, m, d_x, B, chain0, draw);
}
int mlmcpi_dof_update(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *d_x, int B, int ell, int heatbath,
                      uint32_t chain0, uint64_t draw) {
  DeviceGuard device_guard(ctx);
  DISPATCH(m, dof_update, ctx, m, d_x, B, ell, heatbath, chain0, draw);
}
int mlmcpi_prolong(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *d_xc, double *d_x, int B) {
  DeviceGuard device_guard(ctx);
  DISPATCH(m, prolong, ctx, m, d_xc, d_x, B);
}
int mlmcpi_restrict(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *d_xf, double *d_xc, int B) {
  DeviceGuard device_guard(ctx);
  DISPATCH(m, restrict_, ctx, m, d_xf, d_xc, B);
}
int mlmcpi_fill(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *d_x, int B, uint32_t chain0,
                uint64_t draw) {
  DeviceGuard device_guard(ctx);
  DISPATCH(m, fill, ctx, m, d_x, B, chain0, draw);
}
int mlmcpi_prolong_fill(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *d_xc, double *d_x,
                        int B, uint32_t chain0, uint64_t draw) {
  DeviceGuard device_guard(ctx);
  DISPATCH(m, prolong_fill, ctx, m, d_xc, d_x, B, chain0, draw);
}
int mlmcpi_prolong_fill_eval(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *d_xc, double *d_x,
                             int B, uint32_t chain0, uint64_t draw, double *d_S) {
  DeviceGuard device_guard(ctx);
  DISPATCH(m, prolong_fill_eval, ctx, m, d_xc, d_x, B, chain0, draw, d_S);
}
int mlmcpi_cond_action(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *d_x, int B, double *d_S) {
  DeviceGuard device_guard(ctx);
  DISPATCH(m, cond_action, ctx, m, d_x, B, d_S);
}
int mlmcpi_qoi(mlmcpi_ctx *ctx, const mlmcpi_model *m, int qoi, const double *d_x, int B, double *d_q,
               int64_t *d_Qint) {
  DeviceGuard device_guard(ctx);
  DISPATCH(m, qoi, ctx, m, qoi, d_x, B, d_q, d_Qint);
}

int mlmcpi_cluster_update(mlmcpi_ctx *ctx, const mlmcpi_model *rotor, double *d_x, int B, uint32_t chain0,
                          uint64_t update0, int n_updates) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !rotor || B <= 0 || n_updates < 0)
    return MLMCPI_EINVAL;
  return qm::cluster_update(ctx, rotor, d_x, B, chain0, update0, n_updates);
}
int mlmcpi_exact_draw(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *d_x, int B, uint32_t chain0, uint64_t draw) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !m || !d_x || B <= 0)
    return MLMCPI_EINVAL;
  if (m->model == MLMCPI_GFF)
    return gff::exact_draw(ctx, m, d_x, B, chain0, draw);
  if (m->model != MLMCPI_HO)
    return ctx_fail(ctx, MLMCPI_EUNSUPPORTED,
                    "the exact sampler is defined for the harmonic oscillator and the GFF");
  return qm::exact_draw(ctx, m, d_x, B, chain0, draw);
}
int mlmcpi_schwinger_from_cluster(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *d_psi, double *d_x,
                                  int B, uint32_t chain0, uint64_t draw) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !m || B <= 0 || m->model != MLMCPI_SCHWINGER)
    return MLMCPI_EINVAL;
  return schwinger::from_cluster(ctx, m, d_psi, d_x, B, chain0, draw);
}

static int thermal_start(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, uint32_t chain0);
int mlmcpi_thermal_state(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *d_x, int B, uint32_t chain0) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !m || !d_x || B <= 0)
    return MLMCPI_EINVAL;
  return thermal_start(ctx, m, d_x, B, chain0);
}
int mlmcpi_twolevel_step(mlmcpi_ctx *ctx, const mlmcpi_model *fine, const mlmcpi_model *coarse,
                         const double *d_xc, double *d_xf, double *d_Sf, double *d_Scond, int B,
                         uint32_t chain0, uint64_t draw, int32_t *d_accept, double *d_deltas) {
  DeviceGuard device_guard(ctx);
  return twolevel_step_impl(ctx, fine, coarse, d_xc, d_xf, d_Sf, d_Scond, B, chain0, draw, nullptr,
                            d_accept, d_deltas);
}

} // extern "C"

// TwoLevelMetropolisStep::draw, montecarlo/twolevelmetropolisstep.cc:35-89
static int twolevel_step_impl(mlmcpi_ctx *ctx, const mlmcpi_model *fine, const mlmcpi_model *coarse,
                              const double *d_xc, double *d_xf, double *d_Sf, double *d_Scond, int B,
                              uint32_t chain0, uint64_t draw, const int32_t *d_mask,
                              int32_t *d_accept, double *d_deltas, const TwoLevelKnown &known) {
  if (!ctx || !fine || !coarse || B <= 0)
    return MLMCPI_EINVAL;
  const size_t nf = (size_t)mlmcpi_sample_size(fine), nc = (size_t)mlmcpi_sample_size(coarse);
  double *theta_prime = ctx_work(ctx, 4, nf * B);
  double *red = ctx_work(ctx, 6, (size_t)5 * B);
  if (!theta_prime || !red)
    return MLMCPI_ENOMEM;
  int32_t *acc = d_accept ? d_accept : reinterpret_cast<int32_t *>(red + 4 * B);
  int rc;
  // :40-42, :48 and :65-66 in one pass: theta' = fill(prolong(phi_c)), S_f(theta'), S_cond(theta')
  if (fine->model == MLMCPI_SCHWINGER && d_mask) // no proposal for the chains whose cascade has stopped
    rc = schwinger::prolong_fill_eval_masked(ctx, fine, d_xc, theta_prime, B, chain0, draw, red, d_mask);
  else
    rc = mlmcpi_prolong_fill_eval(ctx, fine, d_xc, theta_prime, B, chain0, draw, red);
  if (rc)
    return rc;
  const double *ScC = known.d_ScC, *Scc = known.d_Scc;
  if (!ScC) {
    double *theta_C = ctx_work(ctx, 5, nc * B);
    if (!theta_C)
      return MLMCPI_ENOMEM;
    if ((rc = mlmcpi_restrict(ctx, fine, d_xf, theta_C, B)))                          // :55
      return rc;
    if ((rc = mlmcpi_action(ctx, coarse, theta_C, B, red + 2 * B)))                   // :57
      return rc;
    ScC = red + 2 * B;
  }
  if (!Scc) {
    if ((rc = mlmcpi_action(ctx, coarse, d_xc, B, red + 3 * B)))                      // :58
      return rc;
    Scc = red + 3 * B;
  }
  double *Sf_out = known.d_Sf_out ? known.d_Sf_out : (known.sf_in_place ? d_Sf : nullptr);
  twolevel_accept_kernel<<<cdiv(B, 128), 128, 0, ctx->stream>>>(B, chain0, ctx->seed, draw, red, ScC, Scc, d_Sf,
                                                               Sf_out, d_Scond, known.update_scond ? 1 : 0,
                                                               d_mask, acc, d_deltas);
  MLMCPI_LAUNCHED("twolevel_accept");
  return launch_masked_copy(ctx, d_xf, theta_prime, nf, B, acc);                      // :78-88
}

// ================================================================== samplers
static mlmcpi_ctx *mlmcpi_stats_ctx(mlmcpi_stats *st);
// cached answer of a tau_int query (stats_query_cached below)
struct TauCache {
  double st[6] = {0, 0, 0, 1, 0, 0};
  int age = 0, interval = 1;
  bool valid = false;
};
struct mlmcpi_sampler {
  mlmcpi_ctx *ctx = nullptr;
  mlmcpi_sampler_params prm;
  int B = 0;
  uint32_t chain0 = 0;
  uint64_t draw = 0;
  int L = 1;
  std::vector<mlmcpi_model> model;
  std::vector<double *> state; // [L] device [B][n_l]
  double *Sf = nullptr, *Scond = nullptr, *q = nullptr; // [B]
  // cached S_f / S_cond of the finest state.  state[0] only changes where the two-level step
  // accepts, and the accept kernel updates the caches there, so the values
  // TwoLevelMetropolisStep::set_state would recompute (twolevelmetropolisstep.cc:92-97) are
  // already known -- bit for bit, the reductions being deterministic.
  double *Sf0 = nullptr, *Scond0 = nullptr;
  bool cache0_valid = false;
  // hierarchical cascade: S_l of the level states right after the restriction chain (= S_c(theta_C)
  // of the step one level finer, and S_f(theta) of the step on this level) and after the level's
  // own update (= S_c(phi_c) of the step one level finer); [L][B] each
  double *S_old = nullptr, *S_new = nullptr;
  cudaStream_t copy_stream = nullptr; // H2D stream of mlmcpi_sampler_draw_host, D2H stream of ..._async
  double *stage_x = nullptr, *stage_q = nullptr; // snapshots the asynchronous host copies read
  int32_t *stage_acc = nullptr;
  bool host_copy_pending = false;
  cudaEvent_t ev[9] = {};
  // copy-engine hand-over of mlmcpi_sampler_draw_host_async (MLMCPI_OPT_HOST_COPY_ENGINE): two snapshot buffers, the
  // accept flags of a step in pinned host memory, one cudaMemcpyAsync per run of accepted chains
  double *ce_stage[2] = {nullptr, nullptr};
  int32_t *ce_flags[2] = {nullptr, nullptr};
  cudaEvent_t ce_ev_flags[2] = {}, ce_ev_copied[2] = {};
  bool ce_used[2] = {false, false}, ce_pending = false;
  int ce_cur = 0, ce_pending_stage = 0;
  double *ce_target = nullptr;
  int32_t *acc = nullptr, *acc_step = nullptr;          // [B]
  unsigned long long *counters = nullptr;               // [L] accepted chains per level
  uint64_t n_draws = 0;
  double work[3] = {0, 0, 0};
  // MultilevelSampler (sampler/multilevelsampler.hh): per-level statistics of the sampler
  // QoI, persistent two-level caches, and the independence bookkeeping of the level walk
  std::vector<mlmcpi_stats *> stats_sampler;
  std::vector<double *> SfL, ScondL; // [L-1][B]
  std::vector<double> t_indep;
  std::vector<int> n_indep, t_sampler;
  std::vector<uint64_t> n_steps; // steps taken on every level (acceptance rates of the level walk)
  std::vector<TauCache> tau_cache; // [L] cached tau_int of stats_sampler (level walk decisions)
  // cascade_draw_cached: the coarse levels are only ever changed by a fully accepted cascade, so
  // state[l] == restrict^l(state[0]) holds between draws and S_l(state[l]) / S_cond(state[l]) are cached
  bool cascade_valid = false;
  std::vector<double *> trial;   // [L] theta'_l of the current draw, levels 1 .. L-2 (sampler-owned)
  double *Scond_lvl = nullptr;   // [L][B] S_cond(state[l]), 1 <= l <= L-2
  double *Sprime = nullptr;      // [L][2][B] S_f(theta'_l), S_cond(theta'_l) of the current draw
  // Schwinger model, coarsening both: the fill-in of the finest level also returns the topological charge of theta',
  // so the susceptibility QoI of the chains' states is kept up to date without a pass of its own (mlmcpi_sampler_qoi)
  double *S3 = nullptr;          // [3][B] S_f, S_cond, sum_P mod_2pi(P) of theta'_0
  double *chi_cur = nullptr;     // [B] QOI_SCHWINGER_CHI of state[0]
  bool chi_valid = false;
  // QuenchedSchwingerClusterSampler: the rotor chain psi [B][Mt*Mx] and its action
  double *psi = nullptr;
  mlmcpi_model psi_model = {};
  uint64_t cluster_updates = 0;
};

// tau_int / variance / ... of a device statistics object (host synchronisation)
static int stats_query(mlmcpi_stats *st, int k_max, double out[6]) {
  std::vector<double> packed(mlmcpi_stats_packed_size(k_max));
  int rc = mlmcpi_stats_pack(st, packed.data());
  if (rc)
    return rc;
  // over the chains of ALL processes (statistics.cc:30-35,64-79: mpi_allreduce_avg of the moments)
  if ((rc = ctx_allreduce_host(mlmcpi_stats_ctx(st), packed.data(), packed.size())))
    return rc;
  return mlmcpi_stats_finalize(packed.data(), k_max, out);
}

// The host-side decisions of MultilevelSampler::draw and MonteCarloMultiLevel::draw_coarse_sample need
// tau_int of a device Statistics object: pack kernel + device-to-host copy + stream synchronisation (+ an
// all-reduce over the processes).  The reference asks for it at every sample
// (multilevelsampler.cc:92, montecarlomultilevel.cc:176); tau_int of a running series changes slowly, so
// the answer is reused for a number of calls that doubles from 1 up to MLMCPI_OPT_TAU_REFRESH (default 16;
// 1 = ask every time, the reference's schedule).  Every process follows the same schedule (lockstep).
static int stats_query_cached(mlmcpi_stats *st, int k_max, TauCache &c, double out[6]) {
  mlmcpi_ctx *ctx = mlmcpi_stats_ctx(st);
  if (!c.valid || c.age >= c.interval) {
    const int rc = stats_query(st, k_max, c.st);
    if (rc)
      return rc;
    c.valid = true;
    c.age = 0;
    c.interval = std::min(std::max(1, ctx->tau_refresh), 2 * c.interval);
  }
  c.age++;
  for (int k = 0; k < 6; ++k)
    out[k] = c.st[k];
  return 0;
}

// Start state of a chain that is advanced by two-level Metropolis steps.  The reference starts such
// chains from the zero state (twolevelmetropolisstep.cc:11-22, hierarchicalsampler.cc:43-44) and relies
// on ONE long chain to forget it.  With B chains side by side every chain has to be thermalised, and
// the exactly cold state is metastable under the two-level step (S_f(theta') - S_f(0) grows with the
// volume: 16^2, beta = 4: acceptance 3 % for the first 100 draws, profiles/r01_summary.md 10.2), as is the
// hot state for beta > 8 (10.1).  So: zero state, then MLMCPI_THERMAL_SWEEPS local heat-bath sweeps where
// the action has a heat bath (rotor, GFF, Schwinger); any start state is legitimate, burn-in follows.
#define MLMCPI_THERMAL_SWEEPS 50
static int thermal_start(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, uint32_t chain0) {
  const size_t n = (size_t)mlmcpi_sample_size(m) * B;
  MLMCPI_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * n, ctx->stream));
  const bool sweeps = m->model == MLMCPI_ROTOR || m->model == MLMCPI_GFF ||
                      (m->model == MLMCPI_SCHWINGER && m->Mt_lat % 2 == 0 && m->Mx_lat % 2 == 0);
  if (!sweeps)
    return 0;
  if (m->model == MLMCPI_GFF) {
    const int rc = mlmcpi_init_state(ctx, m, x, B, chain0, 0);
    if (rc)
      return rc;
  }
  for (int k = 0; k < MLMCPI_THERMAL_SWEEPS; ++k) {
    const int rc = mlmcpi_heatbath_sweep(ctx, m, x, B, chain0, (0xFFFFull << 40) + (uint64_t)k);
    if (rc == MLMCPI_EUNSUPPORTED || rc == MLMCPI_EINVAL) // no coloured sweep for this shape: zero state
      return k == 0 ? 0 : rc;
    if (rc)
      return rc;
  }
  return 0;
}

static uint64_t level_draw(uint64_t draw, int level, int rep) {
  return (draw << 12) | ((uint64_t)level << 8) | (uint64_t)(rep & 0xff);
}

static double n_sites(const mlmcpi_model &m) {
  if (m.model == MLMCPI_SCHWINGER)
    return (double)m.Mt_lat * m.Mx_lat;
  return (double)mlmcpi_sample_size(&m);
}

// draws that amount to MLMCPI_CLUSTER_BURNIN_UPDATES single-cluster updates (every update changes the
// topological charge by +-1 with probability ~1/2, so a few thousand updates equilibrate chi_t ~ 10)
#define MLMCPI_CLUSTER_BURNIN_UPDATES 4000
static int cluster_burnin_draws(const mlmcpi_sampler *s) {
  const int n_updates = std::max(1, s->prm.n_updates);
  return (MLMCPI_CLUSTER_BURNIN_UPDATES + n_updates - 1) / n_updates;
}

// the sampler on the coarsest level: HMCSampler::draw (sampler/hmcsampler.cc:8-19) or
// OverrelaxedHeatBathSampler::draw (sampler/overrelaxedheatbathsampler.cc:8-31)
static int coarse_draw(mlmcpi_sampler *s, int c0, int B) {
  mlmcpi_ctx *ctx = s->ctx;
  const int l = s->L - 1;
  const mlmcpi_model *m = &s->model[l];
  double *x = s->state[l] + (size_t)c0 * mlmcpi_sample_size(m);
  int32_t *acc = s->acc + c0, *acc_step = s->acc_step + c0;
  const uint32_t chain0 = s->chain0 + (uint32_t)c0;
  int rc;
  if (s->prm.kind == MLMCPI_SAMPLER_HMC) {
    const int n_rep = std::max(1, s->prm.n_rep);
    for (int r = 0; r < n_rep; ++r) {
      int32_t *a = (r == 0) ? acc : acc_step;
      if ((rc = mlmcpi_hmc_step(ctx, m, s->prm.nt, s->prm.dt, x, B, chain0, level_draw(s->draw, l, r), a,
                                nullptr)))
        return rc;
      if (r > 0) {
        or_accept_kernel<<<cdiv(B, 128), 128, 0, ctx->stream>>>(B, acc, acc_step);
        MLMCPI_LAUNCHED("or_accept");
      }
    }
    s->work[0] += (double)B * n_rep * (s->prm.nt + 1) * n_sites(*m);
  } else if (s->prm.kind == MLMCPI_SAMPLER_HEATBATH) {
    // OverrelaxedHeatBathSampler::draw, overrelaxedheatbathsampler.cc:8-31.  As the coarse sampler
    // of a hierarchy the draw must be a REVERSIBLE kernel (delayed acceptance): every colour update
    // is reversible, so the whole sequence is run forwards or exactly backwards with probability
    // 1/2 (one Philox bit per draw).  A stand-alone sampler keeps the reference's order.
    bool backwards = false;
    if (s->L > 1) {
      // splitmix64 of (seed, draw): the schedule must not depend on the state
      uint64_t h = ctx->seed ^ (0x9E3779B97F4A7C15ull * (s->draw + 1));
      h ^= h >> 30;
      h *= 0xBF58476D1CE4E5B9ull;
      h ^= h >> 27;
      h *= 0x94D049BB133111EBull;
      h ^= h >> 31;
      backwards = (h & 1u) != 0;
    }
    const int saved = ctx->sweep_reverse;
    ctx->sweep_reverse = backwards ? 1 : saved;
    rc = 0;
    const bool gff_sequence = m->model == MLMCPI_GFF && !ctx->sweep_reverse && m->Mt_lat % 2 == 0 && m->Mx_lat % 2 == 0;
    if (gff_sequence) { // all sweeps of the draw ping-pong through the one-pass kernel (one copy back at most)
      std::vector<uint64_t> hb_draws(std::max(0, s->prm.n_sweep_heatbath));
      for (int k = 0; k < (int)hb_draws.size(); ++k)
        hb_draws[k] = level_draw(s->draw, l, k);
      rc = gff::sweep_sequence(ctx, m, x, B, std::max(0, s->prm.n_sweep_overrelax), (int)hb_draws.size(), chain0,
                               hb_draws.data());
    }
    for (int pass = 0; pass < 2 && !rc && !gff_sequence; ++pass) {
      const bool do_hb = backwards ? (pass == 0) : (pass == 1);
      if (do_hb) {
        for (int k = 0; k < s->prm.n_sweep_heatbath && !rc; ++k) {
          const int kk = backwards ? s->prm.n_sweep_heatbath - 1 - k : k;
          rc = mlmcpi_heatbath_sweep(ctx, m, x, B, chain0, level_draw(s->draw, l, kk));
        }
      } else {
        rc = mlmcpi_overrelax_sweeps(ctx, m, x, B, s->prm.n_sweep_overrelax);
      }
    }
    ctx->sweep_reverse = saved;
    if (rc)
      return rc;
    set_i32_kernel<<<cdiv(B, 128), 128, 0, ctx->stream>>>(B, acc, 1);
    MLMCPI_LAUNCHED("set_accept");
    s->work[1] += (double)B * (s->prm.n_sweep_overrelax + s->prm.n_sweep_heatbath) * n_sites(*m);
  } else if (s->prm.kind == MLMCPI_SAMPLER_CLUSTER) {
    // ClusterSampler::draw (clustersampler.cc:38-50): n_updates single-cluster updates
    const int n_updates = std::max(1, s->prm.n_updates);
    if (m->model == MLMCPI_ROTOR) {
      if ((rc = qm::cluster_update(ctx, m, x, B, chain0, s->cluster_updates, n_updates)))
        return rc;
    } else if (m->model == MLMCPI_SCHWINGER) {
      // quenchedschwingerclustersampler.cc:40-86
      double *psi = s->psi + (size_t)c0 * s->psi_model.M_lat;
      if ((rc = qm::cluster_update(ctx, &s->psi_model, psi, B, chain0, s->cluster_updates, n_updates)))
        return rc;
      if ((rc = schwinger::from_cluster(ctx, m, psi, x, B, chain0, level_draw(s->draw, l, 0))))
        return rc;
    } else {
      return ctx_fail(ctx, MLMCPI_EUNSUPPORTED, "no cluster algorithm for this action");
    }
    set_i32_kernel<<<cdiv(B, 128), 128, 0, ctx->stream>>>(B, acc, 1);
    MLMCPI_LAUNCHED("set_accept");
    s->work[1] += (double)B * n_updates; // cluster updates
  } else if (s->prm.kind == MLMCPI_SAMPLER_EXACT) {
    // HarmonicOscillatorAction::draw as a sampler (qm/harmonicoscillatoraction.hh: the action IS a
    // Sampler): independent exact draws, always accepted
    if ((rc = mlmcpi_exact_draw(ctx, m, x, B, chain0, level_draw(s->draw, l, 0))))
      return rc;
    set_i32_kernel<<<cdiv(B, 128), 128, 0, ctx->stream>>>(B, acc, 1);
    MLMCPI_LAUNCHED("set_accept");
    s->work[1] += (double)B * n_sites(*m);
  } else {
    return ctx_fail(ctx, MLMCPI_EINVAL, "unknown sampler kind");
  }
  return 0;
}

// Start state of a hierarchical sampler whose coarse sampler proposes INDEPENDENT states (cluster,
// exact): the two-level steps are then independence samplers with weight w = pi_f / (pi_c q), and a
// chain started from a state of atypically large w never leaves it (measured, profiles/r02_summary.md:
// 512^2, beta = 1024, 3 levels, zero state + 50 heat-bath sweeps: no chain accepts in 130 draws, while
// the very same 256^2 <- 128^2 step accepts 36 % as the top of a 256^2 hierarchy).  The start state
// is therefore drawn the way the proposals are -- a coarse sample, prolonged and filled in level by
// level, i.e. a sample of pi_c q q ... -- whose weight is typical by construction.  Any start state is
// legitimate (burn-in follows); the stationary distribution is untouched.
static int cascade_start(mlmcpi_sampler *s) {
  mlmcpi_ctx *ctx = s->ctx;
  const int L = s->L, B = s->B;
  int rc;
  const uint64_t start_draw = 0xFFFEull << 40;
  const uint64_t saved = s->draw;
  s->draw = start_draw;
  if (s->prm.kind == MLMCPI_SAMPLER_CLUSTER && s->model[L - 1].model != MLMCPI_SCHWINGER) {
    // decorrelate the rotor chain from its hot start: clustersampler.cc:27-31 (burn-in draws);
    // the chain behind the Schwinger cluster sampler was burnt in when it was set up
    const int n_burn = cluster_burnin_draws(s);
    for (int k = 0; k < n_burn; ++k) {
      if ((rc = coarse_draw(s, 0, B)))
        return rc;
      s->cluster_updates += std::max(1, s->prm.n_updates);
    }
  } else if (s->prm.kind == MLMCPI_SAMPLER_HMC || s->prm.kind == MLMCPI_SAMPLER_HEATBATH) {
    // local sampler on the coarsest level: hot start, burnt in there (100 coarse draws cost less than
    // one fine-level sweep)
    if ((rc = mlmcpi_init_state(ctx, &s->model[L - 1], s->state[L - 1], B, s->chain0, 0)))
      return rc;
    for (int k = 0; k < 100; ++k) {
      if ((rc = coarse_draw(s, 0, B)))
        return rc;
      s->draw++;
    }
  } else if ((rc = coarse_draw(s, 0, B))) {
    return rc;
  }
  s->draw = saved;
  for (int l = L - 2; l >= 0; --l)
    if ((rc = mlmcpi_prolong_fill(ctx, &s->model[l], s->state[l + 1], s->state[l], B, s->chain0,
                                  level_draw(start_draw, l, 0))))
      return rc;
  s->work[0] = s->work[1] = s->work[2] = 0.0;
  return 0;
}

// Start state of a GFF hierarchy.  With three or more levels the intermediate two-level steps accept a
// few per cent of the proposals (in the reference as here: Q_hat of a level is paired with the 5-point
// fill-in), the chains are sticky, and an ENSEMBLE of chains started away from the stationary
// distribution approaches it over tens of thousands of draws (measured, profiles/r02_summary.md: 16^2,
// 3 levels, cascade start: <phi^2> 0.3230 -> 0.3354 over 30 000 draws, exact 0.3380; started from an exact
// sample the same kernel keeps 0.3381 +- 0.0004 from the first block on).  One long chain, as the
// reference runs, pays that once; B short ones cannot.  So the chains start from samples of the
// fine-level action itself: the Cholesky sampler where the level is small enough for its dense factor,
// otherwise a hot state thermalised by overrelaxed heat-bath sweeps on the fine level (the massive GFF
// has no slow modes beyond the correlation length 1 / (a m)).
static int gff_equilibrium_start(mlmcpi_sampler *s) {
  mlmcpi_ctx *ctx = s->ctx;
  const mlmcpi_model *m = &s->model[0];
  const int B = s->B;
  const int N = mlmcpi_sample_size(m);
  if (N <= 4096)
    return mlmcpi_exact_draw(ctx, m, s->state[0], B, s->chain0, 0xFFFDull << 40);
  int rc = mlmcpi_init_state(ctx, m, s->state[0], B, s->chain0, 0);
  for (int k = 0; k < 100 && !rc; ++k) {
    rc = mlmcpi_overrelax_sweeps(ctx, m, s->state[0], B, 5);
    if (!rc)
      rc = mlmcpi_heatbath_sweep(ctx, m, s->state[0], B, s->chain0, (0xFFFDull << 40) + (uint64_t)k);
  }
  return rc;
}

// HierarchicalSampler::draw, sampler/hierarchicalsampler.cc:55-81, for the chains
// [c0, c0 + B) of the batch (n_levels == 1: the plain single-level sampler).  Chains are
// independent, so a draw of the whole batch may be issued range by range -- which is what
// lets mlmcpi_sampler_draw_host overlap the upload of one range with the kernels of another.
static int sampler_draw_range(mlmcpi_sampler *s, int c0, int B, bool cache0_valid) {
  mlmcpi_ctx *ctx = s->ctx;
  const int L = s->L;
  const uint32_t chain0 = s->chain0 + (uint32_t)c0;
  int32_t *acc = s->acc + c0;
  auto st = [&](int l) { return s->state[l] + (size_t)c0 * mlmcpi_sample_size(&s->model[l]); };
  int rc;
  // 1-D paths with an HMC coarse sampler: the whole cascade in one kernel, one warp per chain
  if (L > 1 && ctx->fused_qm_hierarchy && s->model[0].model <= MLMCPI_ROTOR && s->prm.kind == MLMCPI_SAMPLER_HMC &&
      std::max(1, s->prm.n_rep) == 1) {
    double *states[16];
    for (int l = 0; l < L && l < 16; ++l)
      states[l] = st(l);
    rc = qm::hierarchical_draw(ctx, s->model.data(), L, s->prm.nt, s->prm.dt, states, B, chain0, s->draw,
                               s->Sf0 + c0, s->Scond0 + c0, cache0_valid, acc, s->counters);
    if (rc <= 0) {
      if (rc == 0) {
        s->work[0] += (double)B * (s->prm.nt + 1) * n_sites(s->model[L - 1]);
        for (int l = L - 2; l >= 0; --l)
          s->work[2] += (double)B * n_sites(s->model[l]);
      }
      return rc;
    }
  }
  // S_l of every coarse level state right after the restriction chain / after its own update.
  // theta_C = restrict(theta_fine) of TwoLevelMetropolisStep::draw line 55 IS the level state
  // produced by the restriction chain, so its action (line 57) is S_old, and S_c(phi_c) (line 58)
  // is S_new of the coarser level: one restriction and two reductions per step are not repeated.
  auto S_old = [&](int l) { return s->S_old + (size_t)l * s->B + c0; };
  auto S_new = [&](int l) { return s->S_new + (size_t)l * s->B + c0; };
  for (int l = 1; l < L; ++l) { // :57-60
    if ((rc = mlmcpi_restrict(ctx, &s->model[l - 1], st(l - 1), st(l), B)))
      return rc;
    if ((rc = mlmcpi_action(ctx, &s->model[l], st(l), B, S_old(l))))
      return rc;
  }
  if ((rc = coarse_draw(s, c0, B))) // :62-66
    return rc;
  count_accept_kernel<<<std::min(cdiv(B, 256), 64), 256, 0, ctx->stream>>>(B, acc, s->counters + (L - 1));
  MLMCPI_LAUNCHED("count_accept");
  if (L > 1 && (rc = mlmcpi_action(ctx, &s->model[L - 1], st(L - 1), B, S_new(L - 1))))
    return rc;
  for (int l = L - 2; l >= 0; --l) {
    // TwoLevelMetropolisStep::set_state, montecarlo/twolevelmetropolisstep.cc:92-97
    double *Sf = (l == 0) ? s->Sf0 + c0 : S_old(l), *Scond = ((l == 0) ? s->Scond0 : s->Scond) + c0;
    if (l > 0) {
      if ((rc = mlmcpi_cond_action(ctx, &s->model[l], st(l), B, Scond)))
        return rc;
    } else if (!cache0_valid) {
      if ((rc = mlmcpi_action(ctx, &s->model[l], st(l), B, Sf)))
        return rc;
      if ((rc = mlmcpi_cond_action(ctx, &s->model[l], st(l), B, Scond)))
        return rc;
    }
    TwoLevelKnown known;
    known.d_ScC = S_old(l + 1);
    known.d_Scc = S_new(l + 1);
    known.d_Sf_out = (l == 0) ? Sf : S_new(l);
    known.update_scond = (l == 0);
    // acc is both the incoming cascade mask and the outgoing accept flag
    if ((rc = twolevel_step_impl(ctx, &s->model[l], &s->model[l + 1], st(l + 1), st(l), Sf, Scond, B, chain0,
                                 level_draw(s->draw, l, 0), acc, acc, nullptr, known)))
      return rc;
    count_accept_kernel<<<std::min(cdiv(B, 256), 64), 256, 0, ctx->stream>>>(B, acc, s->counters + l);
    MLMCPI_LAUNCHED("count_accept");
    s->work[2] += (double)B * n_sites(s->model[l]);
  }
  return 0;
}

// QOI_SCHWINGER_CHI of the accepted trial states from the charge sum the fill-in returned (schwinger::qoi: 0.25 / pi^2 s^2)
__global__ void chi_commit_kernel(int B, const int32_t *acc, double *chi, const double *qsum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < B && acc[c])
    chi[c] = (0.25 / (M_PI * M_PI)) * qsum[c] * qsum[c];
}

// commit of a fully accepted cascade on a coarse level: the cached actions follow the state
__global__ void cascade_commit_kernel(int B, const int32_t *acc, double *S_old, const double *S_prime, double *Scond,
                                      const double *Scond_prime) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= B || !acc[c])
    return;
  S_old[c] = S_prime[c];
  if (Scond)
    Scond[c] = Scond_prime[c];
}

// HierarchicalSampler::draw (hierarchicalsampler.cc:55-81) for the quenched Schwinger model with the HMC
// coarse sampler, WITHOUT the per-draw restriction chain and its reductions.
//
// In the reference every draw starts with x[l] = restrict(x[l-1]) for all levels, because a cascade that
// stopped half-way has left accepted states on the coarse levels that do not belong to the fine state.
// Here a draw never touches the coarse level states until the whole cascade has accepted: the HMC
// trajectory ends in a work buffer, the two-level steps propose into trial buffers, each level passes
// its TRIAL state up as the coarse proposal of the next finer step, and only the chains that accept
// on level 0 copy their trial states into state[l] (l >= 1) at the end.  Since restrict(fill(prolong(
// phi))) == phi bit for bit (tests: restrict o fill o prolong = id), state[l] then IS the restriction of
// the new fine state -- what the reference's restriction chain would produce at the start of the next
// draw -- and for all other chains nothing changed.  With the states the cached actions stay valid:
// S_l(state[l]) (= S_c(theta_C) of step l-1 and S_f(theta) of step l) and S_cond(state[l]).  Per draw this
// removes two restrictions, four full-lattice reductions, the conditioned-action pass of level 1 and the
// level >= 1 commits of rejected cascades (2.0 of 9.4 ms at 512^2 x 512 chains); the draws are the same,
// bit for bit (tests/test_gpu_parity.py::test_hierarchical_draw_equals_explicit_cascade).
static bool cascade_hmc_trial(const mlmcpi_sampler *s);
// d_x_out: the caller's output state of the draw (Sampler::draw(state)), written for the accepted chains by the same
// pass that commits the finest level, or nullptr
static int cascade_draw_cached(mlmcpi_sampler *s, double *d_x_out) {
  mlmcpi_ctx *ctx = s->ctx;
  const int L = s->L, B = s->B;
  const uint32_t chain0 = s->chain0;
  int rc;
  auto S_old = [&](int l) { return s->S_old + (size_t)l * B; };
  auto Scond_l = [&](int l) { return s->Scond_lvl + (size_t)l * B; };
  auto Sp = [&](int l, int k) { return s->Sprime + ((size_t)l * 2 + k) * B; };
  if (!s->cascade_valid) {
    for (int l = 1; l < L; ++l) {
      if ((rc = mlmcpi_restrict(ctx, &s->model[l - 1], s->state[l - 1], s->state[l], B)))
        return rc;
      if ((rc = mlmcpi_action(ctx, &s->model[l], s->state[l], B, S_old(l))))
        return rc;
      if (l <= L - 2 && (rc = mlmcpi_cond_action(ctx, &s->model[l], s->state[l], B, Scond_l(l))))
        return rc;
    }
    s->cascade_valid = true;
  }
  if (!s->cache0_valid) {
    if ((rc = mlmcpi_action(ctx, &s->model[0], s->state[0], B, s->Sf0)))
      return rc;
    if ((rc = mlmcpi_cond_action(ctx, &s->model[0], s->state[0], B, s->Scond0)))
      return rc;
  }
  const bool fused_chi = s->S3 != nullptr;
  if (fused_chi && !s->chi_valid) {
    if ((rc = mlmcpi_qoi(ctx, &s->model[0], MLMCPI_QOI_SCHWINGER_CHI, s->state[0], B, s->chi_cur, nullptr)))
      return rc;
    s->chi_valid = true;
  }
  const double *coarse_trial = nullptr;
  const bool hmc_trial = cascade_hmc_trial(s);
  if (hmc_trial) { // coarsest level: tentative HMC step (hmcsampler.cc:21-69)
    if ((rc = schwinger::hmc_trial(ctx, &s->model[L - 1], s->prm.nt, s->prm.dt, s->state[L - 1], B, chain0,
                                   level_draw(s->draw, L - 1, 0), S_old(L - 1), Sp(L - 1, 0), s->acc, &coarse_trial)))
      return rc;
    s->work[0] += (double)B * (s->prm.nt + 1) * n_sites(s->model[L - 1]);
  } else { // any other coarse sampler advances a COPY of the coarsest state (the sampler works on s->state[L-1])
    double *keep = s->state[L - 1], *trialC = s->trial[L - 1];
    if ((rc = mlmcpi_copy(ctx, trialC, keep, (size_t)mlmcpi_sample_size(&s->model[L - 1]) * B)))
      return rc;
    s->state[L - 1] = trialC;
    rc = coarse_draw(s, 0, B);
    s->state[L - 1] = keep;
    if (rc)
      return rc;
    if ((rc = mlmcpi_action(ctx, &s->model[L - 1], trialC, B, Sp(L - 1, 0))))
      return rc;
    coarse_trial = trialC;
  }
  const double *const hmc_fin = coarse_trial; // (the trajectory's end state in an HMC work buffer / the advanced copy)
  count_accept_kernel<<<std::min(cdiv(B, 256), 64), 256, 0, ctx->stream>>>(B, s->acc, s->counters + (L - 1));
  MLMCPI_LAUNCHED("count_accept");
  for (int l = L - 2; l >= 0; --l) {
    const size_t nf = (size_t)mlmcpi_sample_size(&s->model[l]);
    double *theta_prime = (l == 0) ? ctx_work(ctx, 4, nf * B) : s->trial[l];
    if (!theta_prime)
      return MLMCPI_ENOMEM;
    // theta' = fill(prolong(trial state of level l+1)), S_f(theta'), S_cond(theta')  (twolevelmetropolisstep.cc:40-66)
    double *Sprime_l = (l == 0 && fused_chi) ? s->S3 : Sp(l, 0);
    // (s->acc: the chains whose cascade is still alive; no proposal is made for the others)
    if (l == 0 && fused_chi)
      rc = schwinger::prolong_fill_eval_charge(ctx, &s->model[0], coarse_trial, theta_prime, B, chain0,
                                               level_draw(s->draw, 0, 0), Sprime_l, s->acc);
    else if (s->model[l].model == MLMCPI_SCHWINGER)
      rc = schwinger::prolong_fill_eval_masked(ctx, &s->model[l], coarse_trial, theta_prime, B, chain0,
                                               level_draw(s->draw, l, 0), Sprime_l, s->acc);
    else
      rc = mlmcpi_prolong_fill_eval(ctx, &s->model[l], coarse_trial, theta_prime, B, chain0, level_draw(s->draw, l, 0),
                                    Sprime_l);
    if (rc)
      return rc;
    twolevel_accept_kernel<<<cdiv(B, 128), 128, 0, ctx->stream>>>(
        B, chain0, ctx->seed, level_draw(s->draw, l, 0), Sprime_l, S_old(l + 1), Sp(l + 1, 0),
        (l == 0) ? s->Sf0 : S_old(l), (l == 0) ? s->Sf0 : nullptr, (l == 0) ? s->Scond0 : Scond_l(l),
        (l == 0) ? 1 : 0, s->acc, s->acc, nullptr);
    MLMCPI_LAUNCHED("twolevel_accept");
    count_accept_kernel<<<std::min(cdiv(B, 256), 64), 256, 0, ctx->stream>>>(B, s->acc, s->counters + l);
    MLMCPI_LAUNCHED("count_accept");
    s->work[2] += (double)B * n_sites(s->model[l]);
    if (l == 0) {
      if ((rc = launch_masked_copy(ctx, s->state[0], theta_prime, nf, B, s->acc, false, d_x_out))) // :78-88
        return rc;
      if (fused_chi) {
        chi_commit_kernel<<<cdiv(B, 128), 128, 0, ctx->stream>>>(B, s->acc, s->chi_cur, s->S3 + 2 * (size_t)B);
        MLMCPI_LAUNCHED("chi_commit");
      }
    }
    coarse_trial = theta_prime;
  }
  // the chains whose cascade accepted on every level take their trial states on the coarse levels
  for (int l = 1; l < L; ++l) {
    // (the HMC trajectory does not reduce its angles; the restriction of the accepted fine state, which
    // this copy stands in for, does: quenchedschwingeraction.cc:152-195)
    const double *src = (l == L - 1) ? hmc_fin : s->trial[l];
    if (src != s->state[l] && (rc = launch_masked_copy(ctx, s->state[l], src, (size_t)mlmcpi_sample_size(&s->model[l]),
                                                       B, s->acc, hmc_trial && l == L - 1)))
      return rc;
    cascade_commit_kernel<<<cdiv(B, 128), 128, 0, ctx->stream>>>(B, s->acc, S_old(l), Sp(l, 0),
                                                                (l <= L - 2) ? Scond_l(l) : nullptr, Sp(l, 1));
    MLMCPI_LAUNCHED("cascade_commit");
  }
  return 0;
}

// MultilevelSampler::draw, sampler/multilevelsampler.cc:71-112.  All chains walk the levels
// in lockstep: the decision "independent sample reached on this level" uses tau_int of the
// per-level statistics averaged over the chains, exactly what the reference's Statistics
// does across MPI ranks.
static int multilevel_draw(mlmcpi_sampler *s) {
  mlmcpi_ctx *ctx = s->ctx;
  const int L = s->L, B = s->B;
  const int k_max = s->prm.n_autocorr_window;
  int rc;
  int level = L - 1;
  do {
    if (level == L - 1) {
      if ((rc = coarse_draw(s, 0, B))) // :76-78
        return rc;
      count_accept_kernel<<<std::min(cdiv(B, 256), 64), 256, 0, ctx->stream>>>(B, s->acc, s->counters + level);
      MLMCPI_LAUNCHED("count_accept");
      s->cluster_updates += std::max(1, s->prm.n_updates);
    } else { // :85-86 (the two-level step keeps theta_fine and its cached actions)
      if ((rc = twolevel_step_impl(ctx, &s->model[level], &s->model[level + 1], s->state[level + 1],
                                   s->state[level], s->SfL[level], s->ScondL[level], B, s->chain0,
                                   level_draw(s->draw, level, 0), nullptr, s->acc, nullptr)))
        return rc;
      count_accept_kernel<<<std::min(cdiv(B, 256), 64), 256, 0, ctx->stream>>>(B, s->acc,
                                                                             s->counters + level);
      MLMCPI_LAUNCHED("count_accept");
      s->work[2] += (double)B * n_sites(s->model[level]);
    }
    s->draw++;
    s->n_steps[level]++;
    if ((rc = mlmcpi_qoi(ctx, &s->model[level], s->prm.qoi, s->state[level], B, s->q, nullptr))) // :89
      return rc;
    if ((rc = mlmcpi_stats_record(s->stats_sampler[level], s->q)))
      return rc;
    s->t_sampler[level]++;
    double st[6];
    if (s->tau_cache.size() != (size_t)L)
      s->tau_cache.assign(L, TauCache());
    if ((rc = stats_query_cached(s->stats_sampler[level], k_max, s->tau_cache[level], st)))
      return rc;
    if (s->t_sampler[level] >= std::ceil(st[3])) { // :92-106
      s->t_indep[level] =
          (s->n_indep[level] * s->t_indep[level] + s->t_sampler[level]) / (1.0 + s->n_indep[level]);
      s->n_indep[level]++;
      s->t_sampler[level] = 0;
      level--;
    } else {
      level = L - 1;
    }
  } while (level >= 0);
  return 0;
}

// cascade_draw_cached serves the hierarchical samplers of the two field theories with one coarse draw per cascade:
// the quenched Schwinger model with one HMC trajectory on the coarsest level (BASELINE config C4 / C5: the trajectory
// is run tentatively, schwinger::hmc_trial) and, through a copy of the coarsest state that the coarse sampler then
// advances, every other coarse sampler of the Schwinger model and of the GFF (C3: one dense Q_hat evaluation per level
// and draw instead of three).  The 1-D models have their own one-kernel cascade (fused_qm_hierarchy).
static bool cascade_hmc_trial(const mlmcpi_sampler *s) {
  return s->model[0].model == MLMCPI_SCHWINGER && s->prm.kind == MLMCPI_SAMPLER_HMC && s->prm.nt >= 1;
}
static bool cascade_cache_applies(const mlmcpi_sampler *s) {
  const int model = s->model[0].model;
  return s->ctx->cascade_cache && s->L >= 2 && !s->prm.multilevel && std::max(1, s->prm.n_rep) == 1 &&
         (cascade_hmc_trial(s) || ((model == MLMCPI_SCHWINGER || model == MLMCPI_GFF) &&
                                   (s->prm.kind != MLMCPI_SAMPLER_HMC || s->prm.nt >= 1)));
}

extern "C" {

int mlmcpi_sampler_create(mlmcpi_ctx *ctx, const mlmcpi_model *fine, const mlmcpi_sampler_params *prm,
                          int B, uint32_t chain0, mlmcpi_sampler **out) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !fine || !prm || !out || B <= 0)
    return MLMCPI_EINVAL;
  *out = nullptr;
  if (prm->n_levels < 1 || prm->n_levels > 16)
    return ctx_fail(ctx, MLMCPI_EINVAL, "n_levels out of range");
  mlmcpi_sampler *s = new (std::nothrow) mlmcpi_sampler;
  if (!s)
    return MLMCPI_ENOMEM;
  s->ctx = ctx;
  s->prm = *prm;
  s->B = B;
  s->chain0 = chain0;
  s->L = prm->n_levels;
  s->model.push_back(*fine);
  // the GFF fill-in is an independent conditional draw only under CoarsenRotate (qft/gffconditionedfineaction.hh:
  // 20-37: all four neighbours of a fine-only vertex are coarse); any other coarsening would read unfilled neighbours
  if (s->L > 1 && fine->model == MLMCPI_GFF &&
      (prm->ctype != MLMCPI_COARSEN_ROTATE || fine->coarsening != MLMCPI_COARSEN_ROTATE)) {
    delete s;
    return ctx_fail(ctx, MLMCPI_EUNSUPPORTED,
                    "a GFF hierarchy needs coarsening = rotate (in the sampler parameters and in the fine-level model)");
  }
  // HierarchicalSampler constructor, sampler/hierarchicalsampler.cc:19-29
  for (int l = 0; l + 1 < s->L; ++l) {
    mlmcpi_model c;
    // a 2-D rotated lattice sits on an odd level; everything else starts at level 0
    const int level = (fine->model == MLMCPI_GFF && fine->rotated ? 1 : 0) + l;
    const int rc = mlmcpi_coarse_model(&s->model[l], prm->renorm, level, prm->ctype,
                                       s->model[l].T_final, &c);
    if (rc) {
      delete s;
      return ctx_fail(ctx, rc, "cannot construct the coarse action of a level");
    }
    // GFFAction::coarse_action (gffaction.hh:201-208) gives every coarse level the Gibbs-smoothed
    // action Q_hat, whatever sampler runs on it -- kept (MLMCPI_OPT_GFF_COARSE_SMOOTHING = 0 selects
    // the plain 5-point action on the coarse levels instead)
    if (c.model == MLMCPI_GFF && !ctx->gff_coarse_smoothing)
      c.gff_n_gibbs = 0;
    s->model.push_back(c);
  }
  bool ok = true;
  for (int l = 0; l < s->L && ok; ++l) {
    double *d = nullptr;
    ok = mlmcpi_alloc(ctx, (size_t)mlmcpi_sample_size(&s->model[l]) * B, &d) == 0;
    s->state.push_back(d);
  }
  ok = ok && mlmcpi_alloc(ctx, B, &s->Sf0) == 0 && mlmcpi_alloc(ctx, B, &s->Scond0) == 0;
  ok = ok && mlmcpi_alloc(ctx, (size_t)s->L * B, &s->S_old) == 0 &&
       mlmcpi_alloc(ctx, (size_t)s->L * B, &s->S_new) == 0;
  ok = ok && mlmcpi_alloc(ctx, B, &s->Sf) == 0 && mlmcpi_alloc(ctx, B, &s->Scond) == 0 &&
       mlmcpi_alloc(ctx, B, &s->q) == 0;
  ok = ok && cudaMalloc((void **)&s->acc, sizeof(int32_t) * B) == cudaSuccess &&
       cudaMalloc((void **)&s->acc_step, sizeof(int32_t) * B) == cudaSuccess &&
       cudaMalloc((void **)&s->counters, sizeof(unsigned long long) * s->L) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    mlmcpi_sampler_destroy(s);
    return ctx_fail(ctx, MLMCPI_ENOMEM, "out of device memory for the sampler states");
  }
  cudaMemsetAsync(s->counters, 0, sizeof(unsigned long long) * s->L, ctx->stream);
  if (cascade_cache_applies(s)) {
    s->trial.assign(s->L, nullptr);
    for (int l = 1; l + (cascade_hmc_trial(s) ? 2 : 1) <= s->L && ok; ++l)
      ok = mlmcpi_alloc(ctx, (size_t)mlmcpi_sample_size(&s->model[l]) * B, &s->trial[l]) == 0;
    ok = ok && mlmcpi_alloc(ctx, (size_t)s->L * B, &s->Scond_lvl) == 0 &&
         mlmcpi_alloc(ctx, (size_t)s->L * 2 * B, &s->Sprime) == 0;
    if (ok && fine->model == MLMCPI_SCHWINGER && fine->coarsening == MLMCPI_COARSEN_BOTH &&
        fine->Mt_lat % 2 == 0 && fine->Mx_lat % 2 == 0)
      ok = mlmcpi_alloc(ctx, (size_t)3 * B, &s->S3) == 0 && mlmcpi_alloc(ctx, (size_t)B, &s->chi_cur) == 0;
    if (!ok) {
      cudaGetLastError();
      mlmcpi_sampler_destroy(s);
      return ctx_fail(ctx, MLMCPI_ENOMEM, "out of device memory for the cascade's trial states");
    }
  }
  int rc;
  if (s->prm.kind == MLMCPI_SAMPLER_CLUSTER && s->model[s->L - 1].model == MLMCPI_SCHWINGER) {
    // quenchedschwingerclustersampler.cc:16-23: rotor chain over the Mt*Mx cells, T = 1,
    // m0 = beta * a  (so that m0/a = beta)
    const mlmcpi_model &mc = s->model[s->L - 1];
    mlmcpi_model r = {};
    r.model = MLMCPI_ROTOR;
    r.M_lat = mc.Mt_lat * mc.Mx_lat;
    r.T_final = 1.0;
    r.a_lat = 1.0 / r.M_lat;
    r.m0 = mc.beta * r.a_lat;
    s->psi_model = r;
    // Start of the rotor chain: the reference starts hot (U(-pi,pi) per site) and then runs 10^4
    // draws in the constructors (clustersampler.cc:19-31).  A hot chain of M sites carries O(sqrt(M))
    // windings which single-cluster updates remove only slowly (128^2 cells, beta = 64: chi_t still ten
    // times too large after 2000 updates), so the chain starts cold, is thermalised locally by heat-bath
    // sweeps and the cluster burn-in below only has to spread the topological charge.
    if (mlmcpi_alloc(ctx, (size_t)r.M_lat * B, &s->psi) ||
        thermal_start(ctx, &r, s->psi, B, chain0 + 0x00800000u)) {
      mlmcpi_sampler_destroy(s);
      return ctx_fail(ctx, MLMCPI_ENOMEM, "cannot set up the cluster chain");
    }
    if (qm::cluster_update(ctx, &r, s->psi, B, chain0, s->cluster_updates, MLMCPI_CLUSTER_BURNIN_UPDATES)) {
      mlmcpi_sampler_destroy(s);
      return ctx_fail(ctx, MLMCPI_ECUDA, "cluster burn-in failed");
    }
    s->cluster_updates += MLMCPI_CLUSTER_BURNIN_UPDATES;
  }
  if (s->prm.multilevel) {
    if (s->L < 2) {
      mlmcpi_sampler_destroy(s);
      return ctx_fail(ctx, MLMCPI_EINVAL, "the multilevel sampler needs at least two levels");
    }
    if (s->prm.n_autocorr_window < 1)
      s->prm.n_autocorr_window = 20;
    // multilevelsampler.cc:8-58: the coarsest sampler starts from Action::initialise_state,
    // every TwoLevelMetropolisStep from the zero state (twolevelmetropolisstep.cc:11-22)
    rc = mlmcpi_init_state(ctx, &s->model[s->L - 1], s->state[s->L - 1], B, chain0, 0);
    s->t_indep.assign(s->L, 0.0);
    s->n_indep.assign(s->L, 0);
    s->t_sampler.assign(s->L, 0);
    s->n_steps.assign(s->L, 0);
    for (int l = 0; l < s->L && !rc; ++l) {
      mlmcpi_stats *st = nullptr;
      rc = mlmcpi_stats_create(ctx, s->prm.n_autocorr_window, B, &st);
      s->stats_sampler.push_back(st);
    }
    for (int l = 0; l + 1 < s->L && !rc; ++l) {
      double *a = nullptr, *b = nullptr;
      rc = mlmcpi_alloc(ctx, B, &a) || mlmcpi_alloc(ctx, B, &b);
      s->SfL.push_back(a);
      s->ScondL.push_back(b);
      if (!rc)
        rc = thermal_start(ctx, &s->model[l], s->state[l], B, chain0);
      if (!rc)
        rc = mlmcpi_action(ctx, &s->model[l], s->state[l], B, a);
      if (!rc)
        rc = mlmcpi_cond_action(ctx, &s->model[l], s->state[l], B, b);
    }
  } else {
    // Sampler constructors start from Action::initialise_state (e.g. hmcsampler.hh:99-101); the
    // hierarchical sampler from the zero state (hierarchicalsampler.cc:43-44), here thermalised
    // GFF with the reference's Q_hat on the coarse levels and a heat-bath coarse sampler: the coarse chain
    // samples the 5-point action, not Q_hat, the hierarchy's stationary distribution is NOT the fine-level
    // action (tests/golden/stats.json: the reference's own estimate is 10 % low), and an exact fine-level
    // sample is a metastable start for it (32^2, 4 levels: no level accepts); its chains start the way
    // its proposals are made (cascade start).  Every consistent GFF hierarchy starts in equilibrium.
    const bool gff_inconsistent =
        fine->model == MLMCPI_GFF && s->prm.kind == MLMCPI_SAMPLER_HEATBATH && ctx->gff_coarse_smoothing;
    if (s->L > 1 && fine->model == MLMCPI_GFF && !gff_inconsistent)
      rc = gff_equilibrium_start(s);
    else if (s->L > 1 && (s->prm.kind == MLMCPI_SAMPLER_CLUSTER || s->prm.kind == MLMCPI_SAMPLER_EXACT ||
                          gff_inconsistent))
      rc = cascade_start(s);
    else if (s->L > 1)
      rc = thermal_start(ctx, fine, s->state[0], B, chain0);
    else
      rc = mlmcpi_init_state(ctx, fine, s->state[0], B, chain0, 0);
  }
  if (rc) {
    mlmcpi_sampler_destroy(s);
    return rc;
  }
  *out = s;
  return 0;
}

void mlmcpi_sampler_destroy(mlmcpi_sampler *s) {
  DeviceGuard device_guard(s ? s->ctx : nullptr);
  if (!s)
    return;
  cudaStreamSynchronize(s->ctx->stream);
  if (s->copy_stream) {
    cudaStreamSynchronize(s->copy_stream);
    for (int k = 0; k < 9; ++k)
      cudaEventDestroy(s->ev[k]);
    cudaStreamDestroy(s->copy_stream);
  }
  if (s->psi)
    cudaFree(s->psi);
  if (s->stage_x)
    cudaFree(s->stage_x);
  if (s->stage_q)
    cudaFree(s->stage_q);
  if (s->stage_acc)
    cudaFree(s->stage_acc);
  for (int b = 0; b < 2; ++b) {
    if (s->ce_stage[b])
      cudaFree(s->ce_stage[b]);
    if (s->ce_flags[b]) {
      cudaFreeHost(s->ce_flags[b]);
      cudaEventDestroy(s->ce_ev_flags[b]);
      cudaEventDestroy(s->ce_ev_copied[b]);
    }
  }
  for (double *d : s->trial)
    if (d)
      cudaFree(d);
  if (s->Scond_lvl)
    cudaFree(s->Scond_lvl);
  if (s->Sprime)
    cudaFree(s->Sprime);
  if (s->S3)
    cudaFree(s->S3);
  if (s->chi_cur)
    cudaFree(s->chi_cur);
  for (mlmcpi_stats *st : s->stats_sampler)
    mlmcpi_stats_destroy(st);
  for (double *d : s->SfL)
    if (d)
      cudaFree(d);
  for (double *d : s->ScondL)
    if (d)
      cudaFree(d);
  for (double *d : s->state)
    if (d)
      cudaFree(d);
  if (s->S_old)
    cudaFree(s->S_old);
  if (s->S_new)
    cudaFree(s->S_new);
  if (s->Sf0)
    cudaFree(s->Sf0);
  if (s->Scond0)
    cudaFree(s->Scond0);
  if (s->Sf)
    cudaFree(s->Sf);
  if (s->Scond)
    cudaFree(s->Scond);
  if (s->q)
    cudaFree(s->q);
  if (s->acc)
    cudaFree(s->acc);
  if (s->acc_step)
    cudaFree(s->acc_step);
  if (s->counters)
    cudaFree(s->counters);
  delete s;
}

int mlmcpi_sampler_set_state(mlmcpi_sampler *s, const double *d_x) {
  DeviceGuard device_guard(s ? s->ctx : nullptr);
  // MultilevelSampler::set_state only overwrites the output buffer that the next draw
  // overwrites again (multilevelsampler.cc:115-117): the chains are unaffected
  if (s->prm.multilevel)
    return 0;
  s->cache0_valid = false;
  s->cascade_valid = false;
  s->chi_valid = false;
  return mlmcpi_copy(s->ctx, s->state[0], d_x, (size_t)mlmcpi_sample_size(&s->model[0]) * s->B);
}

// the current state of every chain (what the sampler classes keep in phi_state_cur / phi_sampler_state[0])
int mlmcpi_sampler_get_state(mlmcpi_sampler *s, double *d_x) {
  DeviceGuard device_guard(s ? s->ctx : nullptr);
  if (!s || !d_x)
    return MLMCPI_EINVAL;
  return mlmcpi_copy(s->ctx, d_x, s->state[0], (size_t)mlmcpi_sample_size(&s->model[0]) * s->B);
}

int mlmcpi_sampler_draw(mlmcpi_sampler *s, double *d_x_out, int32_t *d_accept) {
  DeviceGuard device_guard(s ? s->ctx : nullptr);
  mlmcpi_ctx *ctx = s->ctx;
  const int B = s->B;
  int rc;
  s->work[0] = s->work[1] = s->work[2] = 0.0;
  if (s->prm.multilevel) {
    if ((rc = multilevel_draw(s)))
      return rc;
    s->n_draws++;
    set_i32_kernel<<<cdiv(B, 128), 128, 0, ctx->stream>>>(B, s->acc, 1); // multilevelsampler.cc:72
    MLMCPI_LAUNCHED("set_accept");
    if (d_x_out) // :108-109
      if ((rc = mlmcpi_copy(ctx, d_x_out, s->state[0], (size_t)mlmcpi_sample_size(&s->model[0]) * B)))
        return rc;
    if (d_accept)
      MLMCPI_CUDA(cudaMemcpyAsync(d_accept, s->acc, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice,
                                  ctx->stream));
    return 0;
  }
  const bool cached = cascade_cache_applies(s) && !s->trial.empty();
  if (cached) {
    if ((rc = cascade_draw_cached(s, d_x_out)))
      return rc;
  } else {
    s->cascade_valid = false;
    s->chi_valid = false;
    if ((rc = sampler_draw_range(s, 0, B, s->cache0_valid)))
      return rc;
  }
  s->cache0_valid = s->L > 1;
  s->draw++;
  s->cluster_updates += std::max(1, s->prm.n_updates);
  s->n_draws++;
  if (d_x_out && !cached) // hierarchicalsampler.cc:78-80 (the cached cascade wrote it with the commit of level 0)
    if ((rc = launch_masked_copy(ctx, d_x_out, s->state[0], (size_t)mlmcpi_sample_size(&s->model[0]), B,
                                 s->acc)))
      return rc;
  if (d_accept)
    MLMCPI_CUDA(cudaMemcpyAsync(d_accept, s->acc, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice,
                                ctx->stream));
  return 0;
}

// QoI::evaluate of the chains' current states.  For QOI_SCHWINGER_CHI with the cached cascade the values are kept up
// to date by the draws themselves (the fill-in of the finest level returns the charge of the trial state): a copy
// of B doubles instead of a pass over the states.
int mlmcpi_sampler_qoi(mlmcpi_sampler *s, int qoi, double *d_q) {
  DeviceGuard device_guard(s ? s->ctx : nullptr);
  if (!s || !d_q)
    return MLMCPI_EINVAL;
  if (qoi == MLMCPI_QOI_SCHWINGER_CHI && s->chi_cur && s->chi_valid && s->model[0].model == MLMCPI_SCHWINGER)
    return mlmcpi_copy(s->ctx, d_q, s->chi_cur, (size_t)s->B);
  return mlmcpi_qoi(s->ctx, &s->model[0], qoi, s->state[0], s->B, d_q, nullptr);
}

// Host-buffer entry point.  With an input state the batch is cut into ranges of chains: the
// H2D copy of range k+1 runs on a copy stream while the kernels of range k run on the
// context's stream (chains are independent), so a step costs max(copy, compute) + one
// range instead of copy + compute.
int mlmcpi_sampler_draw_host(mlmcpi_sampler *s, const double *h_x_in, int qoi, double *h_q,
                             double *h_x_out) {
  DeviceGuard device_guard(s ? s->ctx : nullptr);
  mlmcpi_ctx *ctx = s->ctx;
  const size_t nd = (size_t)mlmcpi_sample_size(&s->model[0]);
  const size_t n = nd * s->B;
  int rc;
  if (!h_x_in || s->prm.multilevel) {
    if ((rc = mlmcpi_sampler_draw(s, nullptr, nullptr)))
      return rc;
    if (h_q)
      if ((rc = mlmcpi_qoi(ctx, &s->model[0], qoi, s->state[0], s->B, s->q, nullptr)))
        return rc;
  } else {
    // ranges of at least 8 MiB (below that the per-range launches cost more than the overlap gains)
    const int n_ranges = (int)std::max<size_t>(1, std::min<size_t>({(size_t)s->B, (size_t)8, n * sizeof(double) >> 23}));
    if (!s->copy_stream) {
      MLMCPI_CUDA(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
      for (int k = 0; k < 9; ++k)
        MLMCPI_CUDA(cudaEventCreateWithFlags(&s->ev[k], cudaEventDisableTiming));
    }
    // uploads must not overtake earlier work of the compute stream on state[0]
    MLMCPI_CUDA(cudaEventRecord(s->ev[8], ctx->stream));
    MLMCPI_CUDA(cudaStreamWaitEvent(s->copy_stream, s->ev[8], 0));
    s->work[0] = s->work[1] = s->work[2] = 0.0;
    for (int k = 0; k < n_ranges; ++k) {
      const int c0 = (int)((long long)s->B * k / n_ranges), c1 = (int)((long long)s->B * (k + 1) / n_ranges);
      MLMCPI_CUDA(cudaMemcpyAsync(s->state[0] + c0 * nd, h_x_in + c0 * nd, (size_t)(c1 - c0) * nd * sizeof(double),
                                  cudaMemcpyHostToDevice, s->copy_stream));
      MLMCPI_CUDA(cudaEventRecord(s->ev[k], s->copy_stream));
    }
    for (int k = 0; k < n_ranges; ++k) {
      const int c0 = (int)((long long)s->B * k / n_ranges), c1 = (int)((long long)s->B * (k + 1) / n_ranges);
      MLMCPI_CUDA(cudaStreamWaitEvent(ctx->stream, s->ev[k], 0));
      s->cascade_valid = false;
      s->chi_valid = false;
      if ((rc = sampler_draw_range(s, c0, c1 - c0, false)))
        return rc;
      if (h_q)
        if ((rc = mlmcpi_qoi(ctx, &s->model[0], qoi, s->state[0] + c0 * nd, c1 - c0, s->q + c0, nullptr)))
          return rc;
    }
    s->cache0_valid = s->L > 1;
    s->draw++;
    s->cluster_updates += std::max(1, s->prm.n_updates);
    s->n_draws++;
  }
  if (h_q)
    MLMCPI_CUDA(cudaMemcpyAsync(h_q, s->q, sizeof(double) * s->B, cudaMemcpyDeviceToHost, ctx->stream));
  if (h_x_out)
    MLMCPI_CUDA(cudaMemcpyAsync(h_x_out, s->state[0], n * sizeof(double), cudaMemcpyDeviceToHost,
                                ctx->stream));
  MLMCPI_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// Sampler::draw(state) with the state handed back to the HOST, pipelined: the chains stay resident on the
// device (as the reference's samplers keep phi_state_cur), one draw is made, the QoI of the new states is
// evaluated, and the states of the ACCEPTED chains are snapshot into a staging buffer whose device-to-host copy
// runs on a second stream WHILE THE NEXT DRAW COMPUTES (a rejected draw leaves the caller's buffer untouched,
// as Sampler::draw does).  The call returns without synchronising; h_q / h_x_out of this call are
// complete after the next call of this function or after mlmcpi_sampler_wait_host.  (The caller alternates
// two pinned host buffers.)  Per step the device pays the draw and one device-to-device snapshot; the host
// link pays sample_size * B * 8 bytes.
// the copies of the previous step's hand-over: needs that step's accept flags on the host, i.e. blocks until its draw
// has finished; one cudaMemcpyAsync per run of consecutive accepted chains on the copy stream
static int ce_issue(mlmcpi_sampler *s) {
  if (!s->ce_pending)
    return 0;
  mlmcpi_ctx *ctx = s->ctx;
  const int b = s->ce_pending_stage;
  const size_t nd = (size_t)mlmcpi_sample_size(&s->model[0]);
  MLMCPI_CUDA(cudaEventSynchronize(s->ce_ev_flags[b]));
  const int32_t *acc = s->ce_flags[b];
  for (int c = 0; c < s->B;) {
    if (!acc[c]) {
      ++c;
      continue;
    }
    int c1 = c;
    while (c1 < s->B && acc[c1])
      ++c1;
    MLMCPI_CUDA(cudaMemcpyAsync(s->ce_target + (size_t)c * nd, s->ce_stage[b] + (size_t)c * nd,
                                (size_t)(c1 - c) * nd * sizeof(double), cudaMemcpyDeviceToHost, s->copy_stream));
    c = c1;
  }
  MLMCPI_CUDA(cudaEventRecord(s->ce_ev_copied[b], s->copy_stream));
  s->ce_used[b] = true;
  s->ce_pending = false;
  return 0;
}

int mlmcpi_sampler_draw_host_async(mlmcpi_sampler *s, int qoi, double *h_q, double *h_x_out) {
  DeviceGuard device_guard(s ? s->ctx : nullptr);
  mlmcpi_ctx *ctx = s->ctx;
  const size_t n = (size_t)mlmcpi_sample_size(&s->model[0]) * s->B;
  int rc;
  if ((rc = ce_issue(s)))
    return rc;
  {
    cudaPointerAttributes attr;
    const bool pinned = h_x_out && cudaPointerGetAttributes(&attr, h_x_out) == cudaSuccess &&
                        attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned && ctx->host_copy_engine) {
      // Copy-engine hand-over (default): 54 GB/s over the host link against 38 GB/s for stores from the SMs.  The
      // accept flags of this step go to pinned host memory; the NEXT call (or mlmcpi_sampler_wait_host) reads them
      // and issues one copy per run of accepted chains, which then overlaps the draw that call enqueues.  Two
      // snapshot buffers, so that the snapshot of step k + 1 does not wait for the copies of step k.
      if (!s->copy_stream) {
        MLMCPI_CUDA(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 9; ++k)
          MLMCPI_CUDA(cudaEventCreateWithFlags(&s->ev[k], cudaEventDisableTiming));
      }
      for (int b = 0; b < 2; ++b)
        if (!s->ce_stage[b]) {
          if ((rc = mlmcpi_alloc(ctx, n, &s->ce_stage[b])))
            return rc;
          MLMCPI_CUDA(cudaMallocHost((void **)&s->ce_flags[b], sizeof(int32_t) * s->B));
          MLMCPI_CUDA(cudaEventCreateWithFlags(&s->ce_ev_flags[b], cudaEventDisableTiming));
          MLMCPI_CUDA(cudaEventCreateWithFlags(&s->ce_ev_copied[b], cudaEventDisableTiming));
        }
      if ((rc = mlmcpi_sampler_draw(s, nullptr, nullptr)))
        return rc;
      if (h_q) {
        if ((rc = mlmcpi_sampler_qoi(s, qoi, s->q)))
          return rc;
        MLMCPI_CUDA(cudaMemcpyAsync(h_q, s->q, sizeof(double) * s->B, cudaMemcpyDeviceToHost, ctx->stream));
      }
      const int b = s->ce_cur;
      s->ce_cur ^= 1;
      if (s->ce_used[b]) // the copies that read this snapshot buffer two steps ago
        MLMCPI_CUDA(cudaStreamWaitEvent(ctx->stream, s->ce_ev_copied[b], 0));
      if ((rc = launch_masked_copy(ctx, s->ce_stage[b], s->state[0], n / s->B, s->B, s->acc)))
        return rc;
      MLMCPI_CUDA(cudaMemcpyAsync(s->ce_flags[b], s->acc, sizeof(int32_t) * s->B, cudaMemcpyDeviceToHost, ctx->stream));
      MLMCPI_CUDA(cudaEventRecord(s->ce_ev_flags[b], ctx->stream));
      s->ce_pending = true;
      s->ce_pending_stage = b;
      s->ce_target = h_x_out;
      return 0;
    }
  }
  if (!s->copy_stream) {
    MLMCPI_CUDA(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 9; ++k)
      MLMCPI_CUDA(cudaEventCreateWithFlags(&s->ev[k], cudaEventDisableTiming));
  }
  if (h_x_out && !s->stage_x) {
    if ((rc = mlmcpi_alloc(ctx, n, &s->stage_x)))
      return rc;
    MLMCPI_CUDA(cudaMalloc((void **)&s->stage_acc, sizeof(int32_t) * s->B));
  }
  if (h_q && !s->stage_q && (rc = mlmcpi_alloc(ctx, (size_t)s->B, &s->stage_q)))
    return rc;
  if ((rc = mlmcpi_sampler_draw(s, nullptr, nullptr)))
    return rc;
  if (h_q && (rc = mlmcpi_sampler_qoi(s, qoi, s->q)))
    return rc;
  // the snapshot must not overwrite the staging buffers while the previous copy still reads them
  if (s->host_copy_pending)
    MLMCPI_CUDA(cudaStreamWaitEvent(ctx->stream, s->ev[7], 0));
  // pinned (device-addressable) host memory: only the accepted chains' rows are snapshot and sent
  double *d_view = nullptr;
  if (h_x_out) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, h_x_out) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
        attr.devicePointer && (n / s->B) % 2 == 0)
      d_view = static_cast<double *>(attr.devicePointer);
    cudaGetLastError();
    if (d_view) {
      if ((rc = launch_masked_copy(ctx, s->stage_x, s->state[0], n / s->B, s->B, s->acc)))
        return rc;
      MLMCPI_CUDA(cudaMemcpyAsync(s->stage_acc, s->acc, sizeof(int32_t) * s->B, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      MLMCPI_CUDA(cudaMemcpyAsync(s->stage_x, s->state[0], n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    }
  }
  if (h_q)
    MLMCPI_CUDA(cudaMemcpyAsync(s->stage_q, s->q, sizeof(double) * s->B, cudaMemcpyDeviceToDevice, ctx->stream));
  MLMCPI_CUDA(cudaEventRecord(s->ev[6], ctx->stream));
  MLMCPI_CUDA(cudaStreamWaitEvent(s->copy_stream, s->ev[6], 0));
  if (h_q)
    MLMCPI_CUDA(cudaMemcpyAsync(h_q, s->stage_q, sizeof(double) * s->B, cudaMemcpyDeviceToHost, s->copy_stream));
  if (h_x_out) {
    // Sampler::draw(state) leaves `state` untouched when the draw is rejected (hierarchicalsampler.cc:78-80):
    // only the ACCEPTED chains' states cross the host link.  Pinned host memory is addressable from the
    // device, so the masked copy kernel writes the accepted rows straight into the caller's buffer
    // (coalesced 16-byte stores over PCIe; no host round trip to learn which chains accepted).  Pageable
    // buffers get the full copy.
    if (d_view) {
      const size_t n2 = n / s->B / 2;
      const dim3 grid((unsigned)std::min<size_t>(32, (n2 + 1023) / 1024), (unsigned)std::min(s->B, 65535));
      masked_copy_kernel<false><<<grid, 256, 0, s->copy_stream>>>(reinterpret_cast<double2 *>(d_view),
                                                                 reinterpret_cast<const double2 *>(s->stage_x), n2,
                                                                 s->B, s->stage_acc);
      MLMCPI_LAUNCHED("masked_copy_to_host");
    } else {
      MLMCPI_CUDA(cudaMemcpyAsync(h_x_out, s->stage_x, n * sizeof(double), cudaMemcpyDeviceToHost, s->copy_stream));
    }
  }
  MLMCPI_CUDA(cudaEventRecord(s->ev[7], s->copy_stream));
  s->host_copy_pending = true;
  return 0;
}
int mlmcpi_sampler_wait_host(mlmcpi_sampler *s) {
  DeviceGuard device_guard(s ? s->ctx : nullptr);
  mlmcpi_ctx *ctx = s->ctx;
  int rc;
  if ((rc = ce_issue(s)))
    return rc;
  if (s->copy_stream)
    MLMCPI_CUDA(cudaStreamSynchronize(s->copy_stream));
  MLMCPI_CUDA(cudaStreamSynchronize(ctx->stream));
  s->host_copy_pending = false;
  return 0;
}

int mlmcpi_sampler_level_model(const mlmcpi_sampler *s, int level, mlmcpi_model *m) {
  DeviceGuard device_guard(s ? s->ctx : nullptr);
  if (!s || !m || level < 0 || level >= s->L)
    return MLMCPI_EINVAL;
  *m = s->model[level];
  return 0;
}

int mlmcpi_sampler_stats(mlmcpi_sampler *s, double *h_p_accept) {
  DeviceGuard device_guard(s ? s->ctx : nullptr);
  mlmcpi_ctx *ctx = s->ctx;
  std::vector<unsigned long long> c(s->L);
  MLMCPI_CUDA(cudaMemcpyAsync(c.data(), s->counters, sizeof(unsigned long long) * s->L,
                              cudaMemcpyDeviceToHost, ctx->stream));
  MLMCPI_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int l = s->L - 1; l >= 0; --l) {
    // MCMCStep::p_accept = accepted / total of THAT step (montecarlo/mcmcstep.hh).  In the
    // hierarchical cascade a two-level step only runs -- and counts a sample -- when every coarser
    // level accepted (the break at hierarchicalsampler.cc:73-74): the chains that reach level l are
    // the chains level l+1 accepted.  The level walk of the MultilevelSampler takes n_steps[l] steps
    // of all chains on level l.
    double n;
    if (s->prm.multilevel)
      n = (double)s->n_steps[l] * s->B;
    else if (l == s->L - 1)
      n = (double)s->n_draws * s->B;
    else
      n = (double)c[l + 1];
    h_p_accept[l] = n > 0 ? (double)c[l] / n : 0.0;
  }
  return 0;
}

// MCMCStep::reset_stats (montecarlo/mcmcstep.hh) for every level of the sampler
int mlmcpi_sampler_reset_stats(mlmcpi_sampler *s) {
  DeviceGuard device_guard(s ? s->ctx : nullptr);
  mlmcpi_ctx *ctx = s->ctx;
  MLMCPI_CUDA(cudaMemsetAsync(s->counters, 0, sizeof(unsigned long long) * s->L, ctx->stream));
  s->n_draws = 0;
  for (auto &n : s->n_steps)
    n = 0;
  return 0;
}

int mlmcpi_sampler_work(const mlmcpi_sampler *s, double out[3]) {
  DeviceGuard device_guard(s ? s->ctx : nullptr);
  out[0] = s->work[0];
  out[1] = s->work[1];
  out[2] = s->work[2];
  return 0;
}

// HMCSampler::autotune_stepsize, sampler/hmcsampler.cc:72-113: bisection of dt on
// [dt/2, 2 dt] towards the target acceptance, n_samples single_step()s per round.  The
// B chains of the batch supply the samples in parallel: ceil(n_samples / B) steps per
// round.  Acts on the sampler of the coarsest level.  Like the reference it keeps dt
// unchanged if no round came within 1e-2 of the target.
int mlmcpi_sampler_autotune(mlmcpi_sampler *s, double p_accept_target, int n_rounds, int n_samples,
                            double *dt_out, double *p_accept_out) {
  DeviceGuard device_guard(s ? s->ctx : nullptr);
  mlmcpi_ctx *ctx = s->ctx;
  if (s->prm.kind != MLMCPI_SAMPLER_HMC)
    return ctx_fail(ctx, MLMCPI_EINVAL, "autotune is defined for the HMC sampler");
  s->cascade_valid = false; // the tuning steps advance state[L-1]
  const int l = s->L - 1, B = s->B;
  const mlmcpi_model *m = &s->model[l];
  for (int lev = 1; lev < s->L; ++lev) { // tune on the restricted current state
    int rc = mlmcpi_restrict(ctx, &s->model[lev - 1], s->state[lev - 1], s->state[lev], B);
    if (rc)
      return rc;
  }
  const double dt_original = s->prm.dt;
  double dt_min = 0.5 * s->prm.dt, dt_max = 2. * s->prm.dt, dt = s->prm.dt, p_acc = 0.0;
  bool converged = false;
  const int steps = (n_samples + B - 1) / B;
  unsigned long long *cnt = nullptr;
  MLMCPI_CUDA(cudaMalloc((void **)&cnt, sizeof(unsigned long long)));
  for (int k = 0; k < n_rounds; ++k) {
    dt = 0.5 * (dt_min + dt_max);
    cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), ctx->stream);
    for (int j = 0; j < steps; ++j) {
      int rc = mlmcpi_hmc_step(ctx, m, s->prm.nt, dt, s->state[l], B, s->chain0,
                               level_draw(s->draw++, l, 0xff), s->acc, nullptr);
      if (rc) {
        cudaFree(cnt);
        return rc;
      }
      count_accept_kernel<<<std::min(cdiv(B, 256), 64), 256, 0, ctx->stream>>>(B, s->acc, cnt);
    }
    unsigned long long h = 0;
    cudaMemcpyAsync(&h, cnt, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    double acc_tot[2] = {(double)h, (double)steps * B};
    {
      const int rc = ctx_allreduce_host(ctx, acc_tot, 2); // all processes tune to the same dt
      if (rc) {
        cudaFree(cnt);
        return rc;
      }
    }
    p_acc = acc_tot[0] / acc_tot[1];
    if (p_acc > p_accept_target)
      dt_min = dt;
    else
      dt_max = dt;
    if (std::fabs(p_acc - p_accept_target) < 1.E-2)
      converged = true;
  }
  cudaFree(cnt);
  s->prm.dt = converged ? dt : dt_original;
  if (dt_out)
    *dt_out = s->prm.dt;
  if (p_accept_out)
    *p_accept_out = p_acc;
  return converged ? 0 : 1; /* 1: "FAILED to tune, reverting" (hmcsampler.cc:107-110) */
}

int mlmcpi_sampler_set_dt(mlmcpi_sampler *s, double dt) {
  DeviceGuard device_guard(s ? s->ctx : nullptr);
  s->prm.dt = dt;
  return 0;
}

} // extern "C"

// ====================================================== cost and MLMC estimator
extern "C" {

int mlmcpi_sampler_cost(mlmcpi_sampler *s, int n_meas, double *usec_per_sample) {
  DeviceGuard device_guard(s ? s->ctx : nullptr);
  mlmcpi_ctx *ctx = s->ctx;
  if (n_meas < 1 || !usec_per_sample)
    return ctx_fail(ctx, MLMCPI_EINVAL, "bad arguments");
  cudaEvent_t e0, e1;
  MLMCPI_CUDA(cudaEventCreate(&e0));
  MLMCPI_CUDA(cudaEventCreate(&e1));
  MLMCPI_CUDA(cudaEventRecord(e0, ctx->stream));
  for (int k = 0; k < n_meas; ++k) {
    int rc = mlmcpi_sampler_draw(s, nullptr, nullptr);
    if (rc)
      return rc;
  }
  MLMCPI_CUDA(cudaEventRecord(e1, ctx->stream));
  MLMCPI_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  MLMCPI_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *usec_per_sample = 1.0e3 * ms / ((double)n_meas * s->B);
  return 0;
}

int mlmcpi_sampler_indep(const mlmcpi_sampler *s, double *out) {
  DeviceGuard device_guard(s ? s->ctx : nullptr);
  if (!s->prm.multilevel)
    return MLMCPI_EINVAL;
  for (int l = 0; l < s->L; ++l) {
    out[l] = s->t_indep[l];
    out[s->L + l] = s->n_indep[l];
  }
  return 0;
}

} // extern "C"

struct mlmcpi_mlmc {
  mlmcpi_ctx *ctx = nullptr;
  mlmcpi_mlmc_params prm;
  int B = 0, L = 0;
  uint32_t chain0 = 0;
  uint64_t draw = 0;
  std::vector<mlmcpi_model> model;            // [L]
  std::vector<mlmcpi_sampler *> coarse_sampler; // [L-1]: sampler on level l+1
  std::vector<double *> phi_state, phi_coarse_state; // [L] device [B][n_l]
  std::vector<double *> Sf, Scond;            // [L-1] two-level caches
  std::vector<mlmcpi_stats *> stats_qoi;      // [L]   Y_l
  std::vector<mlmcpi_stats *> stats_coarse;   // [L-1] Q_sampler[l]
  std::vector<double> t_indep, cost_twolevel, cost_sampler;
  std::vector<int> n_indep, t_sampler;
  std::vector<double> n_target;
  double *q_fine = nullptr, *q_coarse = nullptr;
  int32_t *acc = nullptr;
  std::vector<TauCache> tau_cache; // [L-1] cached tau_int of stats_coarse (sub-sampling decisions)
};

namespace {

__global__ void diff_kernel(int B, double *a, const double *b) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < B)
    a[c] -= b[c];
}

// MonteCarloMultiLevel::draw_coarse_sample, montecarlomultilevel.cc:170-190 (sub_sample_coarse)
int mlmc_draw_coarse_sample(mlmcpi_mlmc *m, int level, double *state) {
  mlmcpi_ctx *ctx = m->ctx;
  const int k_max = m->prm.n_autocorr_window;
  double st[6];
  int rc;
  if (m->tau_cache.size() != (size_t)(m->L - 1))
    m->tau_cache.assign(m->L - 1, TauCache());
  if ((rc = stats_query_cached(m->stats_coarse[level - 1], k_max, m->tau_cache[level - 1], st)))
    return rc;
  const double tau_int = std::ceil(2. * st[3]);
  while (m->t_sampler[level - 1] < tau_int) {
    if ((rc = mlmcpi_sampler_draw(m->coarse_sampler[level - 1], state, nullptr)))
      return rc;
    if ((rc = mlmcpi_qoi(ctx, &m->model[level], m->prm.qoi, state, m->B, m->q_coarse, nullptr)))
      return rc;
    if ((rc = mlmcpi_stats_record(m->stats_coarse[level - 1], m->q_coarse)))
      return rc;
    m->t_sampler[level - 1]++;
  }
  m->t_indep[level - 1] = (m->n_indep[level - 1] * m->t_indep[level - 1] + m->t_sampler[level - 1]) /
                          (1.0 + m->n_indep[level - 1]);
  m->n_indep[level - 1]++;
  m->t_sampler[level - 1] = 0;
  return 0;
}

// one batched sample of Y_level (B chains); independent = draw_coarse_sample, else plain draw
int mlmc_sample(mlmcpi_mlmc *m, int level, bool independent) {
  mlmcpi_ctx *ctx = m->ctx;
  const int L = m->L, B = m->B;
  int rc;
  if (level == L - 1) { // :86-88, :120-122
    if (independent)
      rc = mlmc_draw_coarse_sample(m, level, m->phi_state[level]);
    else
      rc = mlmcpi_sampler_draw(m->coarse_sampler[level - 1], m->phi_state[level], nullptr);
    if (rc)
      return rc;
    if ((rc = mlmcpi_qoi(ctx, &m->model[level], m->prm.qoi, m->phi_state[level], B, m->q_fine, nullptr)))
      return rc;
  } else { // :90-97, :124-132
    if (independent)
      rc = mlmc_draw_coarse_sample(m, level + 1, m->phi_coarse_state[level + 1]);
    else
      rc = mlmcpi_sampler_draw(m->coarse_sampler[level], m->phi_coarse_state[level + 1], nullptr);
    if (rc)
      return rc;
    if ((rc = twolevel_step_impl(ctx, &m->model[level], &m->model[level + 1], m->phi_coarse_state[level + 1],
                                 m->phi_state[level], m->Sf[level], m->Scond[level], B, m->chain0,
                                 (m->draw++ << 12) | ((uint64_t)level << 8) | 0xfe, nullptr, m->acc, nullptr)))
      return rc;
    if ((rc = mlmcpi_qoi(ctx, &m->model[level], m->prm.qoi, m->phi_state[level], B, m->q_fine, nullptr)))
      return rc;
    if ((rc = mlmcpi_qoi(ctx, &m->model[level + 1], m->prm.qoi, m->phi_coarse_state[level + 1], B, m->q_coarse,
                         nullptr)))
      return rc;
    diff_kernel<<<cdiv(B, 128), 128, 0, ctx->stream>>>(B, m->q_fine, m->q_coarse);
    MLMCPI_LAUNCHED("mlmc_diff");
  }
  return mlmcpi_stats_record(m->stats_qoi[level], m->q_fine);
}

// montecarlomultilevel.cc:193-204
int mlmc_cost_eff(mlmcpi_mlmc *m, int ell, double *cost_out) {
  double cost;
  if (ell == m->L - 1)
    cost = m->t_indep[ell - 1] * m->cost_sampler[ell - 1];
  else
    cost = m->cost_twolevel[ell] + m->t_indep[ell] * m->cost_sampler[ell];
  double st[6];
  int rc = stats_query(m->stats_qoi[ell], m->prm.n_autocorr_window, st);
  if (rc)
    return rc;
  *cost_out = std::ceil(st[3]) * cost;
  return 0;
}

} // namespace

extern "C" {

void mlmcpi_mlmc_destroy(mlmcpi_mlmc *m) {
  DeviceGuard device_guard(m ? m->ctx : nullptr);
  if (!m)
    return;
  cudaStreamSynchronize(m->ctx->stream);
  for (auto *s : m->coarse_sampler)
    mlmcpi_sampler_destroy(s);
  for (auto *st : m->stats_qoi)
    mlmcpi_stats_destroy(st);
  for (auto *st : m->stats_coarse)
    mlmcpi_stats_destroy(st);
  for (auto &v : {m->phi_state, m->phi_coarse_state, m->Sf, m->Scond})
    for (double *d : v)
      if (d)
        cudaFree(d);
  if (m->q_fine)
    cudaFree(m->q_fine);
  if (m->q_coarse)
    cudaFree(m->q_coarse);
  if (m->acc)
    cudaFree(m->acc);
  delete m;
}

// MonteCarloMultiLevel constructor, montecarlomultilevel.cc:7-68
int mlmcpi_mlmc_create(mlmcpi_ctx *ctx, const mlmcpi_model *fine, const mlmcpi_mlmc_params *prm, int B,
                       uint32_t chain0, mlmcpi_mlmc **out) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !fine || !prm || !out || B <= 0)
    return MLMCPI_EINVAL;
  *out = nullptr;
  if (prm->n_level < 2 || prm->n_level > 16)
    return ctx_fail(ctx, MLMCPI_EINVAL, "multilevel MC needs 2 <= n_level <= 16");
  mlmcpi_mlmc *m = new (std::nothrow) mlmcpi_mlmc;
  if (!m)
    return MLMCPI_ENOMEM;
  m->ctx = ctx;
  m->prm = *prm;
  m->B = B;
  m->L = prm->n_level;
  m->chain0 = chain0;
  if (m->prm.n_autocorr_window < 1)
    m->prm.n_autocorr_window = 20;
  const int L = m->L;
  if (L > 1 && fine->model == MLMCPI_GFF &&
      (prm->sampler.ctype != MLMCPI_COARSEN_ROTATE || fine->coarsening != MLMCPI_COARSEN_ROTATE)) {
    delete m;
    return ctx_fail(ctx, MLMCPI_EUNSUPPORTED,
                    "a GFF hierarchy needs coarsening = rotate (in the sampler parameters and in the fine-level model)");
  }
  m->model.push_back(*fine);
  int rc = 0;
  for (int l = 0; l + 1 < L && !rc; ++l) {
    mlmcpi_model c;
    const int level = (fine->model == MLMCPI_GFF && fine->rotated ? 1 : 0) + l;
    rc = mlmcpi_coarse_model(&m->model[l], prm->sampler.renorm, level, prm->sampler.ctype,
                             m->model[l].T_final, &c);
    if (c.model == MLMCPI_GFF && !ctx->gff_coarse_smoothing)
      c.gff_n_gibbs = 0; // see mlmcpi_sampler_create
    m->model.push_back(c);
  }
  if (rc) {
    delete m;
    return ctx_fail(ctx, rc, "cannot construct the coarse action of a level");
  }
  m->t_indep.assign(L, 0.0);
  m->n_indep.assign(L, 0);
  m->t_sampler.assign(L, 0);
  m->cost_twolevel.assign(L, 0.0);
  m->cost_sampler.assign(L, 0.0);
  m->n_target.assign(L, 0.0);
  for (int l = 0; l < L && !rc; ++l) {
    double *a = nullptr, *b = nullptr;
    const size_t n = (size_t)mlmcpi_sample_size(&m->model[l]) * B;
    rc = mlmcpi_alloc(ctx, n, &a) || mlmcpi_alloc(ctx, n, &b);
    m->phi_state.push_back(a);
    m->phi_coarse_state.push_back(b);
    mlmcpi_stats *st = nullptr;
    if (!rc)
      rc = mlmcpi_stats_create(ctx, m->prm.n_autocorr_window, B, &st);
    m->stats_qoi.push_back(st);
  }
  for (int l = 0; l + 1 < L && !rc; ++l) {
    // sampler_factory->get(action[l+1]) (:37-39): the hierarchy below level l+1
    mlmcpi_sampler_params sp = prm->sampler;
    sp.n_levels = std::max(1, prm->sampler.n_levels - (l + 1));
    if (sp.n_levels < 2)
      sp.multilevel = 0;
    sp.qoi = prm->qoi;
    mlmcpi_sampler *s = nullptr;
    // distinct chains for every level: offset the global chain index by the level
    rc = mlmcpi_sampler_create(ctx, &m->model[l + 1], &sp, B, chain0 + (uint32_t)(l + 1) * 0x01000000u, &s);
    m->coarse_sampler.push_back(s);
    mlmcpi_stats *st = nullptr;
    if (!rc)
      rc = mlmcpi_stats_create(ctx, m->prm.n_autocorr_window, B, &st);
    m->stats_coarse.push_back(st);
    double *a = nullptr, *b = nullptr;
    if (!rc)
      rc = mlmcpi_alloc(ctx, B, &a) || mlmcpi_alloc(ctx, B, &b);
    m->Sf.push_back(a);
    m->Scond.push_back(b);
    // The reference's sampler constructors burn in (hmcsampler.hh:99-108, ...); so does this one, from
    // the thermalised zero state, and phi_coarse_state / the coarsest phi_state start as a state of the
    // sampler: a two-level step never sees the all-zero coarse state as a "sample".  (The reference's
    // TwoLevelMetropolisStep constructor times 10 000 draws with a zero coarse state,
    // twolevelmetropolisstep.cc:23-29, which parks theta_fine in a state of very large weight
    // pi_f / (pi_c q); one long chain forgets that, B short ones do not: 16^2, beta = 4, n_burnin = 100
    // left a bias of 9 sigma -- profiles/r01_summary.md 10.2.)
    const int n_therm = std::max(m->prm.n_burnin, 100);
    const size_t n_c = (size_t)mlmcpi_sample_size(&m->model[l + 1]) * B;
    if (!rc && s->L == 1 && s->prm.kind != MLMCPI_SAMPLER_EXACT)
      rc = thermal_start(ctx, &m->model[l + 1], s->state[0], B, chain0);
    for (int k = 0; k < n_therm && !rc; ++k)
      rc = mlmcpi_sampler_draw(s, nullptr, nullptr);
    if (!rc)
      rc = mlmcpi_copy(ctx, m->phi_coarse_state[l + 1], s->state[0], n_c);
    if (!rc && l + 1 == L - 1)
      rc = mlmcpi_copy(ctx, m->phi_state[l + 1], s->state[0], n_c);
    // TwoLevelMetropolisStep constructor (twolevelmetropolisstep.cc:11-22): zero state.  Here EVERY one
    // of the B chains has to forget its start, and the exactly cold state is metastable under the
    // two-level step (16^2, beta = 4: 3 % acceptance over the first 100 draws; 32^2: none in 2000).  The
    // chain of level l starts from a state of the sampler the user configured, run on level l itself:
    // the (burnt-in) coarse sampler of the next finer level for l >= 1, a temporary one for l = 0.
    if (!rc) {
      mlmcpi_sampler *ts = (l == 0) ? nullptr : m->coarse_sampler[l - 1];
      if (l == 0) {
        mlmcpi_sampler_params tp = prm->sampler;
        tp.n_levels = std::max(1, prm->sampler.n_levels);
        tp.multilevel = 0;
        tp.qoi = prm->qoi;
        rc = mlmcpi_sampler_create(ctx, &m->model[0], &tp, B, chain0 + 0x00400000u, &ts);
        if (!rc && ts->L == 1 && ts->prm.kind != MLMCPI_SAMPLER_EXACT)
          rc = thermal_start(ctx, &m->model[0], ts->state[0], B, chain0);
        for (int k = 0; k < n_therm && !rc; ++k)
          rc = mlmcpi_sampler_draw(ts, nullptr, nullptr);
      }
      if (!rc)
        rc = mlmcpi_copy(ctx, m->phi_state[l], ts->state[0], (size_t)mlmcpi_sample_size(&m->model[l]) * B);
      if (l == 0 && ts)
        mlmcpi_sampler_destroy(ts);
    }
    if (!rc)
      rc = mlmcpi_action(ctx, &m->model[l], m->phi_state[l], B, a);
    if (!rc)
      rc = mlmcpi_cond_action(ctx, &m->model[l], m->phi_state[l], B, b);
  }
  if (!rc)
    rc = mlmcpi_alloc(ctx, B, &m->q_fine) || mlmcpi_alloc(ctx, B, &m->q_coarse);
  if (!rc && cudaMalloc((void **)&m->acc, sizeof(int32_t) * B) != cudaSuccess)
    rc = MLMCPI_ENOMEM;
  if (rc) {
    mlmcpi_mlmc_destroy(m);
    return rc;
  }
  *out = m;
  return 0;
}

// MonteCarloMultiLevel::evaluate, montecarlomultilevel.cc:71-167
int mlmcpi_mlmc_evaluate(mlmcpi_mlmc *m) {
  DeviceGuard device_guard(m ? m->ctx : nullptr);
  mlmcpi_ctx *ctx = m->ctx;
  const int L = m->L, B = m->B, k_max = m->prm.n_autocorr_window;
  int rc;
  // cost_per_sample of the samplers and two-level steps (measured in the reference's constructors with 10000
  // draws each, sampler/sampler.hh cost_per_sample): CUDA-event measurement, usec per chain-sample.  One untimed
  // draw first (work buffers, library handles), then batches of 8, 32, 128 draws until a batch lasts 10 ms --
  // with B chains per draw that is 8 B ... 128 B chain-samples per figure.
  for (int l = 0; l + 1 < L; ++l) {
    double unused;
    if ((rc = mlmcpi_sampler_cost(m->coarse_sampler[l], 1, &unused)))
      return rc;
    // (every process must run the same number of batches -- a draw of a MultilevelSampler contains all-reduced
    // decisions -- so the time that ends the loop is the mean over the processes)
    for (int n = 8; n <= 128; n *= 4) {
      if ((rc = mlmcpi_sampler_cost(m->coarse_sampler[l], n, &m->cost_sampler[l])))
        return rc;
      double t_usec = m->cost_sampler[l] * n * B;
      if ((rc = ctx_allreduce_host(ctx, &t_usec, 1)))
        return rc;
      if (t_usec / std::max(1, ctx->world) >= 1.0e4)
        break;
    }
    cudaEvent_t e0, e1;
    MLMCPI_CUDA(cudaEventCreate(&e0));
    MLMCPI_CUDA(cudaEventCreate(&e1));
    for (int n = 1; n <= 128; n = (n == 1) ? 8 : 4 * n) { // n == 1: the untimed first step
      MLMCPI_CUDA(cudaEventRecord(e0, ctx->stream));
      for (int k = 0; k < n; ++k)
        if ((rc = twolevel_step_impl(ctx, &m->model[l], &m->model[l + 1], m->phi_coarse_state[l + 1],
                                     m->phi_state[l], m->Sf[l], m->Scond[l], B, m->chain0,
                                     (m->draw++ << 12) | ((uint64_t)l << 8) | 0xfd, nullptr, m->acc, nullptr))) {
          cudaEventDestroy(e0);
          cudaEventDestroy(e1);
          return rc;
        }
      MLMCPI_CUDA(cudaEventRecord(e1, ctx->stream));
      MLMCPI_CUDA(cudaEventSynchronize(e1));
      float ms = 0.f;
      MLMCPI_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      m->cost_twolevel[l] = 1.0e3 * ms / ((double)n * B);
      double t_ms = ms;
      if ((rc = ctx_allreduce_host(ctx, &t_ms, 1))) {
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return rc;
      }
      if (n > 1 && t_ms / std::max(1, ctx->world) >= 10.0)
        break;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  if (ctx->world > 1) { // every process must allocate samples with the same costs: mean over the processes
    std::vector<double> c(m->cost_sampler.begin(), m->cost_sampler.end());
    c.insert(c.end(), m->cost_twolevel.begin(), m->cost_twolevel.end());
    if ((rc = ctx_allreduce_host(ctx, c.data(), c.size())))
      return rc;
    for (int l = 0; l < L; ++l) {
      m->cost_sampler[l] = c[l] / ctx->world;
      m->cost_twolevel[l] = c[L + l] / ctx->world;
    }
  }
  for (int l = 0; l < L; ++l)
    if ((rc = mlmcpi_stats_hard_reset(m->stats_qoi[l])))
      return rc;
  const double two_epsilon_inv2 = 2. / (m->prm.epsilon * m->prm.epsilon);
  for (int level = L - 1; level >= 0; level--) // burn-in, :83-100
    for (int j = 0; j < m->prm.n_burnin; ++j)
      if ((rc = mlmc_sample(m, level, false)))
        return rc;
  for (int l = 0; l < L; ++l) { // :102-108
    if ((rc = mlmcpi_stats_reset(m->stats_qoi[l])))
      return rc;
    if (l < L - 1)
      if ((rc = mlmcpi_stats_reset(m->stats_coarse[l])))
        return rc;
    m->n_target[l] = m->prm.n_min_samples_qoi;
  }
  bool sufficient = false;
  int iterations = 0;
  do {
    for (int level = L - 1; level >= 0; level--) { // :114-137
      double st[6];
      if ((rc = stats_query(m->stats_qoi[level], k_max, st)))
        return rc;
      // B samples per batched draw and process (st[5] counts the samples of all processes)
      for (double j = st[5]; j < m->n_target[level]; j += (double)B * ctx->world)
        if ((rc = mlmc_sample(m, level, true)))
          return rc;
    }
    sufficient = true;
    double sum_s = 0.0;
    std::vector<double> V(L), C(L), tau(L), ns(L);
    for (int l = 0; l < L; ++l) { // :140-146
      double st[6];
      if ((rc = stats_query(m->stats_qoi[l], k_max, st)))
        return rc;
      V[l] = st[1];
      tau[l] = st[3];
      ns[l] = st[5];
      if ((rc = mlmc_cost_eff(m, l, &C[l])))
        return rc;
      sum_s += std::sqrt(V[l] * C[l]);
    }
    for (int l = 0; l < L; ++l) { // :147-158
      m->n_target[l] = std::ceil(two_epsilon_inv2 * sum_s * std::sqrt(V[l] / C[l]) * tau[l]);
      sufficient = sufficient && (ns[l] >= m->n_target[l]);
    }
    ++iterations;
    if (m->prm.max_iterations > 0 && iterations >= m->prm.max_iterations)
      break;
  } while (!sufficient);
  return sufficient ? 0 : 1;
}

int mlmcpi_mlmc_result(mlmcpi_mlmc *m, double *value, double *error, double *level_out) {
  DeviceGuard device_guard(m ? m->ctx : nullptr);
  double v = 0.0, e2 = 0.0;
  for (int l = 0; l < m->L; ++l) {
    double st[6], cost = 0.0;
    int rc = stats_query(m->stats_qoi[l], m->prm.n_autocorr_window, st);
    if (rc)
      return rc;
    if ((rc = mlmc_cost_eff(m, l, &cost)))
      return rc;
    v += st[0];          // numerical_result(), :255-261
    e2 += st[4] * st[4]; // statistical_error(), :264-271
    if (level_out) {
      double *o = level_out + 6 * l;
      o[0] = st[5];
      o[1] = st[0];
      o[2] = st[1];
      o[3] = st[3];
      o[4] = cost;
      o[5] = m->n_target[l];
    }
  }
  if (value)
    *value = v;
  if (error)
    *error = std::sqrt(e2);
  return 0;
}

} // extern "C"

// ================================================================ statistics
// common/statistics.cc:4-27 for every chain; the chain plays the role of the MPI
// rank of the reference.  Rows of acc ([row][chain]):
//   0            avg            (short-term running mean: cleared by reset())
//   1..4         avg_longterm, avg2_longterm, avg3_longterm, avg4_longterm
//   5..5+k-1     S_k
//   5+k..5+2k-1  ring buffer Q_k
struct mlmcpi_stats {
  mlmcpi_ctx *ctx = nullptr;
  int k_max = 0, B = 0;
  unsigned n_samples = 0, n_samples_longterm = 0; // identical for all chains (lockstep)
  double *acc = nullptr;
  double *packed = nullptr;
};

static mlmcpi_ctx *mlmcpi_stats_ctx(mlmcpi_stats *st) { return st->ctx; }

namespace {

__global__ void stats_record_kernel(int B, int k_max, unsigned n_short, unsigned n, double *acc,
                                    const double *q) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= B)
    return;
  // n / n_short = number of long-term / short-term samples including this one
  const double Q = q[c];
  const double w = (n - 1.0), inv = 1.0 / (1.0 * n);
  double *a0 = acc, *a1 = acc + B, *a2 = acc + 2 * (size_t)B, *a3 = acc + 3 * (size_t)B,
         *a4 = acc + 4 * (size_t)B;
  double *S = acc + 5 * (size_t)B, *ring = acc + (5 + (size_t)k_max) * B;
  a0[c] = ((n_short - 1.0) * a0[c] + Q) / (1.0 * n_short);
  a1[c] = (w * a1[c] + Q) * inv;
  a2[c] = (w * a2[c] + Q * Q) * inv;
  a3[c] = (w * a3[c] + Q * Q * Q) * inv;
  a4[c] = (w * a4[c] + Q * Q * Q * Q) * inv;
  const unsigned slot = (n - 1) % k_max;
  ring[(size_t)slot * B + c] = Q;
  const unsigned window = n < (unsigned)k_max ? n : (unsigned)k_max;
  for (unsigned k = 0; k < window; ++k) {
    const unsigned N_k = n - k;
    const double Qk = ring[(size_t)((n - 1 - k) % k_max) * B + c];
    double *Sk = S + (size_t)k * B + c;
    *Sk = ((N_k - 1.0) * (*Sk) + Q * Qk) / (1.0 * N_k);
  }
}

// packed[r] = sum over chains of row r (rows: avg, avg1..avg4, S_0..S_{k_max-1})
__global__ void stats_pack_kernel(int B, const double *acc, double *packed) {
  const int r = blockIdx.x;
  double v = 0.0;
  for (int c = threadIdx.x; c < B; c += blockDim.x)
    v += acc[(size_t)r * B + c];
  v = block_sum(v);
  if (threadIdx.x == 0)
    packed[r] = v;
}

} // namespace

extern "C" {

int mlmcpi_stats_create(mlmcpi_ctx *ctx, int k_max, int B, mlmcpi_stats **out) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !out || k_max < 1 || B < 1)
    return MLMCPI_EINVAL;
  mlmcpi_stats *st = new (std::nothrow) mlmcpi_stats;
  if (!st)
    return MLMCPI_ENOMEM;
  st->ctx = ctx;
  st->k_max = k_max;
  st->B = B;
  if (mlmcpi_alloc(ctx, (size_t)(5 + 2 * k_max) * B, &st->acc) ||
      mlmcpi_alloc(ctx, 5 + k_max, &st->packed)) {
    mlmcpi_stats_destroy(st);
    return MLMCPI_ENOMEM;
  }
  *out = st;
  return 0;
}

void mlmcpi_stats_destroy(mlmcpi_stats *st) {
  DeviceGuard device_guard(st ? mlmcpi_stats_ctx(st) : nullptr);
  if (!st)
    return;
  cudaStreamSynchronize(st->ctx->stream);
  if (st->acc)
    cudaFree(st->acc);
  if (st->packed)
    cudaFree(st->packed);
  delete st;
}

int mlmcpi_stats_hard_reset(mlmcpi_stats *st) { // Statistics::hard_reset, statistics.hh:125-137
  DeviceGuard device_guard(st ? st->ctx : nullptr);
  mlmcpi_ctx *ctx = st->ctx;
  st->n_samples = st->n_samples_longterm = 0;
  MLMCPI_CUDA(cudaMemsetAsync(st->acc, 0, sizeof(double) * (5 + 2 * st->k_max) * st->B, ctx->stream));
  return 0;
}

int mlmcpi_stats_reset(mlmcpi_stats *st) { // Statistics::reset, statistics.hh:119-122
  DeviceGuard device_guard(st ? st->ctx : nullptr);
  mlmcpi_ctx *ctx = st->ctx;
  st->n_samples = 0;
  MLMCPI_CUDA(cudaMemsetAsync(st->acc, 0, sizeof(double) * st->B, ctx->stream));
  return 0;
}

int mlmcpi_stats_record(mlmcpi_stats *st, const double *d_q) {
  DeviceGuard device_guard(st ? mlmcpi_stats_ctx(st) : nullptr);
  mlmcpi_ctx *ctx = st->ctx;
  st->n_samples++;
  st->n_samples_longterm++;
  stats_record_kernel<<<cdiv(st->B, 128), 128, 0, ctx->stream>>>(st->B, st->k_max, st->n_samples,
                                                                st->n_samples_longterm, st->acc, d_q);
  MLMCPI_LAUNCHED("stats_record");
  return 0;
}

int mlmcpi_stats_packed_size(int k_max) { return 8 + k_max; }

static void stats_head(const mlmcpi_stats *st, double head[3]) {
  head[0] = (double)st->B;
  head[1] = (double)st->n_samples_longterm * st->B;
  head[2] = (double)st->n_samples * st->B;
}

int mlmcpi_stats_pack_device(mlmcpi_stats *st, double *d_packed) {
  DeviceGuard device_guard(st ? mlmcpi_stats_ctx(st) : nullptr);
  mlmcpi_ctx *ctx = st->ctx;
  stats_pack_kernel<<<5 + st->k_max, 256, 0, ctx->stream>>>(st->B, st->acc, d_packed + 3);
  MLMCPI_LAUNCHED("stats_pack");
  double head[3];
  stats_head(st, head);
  // 24 bytes from a stack buffer: the driver stages pageable sources before returning
  MLMCPI_CUDA(cudaMemcpyAsync(d_packed, head, sizeof(head), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

int mlmcpi_stats_pack(mlmcpi_stats *st, double *h_packed) {
  DeviceGuard device_guard(st ? mlmcpi_stats_ctx(st) : nullptr);
  mlmcpi_ctx *ctx = st->ctx;
  stats_pack_kernel<<<5 + st->k_max, 256, 0, ctx->stream>>>(st->B, st->acc, st->packed);
  MLMCPI_LAUNCHED("stats_pack");
  stats_head(st, h_packed);
  MLMCPI_CUDA(cudaMemcpyAsync(h_packed + 3, st->packed, sizeof(double) * (5 + st->k_max),
                              cudaMemcpyDeviceToHost, ctx->stream));
  MLMCPI_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// common/statistics.cc:29-97 with "ranks" = chains.  packed = {n_chains, long-term samples
// (all chains), short-term samples (all chains), sum avg, sum avg_longterm, sum avg2_longterm,
// sum avg3_longterm, sum avg4_longterm, sum S_0 .. S_{k_max-1}}; every entry is additive
// over GPUs.  out = {average, variance, variance_error, tau_int, error, samples}.
int mlmcpi_stats_finalize(const double *p, int k_max, double out[6]) {
  const double n_chains = p[0];
  const double n_tot = p[1]; // mpi_allreduce_sum(n_samples_longterm)
  const double n_short = p[2]; // mpi_allreduce_sum(n_samples)
  if (n_chains < 1)
    return MLMCPI_EINVAL;
  const double avg_short = p[3] / n_chains;   // mpi_allreduce_avg(avg)
  const double avg = p[4] / n_chains;         // mpi_allreduce_avg(avg_longterm)
  const double avg2 = p[5] / n_chains, avg3 = p[6] / n_chains, avg4 = p[7] / n_chains;
  const double S0 = p[8] / n_chains;
  const double variance = n_tot / (n_tot - 1.0) * (S0 - avg * avg);
  const double variance_error =
      std::sqrt(1.0 / n_tot *
                (avg4 - 4 * avg * avg3 + 8 * avg * avg * avg2 - avg2 * avg2 - 4 * avg * avg * avg * avg));
  const double C0 = S0 - avg * avg;
  double tau = 0.0;
  for (int k = 1; k < k_max; ++k)
    tau += (1. - k / n_tot) * (p[8 + k] / n_chains - avg * avg);
  // std::fmax returns the non-NaN argument: tau_int = 1 while there is no variance yet
  const double tau_int = std::fmax(1.0, 1.0 + 2.0 * tau / C0);
  out[0] = avg_short;
  out[1] = variance;
  out[2] = variance_error;
  out[3] = tau_int;
  out[4] = std::sqrt(tau_int * variance / n_short);
  out[5] = n_short;
  return 0;
}

} // extern "C"
