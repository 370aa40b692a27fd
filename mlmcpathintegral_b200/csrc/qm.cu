// qm.cu -- 1-D path models (harmonic oscillator, quartic oscillator, rotor).
//
// One WARP per chain.  A path of M <= 12800 sites is a few KiB, so a whole HMC
// trajectory runs on-chip: x and p live in shared memory for the (nt+1) leapfrog
// steps and HBM sees one read and one write of the path per trajectory
// (SURVEY 7.4 "1-D paths").  Lanes stride over sites (site = lane + 32 k) so that
// global and shared accesses are conflict-free; periodic neighbours come from
// shared memory.  State layout [chain][site] = the reference's SampleState::data.
//
// Reference citations relative to /root/reference/src.
#include <vector>

#include "common.cuh"

namespace {

struct QM {
  int model, M;
  double a, T, m0, mu2, lambda, x0;
};

QM make_qm(const mlmcpi_model *m) {
  QM q;
  q.model = m->model;
  q.M = m->M_lat;
  q.a = m->a_lat;
  q.T = m->T_final;
  q.m0 = m->m0;
  q.mu2 = m->mu2;
  q.lambda = m->lambda;
  q.x0 = m->x0;
  return q;
}

constexpr int WARPS = 8;           // chains per block
constexpr int THREADS = WARPS * 32;

// force at one site.  qm/harmonicoscillatoraction.cc:21-35,
// qm/quarticoscillatoraction.cc:31-53, qm/rotoraction.cc:58-79
template <int MODEL>
__device__ __forceinline__ double site_force(const QM &q, double xm, double x, double xp) {
  if (MODEL == MLMCPI_ROTOR) {
    return (q.m0 / q.a) * (sin_force(x - xm) + sin_force(x - xp));
  } else {
    const double tmp_1 = q.m0 / q.a;
    const double tmp_2 = 2. + q.a * q.a * q.mu2;
    double f = tmp_1 * (tmp_2 * x - xm - xp);
    if (MODEL == MLMCPI_QUARTIC) {
      const double xs = x - q.x0;
      f += (q.a * q.lambda) * xs * xs * xs;
    }
    return f;
  }
}

// action density of site j (needs x_{j-1}); the prefactors are applied by the caller.
// qm/harmonicoscillatoraction.cc:8-18, qm/quarticoscillatoraction.cc:7-28,
// qm/rotoraction.cc:9-18
template <int MODEL>
__device__ __forceinline__ double site_action(const QM &q, double xm, double x) {
  const double d = x - xm;
  if (MODEL == MLMCPI_ROTOR)
    return 1. - cos(d);
  const double ainv2 = 1. / (q.a * q.a);
  if (MODEL == MLMCPI_HO)
    return ainv2 * d * d + q.mu2 * x * x;
  const double xs = x - q.x0;
  const double xs2 = xs * xs;
  return q.m0 * (ainv2 * d * d + q.mu2 * x * x) + 0.5 * q.lambda * xs2 * xs2;
}
template <int MODEL> __device__ __forceinline__ double action_prefactor(const QM &q) {
  if (MODEL == MLMCPI_ROTOR)
    return q.m0 / q.a;
  if (MODEL == MLMCPI_HO)
    return 0.5 * q.a * q.m0;
  return 0.5 * q.a;
}

// qm/harmonicoscillatoraction.hh:171-189, qm/quarticoscillatoraction.hh:160-194,
// qm/rotoraction.hh:195-213
template <int MODEL>
__device__ __forceinline__ void W_min_curv(const QM &q, double x_m, double x_p, double &Wmin,
                                           double &Wcurv) {
  if (MODEL == MLMCPI_ROTOR) {
    Wcurv = 2.0 * q.m0 / q.a * fabs(cos(0.5 * (x_p - x_m)));
    double sp, cp, sm, cm;
    sincos(x_p, &sp, &cp);
    sincos(x_m, &sm, &cm);
    Wmin = atan2(sp + sm, cp + cm);
  } else if (MODEL == MLMCPI_HO) {
    Wcurv = (2. / q.a + q.a * q.mu2) * q.m0;
    Wmin = (0.5 / (1. + 0.5 * q.a * q.a * q.mu2)) * (x_m + x_p);
  } else {
    const double xbar = 0.5 * (x_m + x_p);
    Wcurv = (2. / q.a + q.a * q.mu2) * q.m0 + 3. * q.lambda * q.a * (xbar - q.x0) * (xbar - q.x0);
    const double rho = 1. / (1. + 0.5 * q.a * q.a * q.mu2);
    double x = xbar;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double xs = x - q.x0;
      x = rho * (xbar - 0.5 * q.a * q.a * q.lambda / q.m0 * xs * xs * xs);
    }
    Wmin = x;
  }
}

#define WARP_SETUP                                                                                 \
  const int lane = threadIdx.x & 31;                                                               \
  const int w = threadIdx.x >> 5;                                                                  \
  const long long chain = (long long)blockIdx.x * WARPS + w;                                       \
  const bool active = chain < B;                                                                   \
  const long long c_safe = active ? chain : 0;                                                     \
  (void)lane;                                                                                      \
  (void)w;

// ----------------------------------------------------------------- init_state
__global__ void init_state_kernel(QM q, double *x, int B, uint32_t chain0, uint64_t seed,
                                  uint64_t draw) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int pairs = (q.M + 1) / 2;
  if (t >= (long long)B * pairs)
    return;
  const int chain = (int)(t / pairs), k = (int)(t % pairs);
  double v0 = 0.0, v1 = 0.0;
  if (q.model == MLMCPI_ROTOR) { // qm/rotoraction.cc:82-85
    Rng r = rng_init(seed, MLMCPI_STREAM_INIT, draw, chain0 + chain, k);
    v0 = rng_angle2(r, v1);
  } // HO / quartic: zeros, qm/harmonicoscillatoraction.hh:155-158
  double *xc = x + (size_t)chain * q.M;
  xc[2 * k] = v0;
  if (2 * k + 1 < q.M)
    xc[2 * k + 1] = v1;
}

// --------------------------------------------------------------------- action
template <int MODEL> __global__ void action_kernel(QM q, const double *x, int B, double *S) {
  WARP_SETUP
  const double *xc = x + (size_t)c_safe * q.M;
  double acc = 0.0;
  for (int s = lane; s < q.M; s += 32) {
    const double xm = xc[s == 0 ? q.M - 1 : s - 1];
    acc += site_action<MODEL>(q, xm, xc[s]);
  }
  acc = warp_sum(acc);
  if (active && lane == 0)
    S[chain] = action_prefactor<MODEL>(q) * acc;
}

// ---------------------------------------------------------------------- force
template <int MODEL> __global__ void force_kernel(QM q, const double *x, double *f, int B) {
  WARP_SETUP
  if (!active)
    return;
  const double *xc = x + (size_t)chain * q.M;
  double *fc = f + (size_t)chain * q.M;
  for (int s = lane; s < q.M; s += 32) {
    const double xm = xc[s == 0 ? q.M - 1 : s - 1], xp = xc[s == q.M - 1 ? 0 : s + 1];
    fc[s] = site_force<MODEL>(q, xm, xc[s], xp);
  }
}

// ------------------------------------------------- on-chip trajectory (device)
// xs, ps: this warp's shared arrays.  sampler/hmcsampler.cc:31-46
template <int MODEL>
__device__ __forceinline__ void trajectory(const QM &q, int nt, double dt, double *xs, double *ps,
                                           int lane) {
  const int M = q.M;
  for (int k = 0; k <= nt; ++k) {
    const double dt_p = (k == 0 || k == nt) ? 0.5 * dt : dt;
    const double dt_x = (k == nt) ? 0.0 : dt;
    for (int s = lane; s < M; s += 32) {
      const double xm = xs[s == 0 ? M - 1 : s - 1], xp = xs[s == M - 1 ? 0 : s + 1];
      ps[s] -= dt_p * site_force<MODEL>(q, xm, xs[s], xp);
    }
    __syncwarp();
    for (int s = lane; s < M; s += 32)
      xs[s] += dt_x * ps[s];
    __syncwarp();
  }
}

template <int MODEL>
__global__ void leapfrog_kernel(QM q, int nt, double dt, double *x, double *p, int B) {
  extern __shared__ double smem[];
  WARP_SETUP
  double *xs = smem + (size_t)w * 2 * q.M, *ps = xs + q.M;
  double *xc = x + (size_t)c_safe * q.M, *pc = p + (size_t)c_safe * q.M;
  for (int s = lane; s < q.M; s += 32) {
    xs[s] = xc[s];
    ps[s] = pc[s];
  }
  __syncwarp();
  trajectory<MODEL>(q, nt, dt, xs, ps, lane);
  if (active)
    for (int s = lane; s < q.M; s += 32) {
      xc[s] = xs[s];
      pc[s] = ps[s];
    }
}

__global__ void momentum_kernel(QM q, double *p, int B, uint32_t chain0, uint64_t seed,
                                uint64_t draw) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int pairs = (q.M + 1) / 2;
  if (t >= (long long)B * pairs)
    return;
  const int chain = (int)(t / pairs), k = (int)(t % pairs);
  Rng r = rng_init(seed, MLMCPI_STREAM_HMC_MOMENTUM, draw, chain0 + chain, k);
  double z0, z1;
  rng_normal2(r, z0, z1);
  double *pc = p + (size_t)chain * q.M;
  pc[2 * k] = z0;
  if (2 * k + 1 < q.M)
    pc[2 * k + 1] = z1;
}

// HMCSampler::single_step, sampler/hmcsampler.cc:22-69, one warp per chain, all
// on-chip: momentum refresh, kinetic energies, trajectory, both action
// evaluations, accept/reject and the conditional write-back.
template <int MODEL>
__global__ void hmc_step_kernel(QM q, int nt, double dt, double *x, int B, uint32_t chain0,
                                uint64_t seed, uint64_t draw, int32_t *accept_out, double *diag) {
  extern __shared__ double smem[];
  WARP_SETUP
  const int M = q.M;
  double *xs = smem + (size_t)w * 2 * M, *ps = xs + M;
  double *xc = x + (size_t)c_safe * M;
  const uint32_t gchain = chain0 + (uint32_t)c_safe;
  for (int s = lane; s < M; s += 32)
    xs[s] = xc[s];
  for (int k = lane; 2 * k < M; k += 32) {
    Rng r = rng_init(seed, MLMCPI_STREAM_HMC_MOMENTUM, draw, gchain, k);
    double z0, z1;
    rng_normal2(r, z0, z1);
    ps[2 * k] = z0;
    if (2 * k + 1 < M)
      ps[2 * k + 1] = z1;
  }
  __syncwarp();
  double T_cur = 0.0, S_cur = 0.0;
  for (int s = lane; s < M; s += 32) {
    T_cur += ps[s] * ps[s];
    S_cur += site_action<MODEL>(q, xs[s == 0 ? M - 1 : s - 1], xs[s]);
  }
  T_cur = 0.5 * warp_sum(T_cur);
  S_cur = action_prefactor<MODEL>(q) * warp_sum(S_cur);
  trajectory<MODEL>(q, nt, dt, xs, ps, lane);
  double T_trial = 0.0, S_trial = 0.0;
  for (int s = lane; s < M; s += 32) {
    T_trial += ps[s] * ps[s];
    S_trial += site_action<MODEL>(q, xs[s == 0 ? M - 1 : s - 1], xs[s]);
  }
  T_trial = 0.5 * warp_sum(T_trial);
  S_trial = action_prefactor<MODEL>(q) * warp_sum(S_trial);
  const double deltaH = (S_trial - S_cur) + (T_trial - T_cur);
  bool acc = deltaH < 0.0;
  if (!acc) {
    Rng r = rng_init(seed, MLMCPI_STREAM_HMC_ACCEPT, draw, gchain, 0);
    double u0, u1;
    rng_uniform2(r, u0, u1);
    acc = u0 < exp(-deltaH);
  }
  if (!active)
    return;
  if (acc)
    for (int s = lane; s < M; s += 32)
      xc[s] = xs[s];
  if (lane == 0) {
    if (accept_out)
      accept_out[chain] = acc ? 1 : 0;
    if (diag) {
      double *d = diag + 5 * chain;
      d[0] = deltaH;
      d[1] = S_cur;
      d[2] = S_trial;
      d[3] = T_cur;
      d[4] = T_trial;
    }
  }
}

// Register-resident variant for M = 32 * SPL: lane l owns the sites l + 32 r (r < SPL) with x and p
// in registers for the whole trajectory; neighbours arrive by warp shuffles (site s - 1 is lane
// l - 1's register r, or lane 31's register r - 1 for lane 0), so a leapfrog step needs no shared
// memory and no synchronisation.  For the rotor every bond sine sin(x_s - x_{s-1}) is computed once
// and shared with the left neighbour (sin(x - x_p) = -sin(x_p - x)): one sine per site-step instead
// of two.  Same arithmetic per site as the generic kernel.
template <int SPL> struct RegPath {
  double x[SPL], p[SPL];
};
template <int SPL>
__device__ __forceinline__ void left_neighbours(const double (&v)[SPL], int lane, double (&out)[SPL]) {
  double t[SPL];
#pragma unroll
  for (int r = 0; r < SPL; ++r)
    t[r] = __shfl_sync(0xffffffffu, v[r], (lane + 31) & 31);
#pragma unroll
  for (int r = 0; r < SPL; ++r)
    out[r] = (lane == 0) ? t[(r + SPL - 1) % SPL] : t[r];
}
template <int SPL>
__device__ __forceinline__ void right_neighbours(const double (&v)[SPL], int lane, double (&out)[SPL]) {
  double t[SPL];
#pragma unroll
  for (int r = 0; r < SPL; ++r)
    t[r] = __shfl_sync(0xffffffffu, v[r], (lane + 1) & 31);
#pragma unroll
  for (int r = 0; r < SPL; ++r)
    out[r] = (lane == 31) ? t[(r + 1) % SPL] : t[r];
}

template <int MODEL, int SPL>
__device__ __forceinline__ void path_energies(const QM &q, const RegPath<SPL> &s, int lane, double &T, double &S) {
  double xm[SPL];
  left_neighbours<SPL>(s.x, lane, xm);
  double t = 0.0, a = 0.0;
#pragma unroll
  for (int r = 0; r < SPL; ++r) {
    t += s.p[r] * s.p[r];
    a += site_action<MODEL>(q, xm[r], s.x[r]);
  }
  T = 0.5 * warp_sum(t);
  S = action_prefactor<MODEL>(q) * warp_sum(a);
}

// HMCSampler::single_step (sampler/hmcsampler.cc:22-69) on a register-resident path: momentum
// refresh, trajectory, both energies, accept test.  Returns the accept flag (warp uniform); out[5] =
// {deltaH, S_cur, S_trial, T_cur, T_trial}.  s.x holds the TRIAL state on return.
template <int MODEL, int SPL>
__device__ __forceinline__ bool hmc_step_reg_core(const QM &q, int nt, double dt, RegPath<SPL> &s, int lane,
                                                  uint32_t gchain, uint64_t seed, uint64_t draw, double out[5]) {
#pragma unroll
  for (int r = 0; r < SPL; ++r) {
    const int site = lane + 32 * r;
    Rng rg = rng_init(seed, MLMCPI_STREAM_HMC_MOMENTUM, draw, gchain, site >> 1);
    double z0, z1;
    rng_normal2(rg, z0, z1);
    s.p[r] = (site & 1) ? z1 : z0;
  }
  double T_cur, S_cur;
  path_energies<MODEL, SPL>(q, s, lane, T_cur, S_cur);
  const double pref = q.m0 / q.a;
  for (int k = 0; k <= nt; ++k) { // sampler/hmcsampler.cc:31-46
    const double dt_p = (k == 0 || k == nt) ? 0.5 * dt : dt;
    const double dt_x = (k == nt) ? 0.0 : dt;
    double xm[SPL];
    left_neighbours<SPL>(s.x, lane, xm);
    if (MODEL == MLMCPI_ROTOR) {
      double b[SPL], bn[SPL];
#pragma unroll
      for (int r = 0; r < SPL; ++r)
        b[r] = sin_force(s.x[r] - xm[r]);
      right_neighbours<SPL>(b, lane, bn);
#pragma unroll
      for (int r = 0; r < SPL; ++r)
        s.p[r] -= dt_p * (pref * (b[r] - bn[r]));
    } else {
      double xp[SPL];
      right_neighbours<SPL>(s.x, lane, xp);
#pragma unroll
      for (int r = 0; r < SPL; ++r)
        s.p[r] -= dt_p * site_force<MODEL>(q, xm[r], s.x[r], xp[r]);
    }
#pragma unroll
    for (int r = 0; r < SPL; ++r)
      s.x[r] += dt_x * s.p[r];
  }
  double T_trial, S_trial;
  path_energies<MODEL, SPL>(q, s, lane, T_trial, S_trial);
  const double deltaH = (S_trial - S_cur) + (T_trial - T_cur);
  bool acc = deltaH < 0.0;
  if (!acc) {
    Rng r = rng_init(seed, MLMCPI_STREAM_HMC_ACCEPT, draw, gchain, 0);
    double u0, u1;
    rng_uniform2(r, u0, u1);
    acc = u0 < exp(-deltaH);
  }
  out[0] = deltaH;
  out[1] = S_cur;
  out[2] = S_trial;
  out[3] = T_cur;
  out[4] = T_trial;
  return acc;
}

template <int MODEL, int SPL>
__global__ void __launch_bounds__(THREADS)
    hmc_step_reg_kernel(QM q, int nt, double dt, double *x, int B, uint32_t chain0, uint64_t seed,
                        uint64_t draw, int32_t *accept_out, double *diag) {
  WARP_SETUP
  double *xc = x + (size_t)c_safe * q.M;
  const uint32_t gchain = chain0 + (uint32_t)c_safe;
  RegPath<SPL> s;
#pragma unroll
  for (int r = 0; r < SPL; ++r)
    s.x[r] = xc[lane + 32 * r];
  double out[5];
  const bool acc = hmc_step_reg_core<MODEL, SPL>(q, nt, dt, s, lane, gchain, seed, draw, out);
  if (!active)
    return;
  if (acc) {
#pragma unroll
    for (int r = 0; r < SPL; ++r)
      xc[lane + 32 * r] = s.x[r];
  }
  if (lane == 0) {
    if (accept_out)
      accept_out[chain] = acc ? 1 : 0;
    if (diag) {
      double *d = diag + 5 * chain;
#pragma unroll
      for (int k = 0; k < 5; ++k)
        d[k] = out[k];
    }
  }
}

// ---------------------------------------------------------- fused hierarchical draw (1-D paths)
// HierarchicalSampler::draw (sampler/hierarchicalsampler.cc:55-81) with an HMC sampler on the
// coarsest level, for ONE chain per warp entirely on chip: the restriction chain, the HMC step
// (registers), and for every finer level the two-level Metropolis-Hastings step
// (montecarlo/twolevelmetropolisstep.cc:35-97: prolongation + fill-in, S_f / S_cond of the trial
// state, the three action differences, accept, commit) run in one kernel on shared-memory copies
// of the level states; only the states, the accept flag and the two cached actions of the finest
// level touch global memory.  Arithmetic, summation orders and Philox counters are those of the
// single-purpose kernels above, so the draw is identical to the unfused sequence of launches.
#define QMH_MAX_LEVELS 4
struct QMH {
  QM q[QMH_MAX_LEVELS];
  int L;
};

template <int MODEL>
__device__ __forceinline__ double warp_action(const QM &q, const double *xs, int lane) {
  double acc = 0.0;
  for (int s = lane; s < q.M; s += 32)
    acc += site_action<MODEL>(q, xs[s == 0 ? q.M - 1 : s - 1], xs[s]);
  return action_prefactor<MODEL>(q) * warp_sum(acc);
}
template <int MODEL>
__device__ __forceinline__ double warp_cond_action(const QM &q, const double *xs, int lane) {
  const int Mc = q.M / 2;
  double acc = 0.0;
  for (int j = lane; j < Mc; j += 32) {
    const double x_m = xs[2 * j], x_p = xs[j == Mc - 1 ? 0 : 2 * j + 2];
    double x0, curv;
    W_min_curv<MODEL>(q, x_m, x_p, x0, curv);
    const double dx = xs[2 * j + 1] - x0;
    if (MODEL == MLMCPI_ROTOR)
      acc += -log(expsin2_pdf(dx, 2.0 * curv));
    else
      acc += 0.5 * curv * dx * dx - 0.5 * log(curv);
  }
  return warp_sum(acc);
}

template <int MODEL, int SPL>
__global__ void __launch_bounds__(THREADS)
    hierarchical_draw_kernel(QMH h, int nt, double dt, double *x0_all, double *x1_all, double *x2_all,
                             double *x3_all, int B, uint32_t chain0, uint64_t seed, uint64_t draw,
                             double *Sf0, double *Scond0, int cache0_valid, int32_t *accept_out,
                             unsigned long long *counters) {
  extern __shared__ double smem[];
  WARP_SETUP
  const int L = h.L;
  const int M0 = h.q[0].M;
  // per warp: level states (M0 + M0/2 + ... < 2 M0) and the trial state (M0)
  double *lev[QMH_MAX_LEVELS];
  double *base_w = smem + (size_t)w * 3 * M0;
  {
    int off = 0;
    for (int l = 0; l < L; ++l) {
      lev[l] = base_w + off;
      off += h.q[l].M;
    }
  }
  double *trial = base_w + 2 * M0;
  double *glob[QMH_MAX_LEVELS] = {x0_all, x1_all, x2_all, x3_all};
  const uint32_t gchain = chain0 + (uint32_t)c_safe;
  auto level_draw = [&](int level) { return (draw << 12) | ((uint64_t)level << 8); };
  // restriction chain (hierarchicalsampler.cc:57-60; qm/qmaction.cc:18-26: every second site)
  for (int s = lane; s < M0; s += 32)
    lev[0][s] = x0_all[(size_t)c_safe * M0 + s];
  __syncwarp();
  for (int l = 1; l < L; ++l) {
    for (int s = lane; s < h.q[l].M; s += 32)
      lev[l][s] = lev[l - 1][2 * s];
    __syncwarp();
  }
  double S_old[QMH_MAX_LEVELS], S_new[QMH_MAX_LEVELS];
  for (int l = 1; l + 1 < L; ++l)
    S_old[l] = warp_action<MODEL>(h.q[l], lev[l], lane);
  // coarsest level: HMC step in registers
  bool acc;
  {
    const QM &qc = h.q[L - 1];
    RegPath<SPL> s;
#pragma unroll
    for (int r = 0; r < SPL; ++r)
      s.x[r] = lev[L - 1][lane + 32 * r];
    double out[5];
    acc = hmc_step_reg_core<MODEL, SPL>(qc, nt, dt, s, lane, gchain, seed, level_draw(L - 1), out);
    S_old[L - 1] = out[1];
    if (acc) {
#pragma unroll
      for (int r = 0; r < SPL; ++r)
        lev[L - 1][lane + 32 * r] = s.x[r];
    }
    __syncwarp();
    // the unfused path re-evaluates the action of the (possibly unchanged) coarsest state
    S_new[L - 1] = warp_action<MODEL>(qc, lev[L - 1], lane);
  }
  if (active && lane == 0)
    atomicAdd(counters + (L - 1), (unsigned long long)(acc ? 1 : 0));
  for (int l = L - 2; l >= 0; --l) {
    const QM &q = h.q[l];
    const int M = q.M, Mc = M / 2;
    double Sf, Scond;
    if (l > 0) { // TwoLevelMetropolisStep::set_state, twolevelmetropolisstep.cc:92-97
      Sf = S_old[l];
      Scond = warp_cond_action<MODEL>(q, lev[l], lane);
    } else if (cache0_valid) {
      Sf = Sf0[c_safe];
      Scond = Scond0[c_safe];
    } else {
      Sf = warp_action<MODEL>(q, lev[0], lane);
      Scond = warp_cond_action<MODEL>(q, lev[0], lane);
    }
    // theta' = fill(prolong(phi_c)), :40-42 (qm fill_kernel)
    const uint64_t ldraw = level_draw(l);
    for (int j = lane; j < Mc; j += 32) {
      const double x_m = lev[l + 1][j], x_p = lev[l + 1][j == Mc - 1 ? 0 : j + 1];
      trial[2 * j] = x_m;
      double x0, curv;
      W_min_curv<MODEL>(q, x_m, x_p, x0, curv);
      Rng r = rng_init(seed, MLMCPI_STREAM_FILL1, ldraw, gchain, j);
      if (MODEL == MLMCPI_ROTOR) {
        trial[2 * j + 1] = mod_2pi(x0 + expsin2_draw(r, 2. * curv));
      } else {
        double z0, z1;
        rng_normal2(r, z0, z1);
        trial[2 * j + 1] = x0 + z0 * (1. / sqrt(curv));
      }
    }
    __syncwarp();
    const double Sf_prime = warp_action<MODEL>(q, trial, lane);          // :48
    const double Scond_prime = warp_cond_action<MODEL>(q, trial, lane);  // :65-66
    // :55-58: S_c(theta_C) and S_c(phi_c) are the coarser level's action before / after its update
    const double dS_fine = Sf_prime - Sf;
    const double dS_coarse = S_old[l + 1] - S_new[l + 1];
    const double dS_trial = Scond - Scond_prime;
    const double dS = dS_fine + dS_coarse + dS_trial;
    bool a2 = dS < 0.0;
    if (!a2) {
      Rng r = rng_init(seed, MLMCPI_STREAM_TWOLEVEL_ACCEPT, ldraw, gchain, 0);
      double u0, u1;
      rng_uniform2(r, u0, u1);
      a2 = u0 < exp(-dS);
    }
    if (!acc)
      a2 = false; // the cascade already stopped for this chain (hierarchicalsampler.cc:73-74)
    acc = a2;
    if (acc)
      for (int s = lane; s < M; s += 32)
        lev[l][s] = trial[s];
    __syncwarp();
    S_new[l] = acc ? Sf_prime : Sf;
    if (l == 0 && active && lane == 0) {
      Sf0[chain] = S_new[0];
      Scond0[chain] = acc ? Scond_prime : Scond;
    }
    if (active && lane == 0)
      atomicAdd(counters + l, (unsigned long long)(acc ? 1 : 0));
  }
  if (!active)
    return;
  for (int l = 0; l < L; ++l)
    for (int s = lane; s < h.q[l].M; s += 32)
      glob[l][(size_t)chain * h.q[l].M + s] = lev[l][s];
  if (lane == 0)
    accept_out[chain] = acc ? 1 : 0;
}

// --------------------------------------------------------------------- sweeps
// rotor only (qm/rotoraction.cc:21-56); colours: even sites, then odd sites
template <bool HEATBATH>
__global__ void rotor_sweep_kernel(QM q, double *x, int B, uint32_t chain0, uint64_t seed,
                                   uint64_t draw, int reverse) {
  WARP_SETUP
  const int M = q.M;
  double *xc = x + (size_t)c_safe * M;
  for (int pass = 0; pass < 2; ++pass) {
    const int colour = reverse ? 1 - pass : pass;
    if (active)
      for (int s = 2 * lane + colour; s < M; s += 64) {
        const double x_m = xc[s == 0 ? M - 1 : s - 1], x_p = xc[s == M - 1 ? 0 : s + 1];
        double x0, curv;
        W_min_curv<MLMCPI_ROTOR>(q, x_m, x_p, x0, curv);
        if (HEATBATH) {
          Rng r = rng_init(seed, MLMCPI_STREAM_HEATBATH, draw, chain0 + (uint32_t)chain, s);
          xc[s] = mod_2pi(x0 + expsin2_draw(r, 2. * curv));
        } else {
          xc[s] = mod_2pi(2.0 * x0 - xc[s]);
        }
      }
    __syncwarp();
  }
}

// Action::heatbath_update / overrelaxation_update of ONE site (qm/rotoraction.cc:21-56): the per-dof
// interface of action/action.hh:85-110, one thread per chain
template <bool HEATBATH>
__global__ void rotor_dof_update_kernel(QM q, int s, double *x, int B, uint32_t chain0, uint64_t seed, uint64_t draw) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= B)
    return;
  const int M = q.M;
  double *xc = x + (size_t)chain * M;
  const double x_m = xc[s == 0 ? M - 1 : s - 1], x_p = xc[s == M - 1 ? 0 : s + 1];
  double x0, curv;
  W_min_curv<MLMCPI_ROTOR>(q, x_m, x_p, x0, curv);
  if (HEATBATH) {
    Rng r = rng_init(seed, MLMCPI_STREAM_HEATBATH, draw, chain0 + (uint32_t)chain, s);
    xc[s] = mod_2pi(x0 + expsin2_draw(r, 2. * curv));
  } else {
    xc[s] = mod_2pi(2.0 * x0 - xc[s]);
  }
}

// ------------------------------------------------------- prolong / restrict
// qm/qmaction.cc:7-26
__global__ void prolong_kernel(int M, const double *xc, double *x, int B) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Mc = M / 2;
  if (t >= (long long)B * Mc)
    return;
  const long long chain = t / Mc;
  const int j = (int)(t % Mc);
  x[chain * M + 2 * j] = xc[t];
}
__global__ void restrict_kernel(int M, const double *xf, double *xc, int B) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Mc = M / 2;
  if (t >= (long long)B * Mc)
    return;
  const long long chain = t / Mc;
  const int j = (int)(t % Mc);
  xc[t] = xf[chain * M + 2 * j];
}

// ---------------------------------------------------------------------- fill
// qm/gaussianconditionedfineaction.cc:7-24, qm/rotorconditionedfineaction.cc:7-24.
// xcoarse != nullptr: prolongation fused in (even sites come from the coarse path).
template <int MODEL>
__global__ void fill_kernel(QM q, const double *xcoarse, double *x, int B, uint32_t chain0,
                            uint64_t seed, uint64_t draw) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Mc = q.M / 2;
  if (t >= (long long)B * Mc)
    return;
  const long long chain = t / Mc;
  const int j = (int)(t % Mc);
  double *xc = x + chain * q.M;
  double x_m, x_p;
  if (xcoarse) {
    const double *cc = xcoarse + chain * Mc;
    x_m = cc[j];
    x_p = cc[j == Mc - 1 ? 0 : j + 1];
    xc[2 * j] = x_m;
  } else {
    x_m = xc[2 * j];
    x_p = xc[j == Mc - 1 ? 0 : 2 * j + 2];
  }
  double x0, curv;
  W_min_curv<MODEL>(q, x_m, x_p, x0, curv);
  Rng r = rng_init(seed, MLMCPI_STREAM_FILL1, draw, chain0 + (uint32_t)chain, j);
  if (MODEL == MLMCPI_ROTOR) {
    xc[2 * j + 1] = mod_2pi(x0 + expsin2_draw(r, 2. * curv));
  } else {
    double z0, z1;
    rng_normal2(r, z0, z1);
    xc[2 * j + 1] = x0 + z0 * (1. / sqrt(curv));
  }
}

// --------------------------------------------------------------- cond_action
// qm/gaussianconditionedfineaction.cc:27-43, qm/rotorconditionedfineaction.cc:27-43
template <int MODEL> __global__ void cond_action_kernel(QM q, const double *x, int B, double *S) {
  WARP_SETUP
  const int M = q.M, Mc = q.M / 2;
  const double *xc = x + (size_t)c_safe * M;
  double acc = 0.0;
  for (int j = lane; j < Mc; j += 32) {
    const double x_m = xc[2 * j], x_p = xc[j == Mc - 1 ? 0 : 2 * j + 2];
    double x0, curv;
    W_min_curv<MODEL>(q, x_m, x_p, x0, curv);
    const double dx = xc[2 * j + 1] - x0;
    if (MODEL == MLMCPI_ROTOR)
      acc += -log(expsin2_pdf(dx, 2.0 * curv));
    else
      acc += 0.5 * curv * dx * dx - 0.5 * log(curv);
  }
  acc = warp_sum(acc);
  if (active && lane == 0)
    S[chain] = acc;
}

// ---------------------------------------------------------------------- QoIs
// qoi/qm/qoixsquared.cc:7-20, qoi/qm/qoisusceptibility.cc:7-23
__global__ void qoi_kernel(QM q, int qoi, const double *x, int B, double *out, int64_t *Qint) {
  WARP_SETUP
  const int M = q.M;
  const double *xc = x + (size_t)c_safe * M;
  double acc = 0.0, nwind = 0.0, dsum = 0.0;
  for (int s = lane; s < M; s += 32) {
    if (qoi == MLMCPI_QOI_X2) {
      acc += xc[s] * xc[s];
    } else {
      const double dx = xc[s] - xc[s == 0 ? M - 1 : s - 1];
      acc += mod_2pi(dx);
      nwind += winding(dx);
      dsum += dx;
    }
  }
  acc = warp_sum(acc);
  nwind = warp_sum(nwind);
  if (active && lane == 0) {
    if (qoi == MLMCPI_QOI_X2) {
      out[chain] = acc / M;
    } else {
      const double four_pi2_inv = 0.25 / (M_PI * M_PI);
      out[chain] = four_pi2_inv * (acc * acc) / q.T;
      // sum_j dx_j = 0 exactly in exact arithmetic, so Q / 2 pi = - sum of windings
      if (Qint)
        Qint[chain] = -(int64_t)llrint(nwind);
    }
  }
  (void)dsum;
}

// ------------------------------------------------------------ Wolff cluster
// ClusterSampler::single_cluster_update1d (sampler/clustersampler.cc:88-132) with the rotor's
// S_ell / new_reflection / flip (qm/rotoraction.hh:226-253).
//
// Variates: Philox stream CLUSTER, draw = update number; index 0 = (xbar, start site); index
// 1 + k = the two uniforms of link k (between the sites k and k+1 mod M): the first decides the
// link when the cluster grows FORWARD over it, the second when it grows BACKWARD.  The bond test
// of a link therefore depends only on the link, not on how many links were processed before it,
// and the growth of a cluster -- two runs of bonded links away from the start site, SURVEY 8f-2 --
// is evaluated speculatively: one WARP per chain tests 64 forward and 64 backward links per
// round (the flipped value of the near site is computed on the fly, so the arithmetic of every
// test is that of the sequential walk) and a ballot finds the end of each run.  Expected run
// length is O(m0/a) sites, i.e. one to a few rounds.  If the two runs would meet (the cluster
// wraps the ring: short chains only) the update is walked link by link by lane 0 with the same
// variates, statement for statement as the reference does, including its double flips.
__device__ __forceinline__ double cluster_flip(const double xbar, const double x) {
  return mod_2pi(M_PI + 2. * xbar - x); // RotorAction::flip, rotoraction.hh:245-253
}
__device__ __forceinline__ bool cluster_bond(const double coupling, const double xbar, const double x_near,
                                             const double x_far, const double u) {
  const double Sell = -coupling * cos(x_near - xbar) * cos(x_far - xbar); // S_ell, rotoraction.hh:226-231
  const double p_connect = 1. - exp(fmin(0.0, -Sell));                    // clustersampler.cc:124-126
  return u < p_connect;
}
__device__ __forceinline__ double cluster_link_uniform(uint64_t seed, uint64_t update, uint32_t gchain, int link,
                                                       int backward) {
  Rng r = rng_init(seed, MLMCPI_STREAM_CLUSTER, update, gchain, 1u + (uint32_t)link);
  double uf, ub;
  rng_uniform2(r, uf, ub);
  return backward ? ub : uf;
}
// the reference's walk, one link at a time (lane 0 only; wrap-around case)
__device__ void cluster_update_sequential(double *xc, int M, double coupling, double xbar, int i0, uint64_t seed,
                                          uint64_t update, uint32_t gchain) {
  auto process = [&](int i, int direction, int &i_next) { // process_link1d (:118-132)
    const int nb = (i + direction + M) % M;
    const int link = direction > 0 ? i : nb;
    const bool bonded = cluster_bond(coupling, xbar, xc[i], xc[nb],
                                     cluster_link_uniform(seed, update, gchain, link, direction < 0));
    if (bonded)
      xc[nb] = cluster_flip(xbar, xc[nb]);
    i_next = nb;
    return bonded;
  };
  xc[i0] = cluster_flip(xbar, xc[i0]); // flip(i0)
  int i_p = i0, i_last_p;
  bool bonded;
  do { // forward (:96-103)
    i_last_p = i_p;
    bonded = process(i_p, +1, i_p);
  } while ((i_p != i0) && bonded);
  int i_m = i0;
  do { // backward (:105-110)
    bonded = process(i_m, -1, i_m);
  } while ((i_m != i_last_p) && bonded);
}

constexpr int CLUSTER_LPL = 2; // links per lane, direction and round

__global__ void __launch_bounds__(128) rotor_cluster_kernel(QM q, double *x, int B, uint32_t chain0, uint64_t seed,
                                                             uint64_t update0, int n_updates) {
  const int lane = threadIdx.x & 31;
  const int chain = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (chain >= B) // whole warps leave together
    return;
  const int M = q.M;
  double *xc = x + (size_t)chain * M;
  const double coupling = 2.0 * q.m0 / q.a;
  const uint32_t gchain = chain0 + (uint32_t)chain;
  constexpr int W = 32 * CLUSTER_LPL;
  for (int u = 0; u < n_updates; ++u) {
    const uint64_t update = update0 + u;
    Rng r = rng_init(seed, MLMCPI_STREAM_CLUSTER, update, gchain, 0);
    double u0, u1;
    rng_uniform2(r, u0, u1);
    const double xbar = -M_PI + 2. * M_PI * u0; // new_reflection()
    int i0 = (int)(u1 * M);
    if (i0 >= M)
      i0 = M - 1;
    // t_f = first forward iteration whose link does not bond (iteration t: from site i0+t to
    // i0+t+1), s_b = the same backwards (iteration s: from i0-s to i0-s-1).  Iterations up to
    // M-2 see an unflipped far site as long as the runs do not meet.
    int t_f = -1, s_b = -1;
    for (int base = 0; base < M - 1 && (t_f < 0 || s_b < 0); base += W) {
#pragma unroll
      for (int k = 0; k < CLUSTER_LPL; ++k) {
        const int it = base + 32 * k + lane;
        bool fail_f = false, fail_b = false;
        if (it < M - 1) {
          if (t_f < 0) {
            int a = i0 + it;
            a -= (a >= M) ? M : 0;
            const int b = (a + 1 == M) ? 0 : a + 1;
            fail_f = !cluster_bond(coupling, xbar, cluster_flip(xbar, xc[a]), xc[b],
                                   cluster_link_uniform(seed, update, gchain, a, 0));
          }
          if (s_b < 0) {
            int a = i0 - it;
            a += (a < 0) ? M : 0;
            const int b = (a == 0) ? M - 1 : a - 1;
            fail_b = !cluster_bond(coupling, xbar, cluster_flip(xbar, xc[a]), xc[b],
                                   cluster_link_uniform(seed, update, gchain, b, 1));
          }
        }
        const unsigned mf = __ballot_sync(0xffffffffu, fail_f), mb = __ballot_sync(0xffffffffu, fail_b);
        if (t_f < 0 && mf)
          t_f = base + 32 * k + __ffs(mf) - 1;
        if (s_b < 0 && mb)
          s_b = base + 32 * k + __ffs(mb) - 1;
      }
    }
    if (t_f < 0 || s_b < 0 || t_f + s_b > M - 2) {
      // the runs meet: far sites were flipped before their link is tested -- walk it
      if (lane == 0)
        cluster_update_sequential(xc, M, coupling, xbar, i0, seed, update, gchain);
    } else {
      // the cluster = the ring segment [i0 - s_b, i0 + t_f], every site flipped once
      const int n_flip = 1 + t_f + s_b;
      for (int k = lane; k < n_flip; k += 32) {
        int site = i0 - s_b + k;
        site += (site < 0) ? M : 0;
        site -= (site >= M) ? M : 0;
        xc[site] = cluster_flip(xbar, xc[site]);
      }
    }
    __syncwarp();
  }
}

size_t traj_smem(const QM &q) { return (size_t)WARPS * 2 * q.M * sizeof(double); }

template <typename K> int prepare_smem(mlmcpi_ctx *ctx, K kernel, size_t bytes) {
  if (bytes > 227 * 1024)
    return ctx_fail(ctx, MLMCPI_EUNSUPPORTED, "1-D path too long for the on-chip trajectory kernel");
  if (bytes > 48 * 1024)
    MLMCPI_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

} // namespace

#define QM_DISPATCH(MODEL_VAR, ...)                                                              \
  switch (MODEL_VAR) {                                                                             \
  case MLMCPI_HO: {                                                                                \
    constexpr int MODEL = MLMCPI_HO;                                                               \
    __VA_ARGS__;                                                                                          \
  } break;                                                                                         \
  case MLMCPI_QUARTIC: {                                                                           \
    constexpr int MODEL = MLMCPI_QUARTIC;                                                          \
    __VA_ARGS__;                                                                                          \
  } break;                                                                                         \
  default: {                                                                                       \
    constexpr int MODEL = MLMCPI_ROTOR;                                                            \
    __VA_ARGS__;                                                                                          \
  } break;                                                                                         \
  }

namespace qm {

int init_state(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, uint32_t chain0,
               uint64_t draw) {
  QM q = make_qm(m);
  const long long n = (long long)B * ((q.M + 1) / 2);
  init_state_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(q, x, B, chain0, ctx->seed, draw);
  MLMCPI_LAUNCHED("qm::init_state");
  return 0;
}

int action(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *x, int B, double *S) {
  QM q = make_qm(m);
  QM_DISPATCH(q.model, (action_kernel<MODEL><<<cdiv(B, WARPS), THREADS, 0, ctx->stream>>>(q, x, B, S)));
  MLMCPI_LAUNCHED("qm::action");
  return 0;
}

int force(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *x, double *f, int B) {
  QM q = make_qm(m);
  QM_DISPATCH(q.model, (force_kernel<MODEL><<<cdiv(B, WARPS), THREADS, 0, ctx->stream>>>(q, x, f, B)));
  MLMCPI_LAUNCHED("qm::force");
  return 0;
}

int leapfrog(mlmcpi_ctx *ctx, const mlmcpi_model *m, int nt, double dt, double *x, double *p, int B) {
  QM q = make_qm(m);
  const size_t smem = traj_smem(q);
  QM_DISPATCH(q.model, {
    int rc = prepare_smem(ctx, leapfrog_kernel<MODEL>, smem);
    if (rc)
      return rc;
    leapfrog_kernel<MODEL><<<cdiv(B, WARPS), THREADS, smem, ctx->stream>>>(q, nt, dt, x, p, B);
  });
  MLMCPI_LAUNCHED("qm::leapfrog");
  return 0;
}

int hmc_momentum(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *p, int B, uint32_t chain0,
                 uint64_t draw) {
  QM q = make_qm(m);
  const long long n = (long long)B * ((q.M + 1) / 2);
  momentum_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(q, p, B, chain0, ctx->seed, draw);
  MLMCPI_LAUNCHED("qm::hmc_momentum");
  return 0;
}

int hmc_step(mlmcpi_ctx *ctx, const mlmcpi_model *m, int nt, double dt, double *x, int B,
             uint32_t chain0, uint64_t draw, int32_t *accept, double *diag) {
  QM q = make_qm(m);
  // register-resident trajectory when the path is 32 * {1, 2, 4, 8} sites long
  const int spl = (q.M % 32 == 0) ? q.M / 32 : 0;
  if (spl == 1 || spl == 2 || spl == 4 || spl == 8) {
#define MLMCPI_REG_LAUNCH(SPL)                                                                     \
  QM_DISPATCH(q.model, (hmc_step_reg_kernel<MODEL, SPL><<<cdiv(B, WARPS), THREADS, 0, ctx->stream>>>( \
                           q, nt, dt, x, B, chain0, ctx->seed, draw, accept, diag)))
    if (spl == 1) {
      MLMCPI_REG_LAUNCH(1);
    } else if (spl == 2) {
      MLMCPI_REG_LAUNCH(2);
    } else if (spl == 4) {
      MLMCPI_REG_LAUNCH(4);
    } else {
      MLMCPI_REG_LAUNCH(8);
    }
#undef MLMCPI_REG_LAUNCH
    MLMCPI_LAUNCHED("qm::hmc_step_reg");
    return 0;
  }
  const size_t smem = traj_smem(q);
  QM_DISPATCH(q.model, {
    int rc = prepare_smem(ctx, hmc_step_kernel<MODEL>, smem);
    if (rc)
      return rc;
    hmc_step_kernel<MODEL><<<cdiv(B, WARPS), THREADS, smem, ctx->stream>>>(
        q, nt, dt, x, B, chain0, ctx->seed, draw, accept, diag);
  });
  MLMCPI_LAUNCHED("qm::hmc_step");
  return 0;
}

int overrelax_sweep(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B) {
  // Action::overrelaxation_update is only defined for the rotor among the QM
  // actions (action/action.hh:84-92 default raises)
  if (m->model != MLMCPI_ROTOR)
    return ctx_fail(ctx, MLMCPI_EUNSUPPORTED, "overrelaxation not defined for this action");
  if (m->M_lat % 2)
    return ctx_fail(ctx, MLMCPI_EINVAL, "coloured sweeps need an even number of sites");
  QM q = make_qm(m);
  rotor_sweep_kernel<false><<<cdiv(B, WARPS), THREADS, 0, ctx->stream>>>(q, x, B, 0, 0, 0, ctx->sweep_reverse);
  MLMCPI_LAUNCHED("qm::overrelax_sweep");
  return 0;
}

int dof_update(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, int ell, int heatbath, uint32_t chain0,
               uint64_t draw) {
  if (m->model != MLMCPI_ROTOR)
    return ctx_fail(ctx, MLMCPI_EUNSUPPORTED, "heat bath / overrelaxation not defined for this action");
  QM q = make_qm(m);
  if (ell < 0 || ell >= q.M)
    return ctx_fail(ctx, MLMCPI_EINVAL, "site index out of range");
  if (heatbath)
    rotor_dof_update_kernel<true><<<cdiv(B, 128), 128, 0, ctx->stream>>>(q, ell, x, B, chain0, ctx->seed, draw);
  else
    rotor_dof_update_kernel<false><<<cdiv(B, 128), 128, 0, ctx->stream>>>(q, ell, x, B, 0, 0, 0);
  MLMCPI_LAUNCHED("qm::dof_update");
  return 0;
}

int heatbath_sweep(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, uint32_t chain0,
                   uint64_t draw) {
  if (m->model != MLMCPI_ROTOR)
    return ctx_fail(ctx, MLMCPI_EUNSUPPORTED, "heat bath not defined for this action");
  if (m->M_lat % 2)
    return ctx_fail(ctx, MLMCPI_EINVAL, "coloured sweeps need an even number of sites");
  QM q = make_qm(m);
  rotor_sweep_kernel<true><<<cdiv(B, WARPS), THREADS, 0, ctx->stream>>>(q, x, B, chain0, ctx->seed, draw,
                                                                     ctx->sweep_reverse);
  MLMCPI_LAUNCHED("qm::heatbath_sweep");
  return 0;
}

int prolong(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *xc, double *x, int B) {
  if (m->M_lat % 2)
    return ctx_fail(ctx, MLMCPI_EINVAL, "M_lat must be even");
  const long long n = (long long)B * (m->M_lat / 2);
  prolong_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(m->M_lat, xc, x, B);
  MLMCPI_LAUNCHED("qm::prolong");
  return 0;
}

int restrict_(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *xf, double *xc, int B) {
  if (m->M_lat % 2)
    return ctx_fail(ctx, MLMCPI_EINVAL, "M_lat must be even");
  const long long n = (long long)B * (m->M_lat / 2);
  restrict_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(m->M_lat, xf, xc, B);
  MLMCPI_LAUNCHED("qm::restrict");
  return 0;
}

int prolong_fill(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *xc, double *x, int B,
                 uint32_t chain0, uint64_t draw) {
  if (m->M_lat % 2)
    return ctx_fail(ctx, MLMCPI_EINVAL, "M_lat must be even");
  QM q = make_qm(m);
  const long long n = (long long)B * (q.M / 2);
  QM_DISPATCH(q.model, (fill_kernel<MODEL><<<cdiv(n, 128), 128, 0, ctx->stream>>>(
                           q, xc, x, B, chain0, ctx->seed, draw)));
  MLMCPI_LAUNCHED("qm::fill");
  return 0;
}

int fill(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, uint32_t chain0, uint64_t draw) {
  return prolong_fill(ctx, m, nullptr, x, B, chain0, draw);
}

int cond_action(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *x, int B, double *S) {
  if (m->M_lat % 2)
    return ctx_fail(ctx, MLMCPI_EINVAL, "M_lat must be even");
  QM q = make_qm(m);
  QM_DISPATCH(q.model,
              (cond_action_kernel<MODEL><<<cdiv(B, WARPS), THREADS, 0, ctx->stream>>>(q, x, B, S)));
  MLMCPI_LAUNCHED("qm::cond_action");
  return 0;
}

// prolongation + fill-in followed by S(theta') and S_cond(theta') into S_out[2][B]
int prolong_fill_eval(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *xc, double *x, int B,
                      uint32_t chain0, uint64_t draw, double *S_out) {
  int rc = prolong_fill(ctx, m, xc, x, B, chain0, draw);
  if (rc)
    return rc;
  if ((rc = action(ctx, m, x, B, S_out)))
    return rc;
  return cond_action(ctx, m, x, B, S_out + B);
}

// fused HierarchicalSampler::draw for 1-D paths (HMC coarse sampler); returns 1 if the shape is
// not covered (the caller then uses the sequence of single-purpose kernels)
int hierarchical_draw(mlmcpi_ctx *ctx, const mlmcpi_model *models, int L, int nt, double dt, double *const *states,
                      int B, uint32_t chain0, uint64_t draw, double *Sf0, double *Scond0, bool cache0_valid,
                      int32_t *accept, unsigned long long *counters) {
  if (L < 2 || L > QMH_MAX_LEVELS)
    return 1;
  const int Mc = models[L - 1].M_lat;
  const int spl = (Mc % 32 == 0) ? Mc / 32 : 0;
  if (!(spl == 1 || spl == 2 || spl == 4 || spl == 8))
    return 1;
  QMH h;
  h.L = L;
  for (int l = 0; l < L; ++l) {
    h.q[l] = make_qm(&models[l]);
    if (l > 0 && models[l].M_lat * 2 != models[l - 1].M_lat)
      return 1;
  }
  const size_t smem = (size_t)WARPS * 3 * models[0].M_lat * sizeof(double);
  if (smem > 200 * 1024)
    return 1;
  double *st[QMH_MAX_LEVELS] = {nullptr, nullptr, nullptr, nullptr};
  for (int l = 0; l < L; ++l)
    st[l] = states[l];
#define MLMCPI_HIER_LAUNCH(SPL)                                                                    \
  QM_DISPATCH(h.q[0].model, {                                                                      \
    int rc = prepare_smem(ctx, hierarchical_draw_kernel<MODEL, SPL>, smem);                        \
    if (rc)                                                                                        \
      return rc;                                                                                   \
    hierarchical_draw_kernel<MODEL, SPL><<<cdiv(B, WARPS), THREADS, smem, ctx->stream>>>(          \
        h, nt, dt, st[0], st[1], st[2], st[3], B, chain0, ctx->seed, draw, Sf0, Scond0,            \
        cache0_valid ? 1 : 0, accept, counters);                                                   \
  })
  if (spl == 1) {
    MLMCPI_HIER_LAUNCH(1);
  } else if (spl == 2) {
    MLMCPI_HIER_LAUNCH(2);
  } else if (spl == 4) {
    MLMCPI_HIER_LAUNCH(4);
  } else {
    MLMCPI_HIER_LAUNCH(8);
  }
#undef MLMCPI_HIER_LAUNCH
  MLMCPI_LAUNCHED("qm::hierarchical_draw");
  return 0;
}

int cluster_update(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, uint32_t chain0,
                   uint64_t update0, int n_updates) {
  if (m->model != MLMCPI_ROTOR)
    return ctx_fail(ctx, MLMCPI_EUNSUPPORTED, "cluster updates are defined for the rotor action");
  QM q = make_qm(m);
  rotor_cluster_kernel<<<cdiv((long long)B * 32, 128), 128, 0, ctx->stream>>>(q, x, B, chain0, ctx->seed,
                                                                            update0, n_updates);
  MLMCPI_LAUNCHED("qm::cluster_update");
  return 0;
}

// HarmonicOscillatorAction::build_covariance (qm/harmonicoscillatoraction.cc:38-56) on the host:
// the covariance C = P^{-1} of the cyclic tridiagonal precision matrix P (diagonal
// a m0 mu2 + 2 m0 / a, off-diagonals -m0 / a) and its Cholesky factor C = L L^T, returned
// TRANSPOSED (LT[j][i] = L[i][j]) so that the device reads are coalesced.  (The reference indexes
// the sub-diagonal with `(i - 1) % M_lat` on unsigned integers; the periodic neighbour is meant.)
static bool ho_exact_factor_host(const mlmcpi_model *m, std::vector<double> &LT) {
  const int M = m->M_lat;
  const double d = m->a_lat * m->m0 * m->mu2 + 2.0 * m->m0 / m->a_lat, c = -m->m0 / m->a_lat;
  std::vector<double> P((size_t)M * M, 0.0), C((size_t)M * M, 0.0), L((size_t)M * M, 0.0);
  for (int i = 0; i < M; ++i) {
    P[(size_t)i * M + i] = d;
    P[(size_t)i * M + (i + 1) % M] += c;
    P[(size_t)i * M + (i + M - 1) % M] += c;
    C[(size_t)i * M + i] = 1.0;
  }
  for (int k = 0; k < M; ++k) { // Gauss-Jordan; P is symmetric positive definite
    const double piv = 1.0 / P[(size_t)k * M + k];
    for (int j = 0; j < M; ++j) {
      P[(size_t)k * M + j] *= piv;
      C[(size_t)k * M + j] *= piv;
    }
    for (int i = 0; i < M; ++i) {
      const double f = P[(size_t)i * M + k];
      if (i == k || f == 0.0)
        continue;
      for (int j = 0; j < M; ++j) {
        P[(size_t)i * M + j] -= f * P[(size_t)k * M + j];
        C[(size_t)i * M + j] -= f * C[(size_t)k * M + j];
      }
    }
  }
  for (int j = 0; j < M; ++j) {
    double s = C[(size_t)j * M + j];
    for (int k = 0; k < j; ++k)
      s -= L[(size_t)j * M + k] * L[(size_t)j * M + k];
    if (!(s > 0.0))
      return false;
    const double ljj = std::sqrt(s);
    L[(size_t)j * M + j] = ljj;
    for (int i = j + 1; i < M; ++i) {
      double t = C[(size_t)i * M + j];
      for (int k = 0; k < j; ++k)
        t -= L[(size_t)i * M + k] * L[(size_t)j * M + k];
      L[(size_t)i * M + j] = t / ljj;
    }
  }
  LT.assign((size_t)M * M, 0.0);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j <= i; ++j)
      LT[(size_t)j * M + i] = L[(size_t)i * M + j];
  return true;
}

// HarmonicOscillatorAction::draw (qm/harmonicoscillatoraction.cc:59-66): x = L_cov y, y i.i.d.
// N(0,1).  One block per chain; y is generated into shared memory (one Box-Muller pair per two
// entries), thread i accumulates row i of L_cov against it (LT is shared by all chains: L2 hits).
__global__ void ho_exact_draw_kernel(int M, const double *__restrict__ LT, double *x, uint32_t chain0,
                                     uint64_t seed, uint64_t draw) {
  extern __shared__ double y[];
  const int chain = blockIdx.x;
  for (int k = threadIdx.x; 2 * k < M; k += blockDim.x) {
    Rng r = rng_init(seed, MLMCPI_STREAM_EXACT, draw, chain0 + chain, k);
    double z0, z1;
    rng_normal2(r, z0, z1);
    y[2 * k] = z0;
    if (2 * k + 1 < M)
      y[2 * k + 1] = z1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    double s = 0.0;
    for (int j = 0; j <= i; ++j)
      s += LT[(size_t)j * M + i] * y[j];
    x[(size_t)chain * M + i] = s;
  }
}

int exact_draw(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, uint32_t chain0, uint64_t draw) {
  if (m->model != MLMCPI_HO)
    return ctx_fail(ctx, MLMCPI_EUNSUPPORTED, "the exact sampler is defined for the harmonic oscillator");
  const int M = m->M_lat;
  if (M > 4096)
    return ctx_fail(ctx, MLMCPI_EINVAL, "exact sampler: M_lat too large for the dense Cholesky factor");
  const std::array<double, 4> key = {(double)M, m->a_lat, m->m0, m->mu2};
  auto it = ctx->ho_exact_factor.find(key);
  if (it == ctx->ho_exact_factor.end()) {
    std::vector<double> LT;
    if (!ho_exact_factor_host(m, LT))
      return ctx_fail(ctx, MLMCPI_EINVAL, "exact sampler: covariance matrix is not positive definite");
    double *d = nullptr;
    MLMCPI_CUDA(cudaMalloc((void **)&d, LT.size() * sizeof(double)));
    MLMCPI_CUDA(cudaMemcpyAsync(d, LT.data(), LT.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    MLMCPI_CUDA(cudaStreamSynchronize(ctx->stream)); // LT goes out of scope
    it = ctx->ho_exact_factor.emplace(key, d).first;
  }
  const int threads = std::min(1024, ((M + 31) / 32) * 32);
  ho_exact_draw_kernel<<<B, threads, (size_t)(M + 1) * sizeof(double), ctx->stream>>>(M, it->second, x, chain0,
                                                                                     ctx->seed, draw);
  MLMCPI_LAUNCHED("qm::exact_draw");
  return 0;
}

int qoi(mlmcpi_ctx *ctx, const mlmcpi_model *m, int which, const double *x, int B, double *out,
        int64_t *Qint) {
  if (which != MLMCPI_QOI_X2 && which != MLMCPI_QOI_ROTOR_CHI)
    return ctx_fail(ctx, MLMCPI_EINVAL, "QoI not defined for 1-D paths");
  QM q = make_qm(m);
  qoi_kernel<<<cdiv(B, WARPS), THREADS, 0, ctx->stream>>>(q, which, x, B, out, Qint);
  MLMCPI_LAUNCHED("qm::qoi");
  return 0;
}

} // namespace qm
