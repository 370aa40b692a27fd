// common.cuh -- shared device/host pieces of libmlmcpi.so (sm_100a only).
//
// Context, launch bookkeeping, Philox4x32-10 streams, fp64 helper maths and the
// four fill-in / heat-bath distributions of the reference's distribution/ folder
// as device functions.  Reference citations are relative to /root/reference/src.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <array>
#include <map>
#include <string>
#include <vector>

#include "../../include/mlmcpi.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

// --------------------------------------------------------------------- context
#define MLMCPI_N_WORK 9
struct mlmcpi_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  uint64_t seed = 0;
  int expcos_envelope = 2; // MLMCPI_OPT_EXPCOS_ENVELOPE: 0 reference, 1 chord, 2 chord + Taylor (default)
  int leapfrog_variant = 0; // MLMCPI_OPT_LEAPFROG_VARIANT: 0 TMA row pipeline, 1 register row march, 2 generic
  int leapfrog_rows = 0;    // MLMCPI_OPT_LEAPFROG_ROWS: rows per block (0 = default)
  int leapfrog_fuse = 1;    // MLMCPI_OPT_LEAPFROG_FUSE: leapfrog steps per HBM pass (0: one, 1: auto, 2/3: 2/4 steps, 4: round-1 kernel)
  int sweep_reverse = 0;    // MLMCPI_OPT_SWEEP_REVERSE: colours visited in descending order
  int overrelax_one_pass = 1; // MLMCPI_OPT_OVERRELAX_ONE_PASS: all colours of a Schwinger OR sweep in one HBM pass
  int fused_qm_hierarchy = 1; // MLMCPI_OPT_FUSED_QM_HIERARCHY: 1-D hierarchical draw in one kernel
  int host_copy_engine = 1; // MLMCPI_OPT_HOST_COPY_ENGINE: accepted states to pinned host memory by the copy engine
  int tau_refresh = 16;  // MLMCPI_OPT_TAU_REFRESH: calls a tau_int answer is reused for (host decisions of the level walks)
  int cascade_cache = 1; // MLMCPI_OPT_CASCADE_CACHE: Schwinger / HMC hierarchy without the per-draw restriction chain
  int gff_coarse_smoothing = 1; // MLMCPI_OPT_GFF_COARSE_SMOOTHING: coarse GFF levels carry Q_hat (reference)
  uint64_t launches = 0;
  int n_sm = 148;
  // sum over the processes of a run (mlmcpi_set_allreduce); nullptr: single process
  int (*allreduce)(void *user, double *d_buf, size_t n) = nullptr;
  void *allreduce_user = nullptr;
  int world = 1, rank = 0;
  double *reduce_buf = nullptr; // device staging of ctx_allreduce_host
  size_t reduce_buf_n = 0;
  std::string err;
  // optional timing of the dominant kernel (leapfrog) with CUDA events on ctx->stream
  bool profile = false;
  std::vector<cudaEvent_t> prof_events; // (begin, end) pairs not yet read
  double prof_ms = 0.0, prof_bytes = 0.0;
  uint64_t prof_launches = 0;
  double *scratch = nullptr; // reduction partials / work vectors
  size_t scratch_n = 0;
  double *work[MLMCPI_N_WORK] = {}; // 0-3: HMC work states, 4-6: two-level step
  size_t work_n[MLMCPI_N_WORK] = {};
  // exact sampler of the harmonic oscillator: transposed Cholesky factors of the covariance,
  // device [M][M], keyed by (M, a, m0, mu2)
  std::map<std::array<double, 4>, double *> ho_exact_factor;
  // GFF dense matrices (gffaction.cc:133-174), device [N][N] each, keyed by
  // (Mt, Mx, rotated, mu2, n_gibbs, omega): {Q_hat, transposed inverse of the Cholesky factor U}
  std::map<std::array<double, 6>, std::array<double *, 2>> gff_dense;
  // cuBLAS / cuSOLVER handles of the dense GFF set-up (created on first use; gff.cu)
  void *cublas = nullptr, *cusolver = nullptr;
};

int ctx_fail(mlmcpi_ctx *ctx, int code, const char *what, const char *detail = nullptr);
int ctx_check_launch(mlmcpi_ctx *ctx, const char *what);
double *ctx_scratch(mlmcpi_ctx *ctx, size_t n);          // >= n doubles, nullptr on failure
double *ctx_work(mlmcpi_ctx *ctx, int which, size_t n);  // persistent work vector
// in-place sum over all processes of n host doubles (no-op without a hook); synchronises
int ctx_allreduce_host(mlmcpi_ctx *ctx, double *h, size_t n);
void prof_begin(mlmcpi_ctx *ctx);
void prof_end(mlmcpi_ctx *ctx, uint64_t launches, double algorithmic_bytes);

#define MLMCPI_CUDA(call)                                                      \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess)                                                    \
      return ctx_fail(ctx, MLMCPI_ECUDA, #call, cudaGetErrorString(e__));      \
  } while (0)

#define MLMCPI_LAUNCHED(what)                                                  \
  do {                                                                         \
    int rc__ = ctx_check_launch(ctx, what);                                    \
    if (rc__)                                                                  \
      return rc__;                                                             \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------- RNG
// Philox4x32-10 (Salmon et al. SC'11).  Stream convention: see include/mlmcpi.h.
struct Rng {
  uint32_t c0, c1, c2, a, k0, k1;
};

// MLMCPI_PHILOX_NOINLINE (set per translation unit): one shared copy of the ten rounds instead of one per call
// site.  The Schwinger fill-in kernels run through ~2500 of their 4096 instructions ONCE per thread and stall on
// instruction fetch ("no_instructions" is their top stall reason); a called Philox shrinks the footprint.
#ifdef MLMCPI_PHILOX_NOINLINE
#define MLMCPI_PHILOX_INLINE __noinline__
#else
#define MLMCPI_PHILOX_INLINE __forceinline__
#endif
__device__ MLMCPI_PHILOX_INLINE void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

__device__ __forceinline__ Rng rng_init(uint64_t seed, int stream, uint64_t draw, uint32_t chain,
                                        uint32_t index) {
  Rng r;
  r.c0 = index;
  r.c1 = chain;
  r.c2 = (uint32_t)draw;
  r.a = ((uint32_t)stream) << 24;
  r.k0 = (uint32_t)seed;
  r.k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(draw >> 32);
  return r;
}

__device__ __forceinline__ void rng_uniform2(Rng &r, double &u0, double &u1) {
  uint32_t o[4];
  philox4x32_10(r.c0, r.c1, r.c2, r.a, r.k0, r.k1, o);
  r.a += 1;
  u0 = ((double)(o[0] >> 5) * 67108864.0 + (double)(o[1] >> 6)) * (1.0 / 9007199254740992.0);
  u1 = ((double)(o[2] >> 5) * 67108864.0 + (double)(o[3] >> 6)) * (1.0 / 9007199254740992.0);
}

// Box-Muller: two uniforms in [0, 1) -> two standard normals
__device__ __forceinline__ void box_muller(const double u0, const double u1, double &z0, double &z1) {
  double s, c;
  const double rad = sqrt(-2.0 * log(1.0 - u0));
  sincospi(2.0 * u1, &s, &c);
  z0 = rad * c;
  z1 = rad * s;
}

__device__ __forceinline__ void rng_normal2(Rng &r, double &z0, double &z1) {
  double u0, u1;
  rng_uniform2(r, u0, u1);
  box_muller(u0, u1, z0, z1);
}

__device__ __forceinline__ double rng_angle2(Rng &r, double &second) {
  double u0, u1;
  rng_uniform2(r, u0, u1);
  second = -M_PI + 2. * M_PI * u1;
  return -M_PI + 2. * M_PI * u0;
}

// ----------------------------------------------------------------------- maths
// common/auxilliary.hh:42-44 (same operation order; no FMA-contractible pattern
// feeds the floor, so the branch cut is the reference's)
// u / pi, correctly rounded -- the IEEE quotient, bit for bit -- without the division sequence
// (reciprocal seed, Newton steps and special-case branches: ~20 instructions, and this sits inside
// every compact update): q = RN(u * RN(1/pi)) followed by two residual corrections r = fma(-q, pi, u),
// q += r * RN(1/pi).  Once q is within an ulp the next correction is the correctly rounded quotient
// (Markstein 1990; pi's significand is not all ones, |u| stays far from the over/underflow range).
// scratch check: 350 000 arguments incl. ulp-neighbours of multiples of pi, no difference to u / pi
// already after one correction.
__device__ __forceinline__ double div_pi(const double u) {
  const double c = 0x1.45f306dc9c883p-2; // RN(1 / pi)
  double q = u * c;
  double r = fma(-q, M_PI, u);
  q = fma(r, c, q);
  r = fma(-q, M_PI, u);
  return fma(r, c, q);
}
// sin(x) for the force kernels: straight-line code -- no branch, no special cases, no coefficient-table
// loads, no float <-> int conversion instructions.  (The library routine checks for special values, reduces
// modulo pi/2, branches to a slow path for huge arguments and fetches one of two coefficient sets: ~40
// instructions, and this is the one transcendental of a leapfrog site-step.  A branch inside the routine also
// makes every call its own basic block, which keeps the compiler from interleaving the independent sines of a
// multi-stage leapfrog iteration: ncu showed 12 warps per SM each crawling through one dependent chain.)
//   k = rint(x / pi) by the magic-number addition (the FMA rounds x * RN(1/pi) + 1.5 * 2^52 once; the parity
//   of k is the lowest mantissa bit), r = x - k pi with a two-term Cody-Waite split (exact products through
//   FMA: the reduction error is |k| * 2^-107, i.e. negligible for every |x| < 2^50, beyond which the spacing
//   of doubles exceeds 1/4 and a sine is noise anyway), then r + r^3 q(r^2) with the degree-8 minimax-type
//   polynomial q fitted on |r| <= pi/2 (1 + 1e-7) (scratch fit against 60-digit sines: 2.2e-16 absolute),
//   sign by the parity of k (an XOR on the high word).
// The constants live in the constant bank and enter the DFMAs as c[bank][offset] operands; written as
// literals they are re-materialised with two uniform moves each, per sine and loop iteration.
static __constant__ double SIN_FORCE_C[13] = {
    0x1.45f306dc9c883p-2,  // RN(1 / pi)
    0x1.921fb54442d18p+1,  // pi_hi
    0x1.1a62633145c07p-53, // pi_lo
    -0x1.275f311897998p-57, 0x1.9507ff1c3c031p-49, -0x1.ae7ee3a7e2a24p-41, 0x1.612460b6ab110p-33,
    -0x1.ae64567e733b6p-26, 0x1.71de3a556b9b6p-19, -0x1.a01a01a01a00dp-13, 0x1.1111111111111p-7,
    -0x1.5555555555555p-3,
    0x1.8p+52}; // 1.5 * 2^52
__device__ __forceinline__ double sin_force(const double x) {
  const double kd = fma(x, SIN_FORCE_C[0], SIN_FORCE_C[12]);
  const int parity_bit = __double2loint(kd) << 31; // lowest mantissa bit = parity of k -> sign bit
  const double k = kd - SIN_FORCE_C[12];
  double r = fma(-k, SIN_FORCE_C[1], x);
  r = fma(-k, SIN_FORCE_C[2], r);
  const double t = r * r;
  double q = SIN_FORCE_C[3];
  q = fma(q, t, SIN_FORCE_C[4]);
  q = fma(q, t, SIN_FORCE_C[5]);
  q = fma(q, t, SIN_FORCE_C[6]);
  q = fma(q, t, SIN_FORCE_C[7]);
  q = fma(q, t, SIN_FORCE_C[8]);
  q = fma(q, t, SIN_FORCE_C[9]);
  q = fma(q, t, SIN_FORCE_C[10]);
  q = fma(q, t, SIN_FORCE_C[11]);
  const double s = fma(r * t, q, r);
  return __hiloint2double(__double2hiint(s) ^ parity_bit, __double2loint(s));
}

// cos x = 1 - 2 sin^2(x / 2) through the branch-free sine above (relative accuracy of 1 - cos x for small x,
// 4e-16 absolute otherwise): for the stochastic kernels, whose instruction footprint has to stay inside the
// SM's instruction cache (CUDA's cos() is three times the code, plus a slow path that is never taken)
__device__ __forceinline__ double cos_fast(const double x) {
  const double s = sin_force(0.5 * x);
  return fma(-2.0 * s, s, 1.0);
}

// the same for a divisor b known at run time with c = RN(1 / b) (computed once per thread: the callers'
// divisor is a kernel argument): u / b, correctly rounded
__device__ __forceinline__ double div_exact(const double u, const double b, const double c) {
  double q = u * c;
  double r = fma(-q, b, u);
  q = fma(r, c, q);
  r = fma(-q, b, u);
  return fma(r, c, q);
}
__device__ __forceinline__ double mod_2pi(const double x) {
  return x - 2. * M_PI * floor(div_pi(0.5 * (x + M_PI)));
}
// the same map with the division replaced by a multiplication with 1/(2 pi): may pick the
// other representative when x is within an ulp of an odd multiple of pi (equal modulo
// 2 pi).  Used only by the stochastic kernels, whose outputs are angles modulo 2 pi.
__device__ __forceinline__ double mod_2pi_fast(const double x) {
  return x - 2. * M_PI * floor((x + M_PI) * (0.5 / M_PI));
}
// integer winding number n with x = mod_2pi(x) + 2 pi n
__device__ __forceinline__ double winding(const double x) { return floor(div_pi(0.5 * (x + M_PI))); }

// e^{-z} I0(z): stands in for gsl_sf_bessel_I0_scaled
__device__ __forceinline__ double bessel_I0_scaled(const double z) {
  const double az = fabs(z);
  if (az <= 600.0)
    return exp(-az) * cyl_bessel_i0(az);
  // Hankel asymptotic series, DLMF 10.40.1
  double term = 1.0, sum = 1.0;
  for (int k = 1; k < 12; ++k) {
    term *= (2.0 * k - 1.0) * (2.0 * k - 1.0) / (8.0 * k * az);
    sum += term;
  }
  return sum * rsqrt(2.0 * M_PI * az);
}

// common/fastbessel.hh:38-50
__host__ __device__ constexpr double mbc(int n) {
  return n == 0 ? 1.0 : 0.125 * (2.0 * n - 1.0) * (2.0 * n - 1.0) / n * mbc(n - 1);
}
// common/fastbessel.cc:7-49
__device__ __forceinline__ double fast_bessel_I0_scaled(const double z) {
  // a_n = (2n-1)^2 / (8 n) a_{n-1}, evaluated like the reference's constexpr template
  constexpr double a1 = mbc(1), a2 = mbc(2), a3 = mbc(3), a4 = mbc(4), a5 = mbc(5), a6 = mbc(6),
                   a7 = mbc(7);
  if (z > 100.) {
    const double z_inv = 1. / z;
    double p;
    if (z > 1100.) {
      p = a4;
    } else if (z > 400.) {
      p = a5;
      p = z_inv * p + a4;
    } else if (z > 200.) {
      p = a6;
      p = z_inv * p + a5;
      p = z_inv * p + a4;
    } else {
      p = a7;
      p = z_inv * p + a6;
      p = z_inv * p + a5;
      p = z_inv * p + a4;
    }
    p = z_inv * p + a3;
    p = z_inv * p + a2;
    p = z_inv * p + a1;
    p = z_inv * p + 1.0;
    return p * rsqrt(2. * M_PI * z);
  }
  return bessel_I0_scaled(z);
}

// --------------------------------------------------------------- distributions
// Random numbers are consumed in blocks: one rng_normal2 + one rng_uniform2 serve
// two consecutive attempts (z0,u0), (z1,u1) (stream convention: include/mlmcpi.h).

// distribution/expsin2distribution.hh:44-58
__device__ __forceinline__ double expsin2_draw(Rng &r, const double sigma) {
  const double scale = M_PI / sqrt(2. * sigma);
  for (;;) {
    double z0, z1, u0, u1;
    rng_normal2(r, z0, z1);
    rng_uniform2(r, u0, u1);
    double r_x = scale * z0;
    if (fabs(r_x) < M_PI) {
      const double s = sin(0.5 * r_x);
      if (u0 < exp(-sigma * (s * s - r_x * r_x / (M_PI * M_PI))))
        return r_x;
    }
    r_x = scale * z1;
    if (fabs(r_x) < M_PI) {
      const double s = sin(0.5 * r_x);
      if (u1 < exp(-sigma * (s * s - r_x * r_x / (M_PI * M_PI))))
        return r_x;
    }
  }
}

// distribution/expsin2distribution.cc:7-24
__device__ __forceinline__ double expsin2_pdf(const double x, const double sigma) {
  const double s = sin(0.5 * x);
  const double z = 0.5 * sigma;
  double besselI0;
  if (z > 100.) {
    const double z_inv = 1. / z;
    besselI0 = sqrt(2. * M_PI * z_inv) * (1. + 0.125 * z_inv + 0.0703125 * z_inv * z_inv);
  } else {
    besselI0 = 2. * M_PI * bessel_I0_scaled(z);
  }
  return exp(-sigma * s * s) / besselI0;
}

// distribution/expcosdistribution.hh:50-65.  Target pdf ~ exp(tau cos x) on [-pi, pi).
//   envelope == 0: the reference's Gaussian envelope, variance 2 pi^2 / tau, acceptance
//                  exp(tau (cos x - 1 + x^2 / (4 pi^2)))  (about 22 % for large tau)
//   envelope == 1: the chord bound 1 - cos x >= 2 x^2 / pi^2 on [-pi, pi]: variance
//                  pi^2 / (4 tau), acceptance exp(tau (cos x - 1 + 2 x^2 / pi^2)) (2/pi = 64 %
//                  for large tau); for tau < 1/2 a uniform proposal with acceptance
//                  exp(tau (cos x - 1)).  Same target, fewer rejected attempts (SURVEY 8a-a11
//                  allows a tighter envelope as long as the target pdf is unchanged).
//   envelope == 2: (default) as 1, and for tau >= 64 the Taylor bound
//                  1 - cos x >= (x^2 / 2) (1 - x^2 / 12), which on x^2 <= x1^2 = 160 / tau gives the
//                  Gaussian envelope exp(-a x^2 / 2), a = tau (1 - x1^2 / 12) = tau - 40/3:
//                  acceptance sqrt(1 - 40 / (3 tau)) (97 % at tau = 256, 99.7 % at tau = 2048).
//                  Proposals with x^2 > x1^2 are rejected, i.e. the target is truncated to
//                  |x| <= x1, which removes a probability mass < exp(-76) = 1e-33 -- thirty orders
//                  of magnitude below the 2^-53 resolution of the uniform variates.  The squeeze
//                  u <= 1 - tau delta x^2 / 2 <= exp(log acceptance) accepts most proposals
//                  without evaluating cos or exp.
#define EXPCOS_TIGHT_TAU 64.0
// cosine of the ExpCos sampler: cos_fast has fewer instructions but more of them on the fp64 pipe, and the heat-bath
// sweep (fp64-pipe bound) runs 8.07 instead of 6.54 ms with it
#ifndef EXPCOS_COS
#define EXPCOS_COS cos
#endif
// have_first: the first Gaussian attempt uses the variates (z_first, u_first) handed in by the caller instead of
// a block of the stream r -- the fused fill-in draws ONE normal pair and ONE uniform pair per coarse cell and gives
// a half of each to its two horizontal links, whose further attempts (rare: 0.3 % at tau = 2048) continue on
// their own streams.  Without it the sequence of attempts is (z0, u0), (z1, u1) of block 0, then block 1, ...
// refill(z0, z1, u0, u1): the next block of the stream (one normal pair, one uniform pair); a parameter so that the
// fused fill-in kernel can route it through its one shared copy of the generator code
template <class Refill>
__device__ __forceinline__ double expcos_draw_core_f(Rng &r, Refill refill, const double tau, const int envelope,
                                                     const bool have_first = false, const double z_first = 0.0,
                                                     const double u_first = 0.0) {
  double x = 0.0;
  bool accepted = false;
  if (envelope >= 1 && tau < 0.5) {
    while (!accepted) {
      double a0, a1, u0, u1;
      rng_uniform2(r, a0, a1);
      rng_uniform2(r, u0, u1);
      x = -M_PI + 2. * M_PI * a0;
      accepted = (u0 <= exp(tau * (EXPCOS_COS(x) - 1.)));
      if (!accepted) {
        x = -M_PI + 2. * M_PI * a1;
        accepted = (u1 <= exp(tau * (EXPCOS_COS(x) - 1.)));
      }
    }
    return x;
  }
  // Gaussian proposals N(0, sigma^2) restricted to an interval, log acceptance ratio
  // tau (cos x - 1) + q x^2; one loop for the three envelopes (one copy of the code: the
  // fill-in kernels are instruction-cache bound)
  const bool tight = (envelope == 2 && tau >= EXPCOS_TIGHT_TAU);
  double sigma, q, sq;
  if (tight) {
    const double a = tau - 40. / 3.;
    sigma = rsqrt(a);
    q = 0.5 * a;
    sq = 20. / 3.; // squeeze: log acceptance >= -(20/3) x^2
  } else {
    sigma = envelope >= 1 ? 0.5 * M_PI / sqrt(tau) : M_PI * sqrt(2. / tau);
    q = tau * (envelope >= 1 ? 2. / (M_PI * M_PI) : 1. / (4. * M_PI * M_PI));
    sq = 0.0;
  }
  double z0 = 0.0, z1 = 0.0, u0 = 0.0, u1 = 0.0;
  int t = have_first ? -1 : 0; // attempt counter: -1 = the caller's variates, then (z0,u0), (z1,u1) of block t / 2
#pragma unroll 1
  while (!accepted) {
    double z, u;
    if (t < 0) {
      z = z_first;
      u = u_first;
    } else {
      if ((t & 1) == 0)
        refill(z0, z1, u0, u1);
      z = (t & 1) ? z1 : z0;
      u = (t & 1) ? u1 : u0;
    }
    ++t;
    x = sigma * z;
    const double x2 = x * x;
    const bool inside = tight ? (tau * x2 <= 160.) : ((-M_PI <= x) && (x < M_PI)); // x^2 <= 160 / tau
    if (inside) {
      accepted = tight && (u <= 1. - sq * x2);
      if (!accepted)
        accepted = (u <= exp(tau * (EXPCOS_COS(x) - 1.) + q * x2));
    }
  }
  return x;
}

__device__ __forceinline__ double expcos_draw_core(Rng &r, const double tau, const int envelope,
                                                   const bool have_first = false, const double z_first = 0.0,
                                                   const double u_first = 0.0) {
  return expcos_draw_core_f(
      r,
      [&r](double &z0, double &z1, double &u0, double &u1) {
        rng_normal2(r, z0, z1);
        rng_uniform2(r, u0, u1);
      },
      tau, envelope, have_first, z_first, u_first);
}

// by-products of a draw, from which the fused fill-in kernel evaluates -log pdf of the drawn
// value without re-deriving the angles from the stored state: x is the accepted proposal
// (the drawn link is mod_2pi(x + mean)), tau the concentration
struct ExpCosDrawn {
  double x, tau;
};
template <class Refill>
__device__ __forceinline__ double expcos_draw_f(Rng &r, Refill refill, const double beta, const double x_p,
                                                const double x_m, const int envelope, ExpCosDrawn *info,
                                                const bool have_first, const double z_first, const double u_first) {
  const double dx = x_m - x_p;
  const double tau = 2. * beta * fabs(EXPCOS_COS(0.5 * dx));
  const double x = expcos_draw_core_f(r, refill, tau, envelope, have_first, z_first, u_first);
  if (info) {
    info->x = x;
    info->tau = tau;
  }
  return mod_2pi_fast(x + 0.5 * (x_p + x_m) + (fabs(dx) > M_PI) * M_PI);
}
__device__ __forceinline__ double expcos_draw(Rng &r, const double beta, const double x_p,
                                              const double x_m, const int envelope,
                                              ExpCosDrawn *info = nullptr, const bool have_first = false,
                                              const double z_first = 0.0, const double u_first = 0.0) {
  return expcos_draw_f(
      r,
      [&r](double &z0, double &z1, double &u0, double &u1) {
        rng_normal2(r, z0, z1);
        rng_uniform2(r, u0, u1);
      },
      beta, x_p, x_m, envelope, info, have_first, z_first, u_first);
}

// distribution/expcosdistribution.cc:7-21
__device__ __forceinline__ double expcos_pdf(const double beta, const double x, const double x_p,
                                             const double x_m) {
  double dx = x_p - x_m;
  double z = x - x_m;
  int sign_flip = (dx < 0.0) ? -1 : +1;
  dx *= sign_flip;
  if (dx > M_PI) {
    sign_flip *= -1;
    dx = 2. * M_PI - dx;
  }
  z *= sign_flip;
  const double sigma = 2. * beta * fabs(EXPCOS_COS(0.5 * dx));
  const double Z_norm = 2. * M_PI * fast_bessel_I0_scaled(sigma);
  return 1. / Z_norm * exp(sigma * (cos(z - 0.5 * dx) - 1.0));
}

// constants of distribution/besselproductdistribution.hh:52-80, computed on the
// host once per level (besselproduct_setup in capi.cu) and passed by value
struct BesselProductConst {
  double beta;
  double I0_twobeta;
  double log_I0_twobeta;
  double sigma_beta;
  double alphaZ[17];
};

// distribution/besselproductdistribution.hh:82-142
__device__ __forceinline__ double besselproduct_draw(Rng &r, const BesselProductConst &bp,
                                                     const double x_p, const double x_m) {
  const double beta = bp.beta, sigma_beta = bp.sigma_beta, L = bp.log_I0_twobeta;
  double dx = x_m - x_p;
  const double sign_flip = (dx < 0) ? -1 : +1;
  dx *= sign_flip;
  // pow(I0_twobeta, e) = exp(e * log I0_twobeta)
  const double N_p = erf((M_PI - 0.5 * dx) / sigma_beta);
  const double N_m = erf(0.5 * dx / sigma_beta) * exp(L * (2. * (dx / M_PI - 1.)));
  const double C_gauss_p = exp(L * (2. * (1. - dx * dx / (4. * M_PI * M_PI))));
  const double C_gauss_m =
      exp(L * (2. * (1. - (dx - 2. * M_PI) * (dx - 2. * M_PI) / (4. * M_PI * M_PI))));
  const double sigma = sigma_beta / sqrt(2.);
  const double p_left = N_m / (N_p + N_m);
  double x = 0.0;
  for (;;) {
    double xi, xi_acc, a_min, a_max, mu, C_gauss;
    rng_uniform2(r, xi, xi_acc);
    if (xi >= p_left) {
      a_min = -M_PI + dx;
      a_max = +M_PI;
      mu = 0.5 * dx;
      C_gauss = C_gauss_p;
    } else {
      a_min = -M_PI;
      a_max = -M_PI + dx;
      mu = 0.5 * (dx - 2. * M_PI);
      C_gauss = C_gauss_m;
    }
    bool inside = false;
    while (!inside) {
      double z0, z1;
      rng_normal2(r, z0, z1);
      x = sigma * z0 + mu;
      inside = ((x >= a_min) && (x < a_max));
      if (!inside) {
        x = sigma * z1 + mu;
        inside = ((x >= a_min) && (x < a_max));
      }
    }
    const double I0 = cyl_bessel_i0(fabs(2. * beta * cos(0.5 * x)));
    const double I0_dx = cyl_bessel_i0(fabs(2. * beta * cos(0.5 * (x - dx))));
    const double x_shifted = (x - mu) / sigma_beta;
    const double rho_accept = I0 * I0_dx / C_gauss * exp(x_shifted * x_shifted);
    if (xi_acc <= rho_accept)
      break;
  }
  return mod_2pi(sign_flip * x + x_p);
}

// distribution/besselproductdistribution.cc:15-25 (rescaled = true).  cos(k phi) by the
// Chebyshev recurrence c_k = 2 c_1 c_{k-1} - c_{k-2} (one cosine instead of sixteen; the
// rounding error grows like k^2 eps <= 3e-14 and is weighted by the rapidly decaying alphaZ[k])
__device__ __forceinline__ double besselproduct_Znorm_inv_rescaled(const BesselProductConst &bp,
                                                                   const double phi) {
  const double c1 = cos(phi), two_c1 = 2. * c1;
  double ckm = 1.0, ck = c1;
  double s = 1.0 + bp.alphaZ[1] * c1;
#pragma unroll
  for (int k = 2; k <= 16; ++k) {
    const double cn = two_c1 * ck - ckm;
    ckm = ck;
    ck = cn;
    s += bp.alphaZ[k] * cn;
  }
  return 1.0 / s;
}

// distribution/approximatebesselproductdistribution.cc:38-54 (the reference's rho,
// SURVEY 7.3-8)
__device__ __forceinline__ void approx_N_p_sigma2inv(const double beta, const double x0, double &N_p,
                                                     double &sigma2_p_inv, double &sigma2_m_inv) {
  const double epsilon = 0.125 * M_PI;
  if (x0 < epsilon) {
    sigma2_p_inv = beta;
    sigma2_m_inv = 0.0;
    N_p = 1.0;
  } else {
    sigma2_p_inv = beta * cos_fast(0.25 * x0);
    sigma2_m_inv = beta * sin_force(0.25 * x0);
    const double q = sigma2_p_inv / sigma2_m_inv;
    const double rho = q * sqrt(q) * exp(-4.0 * (sigma2_p_inv - sigma2_m_inv));
    N_p = 1.0 / (1.0 + rho);
  }
}

// by-products of an approximate-Bessel-product draw: w = x - x0/2 (the drawn value relative to
// the main peak, before wrapping), and the mixture parameters
struct ApproxDrawn {
  double w, N_p, s_p, s_m;
};

// distribution/approximatebesselproductdistribution.hh:81-106; xi: the uniform variate that
// selects the mode (supplied by the caller, which gets it for free from the Philox call
// that also yields the step-2 split angle), the normal comes from one rng_normal2
// (approxbessel_draw_z: the same with the normal variate handed in)
__device__ __forceinline__ double approxbessel_draw_z(const double z0, const double beta, const double x_p,
                                                      const double x_m, const double xi,
                                                      ApproxDrawn *info = nullptr) {
  double x0 = x_p - x_m;
  double sign_flip = (x0 < 0) ? -1 : +1;
  x0 *= sign_flip;
  if (x0 > M_PI) {
    x0 = 2. * M_PI - x0;
    sign_flip *= -1;
  }
  double N_p, sigma2_p_inv, sigma2_m_inv;
  approx_N_p_sigma2inv(beta, x0, N_p, sigma2_p_inv, sigma2_m_inv);
  double sigma, xshift;
  if (xi <= N_p) {
    sigma = 1. / sqrt(sigma2_p_inv);
    xshift = 0.0;
  } else {
    sigma = 1. / sqrt(sigma2_m_inv);
    xshift = M_PI;
  }
  if (info) {
    info->w = sigma * z0 - xshift;
    info->N_p = N_p;
    info->s_p = sigma2_p_inv;
    info->s_m = sigma2_m_inv;
  }
  const double x = sigma * z0 + 0.5 * x0 - xshift;
  return mod_2pi_fast(sign_flip * x + x_m);
}
__device__ __forceinline__ double approxbessel_draw(Rng &r, const double beta, const double x_p,
                                                    const double x_m, const double xi,
                                                    ApproxDrawn *info = nullptr) {
  double z0, z1;
  rng_normal2(r, z0, z1);
  return approxbessel_draw_z(z0, beta, x_p, x_m, xi, info);
}

// the mixture pdf of distribution/approximatebesselproductdistribution.cc:17-35 as a function of
// w = z - x0/2 (distance from the main peak)
__device__ __forceinline__ double approxbessel_pdf_w(const double N_p, const double sigma2_p_inv,
                                                     const double sigma2_m_inv, const double w) {
  const double N_m = 1. - N_p;
  const double sq_p = sqrt(sigma2_p_inv), sq_m = sqrt(sigma2_m_inv);
  double s_p = 0.0, s_m = 0.0;
  // images that can contribute: a < 746 needs |w + 2 k pi| (first mode) or |w + (2k + 1) pi| (second mode)
  // below R = sqrt(1492 / sigma2_inv) <= sqrt(1492 / min), so k runs over [(-R - pi - w) / 2 pi, (R - w) / 2 pi]
  // (one image for beta >~ 150; the tests inside the loop still decide term by term)
  int k_lo = -4, k_hi = 4;
  {
    const double smin = (sq_m != 0.0) ? fmin(sigma2_p_inv, sigma2_m_inv) : sigma2_p_inv;
    const double R = 1492.0 * rsqrt(1492.0 * smin) + 1e-6;
    k_lo = max(-4, (int)floor((-R - M_PI - w) * (0.5 / M_PI)));
    k_hi = min(4, (int)ceil((R - w) * (0.5 / M_PI)));
  }
#pragma unroll 1
  for (int k = k_lo; k <= k_hi; ++k) {
    // exp(-a) == +0.0 exactly for a > 746 in IEEE double, and the second mode has weight
    // sq_m == 0 when x0 < pi/8: skipping those terms leaves the sums bit-identical
    double z_shifted = w + 2 * k * M_PI;
    double a = 0.5 * sigma2_p_inv * z_shifted * z_shifted;
    if (a < 746.0)
      s_p += sq_p * exp(-a);
    z_shifted += M_PI;
    a = 0.5 * sigma2_m_inv * z_shifted * z_shifted;
    if (a < 746.0 && sq_m != 0.0)
      s_m += sq_m * exp(-a);
  }
  return sqrt(0.5 / M_PI) * (N_p * s_p + N_m * s_m);
}

// distribution/approximatebesselproductdistribution.cc:7-35
__device__ __forceinline__ double approxbessel_pdf(const double beta, const double x,
                                                   const double x_p, const double x_m) {
  double x0 = x_p - x_m;
  double z = x - x_m;
  double sign_flip = (x0 < 0) ? -1 : +1;
  x0 *= sign_flip;
  if (x0 > M_PI) {
    x0 = 2. * M_PI - x0;
    sign_flip *= -1;
  }
  z *= sign_flip;
  double N_p, sigma2_p_inv, sigma2_m_inv;
  approx_N_p_sigma2inv(beta, x0, N_p, sigma2_p_inv, sigma2_m_inv);
  return approxbessel_pdf_w(N_p, sigma2_p_inv, sigma2_m_inv, z - 0.5 * x0);
}

// ------------------------------------------------------------------ reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum (result valid in thread 0); blockDim.x multiple of 32, <= 1024
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads(); // protect red[] against a previous use
  if (lane == 0)
    red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
  if (w == 0)
    v = warp_sum(v);
  return v;
}

// ------------------------------------------------------- per-model entry points
// (implemented in qm.cu / schwinger.cu / gff.cu; dispatched from capi.cu)
struct FillConst {
  BesselProductConst bp; // valid when model == SCHWINGER, coarsening BOTH, beta <= 8
};

#define DECL_MODEL_API(ns)                                                                         \
  namespace ns {                                                                                   \
  int init_state(mlmcpi_ctx *, const mlmcpi_model *, double *, int, uint32_t, uint64_t);            \
  int action(mlmcpi_ctx *, const mlmcpi_model *, const double *, int, double *);                   \
  int force(mlmcpi_ctx *, const mlmcpi_model *, const double *, double *, int);                    \
  int leapfrog(mlmcpi_ctx *, const mlmcpi_model *, int, double, double *, double *, int);          \
  int hmc_momentum(mlmcpi_ctx *, const mlmcpi_model *, double *, int, uint32_t, uint64_t);          \
  int hmc_step(mlmcpi_ctx *, const mlmcpi_model *, int, double, double *, int, uint32_t, uint64_t, \
               int32_t *, double *);                                                               \
  int overrelax_sweep(mlmcpi_ctx *, const mlmcpi_model *, double *, int);                          \
  int heatbath_sweep(mlmcpi_ctx *, const mlmcpi_model *, double *, int, uint32_t, uint64_t);        \
  int dof_update(mlmcpi_ctx *, const mlmcpi_model *, double *, int, int, int, uint32_t, uint64_t);   \
  int prolong(mlmcpi_ctx *, const mlmcpi_model *, const double *, double *, int);                  \
  int restrict_(mlmcpi_ctx *, const mlmcpi_model *, const double *, double *, int);                \
  int fill(mlmcpi_ctx *, const mlmcpi_model *, double *, int, uint32_t, uint64_t);                  \
  int prolong_fill(mlmcpi_ctx *, const mlmcpi_model *, const double *, double *, int, uint32_t,    \
                   uint64_t);                                                                      \
  int prolong_fill_eval(mlmcpi_ctx *, const mlmcpi_model *, const double *, double *, int,         \
                        uint32_t, uint64_t, double *);                                             \
  int cond_action(mlmcpi_ctx *, const mlmcpi_model *, const double *, int, double *);              \
  int qoi(mlmcpi_ctx *, const mlmcpi_model *, int, const double *, int, double *, int64_t *);      \
  }
DECL_MODEL_API(qm)
DECL_MODEL_API(schwinger)
DECL_MODEL_API(gff)

namespace qm {
int cluster_update(mlmcpi_ctx *, const mlmcpi_model *, double *, int, uint32_t, uint64_t, int);
int exact_draw(mlmcpi_ctx *, const mlmcpi_model *, double *, int, uint32_t, uint64_t);
int hierarchical_draw(mlmcpi_ctx *, const mlmcpi_model *, int, int, double, double *const *, int, uint32_t,
                      uint64_t, double *, double *, bool, int32_t *, unsigned long long *);
}
namespace gff {
int exact_draw(mlmcpi_ctx *, const mlmcpi_model *, double *, int, uint32_t, uint64_t);
void release_dense(mlmcpi_ctx *);
int overrelax_sweeps(mlmcpi_ctx *, const mlmcpi_model *, double *, int, int);
int sweep_sequence(mlmcpi_ctx *, const mlmcpi_model *, double *, int, int, int, uint32_t, const uint64_t *);
}
namespace schwinger {
int overrelax_sweeps(mlmcpi_ctx *, const mlmcpi_model *, double *, int, int);
int from_cluster(mlmcpi_ctx *, const mlmcpi_model *, const double *, double *, int, uint32_t, uint64_t);
int prolong_fill_eval_charge(mlmcpi_ctx *, const mlmcpi_model *, const double *, double *, int, uint32_t, uint64_t,
                             double *, const int32_t *mask);
int prolong_fill_eval_masked(mlmcpi_ctx *, const mlmcpi_model *, const double *, double *, int, uint32_t, uint64_t,
                             double *, const int32_t *mask);
int hmc_trial(mlmcpi_ctx *, const mlmcpi_model *, int, double, const double *, int, uint32_t, uint64_t,
              const double *, double *, int32_t *, const double **);
}

// host helpers shared by the model files (capi.cu)
void besselproduct_setup(double beta, BesselProductConst *bp);
// generic elementwise / accept kernels (capi.cu)
int launch_hmc_accept(mlmcpi_ctx *ctx, int B, uint32_t chain0, uint64_t draw, const double *d_S_cur,
                      const double *d_S_trial, const double *d_T_cur, const double *d_T_trial,
                      int32_t *d_accept, double *d_diag);
int launch_masked_copy(mlmcpi_ctx *ctx, double *d_dst, const double *d_src, size_t n, int B,
                       const int32_t *d_accept, bool wrap_angles = false, double *d_dst2 = nullptr);
int launch_half_sqnorm(mlmcpi_ctx *ctx, const double *d_p, size_t n, int B, double *d_T);

// ------------------------------------------- deterministic two-pass reductions
// pass 1: every (chain, block) pair writes one partial per output; pass 2 sums
// the partials of a chain in a fixed order (no atomics: results are reproducible).
// F: __device__ void operator()(int chain, long long site, double acc[NOUT]) const
template <int NOUT, class F>
__global__ void site_reduce_kernel(F f, long long nsites, int nblk, int B, double *partial) {
  const int chain = blockIdx.x / nblk, blk = blockIdx.x % nblk;
  double acc[NOUT];
#pragma unroll
  for (int k = 0; k < NOUT; ++k)
    acc[k] = 0.0;
  for (long long s = (long long)blk * blockDim.x + threadIdx.x; s < nsites;
       s += (long long)nblk * blockDim.x)
    f(chain, s, acc);
#pragma unroll
  for (int k = 0; k < NOUT; ++k) {
    const double v = block_sum(acc[k]);
    if (threadIdx.x == 0)
      partial[((size_t)k * B + chain) * nblk + blk] = v;
  }
}

enum { EPI_SCALE = 0, EPI_CHI = 1 };
// out[k][chain] = scale_k * sum;  EPI_CHI: out[chain] = scale_0 * sum_0^2 and
// Qint[chain] = -round(sum_1)
int launch_reduce_finish(mlmcpi_ctx *ctx, const double *partial, int nblk, int B, int nout, int epi,
                         double scale0, double scale1, double *out, int64_t *Qint, const int32_t *mask = nullptr);

// number of blocks per chain for a site reduction
static inline int reduce_nblk(const mlmcpi_ctx *ctx, long long nsites, int B, int threads) {
  const long long target = (long long)ctx->n_sm * 8;
  long long nblk = (target + B - 1) / B;
  const long long maxblk = (nsites + threads - 1) / threads;
  if (nblk > maxblk)
    nblk = maxblk;
  if (nblk < 1)
    nblk = 1;
  return (int)nblk;
}

template <int NOUT, class F>
int site_reduce(mlmcpi_ctx *ctx, const char *what, F f, long long nsites, int B, int epi,
                double scale0, double scale1, double *out, int64_t *Qint) {
  const int threads = 256;
  const int nblk = reduce_nblk(ctx, nsites, B, threads);
  double *partial = ctx_scratch(ctx, (size_t)NOUT * B * nblk);
  if (!partial)
    return MLMCPI_ENOMEM;
  site_reduce_kernel<NOUT, F><<<nblk * B, threads, 0, ctx->stream>>>(f, nsites, nblk, B, partial);
  MLMCPI_LAUNCHED(what);
  return launch_reduce_finish(ctx, partial, nblk, B, NOUT, epi, scale0, scale1, out, Qint);
}
