// schwinger.cu -- quenched Schwinger model: U(1) link angles on a periodic 2-D lattice.
//
// State layout [chain][ell], ell = 2*Mt*j + 2*i + mu exactly as the reference
// (lattice/lattice2d.hh:348-353): the two links of a site are one aligned double2,
// a warp reads 32 consecutive sites of a row = 512 contiguous bytes.
//
// Hot kernel: leapfrog_rowmarch_kernel.  One CTA = one lattice row segment of
// Mt threads marching over R consecutive rows.  Each plaquette sine is computed
// exactly once; sin P(i,j-1) is carried in a register from the previous row,
// sin P(i-1,j) and theta(i+1,j,1) come from the neighbouring thread through
// double-buffered shared memory (one __syncthreads per row).  HBM traffic per
// site-step is the algorithmic minimum: R theta, R p, W theta, W p = 64 B.
//
// Reference citations relative to /root/reference/src.
#include <algorithm>

#include <type_traits>

#include "common.cuh"

namespace {

struct SW {
  int Mt, Mx;
  double beta;
  int envelope; // ExpCos envelope: 0 reference, 1 tight
};

SW make_sw(const mlmcpi_ctx *ctx, const mlmcpi_model *m) {
  SW s;
  s.Mt = m->Mt_lat;
  s.Mx = m->Mx_lat;
  s.beta = m->beta;
  s.envelope = ctx->expcos_envelope;
  return s;
}

__device__ __forceinline__ int wrap_inc(int i, int M) { return i + 1 == M ? 0 : i + 1; }
__device__ __forceinline__ int wrap_dec(int i, int M) { return i == 0 ? M - 1 : i - 1; }

#define TH(x, i, j, mu) (x)[2 * ((size_t)Mt * (j) + (i)) + (mu)]

// plaquette angle, qft/quenchedschwingeraction.cc:14-17 (same summation order)
__device__ __forceinline__ double plaq(const double *x, int Mt, int Mx, int i, int j) {
  const int ip = wrap_inc(i, Mt), jp = wrap_inc(j, Mx);
  const double2 own = *reinterpret_cast<const double2 *>(&TH(x, i, j, 0));
  return own.x + TH(x, ip, j, 1) - TH(x, i, jp, 0) - own.y;
}

// qft/quenchedschwingeraction.cc:25-43.  Only the heat bath uses the two angles (as the arguments
// of the ExpCos draw, whose result is an angle modulo 2 pi): the reciprocal-multiply mod_2pi.
__device__ __forceinline__ void staple_angles(const double *x, int Mt, int Mx, int i, int j, int mu,
                                              double &theta_p, double &theta_m) {
  const int ip = wrap_inc(i, Mt), im = wrap_dec(i, Mt), jp = wrap_inc(j, Mx), jm = wrap_dec(j, Mx);
  if (mu == 0) {
    theta_p = mod_2pi_fast(TH(x, i, jp, 0) + TH(x, i, j, 1) - TH(x, ip, j, 1));
    theta_m = mod_2pi_fast(TH(x, i, jm, 0) + TH(x, ip, jm, 1) - TH(x, i, jm, 1));
  } else {
    theta_p = mod_2pi_fast(TH(x, i, j, 0) + TH(x, ip, j, 1) - TH(x, i, jp, 0));
    theta_m = mod_2pi_fast(TH(x, im, jp, 0) + TH(x, im, j, 1) - TH(x, im, j, 0));
  }
}

// ----------------------------------------------------------------- init_state
// qft/quenchedschwingeraction.cc:198-204
__global__ void init_state_kernel(SW sw, double *x, int B, uint32_t chain0, uint64_t seed,
                                  uint64_t draw) {
  const long long nsite = (long long)sw.Mt * sw.Mx;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsite * B)
    return;
  const int chain = (int)(t / nsite);
  const uint32_t k = (uint32_t)(t % nsite);
  Rng r = rng_init(seed, MLMCPI_STREAM_INIT, draw, chain0 + chain, k);
  double v1;
  const double v0 = rng_angle2(r, v1);
  reinterpret_cast<double2 *>(x)[t] = make_double2(v0, v1);
}

// sampler/hmcsampler.cc:24-26
__global__ void momentum_kernel(SW sw, double *p, int B, uint32_t chain0, uint64_t seed,
                                uint64_t draw) {
  const long long nsite = (long long)sw.Mt * sw.Mx;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsite * B)
    return;
  const int chain = (int)(t / nsite);
  const uint32_t k = (uint32_t)(t % nsite);
  Rng r = rng_init(seed, MLMCPI_STREAM_HMC_MOMENTUM, draw, chain0 + chain, k);
  double z0, z1;
  rng_normal2(r, z0, z1);
  reinterpret_cast<double2 *>(p)[t] = make_double2(z0, z1);
}

// ----------------------------------------------------------------- reductions
struct ActionF { // qft/quenchedschwingeraction.cc:7-22
  SW sw;
  const double *x;
  __device__ void operator()(int chain, long long s, double acc[1]) const {
    const int Mt = sw.Mt, Mx = sw.Mx;
    const int j = (int)(s / Mt), i = (int)(s - (long long)j * Mt);
    acc[0] += 1. - cos(plaq(x + (size_t)chain * 2 * Mt * Mx, Mt, Mx, i, j));
  }
};
struct PlaqF { // qoi/qft/qoiavgplaquette.cc:7-27
  SW sw;
  const double *x;
  __device__ void operator()(int chain, long long s, double acc[1]) const {
    const int Mt = sw.Mt, Mx = sw.Mx;
    const int j = (int)(s / Mt), i = (int)(s - (long long)j * Mt);
    acc[0] += cos(plaq(x + (size_t)chain * 2 * Mt * Mx, Mt, Mx, i, j));
  }
};
struct ChiF { // qoi/qft/qoi2dsusceptibility.cc:7-27
  SW sw;
  const double *x;
  __device__ void operator()(int chain, long long s, double acc[2]) const {
    const int Mt = sw.Mt, Mx = sw.Mx;
    const int j = (int)(s / Mt), i = (int)(s - (long long)j * Mt);
    const double th = plaq(x + (size_t)chain * 2 * Mt * Mx, Mt, Mx, i, j);
    acc[0] += mod_2pi(th);
    acc[1] += winding(th); // sum_P theta_P = 0 exactly => Q / 2 pi = - sum of windings
  }
};

// Row-marching plaquette reduction for the action and the two QoIs.  A thread owns one
// column and walks R rows: theta(i,j+1,0) is the next row's load, so a site costs one
// 16-byte and one 8-byte load (the latter an L1 hit) and one transcendental; no integer
// divisions.  MODE 0: sum (1 - cos P); 1: sum cos P; 2: sum mod_2pi(P) and sum of windings.
template <int MODE>
__global__ void plaq_reduce_kernel(SW sw, const double *__restrict__ x_all, int R, int chunks, int strips,
                                   int B, double *partial) {
  const int Mt = sw.Mt, Mx = sw.Mx;
  const int per_chain = chunks * strips;
  const int chain = blockIdx.x / per_chain;
  const int rem = blockIdx.x - chain * per_chain;
  const int chunk = rem / strips, strip = rem - chunk * strips;
  const int i = strip * blockDim.x + threadIdx.x;
  const int j0 = chunk * R, j1 = min(j0 + R, Mx);
  double acc0 = 0.0, acc1 = 0.0;
  if (i < Mt) {
    const double2 *xs = reinterpret_cast<const double2 *>(x_all) + (size_t)chain * Mt * Mx;
    const int ip = wrap_inc(i, Mt);
    double2 cur = xs[(size_t)j0 * Mt + i];
    for (int j = j0; j < j1; ++j) {
      const int jp = wrap_inc(j, Mx);
      const double2 nxt = xs[(size_t)jp * Mt + i];
      const double t1p = xs[(size_t)j * Mt + ip].y;
      const double P = cur.x + t1p - nxt.x - cur.y; // quenchedschwingeraction.cc:14-17
      if (MODE == 0) {
        acc0 += 1. - cos(P);
      } else if (MODE == 1) {
        acc0 += cos(P);
      } else {
        acc0 += mod_2pi(P);
        acc1 += winding(P);
      }
      cur = nxt;
    }
  }
  const double v0 = block_sum(acc0);
  if (threadIdx.x == 0)
    partial[(size_t)chain * per_chain + rem] = v0;
  if (MODE == 2) {
    const double v1 = block_sum(acc1);
    if (threadIdx.x == 0)
      partial[((size_t)B + chain) * per_chain + rem] = v1;
  }
}

template <int MODE>
int plaq_reduce(mlmcpi_ctx *ctx, const char *what, const SW &sw, const double *x, int B, int epi, double scale0,
                double *out, int64_t *Qint) {
  const int threads = std::min(256, ((sw.Mt + 31) / 32) * 32);
  const int strips = cdiv(sw.Mt, threads);
  // enough blocks for a few waves; at least 8 rows per block
  int R = sw.Mx;
  while (R > 8 && (long long)cdiv(sw.Mx, R) * strips * B < (long long)ctx->n_sm * 16)
    R = (R + 1) / 2;
  const int chunks = cdiv(sw.Mx, R);
  const int nblk = chunks * strips;
  const int nout = (MODE == 2) ? 2 : 1;
  double *partial = ctx_scratch(ctx, (size_t)nout * B * nblk);
  if (!partial)
    return MLMCPI_ENOMEM;
  plaq_reduce_kernel<MODE><<<nblk * B, threads, 0, ctx->stream>>>(sw, x, R, chunks, strips, B, partial);
  MLMCPI_LAUNCHED(what);
  return launch_reduce_finish(ctx, partial, nblk, B, nout, epi, scale0, 0.0, out, Qint);
}

// ---------------------------------------------------------------------- force
// One site of sampler/hmcsampler.cc:43-45 given the three plaquette forces F = beta sin P(i,j),
// F_jm = beta sin P(i,j-1), F_im = beta sin P(i-1,j): p -= dt_p dS/dtheta, theta += dt_x p.  Every leapfrog
// kernel goes through this one function, with the roundings pinned by intrinsics (products beta sin P
// rounded once, the force differences rounded, the two axpy's fused), so that all variants produce the
// same bits whatever the compiler would contract in their different loop bodies.
__device__ __forceinline__ double plaq_force(const double beta, const double P) { return __dmul_rn(beta, sin_force(P)); }
__device__ __forceinline__ void leap_site(const double dt_p, const double dt_x, const double F, const double F_jm,
                                          const double F_im, double2 &p, const double2 th, double2 &th_new) {
  p.x = __fma_rn(-dt_p, __dsub_rn(F, F_jm), p.x);
  p.y = __fma_rn(-dt_p, __dsub_rn(F_im, F), p.y);
  th_new.x = __fma_rn(dt_x, p.x, th.x);
  th_new.y = __fma_rn(dt_x, p.y, th.y);
}

// gather form of qft/quenchedschwingeraction.cc:68-89: each link receives +F of
// one plaquette and -F of another (two-term sums: order-independent, so this is
// the reference's value given the same sin)
__global__ void force_kernel(SW sw, const double *x, double *f, int B) {
  const int Mt = sw.Mt, Mx = sw.Mx;
  const long long nsite = (long long)Mt * Mx;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsite * B)
    return;
  const long long chain = t / nsite;
  const int s = (int)(t - chain * nsite);
  const int j = s / Mt, i = s - j * Mt;
  const double *xc = x + chain * 2 * nsite;
  const double F = plaq_force(sw.beta, plaq(xc, Mt, Mx, i, j));
  const double Fjm = plaq_force(sw.beta, plaq(xc, Mt, Mx, i, wrap_dec(j, Mx)));
  const double Fim = plaq_force(sw.beta, plaq(xc, Mt, Mx, wrap_dec(i, Mt), j));
  reinterpret_cast<double2 *>(f)[t] = make_double2(__dsub_rn(F, Fjm), __dsub_rn(Fim, F));
}

// ------------------------------------------------------------------- leapfrog
// generic fallback (any Mt): three sines per site
__global__ void leapfrog_naive_kernel(SW sw, double dt_p, double dt_x, const double *x_in,
                                      double *x_out, double *p, int B) {
  const int Mt = sw.Mt, Mx = sw.Mx;
  const long long nsite = (long long)Mt * Mx;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsite * B)
    return;
  const long long chain = t / nsite;
  const int s = (int)(t - chain * nsite);
  const int j = s / Mt, i = s - j * Mt;
  const double *xc = x_in + chain * 2 * nsite;
  const double F = plaq_force(sw.beta, plaq(xc, Mt, Mx, i, j));
  const double Fjm = plaq_force(sw.beta, plaq(xc, Mt, Mx, i, wrap_dec(j, Mx)));
  const double Fim = plaq_force(sw.beta, plaq(xc, Mt, Mx, wrap_dec(i, Mt), j));
  double2 pp = reinterpret_cast<double2 *>(p)[t];
  const double2 th = reinterpret_cast<const double2 *>(x_in)[t];
  double2 tn;
  leap_site(dt_p, dt_x, F, Fjm, Fim, pp, th, tn);
  reinterpret_cast<double2 *>(p)[t] = pp;
  if (x_out)
    reinterpret_cast<double2 *>(x_out)[t] = tn;
}

// row-marching kernel: blockDim.x == Mt (<= 1024), grid = (chunks per lattice) * B,
// each block marches over R rows of one chain.
//   sampler/hmcsampler.cc:43-45 fused: force, p -= dt_p*dp, theta += dt_x*p
template <bool DRIFT>
__global__ void __launch_bounds__(1024)
    leapfrog_rowmarch_kernel(SW sw, double dt_p, double dt_x, const double *__restrict__ x_in,
                             double *__restrict__ x_out, double *__restrict__ p, int R,
                             int chunks) {
  extern __shared__ double sh[]; // [2][Mt] theta(.,.,1) | [2][Mt] sin P
  const int Mt = sw.Mt, Mx = sw.Mx;
  const int i = threadIdx.x;
  const int ip = wrap_inc(i, Mt), im = wrap_dec(i, Mt);
  const int chain = blockIdx.x / chunks, chunk = blockIdx.x - chain * chunks;
  const int j0 = chunk * R;
  const int j1 = min(j0 + R, Mx);
  const size_t base = (size_t)chain * Mt * Mx;
  const double2 *xin = reinterpret_cast<const double2 *>(x_in) + base;
  double2 *xout = reinterpret_cast<double2 *>(x_out) + base;
  double2 *pp = reinterpret_cast<double2 *>(p) + base;
  double *sh_t1 = sh, *sh_s = sh + 2 * Mt;
  const double beta = sw.beta;

  // prologue: sin P(i, j0-1) and P(i, j0)
  const int jm = wrap_dec(j0, Mx);
  double2 prev = xin[(size_t)jm * Mt + i];
  double2 cur = xin[(size_t)j0 * Mt + i];
  double2 nxt = xin[(size_t)wrap_inc(j0, Mx) * Mt + i];
  sh_t1[i] = prev.y;
  sh_t1[Mt + i] = cur.y;
  __syncthreads();
  double s_prev = plaq_force(beta, prev.x + sh_t1[ip] - cur.x - prev.y);
  double P_cur = cur.x + sh_t1[Mt + ip] - nxt.x - cur.y;
  __syncthreads();
  int b = 0;
  for (int j = j0; j < j1; ++j) {
    // prefetch row j+2 (theta) and row j (p)
    const int jp2 = wrap_inc(wrap_inc(j, Mx), Mx);
    const double2 nxt2 = xin[(size_t)jp2 * Mt + i];
    double2 pj = pp[(size_t)j * Mt + i];
    const double s = plaq_force(beta, P_cur); // (the exchanged quantity is the force beta sin P)
    sh_s[b * Mt + i] = s;
    sh_t1[b * Mt + i] = nxt.y;
    __syncthreads();
    const double s_im = sh_s[b * Mt + im];
    const double t1p = sh_t1[b * Mt + ip];
    double2 tn;
    leap_site(dt_p, dt_x, s, s_prev, s_im, pj, cur, tn);
    pp[(size_t)j * Mt + i] = pj;
    if (DRIFT)
      xout[(size_t)j * Mt + i] = tn;
    // next row
    P_cur = nxt.x + t1p - nxt2.x - nxt.y;
    s_prev = s;
    cur = nxt;
    nxt = nxt2;
    b ^= 1;
  }
}

// ---- TMA (cp.async.bulk) + mbarrier row pipeline -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile("{\n"
               ".reg .pred P1;\n"
               "LAB_WAIT:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
               "@P1 bra DONE;\n"
               "bra LAB_WAIT;\n"
               "DONE:\n"
               "}" ::"r"(smem_u32(bar)), "r"(parity)
               : "memory");
}

// Row-pipelined leapfrog step: the hot kernel.  blockDim.x == Mt; the block marches over R
// rows of one chain.  One elected thread streams the theta and p rows into an S-stage
// shared-memory ring with cp.async.bulk (TMA, 1-D bulk copies: a lattice row is contiguous),
// completion is signalled on one mbarrier per stage, so S-2 rows (S-2)*32*Mt bytes per
// block are in flight without costing registers.  All neighbour accesses -- theta(i+1,j,1),
// theta(i,j+1,0) -- are shared-memory reads of the staged rows; only sin P(i-1,j) is
// exchanged between threads (double-buffered, one __syncthreads per row, which also
// releases the stage of row j-1 for the next bulk copy).
template <bool DRIFT, int S>
__global__ void __launch_bounds__(1024)
    leapfrog_rowpipe_kernel(SW sw, double dt_p, double dt_x, const double *__restrict__ x_in,
                            double *__restrict__ x_out, double *__restrict__ p, int R, int chunks) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int Mt = sw.Mt, Mx = sw.Mx;
  const int i = threadIdx.x;
  const int ip = wrap_inc(i, Mt), im = wrap_dec(i, Mt);
  const int chain = blockIdx.x / chunks, chunk = blockIdx.x - chain * chunks;
  const int j0 = chunk * R;
  const int nrow = min(R, Mx - j0); // rows this block updates
  const size_t base = (size_t)chain * Mt * Mx;
  const double2 *xin = reinterpret_cast<const double2 *>(x_in) + base;
  double2 *xout = reinterpret_cast<double2 *>(x_out) + base;
  double2 *pp = reinterpret_cast<double2 *>(p) + base;
  const uint32_t row_bytes = 16u * Mt;
  // smem: S stages x (theta row | p row), then sin exchange [2][Mt], then S mbarriers
  double2 *st_theta = reinterpret_cast<double2 *>(smem_raw);
  double2 *st_p = st_theta + (size_t)S * Mt;
  double *sh_s = reinterpret_cast<double *>(st_p + (size_t)S * Mt);
  uint64_t *bars = reinterpret_cast<uint64_t *>(sh_s + 2 * Mt);
  const double beta = sw.beta;

  if (i == 0) {
    for (int s = 0; s < S; ++s)
      mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  // logical row q = 0 .. nrow+1 is lattice row j0 - 1 + q; p rows exist for q = 1 .. nrow
  auto issue = [&](int q) {
    const int st = q % S;
    int j = j0 - 1 + q;
    j = j < 0 ? j + Mx : (j >= Mx ? j - Mx : j);
    const bool has_p = (q >= 1 && q <= nrow);
    mbar_expect_tx(&bars[st], has_p ? 2 * row_bytes : row_bytes);
    bulk_g2s(st_theta + (size_t)st * Mt, xin + (size_t)j * Mt, row_bytes, &bars[st]);
    if (has_p)
      bulk_g2s(st_p + (size_t)st * Mt, pp + (size_t)j * Mt, row_bytes, &bars[st]);
  };
  const int nq = nrow + 2;
  if (i == 0)
    for (int q = 0; q < S && q < nq; ++q)
      issue(q);
  // prologue: sin P(i, j0-1) from rows q = 0, 1
  mbar_wait(&bars[0], 0);
  mbar_wait(&bars[1 % S], 0);
  double2 cur = st_theta[(size_t)(1 % S) * Mt + i];
  double s_prev;
  {
    const double2 prev = st_theta[i];
    s_prev = plaq_force(beta, prev.x + st_theta[ip].y - cur.x - prev.y);
  }
  int b = 0;
  for (int r = 1; r <= nrow; ++r) { // updating logical row q = r
    const int st = r % S, stn = (r + 1) % S;
    mbar_wait(&bars[stn], ((r + 1) / S) & 1);
    const double2 nxt = st_theta[(size_t)stn * Mt + i];
    const double t1p = st_theta[(size_t)st * Mt + ip].y;
    double2 pj = st_p[(size_t)st * Mt + i];
    const double s = plaq_force(beta, cur.x + t1p - nxt.x - cur.y);
    sh_s[b * Mt + i] = s;
    __syncthreads(); // force row visible; every thread is done with logical row r-1
    if (i == 0 && r - 1 + S < nq)
      issue(r - 1 + S);
    const double s_im = sh_s[b * Mt + im];
    double2 tn;
    leap_site(dt_p, dt_x, s, s_prev, s_im, pj, cur, tn);
    const size_t g = (size_t)(j0 + r - 1) * Mt + i;
    pp[g] = pj;
    if (DRIFT)
      xout[g] = tn;
    s_prev = s;
    cur = nxt;
    b ^= 1;
  }
}

// Two leapfrog steps per HBM pass (temporal blocking).  Same row pipeline, but every thread
// carries two stages: stage A applies step k to lattice row r+1 while stage B applies step k+1
// to row r-1, fed from registers (own column) and three small exchange arrays (neighbouring
// columns); still one __syncthreads per row.  The block recomputes a halo of one row of step k
// on each side of its R output rows (R+4 theta rows and R+2 p rows are streamed in), so the
// traffic per site is 64 + 96/R bytes for TWO steps.  p is ping-ponged as well, because the
// halo rows of p belong to neighbouring blocks.
template <int S>
__global__ void __launch_bounds__(1024)
    leapfrog_rowpipe2_kernel(SW sw, double dtpA, double dtxA, double dtpB, double dtxB,
                             const double *__restrict__ x_in, double *__restrict__ x_out,
                             const double *__restrict__ p_in, double *__restrict__ p_out, int R,
                             int chunks) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int Mt = sw.Mt, Mx = sw.Mx;
  const int i = threadIdx.x;
  const int ip = wrap_inc(i, Mt), im = wrap_dec(i, Mt);
  const int chain = blockIdx.x / chunks, chunk = blockIdx.x - chain * chunks;
  const int j0 = chunk * R;
  const int nrow = min(R, Mx - j0);
  const size_t base = (size_t)chain * Mt * Mx;
  const double2 *xin = reinterpret_cast<const double2 *>(x_in) + base;
  const double2 *pin = reinterpret_cast<const double2 *>(p_in) + base;
  double2 *xout = reinterpret_cast<double2 *>(x_out) + base;
  double2 *pout = reinterpret_cast<double2 *>(p_out) + base;
  const uint32_t row_bytes = 16u * Mt;
  double2 *st_theta = reinterpret_cast<double2 *>(smem_raw);
  double2 *st_p = st_theta + (size_t)S * Mt;
  double *exA = reinterpret_cast<double *>(st_p + (size_t)S * Mt); // [2][Mt] sin P of step k
  double *exB = exA + 2 * Mt;                                      // [2][Mt] sin P of step k+1
  double *exT = exB + 2 * Mt;                                      // [2][Mt] theta^1(.,.,1)
  uint64_t *bars = reinterpret_cast<uint64_t *>(exT + 2 * Mt);
  const double beta = sw.beta;

  if (i == 0) {
    for (int s = 0; s < S; ++s)
      mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  // logical row q = 0 .. nrow+3 is lattice row j0 - 2 + q; p rows are needed for q = 1 .. nrow+2
  const int nq = nrow + 4;
  auto issue = [&](int q) {
    const int st = q % S;
    int j = (j0 - 2 + q) % Mx;
    if (j < 0)
      j += Mx;
    const bool has_p = (q >= 1 && q <= nrow + 2);
    mbar_expect_tx(&bars[st], has_p ? 2 * row_bytes : row_bytes);
    bulk_g2s(st_theta + (size_t)st * Mt, xin + (size_t)j * Mt, row_bytes, &bars[st]);
    if (has_p)
      bulk_g2s(st_p + (size_t)st * Mt, pin + (size_t)j * Mt, row_bytes, &bars[st]);
  };
  if (i == 0)
    for (int q = 0; q < S && q < nq; ++q)
      issue(q);
  mbar_wait(&bars[0], 0);
  mbar_wait(&bars[1 % S], 0);
  double2 cur0 = st_theta[(size_t)(1 % S) * Mt + i];
  double sA_prev;
  {
    const double2 prev = st_theta[i];
    sA_prev = plaq_force(beta, prev.x + st_theta[ip].y - cur0.x - prev.y);
  }
  __syncthreads(); // row 0 consumed
  if (i == 0 && S < nq)
    issue(S);
  double2 th1_m = make_double2(0., 0.), th1_c = th1_m, th1_p = th1_m;
  double2 pa_m = th1_m, pa_c = th1_m, pa_p = th1_m;
  double sB_prev = 0.0;
  for (int t = 0; t <= nrow + 2; ++t) {
    const bool doA = (t + 1 <= nrow + 2);
    const int qb = t - 1;
    const bool doB = (qb >= 1);
    double2 nxt0 = cur0, pj = pa_p;
    double sA = 0.0, sB = 0.0;
    if (doA) { // stage A, row qa = t + 1: sin of the step-k plaquette
      const int qa = t + 1;
      const int st = qa % S, stn = (qa + 1) % S;
      mbar_wait(&bars[stn], ((qa + 1) / S) & 1);
      nxt0 = st_theta[(size_t)stn * Mt + i];
      const double t1p = st_theta[(size_t)st * Mt + ip].y;
      pj = st_p[(size_t)st * Mt + i];
      sA = plaq_force(beta, cur0.x + t1p - nxt0.x - cur0.y);
      exA[(t & 1) * Mt + i] = sA;
    }
    if (doB) { // stage B, row qb: sin of the step-(k+1) plaquette from theta^1
      const double t1p = exT[(qb & 1) * Mt + ip];
      sB = plaq_force(beta, th1_m.x + t1p - th1_c.x - th1_m.y);
      exB[(t & 1) * Mt + i] = sB;
    }
    __syncthreads();
    // every thread is done with theta^0 / p^0 row t + 1: refill its stage
    if (i == 0 && t + 1 + S < nq)
      issue(t + 1 + S);
    if (doA) {
      const double sA_im = exA[(t & 1) * Mt + im];
      leap_site(dtpA, dtxA, sA, sA_prev, sA_im, pj, cur0, th1_p);
      pa_p = pj;
      exT[((t + 1) & 1) * Mt + i] = th1_p.y;
      sA_prev = sA;
      cur0 = nxt0;
    }
    if (doB) {
      if (qb >= 2) {
        const double sB_im = exB[(t & 1) * Mt + im];
        double2 pb = pa_m, tn;
        leap_site(dtpB, dtxB, sB, sB_prev, sB_im, pb, th1_m, tn);
        const size_t g = (size_t)(j0 + qb - 2) * Mt + i;
        pout[g] = pb;
        xout[g] = tn;
      }
      sB_prev = sB;
    }
    th1_m = th1_c;
    th1_c = th1_p;
    pa_m = pa_c;
    pa_c = pa_p;
  }
}

// K leapfrog steps per HBM pass: the register pipeline of leapfrog_rowpipe2_kernel generalised to K
// stages, with the block size MT = Mt a compile-time constant.  Stage s applies step k+s to logical row
// q = t - 2 s in iteration t (rows q = 0 .. R + 2K - 1 are the lattice rows j0 - K .. j0 + R + K - 1: K halo
// rows per side are recomputed); its inputs theta^s rows q, q+1 and p^s row q are the outputs stage s-1
// left in registers one and two iterations earlier (stage 0: the TMA ring), the neighbouring columns'
// sin P^s(i-1, q) and theta^s(i+1, q, 1) arrive through double-buffered shared-memory rows, and ONE
// __syncthreads per row serves all stages.  HBM traffic per site and K steps: 32 (1 + (2K - 1)/R) B read +
// 32 B written, i.e. 71 B for K = 4, R = 32 against the algorithmic 4 x 64 B.
//
// Why a second kernel rather than more template parameters on the first: the two-step kernel is capped at
// 64 registers by __launch_bounds__(1024), re-materialises the 36 constants of its two sines in every
// iteration and recomputes ring indices with integer divisions (cuobjdump: 290 instructions per thread
// and iteration of which 50 are fp64; profiles/r02_summary.md section 4).  Here: compile-time MT, ring
// stage and mbarrier phase carried incrementally, the producer's row pointer advanced by additions only.
// Every site update is the expression of the other leapfrog kernels on the same operands: bit-identical
// trajectories (tests/test_gpu_parity.py::test_rowmarch_equals_generic_leapfrog, test_leapfrog_variants_*).
template <int K> struct LeapDt {
  double dtp[K], dtx[K];
};

#ifndef MLMCPI_LFK_BLOCKS_128
#define MLMCPI_LFK_BLOCKS_128 4 // resident blocks per SM the register allocation aims at, MT <= 128 (measured: 3 -> 27.5, 4 -> 25.5, 5 (spills) -> 50 us per step)
#endif
template <int K, int MT, int S>
__global__ void __launch_bounds__(MT, (K >= 8 ? 2 : (MT <= 128 ? MLMCPI_LFK_BLOCKS_128 : (MT <= 256 ? 2 : 1))))
    leapfrog_rowpipek_kernel(const double beta, const LeapDt<K> dt, const int Mx, const double *__restrict__ x_in,
                             double *__restrict__ x_out, const double *__restrict__ p_in,
                             double *__restrict__ p_out, const int R, const int chunks) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2 *const st_theta = reinterpret_cast<double2 *>(smem_raw); // [S][MT]
  double2 *const st_p = st_theta + S * MT;                         // [S][MT]
  double *const exS = reinterpret_cast<double *>(st_p + S * MT);   // [K][2][MT] sin P^s
  double *const exT = exS + K * 2 * MT;                            // [K][2][MT] theta^s(., ., 1), s >= 1
  uint64_t *const bars = reinterpret_cast<uint64_t *>(exT + K * 2 * MT);
  const int i = threadIdx.x;
  const int ip = (i + 1 == MT) ? 0 : i + 1, im = (i == 0) ? MT - 1 : i - 1;
  const int chain = blockIdx.x / chunks, chunk = blockIdx.x - chain * chunks;
  const int j0 = chunk * R;
  const int nrow = min(R, Mx - j0);
  const int nq = nrow + 2 * K; // logical rows; p rows exist for q = 1 .. nq - 2
  const size_t base = (size_t)chain * MT * Mx;
  const double2 *const xin = reinterpret_cast<const double2 *>(x_in) + base;
  const double2 *const pin = reinterpret_cast<const double2 *>(p_in) + base;
  // output row of the last stage: logical row q is lattice row j0 - K + q
  double2 *const xout = reinterpret_cast<double2 *>(x_out) + base + (ptrdiff_t)(j0 - K) * MT + i;
  double2 *const pout = reinterpret_cast<double2 *>(p_out) + base + (ptrdiff_t)(j0 - K) * MT + i;
  constexpr uint32_t row_bytes = 16u * MT;

  if (i == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s)
      mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  // producer state (thread 0): next logical row to stream, its lattice row and ring stage
  int iq = 0, ij = j0 - K, ist = 0;
  ij = ij < 0 ? ij + Mx : ij; // (K <= Mx)
  auto issue = [&]() {
    const bool has_p = (iq >= 1 && iq <= nq - 2);
    mbar_expect_tx(&bars[ist], has_p ? 2 * row_bytes : row_bytes);
    bulk_g2s(st_theta + ist * MT, xin + (size_t)ij * MT, row_bytes, &bars[ist]);
    if (has_p)
      bulk_g2s(st_p + ist * MT, pin + (size_t)ij * MT, row_bytes, &bars[ist]);
    ++iq;
    ij = (ij + 1 == Mx) ? 0 : ij + 1;
    ist = (ist + 1 == S) ? 0 : ist + 1;
  };
  if (i == 0)
    for (int q = 0; q < S && q < nq; ++q)
      issue();

  // consumer state: ring stage of row t (st), ring stage / phase of row t + 1 (stn, phn)
  int st = 0, stn = (1 == S) ? 0 : 1;
  uint32_t phn = 0;
  mbar_wait(&bars[0], 0);
  double2 cur0 = st_theta[i];
  // th[s][.], pq[s][.]: theta^{s+1}, p^{s+1} of the two rows stage s finished last.  In an iteration of
  // parity P stage s+1 reads row q from slot P and row q + 1 from slot P ^ 1, and -- the stages being
  // updated in DESCENDING order -- stage s then writes its new row q + 2 into slot P: no register moves.
  double2 th[K][2], pq[K][2];
  double sprev[K];
#pragma unroll
  for (int s = 0; s < K; ++s) {
    sprev[s] = 0.0;
    th[s][0] = th[s][1] = pq[s][0] = pq[s][1] = make_double2(0., 0.);
  }
  // one iteration; PAR = t & 1 and STEADY (every stage does a full update: 3K - 2 <= t <= nq - 2) are
  // compile-time, so the exchange-buffer offsets are constants and the steady state has no range checks
  auto iteration = [&](auto par_c, auto steady_c, const int t) {
    constexpr int PAR = decltype(par_c)::value;
    constexpr bool STEADY = decltype(steady_c)::value;
    // Shared-memory loads first, then the K independent sines, then the stores: a store to one exchange
    // row followed by a load from another must stay in program order (the compiler cannot prove that they
    // do not alias), which would chain the K sines one after the other (seen in the SASS: four back-to-back
    // dependent sequences of 20 fp64 instructions, "wait" the dominant stall reason)
    double ss[K], P[K];
    bool act[K];
    double2 nxt0 = cur0, pj0 = make_double2(0., 0.);
    act[0] = STEADY || (t <= nq - 2);
    if (act[0]) {
      mbar_wait(&bars[stn], phn);
      nxt0 = st_theta[stn * MT + i];
      const double t1p = st_theta[st * MT + ip].y;
      pj0 = st_p[st * MT + i];
      P[0] = cur0.x + t1p - nxt0.x - cur0.y;
    } else {
      P[0] = 0.0;
    }
#pragma unroll
    for (int s = 1; s < K; ++s) {
      const int q = t - 2 * s;
      act[s] = STEADY || (q >= s && q <= nq - 2 - s);
      const double2 c = th[s - 1][PAR], n = th[s - 1][PAR ^ 1];
      const double t1p = exT[(s * 2 + PAR) * MT + ip]; // (finite garbage while the stage is idle)
      P[s] = c.x + t1p - n.x - c.y;
    }
#pragma unroll
    for (int s = 0; s < K; ++s)
      ss[s] = plaq_force(beta, P[s]);
#pragma unroll
    for (int s = 0; s < K; ++s)
      if (act[s])
        exS[(s * 2 + PAR) * MT + i] = ss[s];
    __syncthreads();
    // every thread is done with the staged row t: stream the next row into its stage
    if (i == 0 && iq < nq)
      issue();
    double s_im[K];
#pragma unroll
    for (int s = 0; s < K; ++s)
      s_im[s] = exS[(s * 2 + PAR) * MT + im];
    double2 pn[K], tn[K];
    bool upd[K];
#pragma unroll
    for (int s = 0; s < K; ++s) {
      const int q = t - 2 * s;
      upd[s] = STEADY || (act[s] && q >= s + 1);
      const double2 c = (s == 0) ? cur0 : th[s - 1][PAR];
      pn[s] = (s == 0) ? pj0 : pq[s - 1][PAR];
      leap_site(dt.dtp[s], dt.dtx[s], ss[s], sprev[s], s_im[s], pn[s], c, tn[s]);
    }
    // stores and the register hand-over, last stage first (stage s overwrites the slot stage s + 1 just read)
#pragma unroll
    for (int s = K - 1; s >= 0; --s) {
      const int q = t - 2 * s;
      if (upd[s]) {
        if (s == K - 1) {
          pout[(ptrdiff_t)q * MT] = pn[s];
          xout[(ptrdiff_t)q * MT] = tn[s];
        } else {
          th[s][PAR] = tn[s];
          pq[s][PAR] = pn[s];
          exT[((s + 1) * 2 + PAR) * MT + i] = tn[s].y; // theta^{s+1}(i, q, 1) for the neighbouring column
        }
      }
      if (act[s])
        sprev[s] = ss[s];
    }
    if (act[0])
      cur0 = nxt0;
    st = stn;
    stn = (stn + 1 == S) ? 0 : stn + 1;
    phn ^= (stn == 0) ? 1u : 0u;
  };
  using P0 = std::integral_constant<int, 0>;
  using P1 = std::integral_constant<int, 1>;
  const int t_last = nq + K - 3;
  // ramp-up (t < 3K - 2; 3K - 2 is even for even K), steady state in pairs, ramp-down
  constexpr int T0 = (3 * K - 2 + 1) & ~1; // first even t of the steady state
  int t = 0;
  for (; t < T0 && t <= t_last; t += 2) {
    iteration(P0{}, std::false_type{}, t);
    if (t + 1 <= t_last)
      iteration(P1{}, std::false_type{}, t + 1);
  }
  for (; t + 1 <= nq - 2; t += 2) {
    iteration(P0{}, std::true_type{}, t);
    iteration(P1{}, std::true_type{}, t + 1);
  }
  for (; t <= t_last; t += 2) {
    iteration(P0{}, std::false_type{}, t);
    if (t + 1 <= t_last)
      iteration(P1{}, std::false_type{}, t + 1);
  }
}

// --------------------------------------------------------------------- sweeps
// colours: 0 {mu=0, j even}, 1 {mu=0, j odd}, 2 {mu=1, i even}, 3 {mu=1, i odd}.
// Grid: x = strips of a lattice row, y = the rows that hold links of this colour, z = chains
// (folded when B exceeds the grid limit) -- no integer division anywhere.
// (overrelaxation; the heat bath is heatbath_pair_kernel below)
__global__ void sweep_colour_kernel(SW sw, int colour, double *x, int B) {
  const int Mt = sw.Mt, Mx = sw.Mx;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  int i, j, mu;
  if (colour < 2) {
    mu = 0;
    i = k;
    j = 2 * blockIdx.y + colour;
  } else {
    mu = 1;
    i = 2 * k + (colour - 2);
    j = blockIdx.y;
  }
  if (i >= Mt)
    return;
  const size_t ell = 2 * ((size_t)Mt * j + i) + mu;
  for (int chain = blockIdx.z; chain < B; chain += gridDim.z) {
    double *xc = x + (size_t)chain * 2 * Mt * Mx;
    // qft/quenchedschwingeraction.cc:57-65: theta <- mod_2pi(theta_+ + theta_- - theta).  The reference reduces both
    // staple angles to [-pi, pi) first; the outer mod_2pi makes those two reductions redundant modulo 2 pi
    // (the result differs by rounding in the last bits only), so they are skipped
    const int ip = wrap_inc(i, Mt), im = wrap_dec(i, Mt), jp = wrap_inc(j, Mx), jm = wrap_dec(j, Mx);
    double sp, sm;
    if (mu == 0) {
      sp = TH(xc, i, jp, 0) + TH(xc, i, j, 1) - TH(xc, ip, j, 1);
      sm = TH(xc, i, jm, 0) + TH(xc, ip, jm, 1) - TH(xc, i, jm, 1);
    } else {
      sp = TH(xc, i, j, 0) + TH(xc, ip, j, 1) - TH(xc, i, jp, 0);
      sm = TH(xc, im, jp, 0) + TH(xc, im, j, 1) - TH(xc, im, j, 0);
    }
    xc[ell] = mod_2pi((sp + sm) - xc[ell]);
  }
}

// Heat-bath sweep of one colour with PAIRED variates (stream convention: include/mlmcpi.h).  The links of a colour in a
// lattice row are numbered n = 0, 1, ... (mu = 0: n = i; mu = 1: n = i / 2) and updated in pairs (2p, 2p + 1) by one
// thread: ONE Philox block of the even link -- a normal pair and a uniform pair -- serves the first ExpCos attempt of both
// (z0, u0 for the even, z1, u1 for the odd link); further attempts (3 % of the links at tau = 256, 0.3 % at 2048)
// continue on the link's own stream: calls 2, 3, ... for the even link, calls 0, 1, ... for the odd one.  Half the
// Philox rounds and Box-Muller transforms of a block per link; a last link without a partner behaves like an even one.
__device__ __forceinline__ void heatbath_pair_role(int i, int mu, int &first, int &i_first) {
  const int n = (mu == 0) ? i : (i >> 1);
  first = ((n & 1) == 0);
  i_first = first ? i : ((mu == 0) ? i - 1 : i - 2);
}
__global__ void heatbath_pair_kernel(SW sw, int colour, double *x, int B, uint32_t chain0, uint64_t seed,
                                     uint64_t draw) {
  const int Mt = sw.Mt, Mx = sw.Mx;
  const int p = blockIdx.x * blockDim.x + threadIdx.x; // pair index in the row
  int iA, iB, j, mu, cnt;
  if (colour < 2) {
    mu = 0;
    j = 2 * blockIdx.y + colour;
    cnt = Mt;
    iA = 2 * p;
    iB = 2 * p + 1;
  } else {
    mu = 1;
    j = blockIdx.y;
    cnt = Mt / 2;
    iA = 4 * p + (colour - 2);
    iB = iA + 2;
  }
  const bool inside = 2 * p < cnt;
  const bool hasB = 2 * p + 1 < cnt;
  const unsigned wmask = __ballot_sync(0xffffffffu, inside);
  if (!inside)
    return;
  const size_t ellA = 2 * ((size_t)Mt * j + iA) + mu, ellB = 2 * ((size_t)Mt * j + iB) + mu;
  for (int chain = blockIdx.z; chain < B; chain += gridDim.z) {
    double *xc = x + (size_t)chain * 2 * Mt * Mx;
    const uint32_t gchain = chain0 + (uint32_t)chain;
    Rng rA = rng_init(seed, MLMCPI_STREAM_HEATBATH, draw, gchain, (uint32_t)ellA);
    double z0, z1, u0, u1;
    rng_normal2(rA, z0, z1);
    rng_uniform2(rA, u0, u1);
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      double v = 0.0;
      if (h == 0 || hasB) { // qft/quenchedschwingeraction.cc:46-54
        double theta_p, theta_m;
        staple_angles(xc, Mt, Mx, h ? iB : iA, j, mu, theta_p, theta_m);
        Rng r = rA;
        if (h)
          r = rng_init(seed, MLMCPI_STREAM_HEATBATH, draw, gchain, (uint32_t)ellB);
        v = expcos_draw(r, sw.beta, theta_p, theta_m, sw.envelope, nullptr, true, h ? z1 : z0, h ? u1 : u0);
      }
      __syncwarp(wmask); // reconverge after the rejection loop: one coalesced store
      if (h == 0 || hasB)
        xc[h ? ellB : ellA] = v;
    }
  }
}

// Action::heatbath_update / overrelaxation_update of ONE link ell (qft/quenchedschwingeraction.cc:46-65):
// the per-dof interface of action/action.hh:85-110, one thread per chain.  The overrelaxation follows
// the reference's operation order (both staple angles reduced to [-pi, pi) first), so that a
// lexicographic sweep of single-link calls reproduces the reference's sweep.
template <bool HEATBATH>
__global__ void dof_update_kernel(SW sw, int ell, double *x, int B, uint32_t chain0, uint64_t seed, uint64_t draw) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= B)
    return;
  const int Mt = sw.Mt, Mx = sw.Mx;
  const int mu = ell & 1, site = ell >> 1;
  const int j = site / Mt, i = site - j * Mt;
  double *xc = x + (size_t)chain * 2 * Mt * Mx;
  const int ip = wrap_inc(i, Mt), im = wrap_dec(i, Mt), jp = wrap_inc(j, Mx), jm = wrap_dec(j, Mx);
  double theta_p, theta_m;
  if (mu == 0) { // compute_staple_angles, :25-43
    theta_p = mod_2pi(TH(xc, i, jp, 0) + TH(xc, i, j, 1) - TH(xc, ip, j, 1));
    theta_m = mod_2pi(TH(xc, i, jm, 0) + TH(xc, ip, jm, 1) - TH(xc, i, jm, 1));
  } else {
    theta_p = mod_2pi(TH(xc, i, j, 0) + TH(xc, ip, j, 1) - TH(xc, i, jp, 0));
    theta_m = mod_2pi(TH(xc, im, jp, 0) + TH(xc, im, j, 1) - TH(xc, im, j, 0));
  }
  if (HEATBATH) { // the paired variates of heatbath_pair_kernel: a single-link call reproduces the sweep's draw
    int first, i_first;
    heatbath_pair_role(i, mu, first, i_first);
    const uint32_t gchain = chain0 + (uint32_t)chain;
    Rng rA = rng_init(seed, MLMCPI_STREAM_HEATBATH, draw, gchain, (uint32_t)(2 * (Mt * j + i_first) + mu));
    double z0, z1, u0, u1;
    rng_normal2(rA, z0, z1);
    rng_uniform2(rA, u0, u1);
    Rng rg = rA;
    if (!first)
      rg = rng_init(seed, MLMCPI_STREAM_HEATBATH, draw, gchain, (uint32_t)ell);
    xc[ell] = expcos_draw(rg, sw.beta, theta_p, theta_m, sw.envelope, nullptr, true, first ? z0 : z1, first ? u0 : u1);
  } else {
    xc[ell] = mod_2pi(theta_p + theta_m - xc[ell]);
  }
}

// One overrelaxation sweep (all four colours, ascending order) in ONE pass over HBM, out of place.
// The four colour passes above read the whole lattice four times to update a quarter of the links
// each.  Here a block of Mt threads (one per column) marches over R rows of one chain with the
// colours pipelined over rows: with e an even row, new theta_0 of row e+2 (colour 0: old
// neighbours only), then new theta_0 of row e+1 (colour 1: needs the new rows e and e+2), then new
// theta_1 of rows e and e+1 (colours 2 and 3: need the new theta_0 of rows e, e+1, e+2; even
// columns first, odd columns after a barrier).  Old rows live in a 4-slot ring in shared memory, new
// theta_0 rows in a 3-slot ring.  One row below and two rows above the chunk are read in addition
// (theta_0 of the first and of the row after the last are recomputed, not stored), so the traffic is
// 16 (1 + 3/R) B read + 16 B written per site; the update of every link is the same expression
// on the same operands as in sweep_colour_kernel, i.e. the result is bit-identical.
__global__ void __launch_bounds__(1024)
    overrelax_rowpipe_kernel(SW sw, const double *__restrict__ x_in, double *__restrict__ x_out, int R,
                             int chunks) {
  extern __shared__ __align__(16) double sm_or[];
  const int Mt = sw.Mt, Mx = sw.Mx;
  const int i = threadIdx.x;
  const int ip = wrap_inc(i, Mt), im = wrap_dec(i, Mt);
  const int chain = blockIdx.x / chunks, chunk = blockIdx.x - chain * chunks;
  const int e0 = chunk * R;
  const int nrow = min(R, Mx - e0); // even
  const size_t base = (size_t)chain * Mt * Mx;
  const double2 *xin = reinterpret_cast<const double2 *>(x_in) + base;
  double2 *xout = reinterpret_cast<double2 *>(x_out) + base;
  // Shared memory serves the NEIGHBOURING columns only, and of the old rows the neighbours need
  // theta_1 alone (every old theta_0 an update uses is the thread's own column: registers).  So the
  // old ring holds theta_1 rows as plain doubles: half the footprint, and the 8-byte loads of a warp
  // are contiguous (the former double2 ring cost two wavefronts per load -- 44 % of the shared-memory
  // wavefronts were bank conflicts, ncu).
  double *old1 = sm_or;                  // [4][Mt]: old theta_1 of rows q - 1 .. q + 2 (slot = row & 3)
  double *An = sm_or + (size_t)4 * Mt;   // [3][Mt]: new theta_0 (slot = row % 3)
  double *Bn = An + (size_t)3 * Mt;      // [2][Mt]: new theta_1 of the two rows in flight
  // logical row q = 0, 1, ... is lattice row e0 - 1 + q (periodic); e0 and Mx are even, so inside the
  // loop only the prefetched row e0 + 2p + 4 can wrap, and it wraps to row 0 exactly
  const int jm1 = e0 == 0 ? Mx - 1 : e0 - 1;
  const int j3 = e0 + 2 >= Mx ? e0 + 2 - Mx : e0 + 2; // rows of q = 3, 4 (the pair after the first)
  // colour 0 / 1 on logical row q for this column: `up` / `dn` = theta_0 above / below (old for
  // colour 0, new for colour 1), own = old (theta_0, theta_1) of row q, below = old row q - 1;
  // o / om = old theta_1 rows q and q - 1, a_out = the An row that receives the result
  auto t0_update = [&](const double *o, const double *om, double *a_out, double up, double dn, double2 own,
                       double2 below) {
    const double sp = up + own.y - o[ip];
    const double sm = dn + om[ip] - below.y;
    const double v = mod_2pi((sp + sm) - own.x);
    a_out[i] = v;
    return v;
  };
  // colour 2 (even columns, old theta_1 neighbours in nb) or colour 3 (odd columns, new ones) on a row;
  // A / Aup = new theta_0 of that row and of the row above, a_own / a_up = this column's, b_own = old theta_1
  auto t1_update = [&](const double *nb, const double *A, const double *Aup, double a_own, double a_up,
                       double b_own) {
    const double sp = a_own + nb[ip] - a_up;
    const double sm = Aup[im] + nb[im] - A[im];
    return mod_2pi((sp + sm) - b_own);
  };
  double2 r0 = xin[jm1 * Mt + i]; // old rows q - 1, q, q + 1 of this column
  double2 r1 = xin[e0 * Mt + i];
  double2 r2 = xin[(e0 + 1) * Mt + i];
  old1[i] = r0.y;
  old1[Mt + i] = r1.y;
  old1[2 * Mt + i] = r2.y;
  // rows of the next pair are prefetched into registers one iteration ahead
  double2 pre0 = xin[j3 * Mt + i], pre1 = xin[(j3 + 1) * Mt + i];
  __syncthreads();
  double a_q = t0_update(old1 + Mt, old1, An + Mt, r2.x, r0.x, r1, r0); // colour 0 on the first (even) row
  const bool even_col = (i & 1) == 0;
  __syncthreads(); // the first iteration reuses the slot of logical row 0
  int jin = j3 + 2 >= Mx ? j3 + 2 - Mx : j3 + 2; // lattice row of logical row q + 4
  int jout = e0;                                 // lattice row of logical row q
  int s3 = 1;                                    // q % 3
  for (int p = 0; 2 * p < nrow; ++p) {
    const int q = 1 + 2 * p; // logical index of the even row e; r1 = old row q, r2 = old row q + 1
    // The two ring slots written next held rows q - 2 and q - 1; their last readers were the
    // even-column threads of the previous iteration, before its colour-2/3 barrier.
    const double2 r3 = pre0, r4 = pre1; // old rows q + 2, q + 3
    double *o_q = old1 + (q & 3) * Mt, *o_q1 = old1 + ((q + 1) & 3) * Mt, *o_q2 = old1 + ((q + 2) & 3) * Mt;
    double *a_s = An + s3 * Mt, *a_s1 = An + (s3 == 2 ? 0 : s3 + 1) * Mt, *a_s2 = An + (s3 == 0 ? 2 : s3 - 1) * Mt;
    o_q2[i] = r3.y;
    old1[((q + 3) & 3) * Mt + i] = r4.y;
    if (2 * (p + 1) < nrow) {
      pre0 = xin[jin * Mt + i];
      pre1 = xin[(jin + 1) * Mt + i];
    }
    __syncthreads();
    const double a_q2 = t0_update(o_q2, o_q1, a_s2, r4.x, r2.x, r3, r2); // colour 0, row e + 2
    // (no barrier: colour 1 reads old neighbours and this thread's own new theta_0 values only)
    const double a_q1 = t0_update(o_q1, o_q, a_s1, a_q2, a_q, r2, r1);   // colour 1, row e + 1
    __syncthreads();
    double b0 = 0.0, b1 = 0.0;
    if (even_col) { // colour 2 on rows e and e + 1
      b0 = t1_update(o_q, a_s, a_s1, a_q, a_q1, r1.y);
      b1 = t1_update(o_q1, a_s1, a_s2, a_q1, a_q2, r2.y);
      Bn[i] = b0;
      Bn[Mt + i] = b1;
    }
    __syncthreads();
    if (!even_col) { // colour 3
      b0 = t1_update(Bn, a_s, a_s1, a_q, a_q1, r1.y);
      b1 = t1_update(Bn + Mt, a_s1, a_s2, a_q1, a_q2, r2.y);
    }
    xout[jout * Mt + i] = make_double2(a_q, b0);
    xout[(jout + 1) * Mt + i] = make_double2(a_q1, b1);
    a_q = a_q2;
    r1 = r3;
    r2 = r4;
    jout += 2;
    jin = jin + 2 >= Mx ? 0 : jin + 2;
    s3 = s3 == 0 ? 2 : s3 - 1; // (q + 2) % 3
  }
}

// ------------------------------------------------------- prolong / restrict
// qft/quenchedschwingeraction.cc:92-147; one thread per coarse site
__global__ void prolong_kernel(SW sw, int ctype, const double *xc_all, double *x_all, int B) {
  const int Mt = sw.Mt, Mx = sw.Mx;
  const int Mtc = (ctype == MLMCPI_COARSEN_SPATIAL) ? Mt : Mt / 2;
  const int Mxc = (ctype == MLMCPI_COARSEN_TEMPORAL) ? Mx : Mx / 2;
  const long long nc = (long long)Mtc * Mxc;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * B)
    return;
  const long long chain = t / nc;
  const int s = (int)(t - chain * nc);
  const int j = s / Mtc, i = s - j * Mtc;
  const double2 c = reinterpret_cast<const double2 *>(xc_all)[t];
  double *x = x_all + chain * 2 * (long long)Mt * Mx;
  if (ctype == MLMCPI_COARSEN_BOTH) {
    TH(x, 2 * i, 2 * j, 0) = 0.5 * c.x;
    TH(x, 2 * i + 1, 2 * j, 0) = 0.5 * c.x;
    TH(x, 2 * i, 2 * j, 1) = 0.5 * c.y;
    TH(x, 2 * i, 2 * j + 1, 1) = 0.5 * c.y;
  } else if (ctype == MLMCPI_COARSEN_TEMPORAL) {
    TH(x, 2 * i, j, 0) = 0.5 * c.x;
    TH(x, 2 * i + 1, j, 0) = 0.5 * c.x;
    TH(x, 2 * i, j, 1) = c.y;
  } else {
    TH(x, i, 2 * j, 0) = c.x;
    TH(x, i, 2 * j, 1) = 0.5 * c.y;
    TH(x, i, 2 * j + 1, 1) = 0.5 * c.y;
  }
}

// qft/quenchedschwingeraction.cc:150-195
__global__ void restrict_kernel(SW sw, int ctype, const double *xf_all, double *xc_all, int B) {
  const int Mt = sw.Mt, Mx = sw.Mx;
  const int Mtc = (ctype == MLMCPI_COARSEN_SPATIAL) ? Mt : Mt / 2;
  const int Mxc = (ctype == MLMCPI_COARSEN_TEMPORAL) ? Mx : Mx / 2;
  const long long nc = (long long)Mtc * Mxc;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * B)
    return;
  const long long chain = t / nc;
  const int s = (int)(t - chain * nc);
  const int j = s / Mtc, i = s - j * Mtc;
  const double *x = xf_all + chain * 2 * (long long)Mt * Mx;
  double2 c;
  if (ctype == MLMCPI_COARSEN_BOTH) {
    c.x = mod_2pi(TH(x, 2 * i, 2 * j, 0) + TH(x, 2 * i + 1, 2 * j, 0));
    c.y = mod_2pi(TH(x, 2 * i, 2 * j, 1) + TH(x, 2 * i, 2 * j + 1, 1));
  } else if (ctype == MLMCPI_COARSEN_TEMPORAL) {
    c.x = mod_2pi(TH(x, 2 * i, j, 0) + TH(x, 2 * i + 1, j, 0));
    c.y = mod_2pi(TH(x, 2 * i, j, 1));
  } else {
    c.x = mod_2pi(TH(x, i, 2 * j, 0));
    c.y = mod_2pi(TH(x, i, 2 * j, 1) + TH(x, i, 2 * j + 1, 1));
  }
  reinterpret_cast<double2 *>(xc_all)[t] = c;
}

// -------------------------------------------------------------------- fill-in
// CoarsenBoth, qft/quenchedschwingerconditionedfineaction.cc:7-78.
// STEP 1 (:14-31), in place, one thread per coarse cell
__global__ void fill_both_step1_kernel(SW sw, double *x_all, int B, uint32_t chain0, uint64_t seed,
                                       uint64_t draw) {
  const int Mt = sw.Mt, Mx = sw.Mx, Mtc = Mt / 2, Mxc = Mx / 2;
  const long long nc = (long long)Mtc * Mxc;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * B)
    return;
  const long long chain = t / nc;
  const int cell = (int)(t - chain * nc);
  const int j = cell / Mtc, i = cell - j * Mtc;
  double *x = x_all + chain * 2 * (long long)Mt * Mx;
  Rng r = rng_init(seed, MLMCPI_STREAM_FILL1, draw, chain0 + (uint32_t)chain, cell);
  double dth_s;
  const double dth_t = rng_angle2(r, dth_s);
  TH(x, 2 * i, 2 * j, 0) = mod_2pi_fast(TH(x, 2 * i, 2 * j, 0) + dth_t);
  TH(x, 2 * i + 1, 2 * j, 0) = mod_2pi_fast(TH(x, 2 * i + 1, 2 * j, 0) - dth_t);
  TH(x, 2 * i, 2 * j, 1) = mod_2pi_fast(TH(x, 2 * i, 2 * j, 1) + dth_s);
  TH(x, 2 * i, 2 * j + 1, 1) = mod_2pi_fast(TH(x, 2 * i, 2 * j + 1, 1) - dth_s);
}

// the interior of one coarse cell given its 4 own and 4 neighbouring perimeter
// links: STEP 2 (:33-61) and STEP 3 (:63-77)
struct CellLinks {
  double A0, A1, B0, B1; // own perimeter: (2i,2j,0) (2i+1,2j,0) (2i,2j,1) (2i,2j+1,1)
  double R0, R1;         // cell (i+1,j): (2i+2,2j,1) (2i+2,2j+1,1)
  double T0, T1;         // cell (i,j+1): (2i,2j+2,0) (2i+1,2j+2,0)
  double V0, V1, H0, H1; // interior: (2i+1,2j,1) (2i+1,2j+1,1) (2i,2j+1,0) (2i+1,2j+1,0)
};

// CoarsenBoth, beta <= 8: one cell of qft/quenchedschwingerconditionedfineaction.cc:219-250
__device__ __forceinline__ double cond_both_bessel_cell(const CellLinks &c, const double beta,
                                                        const BesselProductConst &bp) {
  const double phi_12 = +c.B1 + c.T0;
  const double phi_23 = +c.T1 - c.R1;
  const double phi_34 = -c.A1 - c.R0;
  const double phi_41 = -c.A0 + c.B0;
  const double theta_1 = +c.H0;
  const double theta_2 = -c.V1;
  const double theta_3 = -c.H1;
  const double theta_4 = +c.V0;
  const double Phi = phi_12 + phi_23 + phi_34 + phi_41;
  double S = -beta * (cos(theta_1 - theta_2 - phi_12) + cos(theta_2 - theta_3 - phi_23) +
                      cos(theta_3 - theta_4 - phi_34) + cos(theta_4 - theta_1 - phi_41));
  S -= log(besselproduct_Znorm_inv_rescaled(bp, Phi));
  return S;
}

// sum over the four plaquettes of a cell of 1 - cos P (qft/quenchedschwingeraction.cc:14-17,
// same summation order inside a plaquette as plaq())
__device__ __forceinline__ double cell_plaquette_sum(const CellLinks &c) {
  return (1. - cos(c.A0 + c.V0 - c.H0 - c.B0)) + (1. - cos(c.A1 + c.R0 - c.H1 - c.V0)) +
         (1. - cos(c.H0 + c.V1 - c.T0 - c.B1)) + (1. - cos(c.H1 + c.R1 - c.T1 - c.V1));
}

// EVAL: also return sf = sum over the cell's plaquettes of (1 - cos P) and sc = the cell's
// term of ConditionedFineAction::evaluate for the values just drawn.  For beta > 8 both come
// from the by-products of the draws instead of from the stored angles: with x the accepted
// ExpCos proposal of a horizontal interior link and tau its concentration,
//   -log ExpCos.pdf = log(2 pi I0s(tau)) - tau (cos x - 1)      (expcosdistribution.cc:7-21)
//   (1 - cos P) + (1 - cos P') of the two plaquettes it separates = 2 - (tau / beta) cos x
// and the mixture pdf of the vertical pair is evaluated at w (approxbessel_pdf_w); these are the
// reference's formulas with the angle differences taken before the final mod_2pi rounding.
// threads per block of the fused fill-in kernel, measured at 512 chains x 512^2 <- 256^2, beta = 1024:
// 64: 3.81, 128: 3.70, 256: 3.53, 512: 3.57 ms (256 with 4 resident blocks)
#ifndef FILL_THREADS
#define FILL_THREADS 256
#endif
#ifndef FILL_PHASE_SYNC
#define FILL_PHASE_SYNC 0
#endif
// PHASE_SYNC: block-wide barriers between the phases of the draw.  They order nothing -- the phases of a cell
// only touch the cell's registers -- but keep the warps of a block in the same stretch of the instruction
// stream (the fused kernel's threads run through 2500 of its 4500 instructions once and "no_instructions" is
// its top stall reason).  Measured: no gain (3.54 ms with and without); off.  Neither did a called
// (non-inlined) Philox help: 4.29 ms.
// The seven Philox blocks every cell consumes, drawn in ONE rolled loop (one copy of the ten rounds, of the
// conversion and of the Box-Muller transform in the instruction stream instead of seven / two): the fused
// fill-in kernel runs through its code once per thread, and what it executes on the common path has to stay
// inside the 32 KB instruction cache of the SM (ncu: gcc__cache_requests_type_instruction at 99.5 % of peak
// with one copy per call site).  Same counters, same variates as the straight-line version.
//   q = 0, 1, 2: FILL1 of the cell, its right and its upper neighbour (NEIGH only): step-1 split angles
//   q = 3: FILL2 call 0 (split angle of the vertical pair, mode selector)   q = 4: FILL2 call 1 (normal pair)
//   q = 5: FILL3 call 0 of link (2i, 2j+1, 0) (normal pair)                 q = 6: FILL3 call 1 (uniform pair)
//   q = 7, 8: a later block of a FILL3 stream (normal pair, uniform pair: the retry of a rejected horizontal link)
// One function for the first seven blocks and for the retry (taken by a few lanes of every fifth warp); see FILL_GEN_ATTR.
struct FillVariates {
  double v[9][2];
};
// FILL_GEN_ATTR: __noinline__ takes the kernel off the instruction-cache limit (3.17 ms against 3.08 - 3.15 alone, GPC cache
// requests 62 % of peak instead of 99 %) but costs the whole step 4 % (8.5 against 8.1 - 8.2 ms, profiles/r02_summary.md
// section 12: the step runs at the board's power cap, and the faster, denser kernel lowers the clock of the leapfrog launches
// that follow it); inlined is the default.
#ifndef FILL_GEN_ATTR
#define FILL_GEN_ATTR __forceinline__
#endif
__device__ FILL_GEN_ATTR void fill_gen(FillVariates *V, int q_begin, int q_end, uint64_t seed, uint64_t draw,
                                      uint32_t gchain, int cell, int cell_r, int cell_t, int hidx, uint32_t hcall) {
  Rng r = rng_init(seed, MLMCPI_STREAM_FILL1, draw, gchain, cell);
#pragma unroll 1
  for (int q = q_begin; q < q_end; ++q) {
    r.c0 = (uint32_t)((q == 1) ? cell_r : ((q == 2) ? cell_t : ((q >= 5) ? hidx : cell)));
    const uint32_t stream = (q < 3) ? MLMCPI_STREAM_FILL1 : ((q < 5) ? MLMCPI_STREAM_FILL2 : MLMCPI_STREAM_FILL3);
    r.a = (stream << 24) | ((q >= 7) ? hcall + (uint32_t)(q - 7) : ((q == 4 || q == 6) ? 1u : 0u));
    double v0, v1;
    rng_uniform2(r, v0, v1);
    if (q == 4 || q == 5 || q == 7)
      box_muller(v0, v1, v0, v1);
    V->v[q][0] = v0;
    V->v[q][1] = v1;
  }
}
template <bool NEIGH>
__device__ __forceinline__ void fill_variates(FillVariates &V, uint64_t seed, uint64_t draw, uint32_t gchain,
                                              int cell, int cell_r, int cell_t, int hidx) {
  fill_gen(&V, NEIGH ? 0 : 3, 7, seed, draw, gchain, cell, cell_r, cell_t, hidx, 0);
}

template <bool APPROX, bool EVAL, bool PHASE_SYNC = false>
__device__ __forceinline__ void fill_cell_interior(CellLinks &c, const FillVariates &V, const double beta,
                                                   const int envelope,
                                                   const BesselProductConst &bp, uint64_t seed,
                                                   uint64_t draw, uint32_t gchain, int Mt, int i,
                                                   int j, int cell, double &sf, double &sc,
                                                   const unsigned wmask) {
  // wmask: the lanes of this warp that fill a cell.  The rejection loops leave the warp
  // diverged; without an explicit __syncwarp the code after a loop (the next draw, the
  // evaluation, the stores) would run once per group of lanes that left the loop together.
  ApproxDrawn ad;
  {
    const double theta_p = mod_2pi_fast(c.A1 + c.R0 + c.R1 - c.T1);
    const double theta_m = mod_2pi_fast(c.B0 + c.B1 + c.T0 - c.A0);
    // first call of the FILL2 stream: (split angle, mode selector of the approximate distribution); the
    // second: the normal pair of the approximate draw (the exact draw continues on the stream itself)
    const double dtheta = -M_PI + 2. * M_PI * V.v[3][0];
    double theta_tilde;
    if (APPROX) {
      theta_tilde = approxbessel_draw_z(V.v[4][0], beta, theta_p, theta_m, V.v[3][1], EVAL ? &ad : nullptr);
    } else {
      Rng r = rng_init(seed, MLMCPI_STREAM_FILL2, draw, gchain, cell);
      r.a += 1;
      theta_tilde = besselproduct_draw(r, bp, theta_p, theta_m);
    }
    __syncwarp(wmask);
    if (PHASE_SYNC)
      __syncthreads();
    c.V0 = mod_2pi_fast(0.5 * theta_tilde + dtheta);
    c.V1 = mod_2pi_fast(0.5 * theta_tilde - dtheta);
  }
  // STEP 3 for the two horizontal interior links (2i, 2j+1, 0) and (2i+1, 2j+1, 0); a rolled
  // loop, so that the rejection sampler exists once in the instruction stream
  // ONE normal pair and ONE uniform pair (calls 0, 1 of the stream of link (2i, 2j+1, 0)) serve the first attempt
  // of BOTH links: (z0, u0) for h = 0, (z1, u1) for h = 1; further attempts continue on the link's own stream
  // (h = 0: calls 2, 3, ...; h = 1: the stream of link (2i+1, 2j+1, 0) from call 0).  Two Philox calls and one
  // Box-Muller transform fewer per cell than a block per link (15 % of the kernel's instructions).
  double Zprod = 1.0, tsum = 0.0, tcos = 0.0;
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    const double theta_p = mod_2pi_fast(h == 0 ? c.A0 + c.V0 - c.B0 : c.A1 + c.R0 - c.V0);
    const double theta_m = mod_2pi_fast(h == 0 ? c.B1 + c.T0 - c.V1 : c.V1 + c.T1 - c.R1);
    Rng r = rng_init(seed, MLMCPI_STREAM_FILL3, draw, gchain, Mt * j + 2 * i + h);
    if (h == 0)
      r.a += 2;
    ExpCosDrawn e;
    FillVariates *Vp = const_cast<FillVariates *>(&V);
    const double H = expcos_draw_f(
        r,
        [&r, Vp, seed, draw, gchain](double &z0, double &z1, double &u0, double &u1) {
          fill_gen(Vp, 7, 9, seed, draw, gchain, 0, 0, 0, (int)r.c0, r.a & 0xffffffu);
          r.a += 2;
          z0 = Vp->v[7][0];
          z1 = Vp->v[7][1];
          u0 = Vp->v[8][0];
          u1 = Vp->v[8][1];
        },
        beta, theta_p, theta_m, envelope, &e, true, V.v[5][h], V.v[6][h]);
    __syncwarp(wmask);
    if (PHASE_SYNC)
      __syncthreads();
    if (h == 0)
      c.H0 = H;
    else
      c.H1 = H;
    if (EVAL && APPROX) {
      Zprod *= 2. * M_PI * fast_bessel_I0_scaled(e.tau);
      tsum += e.tau;
      tcos += e.tau * cos_fast(e.x);
    }
  }
  if (EVAL) {
    if (APPROX) {
      // log Zprod - log pdf as one logarithm (the pdf at a point that was just drawn is >= exp(-37) of its maximum)
      sc = log(Zprod / approxbessel_pdf_w(ad.N_p, ad.s_p, ad.s_m, ad.w)) + (tsum - tcos);
      sf = 4. - tcos / beta;
    } else {
      sc = cond_both_bessel_cell(c, beta, bp);
      sf = cell_plaquette_sum(c);
    }
  }
}

// STEP 2+3 in place (perimeter links already redistributed by step 1)
template <bool APPROX>
__global__ void fill_both_step23_kernel(SW sw, BesselProductConst bp, double *x_all, int B,
                                        uint32_t chain0, uint64_t seed, uint64_t draw) {
  const int Mt = sw.Mt, Mx = sw.Mx, Mtc = Mt / 2, Mxc = Mx / 2;
  const long long nc = (long long)Mtc * Mxc;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned wmask = __ballot_sync(0xffffffffu, t < nc * B);
  if (t >= nc * B)
    return;
  const long long chain = t / nc;
  const int cell = (int)(t - chain * nc);
  const int j = cell / Mtc, i = cell - j * Mtc;
  double *x = x_all + chain * 2 * (long long)Mt * Mx;
  const int i2 = wrap_inc(2 * i + 1, Mt), j2 = wrap_inc(2 * j + 1, Mx); // 2i+2, 2j+2 (periodic)
  CellLinks c;
  c.A0 = TH(x, 2 * i, 2 * j, 0);
  c.A1 = TH(x, 2 * i + 1, 2 * j, 0);
  c.B0 = TH(x, 2 * i, 2 * j, 1);
  c.B1 = TH(x, 2 * i, 2 * j + 1, 1);
  c.R0 = TH(x, i2, 2 * j, 1);
  c.R1 = TH(x, i2, 2 * j + 1, 1);
  c.T0 = TH(x, 2 * i, j2, 0);
  c.T1 = TH(x, 2 * i + 1, j2, 0);
  double sf_unused, sc_unused;
  FillVariates V;
  fill_variates<false>(V, seed, draw, chain0 + (uint32_t)chain, cell, 0, 0, Mt * j + 2 * i);
  __syncwarp(wmask);
  fill_cell_interior<APPROX, false>(c, V, sw.beta, sw.envelope, bp, seed, draw, chain0 + (uint32_t)chain, Mt, i, j, cell,
                                    sf_unused, sc_unused, wmask);
  TH(x, 2 * i + 1, 2 * j, 1) = c.V0;
  TH(x, 2 * i + 1, 2 * j + 1, 1) = c.V1;
  TH(x, 2 * i, 2 * j + 1, 0) = c.H0;
  TH(x, 2 * i + 1, 2 * j + 1, 0) = c.H1;
}

// prolongation + all three steps in ONE pass from the coarse state: the thread of
// cell (i,j) regenerates the step-1 variates of cells (i+1,j), (i,j+1) from their
// Philox counters instead of waiting for them (counter-based RNG makes the three
// phases of the reference embarrassingly parallel).  80 B of HBM traffic per cell.
// EVAL: the block additionally reduces S_f(theta') / beta and S_cond(theta') of its cells into
// partial[{0,1}][chain][block] (TwoLevelMetropolisStep::draw lines 48 and 65-66 without a second
// and third pass over theta').  Grid: nblk blocks of 128 cells per chain.
// CHARGE (with EVAL): a third partial sum, sum over the cell's four plaquettes of mod_2pi(P) -- the topological charge
// of theta' times 2 pi (qoi/qft/qoi2dsusceptibility.cc:7-27), so that the QoI of an accepted draw needs no pass of its own
template <bool APPROX, bool EVAL, bool CHARGE = false>
#ifndef FILL_MINBLK
#define FILL_MINBLK (1024 / FILL_THREADS)
#endif
// mask: chains whose cascade has already stopped (mask[chain] == 0) are skipped -- no proposal is made for them
// (hierarchicalsampler.cc:73-74 breaks out of the level loop); their rows of x and of the sums are left untouched
__global__ void __launch_bounds__(FILL_THREADS, FILL_MINBLK) prolong_fill_both_kernel(SW sw, BesselProductConst bp, const double *xc_all,
                                         double *x_all, int B, uint32_t chain0, uint64_t seed,
                                         uint64_t draw, int nblk, double *partial, const int32_t *mask = nullptr) {
  const int Mt = sw.Mt, Mx = sw.Mx, Mtc = Mt / 2, Mxc = Mx / 2;
  const int nc = Mtc * Mxc;
  const int chain = blockIdx.x / nblk, blk = blockIdx.x - chain * nblk;
  if (mask && !mask[chain])
    return; // (block-uniform)
  const int cell_raw = blk * blockDim.x + threadIdx.x;
  // (threads beyond the last cell of the chain redo the last cell and store nothing, so that every thread of
  // the block passes the same barriers)
  const bool live = cell_raw < nc;
  const int cell = live ? cell_raw : nc - 1;
  double sf = 0.0, sc = 0.0, qs = 0.0;
  const unsigned wmask = 0xffffffffu;
  {
    const int j = cell / Mtc, i = cell - j * Mtc;
    const uint32_t gchain = chain0 + (uint32_t)chain;
    const double2 *xc = reinterpret_cast<const double2 *>(xc_all) + (size_t)chain * nc;
    double *x = x_all + (size_t)chain * 2 * Mt * Mx;
    const int ipc = wrap_inc(i, Mtc), jpc = wrap_inc(j, Mxc);
    const int cell_r = j * Mtc + ipc, cell_t = jpc * Mtc + i;
    const double2 own = xc[cell];
    const double cr = xc[cell_r].y, ct = xc[cell_t].x;
    CellLinks c;
    FillVariates V;
    fill_variates<true>(V, seed, draw, gchain, cell, cell_r, cell_t, Mt * j + 2 * i);
    {
      const double dth_t = -M_PI + 2. * M_PI * V.v[0][0], dth_s = -M_PI + 2. * M_PI * V.v[0][1];
      c.A0 = mod_2pi_fast(0.5 * own.x + dth_t);
      c.A1 = mod_2pi_fast(0.5 * own.x - dth_t);
      c.B0 = mod_2pi_fast(0.5 * own.y + dth_s);
      c.B1 = mod_2pi_fast(0.5 * own.y - dth_s);
    }
    {
      const double dth_s = -M_PI + 2. * M_PI * V.v[1][1];
      c.R0 = mod_2pi_fast(0.5 * cr + dth_s);
      c.R1 = mod_2pi_fast(0.5 * cr - dth_s);
    }
    {
      const double dth_t = -M_PI + 2. * M_PI * V.v[2][0];
      c.T0 = mod_2pi_fast(0.5 * ct + dth_t);
      c.T1 = mod_2pi_fast(0.5 * ct - dth_t);
    }
    if (FILL_PHASE_SYNC)
      __syncthreads();
    fill_cell_interior<APPROX, EVAL, (FILL_PHASE_SYNC != 0)>(c, V, sw.beta, sw.envelope, bp, seed, draw, gchain, Mt, i, j, cell,
                                                            sf, sc, wmask);
    if (CHARGE) // the cell's four plaquettes, summed in the order of plaq()
      qs = (mod_2pi(c.A0 + c.V0 - c.H0 - c.B0) + mod_2pi(c.A1 + c.R0 - c.H1 - c.V0)) +
           (mod_2pi(c.H0 + c.V1 - c.T0 - c.B1) + mod_2pi(c.H1 + c.R1 - c.T1 - c.V1));
    if (live) {
      // rows 2j and 2j+1, sites 2i and 2i+1: four aligned double2 stores
      double2 *xs = reinterpret_cast<double2 *>(x);
      xs[(size_t)Mt * (2 * j) + 2 * i] = make_double2(c.A0, c.B0);
      xs[(size_t)Mt * (2 * j) + 2 * i + 1] = make_double2(c.A1, c.V0);
      xs[(size_t)Mt * (2 * j + 1) + 2 * i] = make_double2(c.H0, c.B1);
      xs[(size_t)Mt * (2 * j + 1) + 2 * i + 1] = make_double2(c.H1, c.V1);
    } else {
      sf = sc = qs = 0.0;
    }
  }
  if (EVAL) { // one partial sum per WARP: no block-wide barrier at the end of the kernel (ncu: 14 % of the stall samples)
    constexpr int WARPS = FILL_THREADS / 32;
    const double v0 = warp_sum(sf), v1 = warp_sum(sc);
    if ((threadIdx.x & 31) == 0) {
      const size_t slot = (size_t)blk * WARPS + (threadIdx.x >> 5), npart = (size_t)nblk * WARPS;
      partial[(size_t)chain * npart + slot] = v0;
      partial[((size_t)B + chain) * npart + slot] = v1;
    }
    if (CHARGE) {
      const double v2 = warp_sum(qs);
      if ((threadIdx.x & 31) == 0)
        partial[((size_t)2 * B + chain) * ((size_t)nblk * WARPS) + (size_t)blk * WARPS + (threadIdx.x >> 5)] = v2;
    }
  }
}

// semi-coarsening, qft/quenchedschwingerconditionedfineaction.cc:136-209.
// phase 0: uniform redistribution of the coarsened-direction pair; phase 1: ExpCos
// draw of the new transverse link.  One thread per coarse site.
__global__ void fill_semi_kernel(SW sw, int ctype, int phase, double *x_all, int B, uint32_t chain0,
                                 uint64_t seed, uint64_t draw) {
  const int Mt = sw.Mt, Mx = sw.Mx;
  const bool temporal = (ctype == MLMCPI_COARSEN_TEMPORAL);
  const int Mtc = temporal ? Mt / 2 : Mt, Mxc = temporal ? Mx : Mx / 2;
  const long long nc = (long long)Mtc * Mxc;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * B)
    return;
  const long long chain = t / nc;
  const int cell = (int)(t - chain * nc);
  const int j = cell / Mtc, i = cell - j * Mtc;
  const uint32_t gchain = chain0 + (uint32_t)chain;
  double *x = x_all + chain * 2 * (long long)Mt * Mx;
  if (phase == 0) {
    Rng r = rng_init(seed, MLMCPI_STREAM_FILL1, draw, gchain, cell);
    double unused;
    const double dtheta = rng_angle2(r, unused);
    if (temporal) {
      TH(x, 2 * i, j, 0) = mod_2pi_fast(TH(x, 2 * i, j, 0) + dtheta);
      TH(x, 2 * i + 1, j, 0) = mod_2pi_fast(TH(x, 2 * i + 1, j, 0) - dtheta);
    } else {
      TH(x, i, 2 * j, 1) = mod_2pi_fast(TH(x, i, 2 * j, 1) + dtheta);
      TH(x, i, 2 * j + 1, 1) = mod_2pi_fast(TH(x, i, 2 * j + 1, 1) - dtheta);
    }
  } else {
    Rng r = rng_init(seed, MLMCPI_STREAM_FILL3, draw, gchain, cell);
    if (temporal) {
      const int jp = wrap_inc(j, Mx), i2 = wrap_inc(2 * i + 1, Mt);
      const double theta_p = mod_2pi_fast(TH(x, 2 * i, j, 1) + TH(x, 2 * i, jp, 0) - TH(x, 2 * i, j, 0));
      const double theta_m =
          mod_2pi_fast(TH(x, 2 * i + 1, j, 0) + TH(x, i2, j, 1) - TH(x, 2 * i + 1, jp, 0));
      TH(x, 2 * i + 1, j, 1) = expcos_draw(r, sw.beta, theta_p, theta_m, sw.envelope);
    } else {
      const int ip = wrap_inc(i, Mt), j2 = wrap_inc(2 * j + 1, Mx);
      const double theta_p = mod_2pi_fast(TH(x, i, 2 * j, 0) + TH(x, ip, 2 * j, 1) - TH(x, i, 2 * j, 1));
      const double theta_m =
          mod_2pi_fast(TH(x, i, 2 * j + 1, 1) + TH(x, i, j2, 0) - TH(x, ip, 2 * j + 1, 1));
      TH(x, i, 2 * j + 1, 0) = expcos_draw(r, sw.beta, theta_p, theta_m, sw.envelope);
    }
  }
}

// ------------------------------------------------------- conditioned actions
// CoarsenBoth, beta <= 8: qft/quenchedschwingerconditionedfineaction.cc:219-250
struct CondBothBesselF {
  SW sw;
  BesselProductConst bp;
  const double *x_all;
  __device__ void operator()(int chain, long long cell, double acc[1]) const {
    const int Mt = sw.Mt, Mx = sw.Mx, Mtc = Mt / 2;
    const int j = (int)(cell / Mtc), i = (int)(cell - (long long)j * Mtc);
    const double *x = x_all + (size_t)chain * 2 * Mt * Mx;
    const int i2 = wrap_inc(2 * i + 1, Mt), j2 = wrap_inc(2 * j + 1, Mx);
    CellLinks c;
    c.A0 = TH(x, 2 * i, 2 * j, 0);
    c.A1 = TH(x, 2 * i + 1, 2 * j, 0);
    c.B0 = TH(x, 2 * i, 2 * j, 1);
    c.B1 = TH(x, 2 * i, 2 * j + 1, 1);
    c.R0 = TH(x, i2, 2 * j, 1);
    c.R1 = TH(x, i2, 2 * j + 1, 1);
    c.T0 = TH(x, 2 * i, j2, 0);
    c.T1 = TH(x, 2 * i + 1, j2, 0);
    c.V0 = TH(x, 2 * i + 1, 2 * j, 1);
    c.V1 = TH(x, 2 * i + 1, 2 * j + 1, 1);
    c.H0 = TH(x, 2 * i, 2 * j + 1, 0);
    c.H1 = TH(x, 2 * i + 1, 2 * j + 1, 0);
    acc[0] += cond_both_bessel_cell(c, sw.beta, bp);
  }
};

// one ExpCos term of the horizontal interior link (i, 2j+1, 0): :270-287 / :356-372
__device__ __forceinline__ double cond_expcos_term_h(const double *x, int Mt, int Mx, double beta,
                                                     int i, int j) {
  const int ip = wrap_inc(i, Mt), j2 = wrap_inc(2 * j + 1, Mx);
  const double phi_p = mod_2pi(-TH(x, i, 2 * j, 1) + TH(x, i, 2 * j, 0) + TH(x, ip, 2 * j, 1));
  const double phi_m = mod_2pi(+TH(x, i, 2 * j + 1, 1) + TH(x, i, j2, 0) - TH(x, ip, 2 * j + 1, 1));
  const double theta = mod_2pi(+TH(x, i, 2 * j + 1, 0));
  return -log(expcos_pdf(beta, theta, phi_p, phi_m));
}

// CoarsenBoth, beta > 8: :251-288
struct CondBothApproxF {
  SW sw;
  const double *x_all;
  __device__ void operator()(int chain, long long cell, double acc[1]) const {
    const int Mt = sw.Mt, Mx = sw.Mx, Mtc = Mt / 2;
    const int j = (int)(cell / Mtc), i = (int)(cell - (long long)j * Mtc);
    const double *x = x_all + (size_t)chain * 2 * Mt * Mx;
    const int i2 = wrap_inc(2 * i + 1, Mt), j2 = wrap_inc(2 * j + 1, Mx);
    const double phi_p = mod_2pi(+TH(x, 2 * i + 1, 2 * j, 0) + TH(x, i2, 2 * j, 1) +
                                 TH(x, i2, 2 * j + 1, 1) - TH(x, 2 * i + 1, j2, 0));
    const double phi_m = mod_2pi(-TH(x, 2 * i, 2 * j, 0) + TH(x, 2 * i, 2 * j, 1) +
                                 TH(x, 2 * i, 2 * j + 1, 1) + TH(x, 2 * i, j2, 0));
    const double theta = mod_2pi(+TH(x, 2 * i + 1, 2 * j, 1) + TH(x, 2 * i + 1, 2 * j + 1, 1));
    double S = -log(approxbessel_pdf(sw.beta, theta, phi_p, phi_m));
    S += cond_expcos_term_h(x, Mt, Mx, sw.beta, 2 * i, j);
    S += cond_expcos_term_h(x, Mt, Mx, sw.beta, 2 * i + 1, j);
    acc[0] += S;
  }
};

// semi-coarsening: :329-379
struct CondSemiF {
  SW sw;
  int ctype;
  const double *x_all;
  __device__ void operator()(int chain, long long cell, double acc[1]) const {
    const int Mt = sw.Mt, Mx = sw.Mx;
    const double *x = x_all + (size_t)chain * 2 * Mt * Mx;
    if (ctype == MLMCPI_COARSEN_TEMPORAL) {
      const int Mtc = Mt / 2;
      const int j = (int)(cell / Mtc), i = (int)(cell - (long long)j * Mtc);
      const int jp = wrap_inc(j, Mx), i2 = wrap_inc(2 * i + 1, Mt);
      const double phi_p = mod_2pi(-TH(x, 2 * i, j, 0) + TH(x, 2 * i, j, 1) + TH(x, 2 * i, jp, 0));
      const double phi_m =
          mod_2pi(+TH(x, 2 * i + 1, j, 0) + TH(x, i2, j, 1) - TH(x, 2 * i + 1, jp, 0));
      const double theta = mod_2pi(+TH(x, 2 * i + 1, j, 1));
      acc[0] += -log(expcos_pdf(sw.beta, theta, phi_p, phi_m));
    } else {
      const int j = (int)(cell / Mt), i = (int)(cell - (long long)j * Mt);
      acc[0] += cond_expcos_term_h(x, Mt, Mx, sw.beta, i, j);
    }
  }
};

// ------------------------------------------- QuenchedSchwingerClusterSampler
// sampler/quenchedschwingerclustersampler.cc:52-68: the rotor chain psi (length Mt*Mx)
// fixes every plaquette; links are built by running sums.  One thread per column j for the
// vertical links, then one thread per chain for the horizontal links of the last time slice.
__global__ void cluster_links_vertical_kernel(SW sw, const double *psi_all, double *x_all, int B) {
  const int Mt = sw.Mt, Mx = sw.Mx;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)Mx * B)
    return;
  const long long chain = t / Mx;
  const int j = (int)(t - chain * Mx);
  const double *psi = psi_all + chain * (long long)Mt * Mx;
  double *x = x_all + chain * 2 * (long long)Mt * Mx;
  for (int i = 0; i < Mt; ++i) { // zero every link of this column first (:50-51)
    TH(x, i, j, 0) = 0.0;
    TH(x, i, j, 1) = 0.0;
  }
  for (int i = 0; i < Mt - 1; ++i) {
    const long long i_lin = (long long)i * Mx + j;
    TH(x, i + 1, j, 1) = TH(x, i, j, 1) + psi[i_lin + 1] - psi[i_lin];
  }
}
__global__ void cluster_links_horizontal_kernel(SW sw, const double *psi_all, double *x_all, int B) {
  const int Mt = sw.Mt, Mx = sw.Mx;
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= B)
    return;
  const double *psi = psi_all + (long long)chain * Mt * Mx;
  double *x = x_all + (long long)chain * 2 * Mt * Mx;
  long long i_lin = (long long)(Mt - 1) * Mx;
  for (int j = 0; j < Mx - 1; ++j) {
    TH(x, Mt - 1, j + 1, 0) = TH(x, Mt - 1, j, 0) - TH(x, Mt - 1, j, 1) - psi[i_lin + 1] + psi[i_lin];
    i_lin++;
  }
}
// random gauge transformation (:70-82): link (i,j,0) gets +theta(i,j) - theta(i+1,j), link
// (i,j,1) gets +theta(i,j) - theta(i,j+1), each step followed by mod_2pi as in the reference
__device__ __forceinline__ double gauge_angle(uint64_t seed, uint64_t draw, uint32_t gchain, int Mt, int i,
                                              int j) {
  Rng r = rng_init(seed, MLMCPI_STREAM_GAUGE, draw, gchain, (uint32_t)(Mt * j + i));
  double second;
  return rng_angle2(r, second);
}
__global__ void cluster_gauge_kernel(SW sw, double *x_all, int B, uint32_t chain0, uint64_t seed,
                                     uint64_t draw) {
  const int Mt = sw.Mt, Mx = sw.Mx;
  const long long nsite = (long long)Mt * Mx;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsite * B)
    return;
  const long long chain = t / nsite;
  const int s = (int)(t - chain * nsite);
  const int j = s / Mt, i = s - j * Mt;
  const uint32_t gchain = chain0 + (uint32_t)chain;
  const double th = gauge_angle(seed, draw, gchain, Mt, i, j);
  const double th_ip = gauge_angle(seed, draw, gchain, Mt, wrap_inc(i, Mt), j);
  const double th_jp = gauge_angle(seed, draw, gchain, Mt, i, wrap_inc(j, Mx));
  double2 *xs = reinterpret_cast<double2 *>(x_all) + t;
  double2 v = *xs;
  v.x = mod_2pi(mod_2pi(v.x + th) - th_ip);
  v.y = mod_2pi(mod_2pi(v.y + th) - th_jp);
  *xs = v;
}

int check_even(mlmcpi_ctx *ctx, const mlmcpi_model *m) {
  const int c = m->coarsening;
  if ((c == MLMCPI_COARSEN_BOTH && (m->Mt_lat % 2 || m->Mx_lat % 2)) ||
      (c == MLMCPI_COARSEN_TEMPORAL && m->Mt_lat % 2) ||
      (c == MLMCPI_COARSEN_SPATIAL && m->Mx_lat % 2))
    return ctx_fail(ctx, MLMCPI_EINVAL, "lattice cannot be coarsened (odd extent)");
  if (c != MLMCPI_COARSEN_BOTH && c != MLMCPI_COARSEN_TEMPORAL && c != MLMCPI_COARSEN_SPATIAL)
    return ctx_fail(ctx, MLMCPI_EINVAL,
                    "invalid coarsening for quenched Schwinger model (both/temporal/spatial)");
  return 0;
}

long long n_coarse_sites(const mlmcpi_model *m) {
  const int Mtc = (m->coarsening == MLMCPI_COARSEN_SPATIAL) ? m->Mt_lat : m->Mt_lat / 2;
  const int Mxc = (m->coarsening == MLMCPI_COARSEN_TEMPORAL) ? m->Mx_lat : m->Mx_lat / 2;
  return (long long)Mtc * Mxc;
}

// one fused leapfrog step on all chains
int leapfrog_step(mlmcpi_ctx *ctx, const SW &sw, double dt_p, double dt_x, bool drift,
                  const double *x_in, double *x_out, double *p, int B) {
  const long long nsite = (long long)sw.Mt * sw.Mx;
  if (sw.Mt <= 1024 && sw.Mt % 32 == 0 && sw.Mx >= 3 && ctx->leapfrog_variant != 2) {
    // rows per block: small chunks keep the last wave short; the price is one extra
    // theta row read per chunk end (an L2 hit when the neighbouring chunk is in flight)
    int R = ctx->leapfrog_rows > 0 ? ctx->leapfrog_rows : (sw.Mt <= 128 ? 4 : 8);
    if (R > sw.Mx)
      R = sw.Mx;
    const int chunks = cdiv(sw.Mx, R);
    constexpr int S = 6;
    const size_t smem_pipe = (size_t)S * 32 * sw.Mt + 16 * sw.Mt + 8 * S;
    if (ctx->leapfrog_variant == 0 && smem_pipe <= 200 * 1024) {
      auto kern = drift ? leapfrog_rowpipe_kernel<true, S> : leapfrog_rowpipe_kernel<false, S>;
      if (smem_pipe > 48 * 1024)
        MLMCPI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_pipe));
      kern<<<chunks * B, sw.Mt, smem_pipe, ctx->stream>>>(sw, dt_p, dt_x, x_in, x_out, p, R, chunks);
      MLMCPI_LAUNCHED("schwinger::leapfrog_rowpipe");
    } else {
      const size_t smem = (size_t)4 * sw.Mt * sizeof(double);
      if (drift)
        leapfrog_rowmarch_kernel<true><<<chunks * B, sw.Mt, smem, ctx->stream>>>(
            sw, dt_p, dt_x, x_in, x_out, p, R, chunks);
      else
        leapfrog_rowmarch_kernel<false><<<chunks * B, sw.Mt, smem, ctx->stream>>>(
            sw, dt_p, dt_x, x_in, x_out, p, R, chunks);
      MLMCPI_LAUNCHED("schwinger::leapfrog_rowmarch");
    }
  } else {
    leapfrog_naive_kernel<<<cdiv(nsite * B, 256), 256, 0, ctx->stream>>>(
        sw, dt_p, dt_x, x_in, drift ? x_out : nullptr, p, B);
    MLMCPI_LAUNCHED("schwinger::leapfrog_naive");
  }
  return 0;
}

// two fused leapfrog steps (step A then step B) on all chains; false if the shape is not supported
bool leapfrog_pair_supported(const mlmcpi_ctx *ctx, const SW &sw) {
  return ctx->leapfrog_variant == 0 && ctx->leapfrog_fuse && sw.Mt <= 1024 && sw.Mt % 32 == 0 && sw.Mx >= 8;
}
int leapfrog_pair(mlmcpi_ctx *ctx, const SW &sw, double dtpA, double dtxA, double dtpB, double dtxB,
                  const double *x_in, double *x_out, const double *p_in, double *p_out, int B) {
  // halo overhead is 96/R bytes per site and two steps: 32 rows per block by default
  int R = ctx->leapfrog_rows > 0 ? ctx->leapfrog_rows : 32;
  if (R > sw.Mx)
    R = sw.Mx;
  const int chunks = cdiv(sw.Mx, R);
  constexpr int S = 6;
  const size_t smem = (size_t)S * 32 * sw.Mt + 48 * sw.Mt + 8 * S;
  auto kern = leapfrog_rowpipe2_kernel<S>;
  if (smem > 48 * 1024)
    MLMCPI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<chunks * B, sw.Mt, smem, ctx->stream>>>(sw, dtpA, dtxA, dtpB, dtxB, x_in, x_out, p_in, p_out, R,
                                                chunks);
  MLMCPI_LAUNCHED("schwinger::leapfrog_rowpipe2");
  return 0;
}

// K fused leapfrog steps on all chains (leapfrog_rowpipek_kernel); returns MLMCPI_EUNSUPPORTED when the shape
// has no specialisation (the caller then falls back to the pair / single-step kernels)
template <int K, int MT>
int leapfrog_multi_launch(mlmcpi_ctx *ctx, const SW &sw, const double *dtp, const double *dtx, const double *x_in,
                          double *x_out, const double *p_in, double *p_out, int B) {
  constexpr int S = 8;
  // rows per block: the halo costs (2K - 1) / R extra reads and K / R extra arithmetic; 64 rows measured best
  // for K = 4 on 128^2 x 512 chains (32: 28.0, 43: 27.5, 64: 25.5, 128: 28.5 us per step)
  int R = ctx->leapfrog_rows > 0 ? ctx->leapfrog_rows : (K >= 4 ? 64 : 32);
  if (R > sw.Mx)
    R = sw.Mx;
  const int chunks = cdiv(sw.Mx, R);
  LeapDt<K> dt;
  for (int k = 0; k < K; ++k) {
    dt.dtp[k] = dtp[k];
    dt.dtx[k] = dtx[k];
  }
  const size_t smem = (size_t)S * 32 * MT + (size_t)K * 32 * MT + 8 * S;
  auto kern = leapfrog_rowpipek_kernel<K, MT, S>;
  if (smem > 48 * 1024)
    MLMCPI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<chunks * B, MT, smem, ctx->stream>>>(sw.beta, dt, sw.Mx, x_in, x_out, p_in, p_out, R, chunks);
  MLMCPI_LAUNCHED("schwinger::leapfrog_rowpipek");
  return 0;
}
template <int K>
int leapfrog_multi(mlmcpi_ctx *ctx, const SW &sw, const double *dtp, const double *dtx, const double *x_in,
                   double *x_out, const double *p_in, double *p_out, int B) {
  if (sw.Mx < 2 * K + 2)
    return MLMCPI_EUNSUPPORTED;
  switch (sw.Mt) {
  case 64:
    return leapfrog_multi_launch<K, 64>(ctx, sw, dtp, dtx, x_in, x_out, p_in, p_out, B);
  case 128:
    return leapfrog_multi_launch<K, 128>(ctx, sw, dtp, dtx, x_in, x_out, p_in, p_out, B);
  case 256:
    return leapfrog_multi_launch<K, 256>(ctx, sw, dtp, dtx, x_in, x_out, p_in, p_out, B);
  case 512:
    return leapfrog_multi_launch<K, 512>(ctx, sw, dtp, dtx, x_in, x_out, p_in, p_out, B);
  }
  return MLMCPI_EUNSUPPORTED;
}
// eight steps per pass: Mt <= 128 only (the register queues of seven stages)
int leapfrog_multi8(mlmcpi_ctx *ctx, const SW &sw, const double *dtp, const double *dtx, const double *x_in,
                    double *x_out, const double *p_in, double *p_out, int B) {
  if (sw.Mx < 18)
    return MLMCPI_EUNSUPPORTED;
  if (sw.Mt == 64)
    return leapfrog_multi_launch<8, 64>(ctx, sw, dtp, dtx, x_in, x_out, p_in, p_out, B);
  if (sw.Mt == 128)
    return leapfrog_multi_launch<8, 128>(ctx, sw, dtp, dtx, x_in, x_out, p_in, p_out, B);
  return MLMCPI_EUNSUPPORTED;
}
// steps per HBM pass for this shape: MLMCPI_OPT_LEAPFROG_FUSE = 1 picks 4 where the register pipeline
// fits three blocks per SM (Mt <= 128) and 2 otherwise; 2, 3 (= 4 steps) select explicitly
int leapfrog_steps_per_pass(const mlmcpi_ctx *ctx, const SW &sw) {
  if (ctx->leapfrog_variant != 0 || !ctx->leapfrog_fuse)
    return 1;
  const bool specialised = sw.Mt == 64 || sw.Mt == 128 || sw.Mt == 256 || sw.Mt == 512;
  if (!specialised)
    return 0; // the generic pair kernel
  if (ctx->leapfrog_fuse == 4)
    return 0; // the round-1 two-step kernel
  if (ctx->leapfrog_fuse == 2)
    return 2;
  if (ctx->leapfrog_fuse == 3)
    return 4;
  if (ctx->leapfrog_fuse == 5)
    return 8;
  return sw.Mt <= 256 ? 4 : 2;
}

// trajectory of sampler/hmcsampler.cc:31-46.  x_first: state read by the first step;
// bufA/bufB: ping-pong trial buffers; returns the buffer holding the final state.
int trajectory(mlmcpi_ctx *ctx, const SW &sw, int nt, double dt, const double *x_first,
               double *bufA, double *bufB, double *p, int B, double **x_final) {
  const double *in = x_first;
  double *out = bufA;
  double *last = nullptr;
  const size_t n = (size_t)2 * sw.Mt * sw.Mx * B;
  // the fused two-step kernel ping-pongs p as well
  double *p_alt = nullptr, *p_cur = p;
  const bool fuse = leapfrog_pair_supported(ctx, sw) && nt >= 2 && sw.Mt * 240 + 64 <= 200 * 1024;
  if (fuse) {
    p_alt = ctx_work(ctx, 7, n);
    if (!p_alt)
      return MLMCPI_ENOMEM;
  }
  auto dtp = [&](int k) { return (k == 0 || k == nt) ? 0.5 * dt : dt; };
  auto dtx = [&](int k) { return (k == nt) ? 0.0 : dt; };
  prof_begin(ctx);
  uint64_t launches = 0;
  int k = 0;
  const int per_pass = fuse ? leapfrog_steps_per_pass(ctx, sw) : 1;
  while (k <= nt) {
    int rc;
    const int K = (per_pass >= 8 && k + 7 <= nt && sw.Mt <= 128)
                      ? 8
                      : ((per_pass >= 4 && k + 3 <= nt) ? 4 : ((per_pass >= 2 && k + 1 <= nt) ? 2 : 0));
    if (K) { // steps k .. k+K-1 in one pass (compile-time block size, K-stage register pipeline)
      double a[8], b[8];
      for (int q = 0; q < K; ++q) {
        a[q] = dtp(k + q);
        b[q] = dtx(k + q);
      }
      double *p_next = (p_cur == p) ? p_alt : p;
      rc = (K == 8) ? leapfrog_multi8(ctx, sw, a, b, in, out, p_cur, p_next, B)
                    : ((K == 4) ? leapfrog_multi<4>(ctx, sw, a, b, in, out, p_cur, p_next, B)
                                : leapfrog_multi<2>(ctx, sw, a, b, in, out, p_cur, p_next, B));
      if (rc == 0) {
        p_cur = p_next;
        last = out;
        in = out;
        out = (out == bufA) ? bufB : bufA;
        k += K;
        ++launches;
        continue;
      }
      if (rc != MLMCPI_EUNSUPPORTED)
        return rc;
    }
    if (fuse && k + 1 <= nt) { // steps k and k+1 in one pass
      double *p_next = (p_cur == p) ? p_alt : p;
      rc = leapfrog_pair(ctx, sw, dtp(k), dtx(k), dtp(k + 1), dtx(k + 1), in, out, p_cur, p_next, B);
      if (rc)
        return rc;
      p_cur = p_next;
      last = out;
      in = out;
      out = (out == bufA) ? bufB : bufA;
      k += 2;
    } else {
      const bool drift = (k != nt);
      // single step; p updated in place in whichever buffer currently holds it
      rc = leapfrog_step(ctx, sw, dtp(k), dtx(k), drift, in, out, p_cur, B);
      if (rc)
        return rc;
      if (drift) {
        last = out;
        in = out;
        out = (out == bufA) ? bufB : bufA;
      }
      k += 1;
    }
    ++launches;
  }
  if (p_cur != p) // hand the momenta back in the caller's buffer
    MLMCPI_CUDA(cudaMemcpyAsync(p, p_cur, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  // algorithmic bytes: R theta, R p, W theta, W p per site-step; the final kick writes no theta
  const double site_bytes = 16.0 * (double)sw.Mt * sw.Mx * B;
  prof_end(ctx, launches, site_bytes * (4.0 * nt + 3.0));
  *x_final = last; // nullptr when nt == 0 (state unchanged)
  return 0;
}

} // namespace

namespace schwinger {

int init_state(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, uint32_t chain0,
               uint64_t draw) {
  SW sw = make_sw(ctx, m);
  const long long n = (long long)sw.Mt * sw.Mx * B;
  init_state_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(sw, x, B, chain0, ctx->seed, draw);
  MLMCPI_LAUNCHED("schwinger::init_state");
  return 0;
}

int action(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *x, int B, double *S) {
  SW sw = make_sw(ctx, m);
  return plaq_reduce<0>(ctx, "schwinger::action", sw, x, B, EPI_SCALE, sw.beta, S, nullptr);
}

int force(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *x, double *f, int B) {
  SW sw = make_sw(ctx, m);
  const long long n = (long long)sw.Mt * sw.Mx * B;
  force_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(sw, x, f, B);
  MLMCPI_LAUNCHED("schwinger::force");
  return 0;
}

int leapfrog(mlmcpi_ctx *ctx, const mlmcpi_model *m, int nt, double dt, double *x, double *p,
             int B) {
  SW sw = make_sw(ctx, m);
  const size_t n = (size_t)2 * sw.Mt * sw.Mx * B;
  double *bufA = ctx_work(ctx, 1, n), *bufB = ctx_work(ctx, 2, n);
  if (!bufA || !bufB)
    return MLMCPI_ENOMEM;
  double *fin = nullptr;
  int rc = trajectory(ctx, sw, nt, dt, x, bufA, bufB, p, B, &fin);
  if (rc)
    return rc;
  if (fin)
    MLMCPI_CUDA(cudaMemcpyAsync(x, fin, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}

int hmc_momentum(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *p, int B, uint32_t chain0,
                 uint64_t draw) {
  SW sw = make_sw(ctx, m);
  const long long n = (long long)sw.Mt * sw.Mx * B;
  momentum_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(sw, p, B, chain0, ctx->seed, draw);
  MLMCPI_LAUNCHED("schwinger::hmc_momentum");
  return 0;
}

// HMCSampler::single_step, sampler/hmcsampler.cc:22-69
int hmc_step(mlmcpi_ctx *ctx, const mlmcpi_model *m, int nt, double dt, double *x, int B,
             uint32_t chain0, uint64_t draw, int32_t *accept, double *diag) {
  SW sw = make_sw(ctx, m);
  const size_t nd = (size_t)2 * sw.Mt * sw.Mx;
  const size_t n = nd * B;
  double *p = ctx_work(ctx, 0, n), *bufA = ctx_work(ctx, 1, n), *bufB = ctx_work(ctx, 2, n);
  double *red = ctx_work(ctx, 3, (size_t)5 * B); // S_cur S_trial T_cur T_trial | accept flags
  if (!p || !bufA || !bufB || !red)
    return MLMCPI_ENOMEM;
  double *S_cur = red, *S_trial = red + B, *T_cur = red + 2 * B, *T_trial = red + 3 * B;
  int32_t *acc = accept ? accept : reinterpret_cast<int32_t *>(red + 4 * B);
  int rc;
  if ((rc = hmc_momentum(ctx, m, p, B, chain0, draw)))
    return rc;
  if ((rc = launch_half_sqnorm(ctx, p, nd, B, T_cur)))
    return rc;
  if ((rc = action(ctx, m, x, B, S_cur)))
    return rc;
  double *fin = nullptr;
  if ((rc = trajectory(ctx, sw, nt, dt, x, bufA, bufB, p, B, &fin)))
    return rc;
  if ((rc = launch_half_sqnorm(ctx, p, nd, B, T_trial)))
    return rc;
  if ((rc = action(ctx, m, fin ? fin : x, B, S_trial)))
    return rc;
  if ((rc = launch_hmc_accept(ctx, B, chain0, draw, S_cur, S_trial, T_cur, T_trial, acc, diag)))
    return rc;
  if (fin)
    if ((rc = launch_masked_copy(ctx, x, fin, nd, B, acc)))
      return rc;
  return 0;
}

static int sweep(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, bool heatbath,
                 uint32_t chain0, uint64_t draw) {
  if (m->Mt_lat % 2 || m->Mx_lat % 2)
    return ctx_fail(ctx, MLMCPI_EINVAL, "coloured sweeps need even lattice extents");
  SW sw = make_sw(ctx, m);
  for (int pass = 0; pass < 4; ++pass) {
    const int colour = ctx->sweep_reverse ? 3 - pass : pass;
    const int per_row = (colour < 2) ? sw.Mt : sw.Mt / 2; // links of this colour in one row
    const int rows = (colour < 2) ? sw.Mx / 2 : sw.Mx;
    const int work = heatbath ? (per_row + 1) / 2 : per_row; // the heat bath updates the links in pairs
    const int threads = std::min(heatbath ? 128 : 256, ((work + 31) / 32) * 32);
    const dim3 grid(cdiv(work, threads), rows, std::min(B, 32768));
    if (rows > 65535)
      return ctx_fail(ctx, MLMCPI_EINVAL, "lattice too large for the sweep kernel");
    if (heatbath)
      heatbath_pair_kernel<<<grid, threads, 0, ctx->stream>>>(sw, colour, x, B, chain0, ctx->seed, draw);
    else
      sweep_colour_kernel<<<grid, threads, 0, ctx->stream>>>(sw, colour, x, B);
    MLMCPI_LAUNCHED("schwinger::sweep_colour");
  }
  return 0;
}

// n_sweeps overrelaxation sweeps.  Ascending colour order on lattices with even extents and at most
// 1024 columns: the one-pass kernel, ping-ponging between x and a work buffer (copied back when
// n_sweeps is odd); otherwise four colour passes per sweep, in place.
// HMCSampler::single_step without the commit: the trajectory's end state stays in a work buffer (*fin),
// d_accept tells which chains would take it, d_S_trial is its action.  d_S_cur: the action of x, known to the
// caller (the hierarchical cascade caches it: csrc/capi.cu:cascade_draw_cached).  Same variates, same
// arithmetic, same accept flags as hmc_step.
int hmc_trial(mlmcpi_ctx *ctx, const mlmcpi_model *m, int nt, double dt, const double *x, int B, uint32_t chain0,
              uint64_t draw, const double *d_S_cur, double *d_S_trial, int32_t *d_accept, const double **fin_out) {
  SW sw = make_sw(ctx, m);
  const size_t nd = (size_t)2 * sw.Mt * sw.Mx;
  const size_t n = nd * B;
  double *p = ctx_work(ctx, 0, n), *bufA = ctx_work(ctx, 1, n), *bufB = ctx_work(ctx, 2, n);
  double *red = ctx_work(ctx, 3, (size_t)5 * B);
  if (!p || !bufA || !bufB || !red)
    return MLMCPI_ENOMEM;
  double *T_cur = red + 2 * B, *T_trial = red + 3 * B;
  int rc;
  if ((rc = hmc_momentum(ctx, m, p, B, chain0, draw)))
    return rc;
  if ((rc = launch_half_sqnorm(ctx, p, nd, B, T_cur)))
    return rc;
  double *fin = nullptr;
  if ((rc = trajectory(ctx, sw, nt, dt, x, bufA, bufB, p, B, &fin)))
    return rc;
  if ((rc = launch_half_sqnorm(ctx, p, nd, B, T_trial)))
    return rc;
  if ((rc = action(ctx, m, fin ? fin : x, B, d_S_trial)))
    return rc;
  if ((rc = launch_hmc_accept(ctx, B, chain0, draw, d_S_cur, d_S_trial, T_cur, T_trial, d_accept, nullptr)))
    return rc;
  *fin_out = fin ? fin : x;
  return 0;
}

int dof_update(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, int ell, int heatbath, uint32_t chain0,
               uint64_t draw) {
  SW sw = make_sw(ctx, m);
  if (ell < 0 || ell >= 2 * sw.Mt * sw.Mx)
    return ctx_fail(ctx, MLMCPI_EINVAL, "link index out of range");
  if (heatbath)
    dof_update_kernel<true><<<cdiv(B, 128), 128, 0, ctx->stream>>>(sw, ell, x, B, chain0, ctx->seed, draw);
  else
    dof_update_kernel<false><<<cdiv(B, 128), 128, 0, ctx->stream>>>(sw, ell, x, B, 0, 0, 0);
  MLMCPI_LAUNCHED("schwinger::dof_update");
  return 0;
}

int overrelax_sweeps(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, int n_sweeps) {
  SW sw = make_sw(ctx, m);
  const bool one_pass = !ctx->sweep_reverse && ctx->overrelax_one_pass && sw.Mt % 2 == 0 && sw.Mx % 2 == 0 &&
                        sw.Mt <= 1024;
  if (!one_pass) {
    for (int k = 0; k < n_sweeps; ++k) {
      int rc = sweep(ctx, m, x, B, false, 0, 0);
      if (rc)
        return rc;
    }
    return 0;
  }
  const size_t n = (size_t)2 * sw.Mt * sw.Mx * B;
  double *tmp = ctx_work(ctx, 1, n);
  if (!tmp)
    return MLMCPI_ENOMEM;
  // rows per block: 3 halo rows are re-read per chunk (64 rows: 5 % extra reads; measured best at
  // 512^2 and 1024^2, profiles/r01_summary.md section 11), fewer when the grid would not fill the SMs
  int R = 64;
  while (R > 2 && (sw.Mx % R != 0))
    R /= 2;
  while (R > 8 && (long long)(sw.Mx / R) * B < 4LL * ctx->n_sm)
    R /= 2;
  if (sw.Mx % R != 0)
    R = sw.Mx;
  const int chunks = sw.Mx / R;
  const size_t smem = (size_t)9 * sw.Mt * sizeof(double);
  if (smem > 48 * 1024)
    MLMCPI_CUDA(cudaFuncSetAttribute(overrelax_rowpipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
  double *src = x, *dst = tmp;
  for (int k = 0; k < n_sweeps; ++k) {
    overrelax_rowpipe_kernel<<<chunks * B, sw.Mt, smem, ctx->stream>>>(sw, src, dst, R, chunks);
    MLMCPI_LAUNCHED("schwinger::overrelax_rowpipe");
    std::swap(src, dst);
  }
  if (src != x)
    MLMCPI_CUDA(cudaMemcpyAsync(x, src, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}

int overrelax_sweep(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B) {
  if (m->Mt_lat % 2 || m->Mx_lat % 2)
    return ctx_fail(ctx, MLMCPI_EINVAL, "coloured sweeps need even lattice extents");
  return overrelax_sweeps(ctx, m, x, B, 1);
}
int heatbath_sweep(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, uint32_t chain0,
                   uint64_t draw) {
  return sweep(ctx, m, x, B, true, chain0, draw);
}

int prolong(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *xc, double *x, int B) {
  int rc = check_even(ctx, m);
  if (rc)
    return rc;
  SW sw = make_sw(ctx, m);
  const long long n = n_coarse_sites(m) * B;
  prolong_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(sw, m->coarsening, xc, x, B);
  MLMCPI_LAUNCHED("schwinger::prolong");
  return 0;
}

int restrict_(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *xf, double *xc, int B) {
  int rc = check_even(ctx, m);
  if (rc)
    return rc;
  SW sw = make_sw(ctx, m);
  const long long n = n_coarse_sites(m) * B;
  restrict_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(sw, m->coarsening, xf, xc, B);
  MLMCPI_LAUNCHED("schwinger::restrict");
  return 0;
}

int fill(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, uint32_t chain0, uint64_t draw) {
  int rc = check_even(ctx, m);
  if (rc)
    return rc;
  SW sw = make_sw(ctx, m);
  const long long n = n_coarse_sites(m) * B;
  if (m->coarsening == MLMCPI_COARSEN_BOTH) {
    fill_both_step1_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(sw, x, B, chain0, ctx->seed, draw);
    MLMCPI_LAUNCHED("schwinger::fill_step1");
    BesselProductConst bp;
    if (sw.beta > 8.0) { // quenchedschwingerconditionedfineaction.hh:39-45
      bp.beta = sw.beta;
      fill_both_step23_kernel<true><<<cdiv(n, 128), 128, 0, ctx->stream>>>(sw, bp, x, B, chain0,
                                                                          ctx->seed, draw);
    } else {
      besselproduct_setup(sw.beta, &bp);
      fill_both_step23_kernel<false><<<cdiv(n, 128), 128, 0, ctx->stream>>>(sw, bp, x, B, chain0,
                                                                           ctx->seed, draw);
    }
    MLMCPI_LAUNCHED("schwinger::fill_step23");
  } else {
    for (int phase = 0; phase < 2; ++phase) {
      fill_semi_kernel<<<cdiv(n, 128), 128, 0, ctx->stream>>>(sw, m->coarsening, phase, x, B, chain0,
                                                             ctx->seed, draw);
      MLMCPI_LAUNCHED("schwinger::fill_semi");
    }
  }
  return 0;
}

// S_out: nullptr, or [2][B] receiving S_f(theta') and S_cond(theta') (fused evaluation); charge: [3][B], the third row
// sum_P mod_2pi(P) of theta' (coarsening both only)
static int prolong_fill_impl(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *xc, double *x, int B,
                             uint32_t chain0, uint64_t draw, double *S_out, bool charge = false,
                             const int32_t *mask = nullptr) {
  int rc = check_even(ctx, m);
  if (rc)
    return rc;
  if (m->coarsening != MLMCPI_COARSEN_BOTH) {
    if (charge)
      return ctx_fail(ctx, MLMCPI_EUNSUPPORTED, "fused topological charge: coarsening both only");
    if ((rc = prolong(ctx, m, xc, x, B)))
      return rc;
    if ((rc = fill(ctx, m, x, B, chain0, draw)))
      return rc;
    if (S_out) {
      if ((rc = action(ctx, m, x, B, S_out)))
        return rc;
      return cond_action(ctx, m, x, B, S_out + B);
    }
    return 0;
  }
  SW sw = make_sw(ctx, m);
  const int nblk = cdiv(n_coarse_sites(m), FILL_THREADS);
  const int grid = nblk * B;
  double *partial = nullptr;
  const int npart = nblk * (FILL_THREADS / 32); // one partial sum per warp
  if (S_out && !(partial = ctx_scratch(ctx, (size_t)(charge ? 3 : 2) * B * npart)))
    return MLMCPI_ENOMEM;
  BesselProductConst bp;
  if (sw.beta > 8.0) {
    bp.beta = sw.beta;
    if (S_out && charge)
      prolong_fill_both_kernel<true, true, true><<<grid, FILL_THREADS, 0, ctx->stream>>>(sw, bp, xc, x, B, chain0,
                                                                                       ctx->seed, draw, nblk, partial, mask);
    else if (S_out)
      prolong_fill_both_kernel<true, true><<<grid, FILL_THREADS, 0, ctx->stream>>>(sw, bp, xc, x, B, chain0, ctx->seed,
                                                                         draw, nblk, partial, mask);
    else
      prolong_fill_both_kernel<true, false><<<grid, FILL_THREADS, 0, ctx->stream>>>(sw, bp, xc, x, B, chain0, ctx->seed,
                                                                          draw, nblk, nullptr, mask);
  } else {
    besselproduct_setup(sw.beta, &bp);
    if (S_out && charge)
      prolong_fill_both_kernel<false, true, true><<<grid, FILL_THREADS, 0, ctx->stream>>>(sw, bp, xc, x, B, chain0,
                                                                                        ctx->seed, draw, nblk, partial, mask);
    else if (S_out)
      prolong_fill_both_kernel<false, true><<<grid, FILL_THREADS, 0, ctx->stream>>>(sw, bp, xc, x, B, chain0, ctx->seed,
                                                                          draw, nblk, partial, mask);
    else
      prolong_fill_both_kernel<false, false><<<grid, FILL_THREADS, 0, ctx->stream>>>(sw, bp, xc, x, B, chain0, ctx->seed,
                                                                           draw, nblk, nullptr, mask);
  }
  MLMCPI_LAUNCHED("schwinger::prolong_fill");
  if (S_out)
    return launch_reduce_finish(ctx, partial, npart, B, charge ? 3 : 2, EPI_SCALE, sw.beta, 1.0, S_out, nullptr, mask);
  return 0;
}

// prolong_fill_eval with a third output row: sum over the plaquettes of theta' of mod_2pi(P)  (S_out: [3][B])
int prolong_fill_eval_charge(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *xc, double *x, int B, uint32_t chain0,
                             uint64_t draw, double *S_out, const int32_t *mask) {
  return prolong_fill_impl(ctx, m, xc, x, B, chain0, draw, S_out, true, mask);
}
// prolong_fill_eval for the chains with mask[chain] != 0 only (nullptr: all); the S_out entries of the others are not written
int prolong_fill_eval_masked(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *xc, double *x, int B, uint32_t chain0,
                             uint64_t draw, double *S_out, const int32_t *mask) {
  return prolong_fill_impl(ctx, m, xc, x, B, chain0, draw, S_out, false,
                           m->coarsening == MLMCPI_COARSEN_BOTH ? mask : nullptr);
}

int prolong_fill(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *xc, double *x, int B,
                 uint32_t chain0, uint64_t draw) {
  return prolong_fill_impl(ctx, m, xc, x, B, chain0, draw, nullptr);
}

int prolong_fill_eval(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *xc, double *x, int B,
                      uint32_t chain0, uint64_t draw, double *S_out) {
  return prolong_fill_impl(ctx, m, xc, x, B, chain0, draw, S_out);
}

int cond_action(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *x, int B, double *S) {
  int rc = check_even(ctx, m);
  if (rc)
    return rc;
  SW sw = make_sw(ctx, m);
  const long long n = n_coarse_sites(m);
  if (m->coarsening == MLMCPI_COARSEN_BOTH) {
    if (sw.beta > 8.0)
      return site_reduce<1>(ctx, "schwinger::cond_action", CondBothApproxF{sw, x}, n, B, EPI_SCALE,
                            1.0, 0.0, S, nullptr);
    CondBothBesselF f;
    f.sw = sw;
    f.x_all = x;
    besselproduct_setup(sw.beta, &f.bp);
    return site_reduce<1>(ctx, "schwinger::cond_action", f, n, B, EPI_SCALE, 1.0, 0.0, S, nullptr);
  }
  return site_reduce<1>(ctx, "schwinger::cond_action", CondSemiF{sw, m->coarsening, x}, n, B,
                        EPI_SCALE, 1.0, 0.0, S, nullptr);
}

int from_cluster(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *psi, double *x, int B,
                 uint32_t chain0, uint64_t draw) {
  SW sw = make_sw(ctx, m);
  cluster_links_vertical_kernel<<<cdiv((long long)sw.Mx * B, 128), 128, 0, ctx->stream>>>(sw, psi, x, B);
  MLMCPI_LAUNCHED("schwinger::cluster_links_vertical");
  cluster_links_horizontal_kernel<<<cdiv(B, 64), 64, 0, ctx->stream>>>(sw, psi, x, B);
  MLMCPI_LAUNCHED("schwinger::cluster_links_horizontal");
  cluster_gauge_kernel<<<cdiv((long long)sw.Mt * sw.Mx * B, 256), 256, 0, ctx->stream>>>(
      sw, x, B, chain0, ctx->seed, draw);
  MLMCPI_LAUNCHED("schwinger::cluster_gauge");
  return 0;
}

int qoi(mlmcpi_ctx *ctx, const mlmcpi_model *m, int which, const double *x, int B, double *out,
        int64_t *Qint) {
  SW sw = make_sw(ctx, m);
  const long long n = (long long)sw.Mt * sw.Mx;
  if (which == MLMCPI_QOI_SCHWINGER_CHI)
    return plaq_reduce<2>(ctx, "schwinger::qoi_chi", sw, x, B, EPI_CHI, 0.25 / (M_PI * M_PI), out, Qint);
  if (which == MLMCPI_QOI_AVG_PLAQUETTE)
    return plaq_reduce<1>(ctx, "schwinger::qoi_plaq", sw, x, B, EPI_SCALE, 1.0 / ((double)sw.Mx * sw.Mt), out,
                          nullptr);
  return ctx_fail(ctx, MLMCPI_EINVAL, "QoI not defined for the Schwinger model");
}

} // namespace schwinger
