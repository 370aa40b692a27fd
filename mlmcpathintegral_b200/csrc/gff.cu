// gff.cu -- Gaussian free field on the 2-D vertex lattice (5-point action).
//
// State layout [chain][ell] with ell the reference's vertex index
// (lattice/lattice2d.hh:230-245: row-major on unrotated levels; even-even block
// followed by odd-odd block on the rotated levels of CoarsenRotate).  Neighbour
// indices are recomputed from (i,j) in registers instead of the reference's
// vector<vector<unsigned>> gather lists.
//
// The sweeps, force, fill-in and reductions implement the n_gibbs_smooth == 0 branch of GFFAction
// (the 5-point action).  The reference's coarse levels carry a dense N x N precision matrix Q_hat
// (qft/gffaction.cc:25-28, 133-174) and an exact Cholesky sampler (:200-213): these are built on
// the host for levels of at most MLMCPI_GFF_DENSE_MAX vertices (second half of this file); at the
// named 256^2 size neither the reference nor this library can form them.
//
// Reference citations relative to /root/reference/src.
#include <algorithm>
#include <cmath>
#include <vector>

#include <cublas_v2.h>
#include <cusolverDn.h>

#include "common.cuh"

namespace {

struct GF {
  int Mt, Mx, rotated, ctype, N;
  double mu2;
};

GF make_gf(const mlmcpi_model *m) {
  GF g;
  g.Mt = m->Mt_lat;
  g.Mx = m->Mx_lat;
  g.rotated = m->rotated;
  g.ctype = m->coarsening;
  g.N = m->rotated ? m->Mt_lat * m->Mx_lat / 2 : m->Mt_lat * m->Mx_lat;
  g.mu2 = m->gff_mu2;
  return g;
}

// lattice/lattice2d.hh:230-245 (i, j may be out of range by less than one period)
// (the periodic wraps are conditional subtractions, not integer remainders: i and j are at most one
// lattice spacing outside [0, Mt) x [0, Mx) -- a remainder by a run-time divisor costs ~25 instructions
// and nn_sum needs eight of them per vertex)
__device__ __forceinline__ int wrap_period(int v, int M) { return v < 0 ? v + M : (v >= M ? v - M : v); }
__device__ __forceinline__ int v_cart2lin(int Mt, int Mx, int rotated, int i, int j) {
  if (rotated) {
    const int Mth = Mt / 2, Mxh = Mx / 2;
    int is = ((i + Mt) - (i & 1)) >> 1; // in [Mth - 1, 2 Mth]
    int js = ((j + Mx) - (j & 1)) >> 1;
    is = is >= 2 * Mth ? is - 2 * Mth : (is >= Mth ? is - Mth : is);
    js = js >= 2 * Mxh ? js - 2 * Mxh : (js >= Mxh ? js - Mxh : js);
    return Mth * js + is + (Mt * Mx / 4) * (i & 1);
  }
  return Mt * wrap_period(j, Mx) + wrap_period(i, Mt);
}
// lattice/lattice2d.hh:255-268
__device__ __forceinline__ void v_lin2cart(int Mt, int Mx, int rotated, int ell, int &i, int &j) {
  if (rotated) {
    const int Mth = Mt / 2, quarter = Mt * Mx / 4;
    const int parity = ell / quarter;
    const int eh = ell - quarter * parity;
    const int jh = eh / Mth;
    j = 2 * jh + parity;
    i = 2 * (eh - Mth * jh) + parity;
  } else {
    j = ell / Mt;
    i = ell - Mt * j;
  }
}

// sum over the four nearest neighbours in the reference's order
// (lattice/lattice2d.cc:138-146), left-to-right accumulation as gffaction.cc:37-40
__device__ __forceinline__ double nn_sum(const GF &g, const double *x, int i, int j) {
  double d = 0.0;
  if (g.rotated) {
    d += x[v_cart2lin(g.Mt, g.Mx, 1, i + 1, j + 1)];
    d += x[v_cart2lin(g.Mt, g.Mx, 1, i + 1, j - 1)];
    d += x[v_cart2lin(g.Mt, g.Mx, 1, i - 1, j + 1)];
    d += x[v_cart2lin(g.Mt, g.Mx, 1, i - 1, j - 1)];
  } else {
    d += x[v_cart2lin(g.Mt, g.Mx, 0, i + 1, j)];
    d += x[v_cart2lin(g.Mt, g.Mx, 0, i - 1, j)];
    d += x[v_cart2lin(g.Mt, g.Mx, 0, i, j + 1)];
    d += x[v_cart2lin(g.Mt, g.Mx, 0, i, j - 1)];
  }
  return d;
}

// colour of a vertex for the checkerboard sweeps
__device__ __forceinline__ int colour_of(const GF &g, int i, int j) {
  return g.rotated ? (i & 1) : ((i + j) & 1);
}

// is (i,j) a coarse vertex, and what are the coarsening factors: lattice2d.cc:20-110
__device__ __forceinline__ bool is_coarse(const GF &g, int i, int j, int &rho_t, int &rho_x) {
  switch (g.ctype) {
  case MLMCPI_COARSEN_BOTH:
    rho_t = 2;
    rho_x = 2;
    break;
  case MLMCPI_COARSEN_TEMPORAL:
    rho_t = 2;
    rho_x = 1;
    break;
  case MLMCPI_COARSEN_SPATIAL:
    rho_t = 1;
    rho_x = 2;
    break;
  default: // ROTATE
    if (g.rotated) {
      rho_t = 2;
      rho_x = 2;
    } else {
      rho_t = 1;
      rho_x = 1;
      return ((i + j) & 1) == 0;
    }
  }
  return (i % rho_t == 0) && (j % rho_x == 0);
}

// One thread per vertex without integer divisions: grid (strips of a row, rows, chains).  On an
// unrotated level row y is lattice row j = y (ell = Mt j + i); on a rotated level the reference's index
// range is the even-even block followed by the odd-odd block (lattice2d.hh:230-245), each made of Mx/2
// rows of Mt/2 vertices: row y has parity p = (y >= Mx/2), j = 2 (y - p Mx/2) + p, i = 2 k + p.
// The body runs once per chain of the thread's z-slice; `t` is the index into a [chain][ell] array.
__device__ __forceinline__ bool vertex_of(const GF &g, int k, int y, int &i, int &j, int &ell) {
  if (g.rotated) {
    const int Mth = g.Mt / 2, Mxh = g.Mx / 2;
    const int par = y >= Mxh ? 1 : 0, jh = y - par * Mxh;
    i = 2 * k + par;
    j = 2 * jh + par;
    ell = (g.Mt * g.Mx / 4) * par + Mth * jh + k;
    return k < Mth;
  }
  i = k;
  j = y;
  ell = g.Mt * y + k;
  return k < g.Mt;
}
// the FINE-ONLY vertices of a level under CoarsenRotate (lattice2d.cc:20-110): on an unrotated level the
// vertices with i + j odd (half of every row), on a rotated level the odd-odd block (the second half of
// the index range).  Rows of Mt/2 vertices each; Mx (unrotated) or Mx/2 (rotated) rows.
__device__ __forceinline__ bool fine_vertex_of(const GF &g, int k, int y, int &i, int &j, int &ell) {
  const int Mth = g.Mt / 2;
  if (g.rotated) {
    i = 2 * k + 1;
    j = 2 * y + 1;
    ell = g.Mt * g.Mx / 4 + Mth * y + k;
  } else {
    i = 2 * k + ((y + 1) & 1);
    j = y;
    ell = g.Mt * y + i;
  }
  return k < Mth;
}
#define VERTEX_LOOP_BEGIN                                                                          \
  int i, j, ell;                                                                                   \
  if (!vertex_of(g, blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y, i, j, ell))                 \
    return;                                                                                        \
  for (long long chain = blockIdx.z; chain < B; chain += gridDim.z) {                              \
    const long long t = chain * g.N + ell;
#define VERTEX_LOOP_END }
#define VERTEX_THREADS 128
static inline dim3 vertex_grid(const GF &g, int B) {
  const int rowlen = g.rotated ? g.Mt / 2 : g.Mt;
  return dim3((unsigned)cdiv(rowlen, VERTEX_THREADS), (unsigned)g.Mx, (unsigned)std::min(B, 65535));
}

// start state: i.i.d. N(0, 1/(4+mu2)) (the reference draws exactly from the
// Gaussian by sparse Cholesky, qft/gffaction.cc:121-123, 200-213: SURVEY 8f-3)
__global__ void init_state_kernel(GF g, double *x, int B, uint32_t chain0, uint64_t seed,
                                  uint64_t draw) {
  const long long pairs = (g.N + 1) / 2;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= pairs * B)
    return;
  const long long chain = t / pairs;
  const int k = (int)(t - chain * pairs);
  Rng r = rng_init(seed, MLMCPI_STREAM_INIT, draw, chain0 + (uint32_t)chain, k);
  double z0, z1;
  rng_normal2(r, z0, z1);
  const double s = 1. / sqrt(4. + g.mu2);
  double *xc = x + chain * g.N;
  xc[2 * k] = s * z0;
  if (2 * k + 1 < g.N)
    xc[2 * k + 1] = s * z1;
}

__global__ void momentum_kernel(GF g, double *p, int B, uint32_t chain0, uint64_t seed,
                                uint64_t draw) {
  const long long pairs = (g.N + 1) / 2;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= pairs * B)
    return;
  const long long chain = t / pairs;
  const int k = (int)(t - chain * pairs);
  Rng r = rng_init(seed, MLMCPI_STREAM_HMC_MOMENTUM, draw, chain0 + (uint32_t)chain, k);
  double z0, z1;
  rng_normal2(r, z0, z1);
  double *pc = p + chain * g.N;
  pc[2 * k] = z0;
  if (2 * k + 1 < g.N)
    pc[2 * k + 1] = z1;
}

struct ActionF { // qft/gffaction.cc:7-24
  GF g;
  const double *x;
  __device__ void operator()(int chain, int i, int j, int ell, double acc[1]) const {
    const double *xc = x + (size_t)chain * g.N;
    const double phi = xc[ell];
    acc[0] += phi * ((4. + g.mu2) * phi - nn_sum(g, xc, i, j));
  }
};
struct Phi2F { // qoi/qft/qoi2dphisquared.cc:7-15
  GF g;
  const double *x;
  __device__ void operator()(int chain, int, int, int ell, double acc[1]) const {
    const double phi = x[(size_t)chain * g.N + ell];
    acc[0] += phi * phi;
  }
};
struct CondF { // qft/gffconditionedfineaction.cc:28-50
  GF g;
  const double *x;
  __device__ void operator()(int chain, int i, int j, int ell, double acc[1]) const {
    int rt, rx;
    if (is_coarse(g, i, j, rt, rx))
      return;
    const double *xc = x + (size_t)chain * g.N;
    const double sigma2 = 1. / (4. + g.mu2);
    const double dphi = xc[ell] - sigma2 * nn_sum(g, xc, i, j);
    acc[0] += 0.5 * (1. / sigma2) * dphi * dphi;
  }
};

// pass 1 of the deterministic two-pass reductions (common.cuh) over the vertices of a level, row by row
// without integer divisions (vertex_of): block (blk, chain) sums the rows blk, blk + nblk, ...
// FINE_ONLY: the fine-only vertices of CoarsenRotate (fine_vertex_of)
template <class F, bool FINE_ONLY>
__global__ void vertex_reduce_kernel(F f, GF g, int nblk, int B, double *partial) {
  const int blk = blockIdx.x;
  const int rowlen = (g.rotated || FINE_ONLY) ? g.Mt / 2 : g.Mt;
  const int nrows = (FINE_ONLY && g.rotated) ? g.Mx / 2 : g.Mx;
  for (int chain = blockIdx.y; chain < B; chain += gridDim.y) { // (gridDim.y = min(B, 65535))
    double acc[1] = {0.0};
    for (int y = blk; y < nrows; y += nblk)
      for (int k = threadIdx.x; k < rowlen; k += blockDim.x) {
        int i, j, ell;
        if (FINE_ONLY)
          fine_vertex_of(g, k, y, i, j, ell);
        else
          vertex_of(g, k, y, i, j, ell);
        f(chain, i, j, ell, acc);
      }
    const double v = block_sum(acc[0]);
    if (threadIdx.x == 0)
      partial[(size_t)chain * nblk + blk] = v;
  }
}
template <class F, bool FINE_ONLY = false>
int vertex_reduce(mlmcpi_ctx *ctx, const char *what, F f, const GF &g, int B, double scale, double *out) {
  const int rowlen = (g.rotated || FINE_ONLY) ? g.Mt / 2 : g.Mt;
  const int threads = std::min(256, std::max(32, ((rowlen + 31) / 32) * 32));
  int nblk = std::min(g.Mx, std::max(1, cdiv((long long)ctx->n_sm * 8, B)));
  double *partial = ctx_scratch(ctx, (size_t)B * nblk);
  if (!partial)
    return MLMCPI_ENOMEM;
  vertex_reduce_kernel<F, FINE_ONLY><<<dim3(nblk, std::min(B, 65535)), threads, 0, ctx->stream>>>(f, g, nblk, B, partial);
  MLMCPI_LAUNCHED(what);
  return launch_reduce_finish(ctx, partial, nblk, B, 1, EPI_SCALE, scale, 1.0, out, nullptr);
}

// qft/gffaction.cc:82-94
__global__ void force_kernel(GF g, const double *x, double *f, int B) {
  VERTEX_LOOP_BEGIN
  const double *xc = x + chain * g.N;
  f[t] = (4. + g.mu2) * xc[ell] - nn_sum(g, xc, i, j);
  VERTEX_LOOP_END
}

// sampler/hmcsampler.cc:43-45 fused (ping-pong phi buffers)
__global__ void leapfrog_kernel(GF g, double dt_p, double dt_x, const double *x_in, double *x_out,
                                double *p, int B) {
  VERTEX_LOOP_BEGIN
  const double *xc = x_in + chain * g.N;
  const double phi = xc[ell];
  const double F = (4. + g.mu2) * phi - nn_sum(g, xc, i, j);
  const double pn = p[t] - dt_p * F;
  p[t] = pn;
  if (x_out)
    x_out[t] = phi + dt_x * pn;
  VERTEX_LOOP_END
}

// the update of one vertex given the sum Delta over its four neighbours: heat bath
// (qft/gffaction.cc:32-42) or overrelaxation (:68-79); one function for every sweep kernel, so that they
// produce the same bits
// Heat-bath variates (stream convention, include/mlmcpi.h): the vertices ell = 2q and 2q + 1 share ONE Philox block,
// index 2q of the HEATBATH stream: its two Box-Muller normals are z0 for the even and z1 for the odd vertex.  The
// one-pass kernel, whose threads own such a pair in every row, draws the block once per pair.
__device__ __forceinline__ void heatbath_pair_normals(uint64_t seed, uint64_t draw, uint32_t gchain, int ell_even,
                                                      double &z0, double &z1) {
  Rng r = rng_init(seed, MLMCPI_STREAM_HEATBATH, draw, gchain, ell_even);
  rng_normal2(r, z0, z1);
}
__device__ __forceinline__ double site_heatbath_z(const GF &g, const double Delta, const double z) {
  return (1. / sqrt(4. + g.mu2)) * z + div_exact(Delta, 4. + g.mu2, 1. / (4. + g.mu2));
}
template <bool HEATBATH>
__device__ __forceinline__ double site_update(const GF &g, const double Delta, const double phi, uint64_t seed,
                                              uint64_t draw, uint32_t gchain, int ell) {
  if (HEATBATH) {
    double z0, z1;
    heatbath_pair_normals(seed, draw, gchain, ell & ~1, z0, z1);
    return site_heatbath_z(g, Delta, (ell & 1) ? z1 : z0);
  }
  return div_exact(2. * Delta, 4. + g.mu2, 1. / (4. + g.mu2)) - phi;
}

// qft/gffaction.cc:32-42, 68-79; one colour per launch, one thread per vertex OF THAT COLOUR.
// Unrotated level: grid (strips of half a row, rows, chains), vertex i = 2 k + ((colour + j) & 1).
// Rotated level: the vertices of colour c (= parity of i) are the contiguous half
// [c N/2, (c+1) N/2) of the reference's index range; grid (strips of that half, 1, chains).
template <bool HEATBATH>
__global__ void sweep_colour_kernel(GF g, int colour, double *x, int B, uint32_t chain0,
                                    uint64_t seed, uint64_t draw) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  int i, j, ell;
  if (g.rotated) {
    if (k >= g.N / 2)
      return;
    ell = colour * (g.N / 2) + k;
    v_lin2cart(g.Mt, g.Mx, 1, ell, i, j);
  } else {
    if (k >= g.Mt / 2)
      return;
    j = blockIdx.y;
    i = 2 * k + ((colour + j) & 1);
    ell = g.Mt * j + i;
  }
  for (int chain = blockIdx.z; chain < B; chain += gridDim.z) {
    double *xc = x + (size_t)chain * g.N;
    const double Delta = nn_sum(g, xc, i, j);
    xc[ell] = site_update<HEATBATH>(g, Delta, xc[ell], seed, draw, chain0 + (uint32_t)chain, ell);
  }
}

// Action::heatbath_update / overrelaxation_update of ONE vertex (qft/gffaction.cc:32-42, 68-79): the
// per-dof interface of action/action.hh:85-110, one thread per chain
template <bool HEATBATH>
__global__ void dof_update_kernel(GF g, int ell, double *x, int B, uint32_t chain0, uint64_t seed, uint64_t draw) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= B)
    return;
  int i, j;
  v_lin2cart(g.Mt, g.Mx, g.rotated, ell, i, j);
  double *xc = x + (size_t)chain * g.N;
  const double Delta = nn_sum(g, xc, i, j);
  xc[ell] = site_update<HEATBATH>(g, Delta, xc[ell], seed, draw, chain0 + (uint32_t)chain, ell);
}

// One sweep (both colours, ascending order; overrelaxation or heat bath) of an UNROTATED level in ONE pass
// over HBM, out of place.  The two colour passes above read the whole field twice to update half of it each
// (24 B per site against the algorithmic 16 B).  Here a block of Mt/2 threads marches over R rows of one
// chain; thread t owns the columns 2t and 2t+1 (one double2 per row: in every row one of the two is a
// colour-0 site, the other a colour-1 site, so all lanes work in both stages).  With r the row being
// finished: stage A = colour 0 on row r+1 (old neighbours: rows r, r+1, r+2), stage B = colour 1 on row r
// (new colour-0 neighbours: rows r-1, r, r+1).  The column neighbours of a site are the thread's own other
// column (register) and ONE value of an adjacent thread, exchanged through two small shared-memory rings
// (4 x Mt/2 old colour-1 values, 2 x Mt/2 new colour-0 values); row neighbours are registers.  One barrier
// per row.  Two rows below and two above the chunk are read in addition (their colour-0 values are
// recomputed, not stored): 8 (1 + 4/R) B read + 8 B written per site.  Every site is updated by
// site_update on the same operands in the same order as in sweep_colour_kernel: bit-identical.
template <bool HEATBATH>
__global__ void __launch_bounds__(1024)
    sweep_rowpipe_kernel(GF g, const double *__restrict__ x_in, double *__restrict__ x_out, int R, int chunks,
                         uint32_t chain0, uint64_t seed, uint64_t draw) {
  extern __shared__ __align__(16) double sm_gff[];
  const int Mt = g.Mt, Mx = g.Mx, H = Mt / 2;
  const int t = threadIdx.x;
  const int tp = t + 1 == H ? 0 : t + 1, tm = t == 0 ? H - 1 : t - 1;
  const int chain = blockIdx.x / chunks, chunk = blockIdx.x - chain * chunks;
  const int e0 = chunk * R; // even
  const int nrow = min(R, Mx - e0);
  const double2 *xin = reinterpret_cast<const double2 *>(x_in + (size_t)chain * g.N);
  double2 *xout = reinterpret_cast<double2 *>(x_out + (size_t)chain * g.N);
  double *c1val = sm_gff;                // [4][H]: old colour-1 value of this thread's pair, rows by (row & 3)
  double *n0row = sm_gff + (size_t)4 * H; // [2][H]: new colour-0 value, rows by (row & 1)
  const uint32_t gchain = chain0 + (uint32_t)chain;
  auto wrap = [&](int j) { return j < 0 ? j + Mx : (j >= Mx ? j - Mx : j); };
  // colour 0 on lattice row j (parity par = j & 1; Mx is even, so the parity survives the wrap):
  // below / own / above = old rows j-1, j, j+1 of this thread's pair
  // (slot = ring slot of that row: the UNWRAPPED row number & 3)
  // heat bath: the thread's two sites of row j, (2t, j) and (2t+1, j), are the vertex pair of one Philox block
  // (heatbath_pair_normals): stage A draws it for its colour-0 site and hands the other normal (zother) to stage B
  // of the same row, one loop iteration later
  auto stage_a = [&](int j, int slot, int par, double2 below, double2 own, double2 above, double &zother) {
    const double *cv = c1val + (size_t)slot * H;
    double z0 = 0.0, z1 = 0.0;
    if (HEATBATH)
      heatbath_pair_normals(seed, draw, gchain, Mt * j + 2 * t, z0, z1);
    double d = 0.0;
    if (par == 0) { // site 2t: neighbours (2t+1, j) own .y, (2t-1, j) of thread t-1, (2t, j+1), (2t, j-1)
      d += own.y;
      d += cv[tm];
      d += above.x;
      d += below.x;
      zother = z1;
      return HEATBATH ? site_heatbath_z(g, d, z0) : site_update<false>(g, d, own.x, 0, 0, 0, 0);
    }
    d += cv[tp]; // site 2t+1: neighbours (2t+2, j) of thread t+1, (2t, j) own .x
    d += own.x;
    d += above.y;
    d += below.y;
    zother = z0;
    return HEATBATH ? site_heatbath_z(g, d, z1) : site_update<false>(g, d, own.y, 0, 0, 0, 0);
  };
  int jl = wrap(e0 - 2);
  double2 o0 = xin[(size_t)jl * H + t]; // old rows e0-2, e0-1, e0, e0+1
  jl = wrap(e0 - 1);
  double2 o1 = xin[(size_t)jl * H + t];
  double2 o2 = xin[(size_t)e0 * H + t];
  double2 o3 = xin[(size_t)(e0 + 1) * H + t];
  // (e0 - 1) is odd: its colour-1 column is 2t (.x); e0 is even: 2t+1 (.y); e0 + 1 odd: .x
  c1val[(size_t)((e0 - 1) & 3) * H + t] = o1.x;
  c1val[(size_t)(e0 & 3) * H + t] = o2.y;
  c1val[(size_t)((e0 + 1) & 3) * H + t] = o3.x;
  // two rows are kept in flight per thread
  int jn = wrap(e0 + 2); // lattice row of the last row requested (e0 + 2 wraps to 0 at most)
  double2 pre = xin[(size_t)jn * H + t];
  jn = jn + 1 == Mx ? 0 : jn + 1;
  double2 pre2 = xin[(size_t)jn * H + t]; // row e0 + 3
  __syncthreads();
  double zc0, zc1, zunused;
  double nm1 = stage_a(wrap(e0 - 1), (e0 - 1) & 3, 1, o0, o1, o2, zunused); // new colour-0 values of rows e0-1 and e0
  double n0 = stage_a(e0, e0 & 3, 0, o1, o2, o3, zc0);                      // zc0: the normal of row e0's colour-1 site
  n0row[(size_t)(e0 & 1) * H + t] = n0;
  __syncthreads();
  // loop invariant for output row r: o2 = old(r), o3 = old(r+1), pre = old(r+2), nm1 / n0 = new colour 0 of
  // rows r-1 / r; c1val holds rows r-1 .. r+1, n0row row r
  for (int k = 0; k < nrow; ++k) {
    const int r = e0 + k, par = k & 1; // e0 even: parity of r
    const double2 o4 = pre;            // old row r + 2
    c1val[(size_t)((r + 2) & 3) * H + t] = par == 0 ? o4.y : o4.x;
    pre = pre2;
    if (k + 2 < nrow) { // row r + 4, the "row r + 2" of iteration k + 2
      jn = jn + 1 == Mx ? 0 : jn + 1;
      pre2 = xin[(size_t)jn * H + t];
    }
    const int r1 = r + 1 == Mx ? 0 : r + 1;
    const double n1 = stage_a(r1, (r + 1) & 3, par ^ 1, o2, o3, o4, zc1); // colour 0 on row r + 1
    n0row[(size_t)((r + 1) & 1) * H + t] = n1;
    // colour 1 on row r: the new colour-0 neighbours
    const double *nr = n0row + (size_t)(r & 1) * H;
    double d = 0.0, m;
    if (par == 0) { // site 2t+1: (2t+2, r) of thread t+1, (2t, r) own, (2t+1, r+1), (2t+1, r-1)
      d += nr[tp];
      d += n0;
      d += n1;
      d += nm1;
      m = HEATBATH ? site_heatbath_z(g, d, zc0) : site_update<false>(g, d, o2.y, 0, 0, 0, 0);
      xout[(size_t)r * H + t] = make_double2(n0, m);
    } else { // site 2t: (2t+1, r) own, (2t-1, r) of thread t-1
      d += n0;
      d += nr[tm];
      d += n1;
      d += nm1;
      m = HEATBATH ? site_heatbath_z(g, d, zc0) : site_update<false>(g, d, o2.x, 0, 0, 0, 0);
      xout[(size_t)r * H + t] = make_double2(m, n0);
    }
    __syncthreads();
    o2 = o3;
    o3 = o4;
    nm1 = n0;
    n0 = n1;
    zc0 = zc1;
  }
}

// qft/gffaction.cc:97-118 through the fine->coarse map of lattice2d.cc:121-130
template <bool TO_FINE>
__global__ void transfer_kernel(GF g, int Mtc, int Mxc, int rotc, int Nc, const double *src,
                                double *dst, int B) {
  VERTEX_LOOP_BEGIN
  int rt, rx;
  if (!is_coarse(g, i, j, rt, rx))
    return; // (the same for every chain)
  const int ellc = v_cart2lin(Mtc, Mxc, rotc, rt == 2 ? i >> 1 : i, rx == 2 ? j >> 1 : j);
  if (TO_FINE)
    dst[t] = src[chain * Nc + ellc];
  else
    dst[chain * Nc + ellc] = src[t];
  VERTEX_LOOP_END
}

// qft/gffconditionedfineaction.cc:7-25 for CoarsenRotate: one thread per FINE vertex (the generic kernel
// below leaves every other lane idle on the checkerboard)
__global__ void fill_rotate_kernel(GF g, double *x, int B, uint32_t chain0, uint64_t seed, uint64_t draw) {
  int i, j, ell;
  if (!fine_vertex_of(g, blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y, i, j, ell))
    return;
  const double sigma = 1. / sqrt(4. + g.mu2);
  for (long long chain = blockIdx.z; chain < B; chain += gridDim.z) {
    double *xc = x + chain * g.N;
    const double Delta = nn_sum(g, xc, i, j);
    Rng r = rng_init(seed, MLMCPI_STREAM_FILL1, draw, chain0 + (uint32_t)chain, ell);
    double z0, z1;
    rng_normal2(r, z0, z1);
    xc[ell] = sigma * (z0 + sigma * Delta);
  }
}

// qft/gffconditionedfineaction.cc:7-25
__global__ void fill_kernel(GF g, double *x, int B, uint32_t chain0, uint64_t seed, uint64_t draw) {
  VERTEX_LOOP_BEGIN
  int rt, rx;
  if (is_coarse(g, i, j, rt, rx))
    return; // (the same for every chain)
  double *xc = x + chain * g.N;
  const double Delta = nn_sum(g, xc, i, j);
  Rng r = rng_init(seed, MLMCPI_STREAM_FILL1, draw, chain0 + (uint32_t)chain, ell);
  double z0, z1;
  rng_normal2(r, z0, z1);
  const double sigma = 1. / sqrt(4. + g.mu2);
  xc[ell] = sigma * (z0 + sigma * Delta);
  VERTEX_LOOP_END
}

int coarse_dims(mlmcpi_ctx *ctx, const mlmcpi_model *m, int *Mtc, int *Mxc, int *rotc) {
  // level parity only matters for ROTATE, where rotated <=> odd level
  if (!mlmcpi_coarse_shape(m->Mt_lat, m->Mx_lat, m->coarsening, m->rotated ? 1 : 0, Mtc, Mxc, rotc))
    return ctx_fail(ctx, MLMCPI_EINVAL, "cannot coarsen 2d lattice");
  return 0;
}

int check_rotate(mlmcpi_ctx *ctx, const mlmcpi_model *m) {
  // the fill-in is an independent conditional only if all four nearest neighbours
  // of a fine-only vertex are coarse (qft/gffconditionedfineaction.hh:20-37)
  if (m->coarsening != MLMCPI_COARSEN_ROTATE)
    return ctx_fail(ctx, MLMCPI_EUNSUPPORTED, "GFF fill-in requires coarsening = rotate");
  return 0;
}

int step(mlmcpi_ctx *ctx, const GF &g, double dt_p, double dt_x, bool drift, const double *in,
         double *out, double *p, int B) {
  leapfrog_kernel<<<vertex_grid(g, B), VERTEX_THREADS, 0, ctx->stream>>>(g, dt_p, dt_x, in,
                                                                         drift ? out : nullptr, p, B);
  MLMCPI_LAUNCHED("gff::leapfrog");
  return 0;
}

int trajectory(mlmcpi_ctx *ctx, const GF &g, int nt, double dt, const double *x_first, double *bufA,
               double *bufB, double *p, int B, double **x_final) {
  const double *in = x_first;
  double *out = bufA, *last = nullptr;
  for (int k = 0; k <= nt; ++k) {
    const double dt_p = (k == 0 || k == nt) ? 0.5 * dt : dt;
    const bool drift = (k != nt);
    int rc = step(ctx, g, dt_p, drift ? dt : 0.0, drift, in, out, p, B);
    if (rc)
      return rc;
    if (drift) {
      last = out;
      in = out;
      out = (out == bufA) ? bufB : bufA;
    }
  }
  *x_final = last;
  return 0;
}

} // namespace


// ===================================================================== dense coarse levels
// GFFAction::buildMatrices, qft/gffaction.cc:133-174, ON THE DEVICE: the reference's coarse action is
// S = phi^T Q_hat phi / 2 with the dense N x N matrix
//     Q_hat = (Sigma_eff + G (Sigma - Sigma_eff) G^T)^{-1},  Sigma = Q^{-1}, Sigma_eff = Q_eff^{-1},
//     G = (1 - M^{-1} Q_eff)^{n_gibbs},  M = lower triangle of Q_eff (+ (1/omega - 1) diag),
// Q the 5-point and Q_eff the 9-point effective precision matrix.  G is a lexicographic Gauss-Seidel
// iteration matrix, so Q_hat is not translation invariant (no FFT shortcut); it is formed with dense
// linear algebra -- three Cholesky inverses (cuSOLVER potrf/potri), one triangular solve and three
// products (cuBLAS), ~10 N^3 flops: 0.3 s for 8192 vertices, ~15 s for the 32768 vertices of level 1 of
// BASELINE config C3 (four N x N work matrices: 34 GB of the 180 GB HBM) -- once per level and context.
// The reference needs the same matrices through Eigen on one core, O(hours) at that size.
// All matrices are symmetric or handled in column-major order (element (i, j) at j * N + i).
namespace {

__device__ __forceinline__ void neighbour8(const GF &g, int ell, int nb[8]) {
  int i, j;
  v_lin2cart(g.Mt, g.Mx, g.rotated, ell, i, j);
  if (g.rotated) { // neighbour order of lattice/lattice2d.cc:138-155
    nb[0] = v_cart2lin(g.Mt, g.Mx, 1, i + 1, j + 1);
    nb[1] = v_cart2lin(g.Mt, g.Mx, 1, i + 1, j - 1);
    nb[2] = v_cart2lin(g.Mt, g.Mx, 1, i - 1, j + 1);
    nb[3] = v_cart2lin(g.Mt, g.Mx, 1, i - 1, j - 1);
    nb[4] = v_cart2lin(g.Mt, g.Mx, 1, i + 2, j);
    nb[5] = v_cart2lin(g.Mt, g.Mx, 1, i - 2, j);
    nb[6] = v_cart2lin(g.Mt, g.Mx, 1, i, j + 2);
    nb[7] = v_cart2lin(g.Mt, g.Mx, 1, i, j - 2);
  } else {
    nb[0] = v_cart2lin(g.Mt, g.Mx, 0, i + 1, j);
    nb[1] = v_cart2lin(g.Mt, g.Mx, 0, i - 1, j);
    nb[2] = v_cart2lin(g.Mt, g.Mx, 0, i, j + 1);
    nb[3] = v_cart2lin(g.Mt, g.Mx, 0, i, j - 1);
    nb[4] = v_cart2lin(g.Mt, g.Mx, 0, i + 1, j + 1);
    nb[5] = v_cart2lin(g.Mt, g.Mx, 0, i + 1, j - 1);
    nb[6] = v_cart2lin(g.Mt, g.Mx, 0, i - 1, j + 1);
    nb[7] = v_cart2lin(g.Mt, g.Mx, 0, i - 1, j - 1);
  }
}

// GFFAction::buildPrecisionMatrix, gffaction.cc:177-197 (A zeroed beforehand; one thread owns one row,
// entries of coinciding neighbours on tiny lattices accumulate as setFromTriplets does)
__global__ void precision_matrix_kernel(GF g, double s0, double s1, double s2, int n_stencil, double *A) {
  const int ell = blockIdx.x * blockDim.x + threadIdx.x;
  if (ell >= g.N)
    return;
  int nb[8];
  neighbour8(g, ell, nb);
  double *row = A + (size_t)ell * g.N;
  row[ell] += s0;
  for (int k = 0; k < 4; ++k)
    row[nb[k]] += s1;
  if (n_stencil > 2)
    for (int k = 4; k < 8; ++k)
      row[nb[k]] += s2;
}
// potrf / potri work on the lower triangle: complete the symmetric matrix
__global__ void mirror_lower_kernel(int N, double *A) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)N * N)
    return;
  const int i = (int)(t % N), j = (int)(t / N);
  if (i < j) // (i, j) in the upper triangle <- (j, i)
    A[t] = A[(size_t)i * N + j];
}
__global__ void one_minus_kernel(int N, double *A) { // A <- 1 - A
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)N * N)
    return;
  const int i = (int)(t % N), j = (int)(t / N);
  A[t] = (i == j ? 1.0 : 0.0) - A[t];
}
__global__ void sub_kernel(size_t n, double *A, const double *B) { // A <- A - B
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n)
    A[t] -= B[t];
}
__global__ void add_diag_kernel(int N, double *A, const double *Dsrc, double f) { // A_ii += f * Dsrc_ii
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N)
    A[(size_t)i * N + i] += f * Dsrc[(size_t)i * N + i];
}
__global__ void symmetrise_kernel(int N, double *A) { // A <- (A + A^T) / 2, lower and upper at once
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)N * N)
    return;
  const int i = (int)(t % N), j = (int)(t / N);
  if (i > j) {
    const size_t u = (size_t)i * N + j;
    const double v = 0.5 * (A[t] + A[u]);
    A[t] = v;
    A[u] = v;
  }
}
__global__ void identity_kernel(int N, double *A) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)N * N)
    return;
  A[t] = (t % N == t / N) ? 1.0 : 0.0;
}

struct DenseLA {
  mlmcpi_ctx *ctx;
  int N;
  cublasHandle_t blas;
  cusolverDnHandle_t solver;
  int *info = nullptr;
  double *work = nullptr;
  int lwork = 0;
  std::vector<double *> owned;
  const char *fail = nullptr;

  bool init() {
    if (!ctx->cublas) {
      cublasHandle_t h;
      if (cublasCreate(&h) != CUBLAS_STATUS_SUCCESS)
        return false;
      ctx->cublas = h;
    }
    if (!ctx->cusolver) {
      cusolverDnHandle_t h;
      if (cusolverDnCreate(&h) != CUSOLVER_STATUS_SUCCESS)
        return false;
      ctx->cusolver = h;
    }
    blas = (cublasHandle_t)ctx->cublas;
    solver = (cusolverDnHandle_t)ctx->cusolver;
    cublasSetStream(blas, ctx->stream);
    cusolverDnSetStream(solver, ctx->stream);
    return cudaMalloc((void **)&info, sizeof(int)) == cudaSuccess;
  }
  double *matrix() {
    double *d = nullptr;
    if (cudaMalloc((void **)&d, (size_t)N * N * sizeof(double)) != cudaSuccess) {
      cudaGetLastError();
      fail = "out of device memory for the dense GFF matrices";
      return nullptr;
    }
    owned.push_back(d);
    return d;
  }
  void release(double *keep0 = nullptr, double *keep1 = nullptr) {
    for (double *d : owned)
      if (d != keep0 && d != keep1)
        cudaFree(d);
    owned.clear();
    if (info)
      cudaFree(info);
    if (work)
      cudaFree(work);
    info = nullptr;
    work = nullptr;
  }
  int grid2() const { return (int)(((size_t)N * N + 255) / 256); }
  bool ensure_work(int need) {
    if (need <= lwork)
      return true;
    if (work)
      cudaFree(work);
    work = nullptr;
    lwork = 0;
    if (cudaMalloc((void **)&work, (size_t)need * sizeof(double)) != cudaSuccess) {
      cudaGetLastError();
      fail = "out of device memory for the cuSOLVER workspace";
      return false;
    }
    lwork = need;
    return true;
  }
  bool check_info(const char *what) {
    int h = 0;
    cudaMemcpyAsync(&h, info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    if (h != 0) {
      fail = what;
      return false;
    }
    return true;
  }
  bool precision(double *A, const GF &g, double s0, double s1, double s2, int n_stencil) {
    cudaMemsetAsync(A, 0, (size_t)N * N * sizeof(double), ctx->stream);
    precision_matrix_kernel<<<cdiv(N, 128), 128, 0, ctx->stream>>>(g, s0, s1, s2, n_stencil, A);
    return cudaGetLastError() == cudaSuccess;
  }
  bool cholesky(double *A) { // lower factor in place
    int need = 0;
    if (cusolverDnDpotrf_bufferSize(solver, CUBLAS_FILL_MODE_LOWER, N, A, N, &need) != CUSOLVER_STATUS_SUCCESS ||
        !ensure_work(need))
      return false;
    if (cusolverDnDpotrf(solver, CUBLAS_FILL_MODE_LOWER, N, A, N, work, lwork, info) != CUSOLVER_STATUS_SUCCESS)
      return false;
    return check_info("GFF dense matrices: matrix not positive definite");
  }
  bool spd_inverse(double *A) { // symmetric positive definite inverse in place (full matrix on return)
    if (!cholesky(A))
      return false;
    int need = 0;
    if (cusolverDnDpotri_bufferSize(solver, CUBLAS_FILL_MODE_LOWER, N, A, N, &need) != CUSOLVER_STATUS_SUCCESS ||
        !ensure_work(need))
      return false;
    if (cusolverDnDpotri(solver, CUBLAS_FILL_MODE_LOWER, N, A, N, work, lwork, info) != CUSOLVER_STATUS_SUCCESS)
      return false;
    if (!check_info("GFF dense matrices: inversion failed"))
      return false;
    mirror_lower_kernel<<<grid2(), 256, 0, ctx->stream>>>(N, A);
    return cudaGetLastError() == cudaSuccess;
  }
  bool gemm(const double *A, const double *B, bool transpose_b, double beta, double *C) { // C = A op(B) + beta C
    const double one = 1.0;
    return cublasDgemm(blas, CUBLAS_OP_N, transpose_b ? CUBLAS_OP_T : CUBLAS_OP_N, N, N, N, &one, A, N, B, N,
                       &beta, C, N) == CUBLAS_STATUS_SUCCESS;
  }
};

// Q_hat of a coarse level (n_gibbs > 0)
int build_qhat(mlmcpi_ctx *ctx, const GF &g, int n_gibbs, double omega, double **out) {
  DenseLA la{ctx, g.N};
  *out = nullptr;
  if (!la.init())
    return ctx_fail(ctx, MLMCPI_ECUDA, "cannot create the cuBLAS / cuSOLVER handles");
  const int N = g.N;
  const size_t nn = (size_t)N * N;
  double *A = la.matrix(), *E = la.matrix(), *X = la.matrix(), *G = nullptr;
  bool ok = A && E && X;
  const double d = 4. + 0.5 * g.mu2;
  ok = ok && la.precision(A, g, 4. + g.mu2, -1., 0., 2);          // Q          (gffaction.cc:137-140)
  ok = ok && la.precision(E, g, d - 4. / d, -2. / d, -1. / d, 3); // Q_eff      (:143-148)
  // G1 = 1 - M^{-1} Q_eff (:150-159): X <- Q_eff, solve M X = Q_eff with M = the lower triangle of E
  if (ok) {
    cudaMemcpyAsync(X, E, nn * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream);
    const double *Mtri = E;
    double *Mmod = nullptr;
    if (std::fabs(omega - 1.0) > 1.E-14) { // M += (1/omega - 1) diag(Q_eff)
      Mmod = la.matrix();
      ok = Mmod != nullptr;
      if (ok) {
        cudaMemcpyAsync(Mmod, E, nn * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream);
        add_diag_kernel<<<cdiv(N, 128), 128, 0, ctx->stream>>>(N, Mmod, E, 1. / omega - 1.);
        Mtri = Mmod;
      }
    }
    const double one = 1.0;
    ok = ok && cublasDtrsm(la.blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, N, N,
                           &one, Mtri, N, X, N) == CUBLAS_STATUS_SUCCESS;
    if (ok)
      one_minus_kernel<<<la.grid2(), 256, 0, ctx->stream>>>(N, X);
    if (Mmod) {
      cudaStreamSynchronize(ctx->stream);
      cudaFree(Mmod);
      la.owned.pop_back();
    }
  }
  // G = G1^{n_gibbs} (:158-160); for the reference's n_gibbs = 2 this is one product into one new matrix
  G = X;
  {
    double *P[2] = {nullptr, nullptr};
    for (int k = 1; k < n_gibbs && ok; ++k) {
      const int dst = (G == P[0]) ? 1 : 0;
      if (!P[dst])
        P[dst] = la.matrix();
      ok = P[dst] != nullptr && la.gemm(G, X, false, 0.0, P[dst]);
      G = P[dst];
    }
    for (double *f : P)
      if (f && f != G) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(f);
        la.owned.erase(std::find(la.owned.begin(), la.owned.end(), f));
      }
  }
  // Sigma, Sigma_eff (:141, :149), D = Sigma - Sigma_eff in A
  ok = ok && la.spd_inverse(A) && la.spd_inverse(E);
  if (ok)
    sub_kernel<<<la.grid2(), 256, 0, ctx->stream>>>(nn, A, E);
  // Sigma_hat = Sigma_eff + G D G^T (:164-165), accumulated into E; T = G D
  if (ok) {
    double *T = (G == X) ? la.matrix() : X;
    ok = T != nullptr;
    ok = ok && la.gemm(G, A, false, 0.0, T) && la.gemm(T, G, true, 1.0, E);
  }
  if (ok) // symmetrise the rounding before the Cholesky inverse
    symmetrise_kernel<<<la.grid2(), 256, 0, ctx->stream>>>(N, E);
  ok = ok && la.spd_inverse(E); // Q_hat
  ok = ok && cudaStreamSynchronize(ctx->stream) == cudaSuccess;
  if (!ok) {
    const char *why = la.fail ? la.fail : "cuBLAS / cuSOLVER call failed";
    la.release();
    return ctx_fail(ctx, MLMCPI_EINVAL, "GFF dense matrices", why);
  }
  la.release(E);
  *out = E;
  return 0;
}

// U^{-1} with Q = L L^T = U^T U (the exact sampler solves U phi = psi, gffaction.cc:166-173,200-208):
// stored so that element [j * N + i] = (L^{-1})(j, i) = (U^{-1})(i, j)
int build_uinv(mlmcpi_ctx *ctx, const GF &g, double **out) {
  DenseLA la{ctx, g.N};
  *out = nullptr;
  if (!la.init())
    return ctx_fail(ctx, MLMCPI_ECUDA, "cannot create the cuBLAS / cuSOLVER handles");
  const int N = g.N;
  double *A = la.matrix(), *Z = la.matrix();
  bool ok = A && Z;
  ok = ok && la.precision(A, g, 4. + g.mu2, -1., 0., 2) && la.cholesky(A);
  if (ok) {
    identity_kernel<<<la.grid2(), 256, 0, ctx->stream>>>(N, Z);
    // column-major Z = L^{-T}: solve L^T Z = 1
    const double one = 1.0;
    ok = cublasDtrsm(la.blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, N, N, &one,
                     A, N, Z, N) == CUBLAS_STATUS_SUCCESS;
  }
  ok = ok && cudaStreamSynchronize(ctx->stream) == cudaSuccess;
  if (!ok) {
    const char *why = la.fail ? la.fail : "cuBLAS / cuSOLVER call failed";
    la.release();
    return ctx_fail(ctx, MLMCPI_EINVAL, "GFF exact sampler", why);
  }
  la.release(Z);
  *out = Z;
  return 0;
}

// S = phi^T Q_hat phi / 2 (gffaction.cc:25-28): one block per chain, phi staged in shared
// memory, thread i forms row i of Q_hat phi (Q_hat is symmetric: column reads are coalesced)
__global__ void dense_action_kernel(int N, const double *__restrict__ Qhat, const double *x, double *S) {
  extern __shared__ double phi[];
  const int chain = blockIdx.x;
  for (int k = threadIdx.x; k < N; k += blockDim.x)
    phi[k] = x[(size_t)chain * N + k];
  __syncthreads();
  double acc = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    double s = 0.0;
    for (int j = 0; j < N; ++j)
      s += Qhat[(size_t)j * N + i] * phi[j];
    acc += phi[i] * s;
  }
  const double v = block_sum(acc);
  if (threadIdx.x == 0)
    S[chain] = 0.5 * v;
}

// phi = U^{-1} psi (gffaction.cc:200-208), psi i.i.d. N(0,1): phi_i = sum_{j >= i} Uinv[i][j] psi_j
__global__ void dense_draw_kernel(int N, const double *__restrict__ UinvT, double *x, uint32_t chain0,
                                  uint64_t seed, uint64_t draw) {
  extern __shared__ double psi[];
  const int chain = blockIdx.x;
  for (int k = threadIdx.x; 2 * k < N; k += blockDim.x) {
    Rng r = rng_init(seed, MLMCPI_STREAM_EXACT, draw, chain0 + chain, k);
    double z0, z1;
    rng_normal2(r, z0, z1);
    psi[2 * k] = z0;
    if (2 * k + 1 < N)
      psi[2 * k + 1] = z1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    double s = 0.0;
    for (int j = i; j < N; ++j)
      s += UinvT[(size_t)j * N + i] * psi[j];
    x[(size_t)chain * N + i] = s;
  }
}

// GFFAction::global_heatbath_update_eff, gffaction.cc:45-65: a lexicographic (sequential) Gibbs
// sweep with the 9-point effective action -- one thread walks one chain
__global__ void gibbs_eff_kernel(GF g, double omega, int sweep, double *x_all, int B, uint32_t chain0,
                                 uint64_t seed, uint64_t draw) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= B)
    return;
  double *x = x_all + (size_t)chain * g.N;
  const double d = 4. + 0.5 * g.mu2;
  const double diag_eff = d - 4. / d;
  const double sigma_eff = 1. / sqrt(diag_eff);
  const double kappa = omega / d;
  const double gamma = sqrt(omega * (2. - omega));
  for (int ell = 0; ell < g.N; ++ell) {
    int i, j;
    v_lin2cart(g.Mt, g.Mx, g.rotated, ell, i, j);
    double Delta = (1. - omega) * diag_eff * x[ell];
    double nn = 0.0, nnn = 0.0;
    if (g.rotated) { // neighbour order of lattice/lattice2d.cc:138-155
      nn += x[v_cart2lin(g.Mt, g.Mx, 1, i + 1, j + 1)];
      nn += x[v_cart2lin(g.Mt, g.Mx, 1, i + 1, j - 1)];
      nn += x[v_cart2lin(g.Mt, g.Mx, 1, i - 1, j + 1)];
      nn += x[v_cart2lin(g.Mt, g.Mx, 1, i - 1, j - 1)];
      nnn += x[v_cart2lin(g.Mt, g.Mx, 1, i + 2, j)];
      nnn += x[v_cart2lin(g.Mt, g.Mx, 1, i - 2, j)];
      nnn += x[v_cart2lin(g.Mt, g.Mx, 1, i, j + 2)];
      nnn += x[v_cart2lin(g.Mt, g.Mx, 1, i, j - 2)];
    } else {
      nn += x[v_cart2lin(g.Mt, g.Mx, 0, i + 1, j)];
      nn += x[v_cart2lin(g.Mt, g.Mx, 0, i - 1, j)];
      nn += x[v_cart2lin(g.Mt, g.Mx, 0, i, j + 1)];
      nn += x[v_cart2lin(g.Mt, g.Mx, 0, i, j - 1)];
      nnn += x[v_cart2lin(g.Mt, g.Mx, 0, i + 1, j + 1)];
      nnn += x[v_cart2lin(g.Mt, g.Mx, 0, i + 1, j - 1)];
      nnn += x[v_cart2lin(g.Mt, g.Mx, 0, i - 1, j + 1)];
      nnn += x[v_cart2lin(g.Mt, g.Mx, 0, i - 1, j - 1)];
    }
    Delta += 2. * kappa * nn + kappa * nnn;
    Rng r = rng_init(seed, MLMCPI_STREAM_EXACT, draw, chain0 + chain, (uint32_t)((sweep + 1) * g.N + ell));
    double z0, z1;
    rng_normal2(r, z0, z1);
    x[ell] = sigma_eff * (gamma * z0 + sigma_eff * Delta);
  }
}

// lazily built, cached per (lattice, mu2, n_gibbs, omega) in the context: Q_hat and / or U^{-1}
int dense_get(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double **Qhat, const double **UinvT) {
  const GF g = make_gf(m);
  if (g.N > MLMCPI_GFF_DENSE_MAX)
    return ctx_fail(ctx, MLMCPI_EUNSUPPORTED,
                    "GFF dense matrices: more than MLMCPI_GFF_DENSE_MAX vertices on this level");
  const std::array<double, 6> key = {(double)g.Mt,          (double)g.Mx, (double)g.rotated, g.mu2,
                                     (double)m->gff_n_gibbs, m->gff_omega};
  auto it = ctx->gff_dense.find(key);
  if (it == ctx->gff_dense.end())
    it = ctx->gff_dense.emplace(key, std::array<double *, 2>{nullptr, nullptr}).first;
  int rc;
  if (Qhat) {
    if (m->gff_n_gibbs <= 0)
      return ctx_fail(ctx, MLMCPI_EINVAL, "GFF dense action requested for a level without Gibbs smoothing");
    if (!it->second[0] && (rc = build_qhat(ctx, g, m->gff_n_gibbs, m->gff_omega, &it->second[0])))
      return rc;
    *Qhat = it->second[0];
  }
  if (UinvT) {
    if (!it->second[1] && (rc = build_uinv(ctx, g, &it->second[1])))
      return rc;
    *UinvT = it->second[1];
  }
  return 0;
}

// levels above this size evaluate Q_hat Phi / U^{-1} Psi for all chains with one DGEMM
constexpr int DENSE_GEMM_MIN = 2048;

// S[b] = 1/2 sum_i phi[b][i] y[b][i]
__global__ void half_dot_kernel(int N, const double *x, const double *y, double *S) {
  const int chain = blockIdx.x;
  double acc = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    acc += x[(size_t)chain * N + i] * y[(size_t)chain * N + i];
  const double v = block_sum(acc);
  if (threadIdx.x == 0)
    S[chain] = 0.5 * v;
}
// psi i.i.d. N(0,1), the variates of dense_draw_kernel
__global__ void normal_fill_kernel(int N, double *psi, int B, uint32_t chain0, uint64_t seed, uint64_t draw) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int half = (N + 1) / 2;
  if (t >= (long long)half * B)
    return;
  const int chain = (int)(t / half), k = (int)(t - (long long)chain * half);
  Rng r = rng_init(seed, MLMCPI_STREAM_EXACT, draw, chain0 + chain, k);
  double z0, z1;
  rng_normal2(r, z0, z1);
  psi[(size_t)chain * N + 2 * k] = z0;
  if (2 * k + 1 < N)
    psi[(size_t)chain * N + 2 * k + 1] = z1;
}

} // namespace

namespace gff {

void release_dense(mlmcpi_ctx *ctx) {
  for (auto &kv : ctx->gff_dense)
    for (double *d : kv.second)
      if (d)
        cudaFree(d);
  ctx->gff_dense.clear();
  if (ctx->cublas)
    cublasDestroy((cublasHandle_t)ctx->cublas);
  if (ctx->cusolver)
    cusolverDnDestroy((cusolverDnHandle_t)ctx->cusolver);
  ctx->cublas = ctx->cusolver = nullptr;
}

int init_state(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, uint32_t chain0,
               uint64_t draw) {
  GF g = make_gf(m);
  const long long n = (long long)((g.N + 1) / 2) * B;
  init_state_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(g, x, B, chain0, ctx->seed, draw);
  MLMCPI_LAUNCHED("gff::init_state");
  return 0;
}

int action(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *x, int B, double *S) {
  GF g = make_gf(m);
  if (m->gff_n_gibbs > 0) { // coarse level: S = phi^T Q_hat phi / 2, gffaction.cc:25-28
    const double *Qhat = nullptr;
    int rc = dense_get(ctx, m, &Qhat, nullptr);
    if (rc)
      return rc;
    if (g.N >= DENSE_GEMM_MIN) { // Y = Q_hat Phi for all chains (states [B][N] = column-major N x B)
      double *Y = ctx_work(ctx, 8, (size_t)g.N * B);
      if (!Y)
        return ctx_fail(ctx, MLMCPI_ENOMEM, "out of device memory for the dense GFF action");
      const double one = 1.0, zero = 0.0;
      cublasHandle_t blas = (cublasHandle_t)ctx->cublas;
      cublasSetStream(blas, ctx->stream);
      if (cublasDgemm(blas, CUBLAS_OP_N, CUBLAS_OP_N, g.N, B, g.N, &one, Qhat, g.N, x, g.N, &zero, Y, g.N) !=
          CUBLAS_STATUS_SUCCESS)
        return ctx_fail(ctx, MLMCPI_ECUDA, "cublasDgemm (dense GFF action)");
      ctx->launches++;
      half_dot_kernel<<<B, 256, 0, ctx->stream>>>(g.N, x, Y, S);
      MLMCPI_LAUNCHED("gff::dense_action_dot");
      return 0;
    }
    const int threads = std::min(512, ((g.N + 31) / 32) * 32);
    dense_action_kernel<<<B, threads, (size_t)g.N * sizeof(double), ctx->stream>>>(g.N, Qhat, x, S);
    MLMCPI_LAUNCHED("gff::dense_action");
    return 0;
  }
  return vertex_reduce(ctx, "gff::action", ActionF{g, x}, g, B, 0.5, S);
}

// GFFAction::draw, qft/gffaction.cc:200-213
int exact_draw(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, uint32_t chain0, uint64_t draw) {
  GF g = make_gf(m);
  const double *UinvT = nullptr;
  int rc = dense_get(ctx, m, nullptr, &UinvT);
  if (rc)
    return rc;
  if (g.N >= DENSE_GEMM_MIN) { // phi = U^{-1} psi for all chains with one DGEMM
    double *Psi = ctx_work(ctx, 8, (size_t)g.N * B);
    if (!Psi)
      return ctx_fail(ctx, MLMCPI_ENOMEM, "out of device memory for the exact GFF draw");
    normal_fill_kernel<<<cdiv((long long)((g.N + 1) / 2) * B, 256), 256, 0, ctx->stream>>>(g.N, Psi, B, chain0,
                                                                                        ctx->seed, draw);
    MLMCPI_LAUNCHED("gff::normal_fill");
    const double one = 1.0, zero = 0.0;
    cublasHandle_t blas = (cublasHandle_t)ctx->cublas;
    cublasSetStream(blas, ctx->stream);
    // UinvT[j * N + i] = U^{-1}(i, j): as a column-major matrix it IS U^{-1}
    if (cublasDgemm(blas, CUBLAS_OP_N, CUBLAS_OP_N, g.N, B, g.N, &one, UinvT, g.N, Psi, g.N, &zero, x, g.N) !=
        CUBLAS_STATUS_SUCCESS)
      return ctx_fail(ctx, MLMCPI_ECUDA, "cublasDgemm (exact GFF draw)");
    ctx->launches++;
  } else {
    const int threads = std::min(512, ((g.N + 31) / 32) * 32);
    dense_draw_kernel<<<B, threads, (size_t)(g.N + 1) * sizeof(double), ctx->stream>>>(g.N, UinvT, x, chain0,
                                                                                      ctx->seed, draw);
    MLMCPI_LAUNCHED("gff::dense_draw");
  }
  for (int k = 0; k < m->gff_n_gibbs; ++k) {
    gibbs_eff_kernel<<<cdiv(B, 32), 32, 0, ctx->stream>>>(g, m->gff_omega, k, x, B, chain0, ctx->seed, draw);
    MLMCPI_LAUNCHED("gff::gibbs_eff");
  }
  return 0;
}

int force(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *x, double *f, int B) {
  GF g = make_gf(m);
  force_kernel<<<vertex_grid(g, B), VERTEX_THREADS, 0, ctx->stream>>>(g, x, f, B);
  MLMCPI_LAUNCHED("gff::force");
  return 0;
}

int leapfrog(mlmcpi_ctx *ctx, const mlmcpi_model *m, int nt, double dt, double *x, double *p,
             int B) {
  GF g = make_gf(m);
  const size_t n = (size_t)g.N * B;
  double *bufA = ctx_work(ctx, 1, n), *bufB = ctx_work(ctx, 2, n);
  if (!bufA || !bufB)
    return MLMCPI_ENOMEM;
  double *fin = nullptr;
  int rc = trajectory(ctx, g, nt, dt, x, bufA, bufB, p, B, &fin);
  if (rc)
    return rc;
  if (fin)
    MLMCPI_CUDA(cudaMemcpyAsync(x, fin, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}

int hmc_momentum(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *p, int B, uint32_t chain0,
                 uint64_t draw) {
  GF g = make_gf(m);
  const long long n = (long long)((g.N + 1) / 2) * B;
  momentum_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(g, p, B, chain0, ctx->seed, draw);
  MLMCPI_LAUNCHED("gff::hmc_momentum");
  return 0;
}

int hmc_step(mlmcpi_ctx *ctx, const mlmcpi_model *m, int nt, double dt, double *x, int B,
             uint32_t chain0, uint64_t draw, int32_t *accept, double *diag) {
  GF g = make_gf(m);
  const size_t nd = g.N, n = nd * B;
  double *p = ctx_work(ctx, 0, n), *bufA = ctx_work(ctx, 1, n), *bufB = ctx_work(ctx, 2, n);
  double *red = ctx_work(ctx, 3, (size_t)5 * B);
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
  if ((rc = trajectory(ctx, g, nt, dt, x, bufA, bufB, p, B, &fin)))
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
  GF g = make_gf(m);
  const int per_x = g.rotated ? g.N / 2 : g.Mt / 2;
  const int threads = std::min(256, ((per_x + 31) / 32) * 32);
  const dim3 grid(cdiv(per_x, threads), g.rotated ? 1 : g.Mx, std::min(B, 32768));
  if (grid.y > 65535)
    return ctx_fail(ctx, MLMCPI_EINVAL, "lattice too large for the sweep kernel");
  for (int pass = 0; pass < 2; ++pass) {
    const int colour = ctx->sweep_reverse ? 1 - pass : pass;
    if (heatbath)
      sweep_colour_kernel<true><<<grid, threads, 0, ctx->stream>>>(g, colour, x, B, chain0, ctx->seed, draw);
    else
      sweep_colour_kernel<false><<<grid, threads, 0, ctx->stream>>>(g, colour, x, B, 0, 0, 0);
    MLMCPI_LAUNCHED("gff::sweep_colour");
  }
  return 0;
}
// n_or overrelaxation sweeps followed by n_hb heat-bath sweeps (draw counters hb_draws[k]): the one-pass
// kernel on unrotated levels with even extents (ascending colour order), ping-ponging between x and a
// work buffer, one copy back when the number of sweeps is odd; otherwise two colour passes per sweep.
int sweep_sequence(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, int n_or, int n_hb, uint32_t chain0,
                   const uint64_t *hb_draws) {
  GF g = make_gf(m);
  const bool one_pass = !ctx->sweep_reverse && ctx->overrelax_one_pass && !g.rotated && g.Mt % 2 == 0 &&
                        g.Mx % 2 == 0 && g.Mt >= 64 && g.Mt <= 2048 && g.Mx >= 4;
  if (!one_pass) {
    int rc = 0;
    for (int k = 0; k < n_or && !rc; ++k)
      rc = sweep(ctx, m, x, B, false, 0, 0);
    for (int k = 0; k < n_hb && !rc; ++k)
      rc = sweep(ctx, m, x, B, true, chain0, hb_draws[k]);
    return rc;
  }
  if (n_or + n_hb == 0)
    return 0;
  const size_t n = (size_t)g.N * B;
  double *tmp = ctx_work(ctx, 1, n);
  if (!tmp)
    return MLMCPI_ENOMEM;
  int R = 64;
  while (R > 2 && g.Mx % R != 0)
    R /= 2;
  while (R > 8 && (long long)(g.Mx / R) * B < 4LL * ctx->n_sm)
    R /= 2;
  const int chunks = g.Mx / R, H = g.Mt / 2;
  const size_t smem = (size_t)6 * H * sizeof(double);
  double *src = x, *dst = tmp;
  for (int k = 0; k < n_or + n_hb; ++k) {
    if (k < n_or)
      sweep_rowpipe_kernel<false><<<chunks * B, H, smem, ctx->stream>>>(g, src, dst, R, chunks, 0, 0, 0);
    else
      sweep_rowpipe_kernel<true><<<chunks * B, H, smem, ctx->stream>>>(g, src, dst, R, chunks, chain0, ctx->seed,
                                                                         hb_draws[k - n_or]);
    MLMCPI_LAUNCHED("gff::sweep_rowpipe");
    std::swap(src, dst);
  }
  if (src != x)
    MLMCPI_CUDA(cudaMemcpyAsync(x, src, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}
int dof_update(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, int ell, int heatbath, uint32_t chain0,
               uint64_t draw) {
  GF g = make_gf(m);
  if (ell < 0 || ell >= g.N)
    return ctx_fail(ctx, MLMCPI_EINVAL, "vertex index out of range");
  if (heatbath)
    dof_update_kernel<true><<<cdiv(B, 128), 128, 0, ctx->stream>>>(g, ell, x, B, chain0, ctx->seed, draw);
  else
    dof_update_kernel<false><<<cdiv(B, 128), 128, 0, ctx->stream>>>(g, ell, x, B, 0, 0, 0);
  MLMCPI_LAUNCHED("gff::dof_update");
  return 0;
}

int overrelax_sweeps(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, int n_sweeps) {
  if (m->Mt_lat % 2 || m->Mx_lat % 2)
    return ctx_fail(ctx, MLMCPI_EINVAL, "coloured sweeps need even lattice extents");
  return sweep_sequence(ctx, m, x, B, n_sweeps, 0, 0, nullptr);
}
int overrelax_sweep(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B) {
  return sweep(ctx, m, x, B, false, 0, 0);
}
int heatbath_sweep(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, uint32_t chain0,
                   uint64_t draw) {
  return sweep(ctx, m, x, B, true, chain0, draw);
}

int prolong(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *xc, double *x, int B) {
  int Mtc, Mxc, rotc, rc;
  if ((rc = coarse_dims(ctx, m, &Mtc, &Mxc, &rotc)))
    return rc;
  GF g = make_gf(m);
  const int Nc = rotc ? Mtc * Mxc / 2 : Mtc * Mxc;
  transfer_kernel<true><<<vertex_grid(g, B), VERTEX_THREADS, 0, ctx->stream>>>(g, Mtc, Mxc, rotc, Nc,
                                                                               xc, x, B);
  MLMCPI_LAUNCHED("gff::prolong");
  return 0;
}

int restrict_(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *xf, double *xc, int B) {
  int Mtc, Mxc, rotc, rc;
  if ((rc = coarse_dims(ctx, m, &Mtc, &Mxc, &rotc)))
    return rc;
  GF g = make_gf(m);
  const int Nc = rotc ? Mtc * Mxc / 2 : Mtc * Mxc;
  transfer_kernel<false><<<vertex_grid(g, B), VERTEX_THREADS, 0, ctx->stream>>>(g, Mtc, Mxc, rotc,
                                                                                Nc, xf, xc, B);
  MLMCPI_LAUNCHED("gff::restrict");
  return 0;
}

int fill(mlmcpi_ctx *ctx, const mlmcpi_model *m, double *x, int B, uint32_t chain0, uint64_t draw) {
  int rc = check_rotate(ctx, m);
  if (rc)
    return rc;
  GF g = make_gf(m);
  if (g.ctype == MLMCPI_COARSEN_ROTATE) {
    const dim3 grid((unsigned)cdiv(g.Mt / 2, VERTEX_THREADS), (unsigned)(g.rotated ? g.Mx / 2 : g.Mx),
                    (unsigned)std::min(B, 65535));
    fill_rotate_kernel<<<grid, VERTEX_THREADS, 0, ctx->stream>>>(g, x, B, chain0, ctx->seed, draw);
  } else {
    fill_kernel<<<vertex_grid(g, B), VERTEX_THREADS, 0, ctx->stream>>>(g, x, B, chain0, ctx->seed, draw);
  }
  MLMCPI_LAUNCHED("gff::fill");
  return 0;
}

int prolong_fill(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *xc, double *x, int B,
                 uint32_t chain0, uint64_t draw) {
  int rc = prolong(ctx, m, xc, x, B);
  if (rc)
    return rc;
  return fill(ctx, m, x, B, chain0, draw);
}

int cond_action(mlmcpi_ctx *ctx, const mlmcpi_model *m, const double *x, int B, double *S) {
  int rc = check_rotate(ctx, m);
  if (rc)
    return rc;
  GF g = make_gf(m);
  if (g.ctype == MLMCPI_COARSEN_ROTATE) // every other vertex is coarse: visit the fine ones only
    return vertex_reduce<CondF, true>(ctx, "gff::cond_action", CondF{g, x}, g, B, 1.0, S);
  return vertex_reduce(ctx, "gff::cond_action", CondF{g, x}, g, B, 1.0, S);
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

int qoi(mlmcpi_ctx *ctx, const mlmcpi_model *m, int which, const double *x, int B, double *out,
        int64_t *Qint) {
  (void)Qint;
  if (which != MLMCPI_QOI_PHI2)
    return ctx_fail(ctx, MLMCPI_EINVAL, "QoI not defined for the GFF model");
  GF g = make_gf(m);
  return vertex_reduce(ctx, "gff::qoi_phi2", Phi2F{g, x}, g, B, 1.0 / g.N, out);
}

} // namespace gff
