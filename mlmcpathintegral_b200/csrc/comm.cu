// comm.cu -- libmlmcpi_comm.so: NCCL all-reduce of the packed Statistics moments (the only
// inter-GPU exchange of the path, include/mlmcpi_comm.h).  Links libmlmcpi.so and libnccl.
#include <cuda_runtime.h>
#include <nccl.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mlmcpi_comm.h"

static_assert(sizeof(ncclUniqueId) <= MLMCPI_COMM_ID_BYTES, "ncclUniqueId does not fit the id buffer");

struct mlmcpi_comm {
  mlmcpi_ctx *ctx = nullptr;
  ncclComm_t nccl = nullptr;
  int rank = 0, world = 1;
  double *d_packed = nullptr;
  size_t packed_n = 0;
};

extern "C" {

int mlmcpi_comm_unique_id(char id[MLMCPI_COMM_ID_BYTES]) {
  ncclUniqueId u;
  if (ncclGetUniqueId(&u) != ncclSuccess)
    return MLMCPI_ECUDA;
  std::memset(id, 0, MLMCPI_COMM_ID_BYTES);
  std::memcpy(id, &u, sizeof(u));
  return 0;
}

int mlmcpi_comm_create(mlmcpi_ctx *ctx, int rank, int world_size, const char id[MLMCPI_COMM_ID_BYTES],
                       mlmcpi_comm **out) {
  if (!ctx || !out || world_size < 1 || rank < 0 || rank >= world_size)
    return MLMCPI_EINVAL;
  mlmcpi_comm *c = new mlmcpi_comm;
  c->ctx = ctx;
  c->rank = rank;
  c->world = world_size;
  if (world_size > 1) {
    if (!id) {
      delete c;
      return MLMCPI_EINVAL;
    }
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    if (cudaSetDevice(mlmcpi_device(ctx)) != cudaSuccess ||
        ncclCommInitRank(&c->nccl, world_size, u, rank) != ncclSuccess) {
      delete c;
      return MLMCPI_ECUDA;
    }
  }
  *out = c;
  return 0;
}

int mlmcpi_comm_create_from_env(mlmcpi_ctx *ctx, mlmcpi_comm **out) {
  const char *r = std::getenv("MLMCPI_RANK"), *w = std::getenv("MLMCPI_WORLD_SIZE"),
             *f = std::getenv("MLMCPI_COMM_FILE");
  const int rank = r ? std::atoi(r) : 0, world = w ? std::atoi(w) : 1;
  if (world <= 1)
    return mlmcpi_comm_create(ctx, 0, 1, nullptr, out);
  if (!f)
    return MLMCPI_EINVAL;
  // Rendezvous file = 8-byte run nonce (MLMCPI_COMM_NONCE, set by the launcher) + the NCCL unique id.
  // Rank 0 removes whatever a killed run left under the path before it publishes (atomically, by rename);
  // the other ranks ignore a file whose nonce is not this run's, so a stale id can never reach
  // ncclCommInitRank, and they give up after MLMCPI_COMM_TIMEOUT_S seconds (default 120).
  char id[MLMCPI_COMM_ID_BYTES];
  const std::string path(f), tmp = path + ".tmp";
  const char *nv = std::getenv("MLMCPI_COMM_NONCE"), *tv = std::getenv("MLMCPI_COMM_TIMEOUT_S");
  const unsigned long long nonce = nv ? std::strtoull(nv, nullptr, 10) : 0ull;
  const int timeout_s = tv ? std::max(1, std::atoi(tv)) : 120;
  if (rank == 0) {
    int rc = mlmcpi_comm_unique_id(id);
    if (rc)
      return rc;
    std::remove(path.c_str());
    FILE *fp = std::fopen(tmp.c_str(), "wb");
    if (!fp)
      return MLMCPI_EINVAL;
    std::fwrite(&nonce, 1, sizeof(nonce), fp);
    std::fwrite(id, 1, sizeof(id), fp);
    std::fclose(fp);
    if (std::rename(tmp.c_str(), path.c_str()) != 0) // atomic publish
      return MLMCPI_EINVAL;
  } else {
    bool ok = false;
    for (int tries = 0; tries < 10 * timeout_s && !ok; ++tries) {
      FILE *fp = std::fopen(path.c_str(), "rb");
      if (fp) {
        unsigned long long got = 0;
        ok = std::fread(&got, 1, sizeof(got), fp) == sizeof(got) && got == nonce &&
             std::fread(id, 1, sizeof(id), fp) == sizeof(id);
        std::fclose(fp);
      }
      if (!ok)
        usleep(100000);
    }
    if (!ok)
      return MLMCPI_EINVAL;
  }
  return mlmcpi_comm_create(ctx, rank, world, id, out);
}

void mlmcpi_comm_destroy(mlmcpi_comm *c) {
  if (!c)
    return;
  if (c->d_packed)
    mlmcpi_free(c->ctx, c->d_packed);
  if (c->nccl)
    ncclCommDestroy(c->nccl);
  delete c;
}

int mlmcpi_comm_rank(const mlmcpi_comm *c) { return c ? c->rank : 0; }
int mlmcpi_comm_world_size(const mlmcpi_comm *c) { return c ? c->world : 1; }

int mlmcpi_comm_allreduce_sum(mlmcpi_comm *c, double *d_buf, size_t n) {
  if (!c || !d_buf)
    return MLMCPI_EINVAL;
  if (c->world == 1 || n == 0)
    return 0;
  // (the library's entry points restore the caller's current device; the collective is issued on the context's)
  int prev = -1;
  const int dev = mlmcpi_device(c->ctx);
  if (cudaGetDevice(&prev) == cudaSuccess && prev != dev)
    cudaSetDevice(dev);
  else
    prev = -1;
  const ncclResult_t rc =
      ncclAllReduce(d_buf, d_buf, n, ncclDouble, ncclSum, c->nccl, (cudaStream_t)mlmcpi_stream(c->ctx));
  if (prev >= 0)
    cudaSetDevice(prev);
  return rc == ncclSuccess ? 0 : MLMCPI_ECUDA;
}

static int comm_allreduce_thunk(void *user, double *d_buf, size_t n) {
  return mlmcpi_comm_allreduce_sum((mlmcpi_comm *)user, d_buf, n);
}

int mlmcpi_comm_attach(mlmcpi_comm *c) {
  if (!c)
    return MLMCPI_EINVAL;
  if (c->world == 1)
    return mlmcpi_set_allreduce(c->ctx, nullptr, nullptr, 1, 0);
  return mlmcpi_set_allreduce(c->ctx, comm_allreduce_thunk, c, c->world, c->rank);
}

int mlmcpi_comm_stats(mlmcpi_comm *c, mlmcpi_stats *st, int k_max, double out[6]) {
  if (!c || !st || !out)
    return MLMCPI_EINVAL;
  const size_t n = (size_t)mlmcpi_stats_packed_size(k_max);
  if (c->packed_n < n) {
    if (c->d_packed)
      mlmcpi_free(c->ctx, c->d_packed);
    int rc = mlmcpi_alloc(c->ctx, n, &c->d_packed);
    if (rc)
      return rc;
    c->packed_n = n;
  }
  int rc = mlmcpi_stats_pack_device(st, c->d_packed);
  if (rc)
    return rc;
  if ((rc = mlmcpi_comm_allreduce_sum(c, c->d_packed, n)))
    return rc;
  std::vector<double> h(n);
  if ((rc = mlmcpi_download(c->ctx, h.data(), c->d_packed, n))) // synchronises
    return rc;
  return mlmcpi_stats_finalize(h.data(), k_max, out);
}

} // extern "C"
