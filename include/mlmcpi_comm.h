/* mlmcpi_comm.h -- C-ABI of libmlmcpi_comm.so: the one inter-GPU exchange of the path.
 *
 * The reference averages the moments of its per-rank Statistics objects over MPI ranks
 * (common/statistics.cc:30-35, 38-47, 64-79 through mpi/mpi_wrapper.cc:53-60) and ANDs the
 * stopping test (montecarlo/montecarlosinglelevel.cc:85-86).  Here one process drives one GPU with
 * a batch of chains; the packed moment vector of a device Statistics object
 * (mlmcpi_stats_pack_device, every entry additive) is summed over the processes by ncclAllReduce
 * over NVLink, after which every process holds the same numbers and takes the same decisions.
 * No other data crosses GPUs.
 *
 * NCCL lives in this separate library so that libmlmcpi.so itself has no NCCL dependency (the
 * Python host layer brings its own through torch.distributed).
 */
#ifndef MLMCPI_COMM_H
#define MLMCPI_COMM_H
#include "mlmcpi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mlmcpi_comm mlmcpi_comm;

#define MLMCPI_COMM_ID_BYTES 128
/* ncclGetUniqueId: rank 0 creates the id and hands it to the other ranks (file, socket, ...) */
int mlmcpi_comm_unique_id(char id[MLMCPI_COMM_ID_BYTES]);
/* ncclCommInitRank on the device of ctx; collectives are issued on the stream of ctx */
int mlmcpi_comm_create(mlmcpi_ctx *ctx, int rank, int world_size, const char id[MLMCPI_COMM_ID_BYTES],
                       mlmcpi_comm **comm);
/* rank / world size / rendezvous from the environment: MLMCPI_RANK, MLMCPI_WORLD_SIZE and
 * MLMCPI_COMM_FILE (a path on a file system all ranks see: rank 0 writes the id there, the others
 * wait for it).  World size 1 (or unset) gives a communicator whose collectives are no-ops. */
int mlmcpi_comm_create_from_env(mlmcpi_ctx *ctx, mlmcpi_comm **comm);
void mlmcpi_comm_destroy(mlmcpi_comm *comm);
int mlmcpi_comm_rank(const mlmcpi_comm *comm);
int mlmcpi_comm_world_size(const mlmcpi_comm *comm);
/* in-place sum over the ranks of n doubles in device memory (asynchronous, stream of ctx) */
int mlmcpi_comm_allreduce_sum(mlmcpi_comm *comm, double *d_buf, size_t n);
/* install the communicator as the all-reduce of its context (mlmcpi_set_allreduce): from then on the
 * library's own Statistics queries -- MultilevelSampler, MonteCarloMultiLevel, autotune -- run over
 * the chains of all ranks */
int mlmcpi_comm_attach(mlmcpi_comm *comm);
/* Statistics over ALL chains of ALL ranks: pack on the device, all-reduce, finalize;
 * out = {average, variance, variance_error, tau_int, error, samples} as mlmcpi_stats_finalize
 * (synchronises; identical on every rank) */
int mlmcpi_comm_stats(mlmcpi_comm *comm, mlmcpi_stats *st, int k_max, double out[6]);

#ifdef __cplusplus
}
#endif
#endif /* MLMCPI_COMM_H */
