#!/bin/bash
# One process per GPU: run_multi_gpu.sh N ./driver_qft PARAMETERFILE [CHAINS]
# Chains are sharded over the processes (CHAINS per GPU); the packed Statistics moments are
# all-reduced with NCCL (libmlmcpi_comm.so).  Rank 0 prints.
set -e
N=$1
shift
F=$(mktemp -u /tmp/mlmcpi_comm.XXXXXX)
pids=()
for r in $(seq 0 $((N - 1))); do
  MLMCPI_RANK=$r MLMCPI_WORLD_SIZE=$N MLMCPI_COMM_FILE=$F "$@" &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait $p || rc=$?; done
rm -f "$F"
exit $rc
