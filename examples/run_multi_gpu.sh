#!/bin/bash
# One process per GPU: run_multi_gpu.sh N ./driver_qft PARAMETERFILE [CHAINS]
# Chains are sharded over the processes (CHAINS per GPU); the packed Statistics moments are
# all-reduced with NCCL (libmlmcpi_comm.so).  Rank 0 prints.
set -e
N=$1
shift
F=$(mktemp -u /tmp/mlmcpi_comm.XXXXXX)
NONCE=$(( ($$ << 16) ^ RANDOM ^ $(date +%s) ))   # readers ignore a rendezvous file that is not this run's
trap 'rm -f "$F" "$F.tmp"' EXIT INT TERM
pids=()
for r in $(seq 0 $((N - 1))); do
  MLMCPI_RANK=$r MLMCPI_WORLD_SIZE=$N MLMCPI_COMM_FILE=$F MLMCPI_COMM_NONCE=$NONCE "$@" &
  pids+=($!)
done
# a process that fails must not leave its peers waiting in a collective: the first non-zero exit stops the run
rc=0
left=${#pids[@]}
while [ "$left" -gt 0 ]; do
  wait -n
  s=$?
  left=$((left - 1))
  if [ "$s" -ne 0 ] && [ "$rc" -eq 0 ]; then
    rc=$s
    kill "${pids[@]}" 2>/dev/null
  fi
done
exit $rc
