#!/bin/bash
# diagnostic: the two-process C++ driver runs of tests/test_drivers.py with a hard time limit per run
cd "$(dirname "$0")/.."
python - <<'PY'
import sys
sys.path.insert(0, ".")
from tests.test_drivers import QFT, QFT_DEFAULTS
open("/tmp/p_single.in", "w").write(QFT.format(**dict(QFT_DEFAULTS, n_samples=200000)))
open("/tmp/p_mlmc.in", "w").write(QFT.format(**dict(QFT_DEFAULTS, method="multilevel", n_max_level=2, epsilon=0.05)))
PY
run() { # name limit_s command...
  name=$1; limit=$2; shift 2
  setsid "$@" > gpurun_out/two_gpu_$name.log 2>&1 &
  pg=$!
  t0=$(date +%s)
  while kill -0 $pg 2>/dev/null; do
    if [ $(( $(date +%s) - t0 )) -ge $limit ]; then echo "$name: TIME LIMIT, killing"; kill -- -$pg 2>/dev/null; sleep 1; kill -9 -- -$pg 2>/dev/null; break; fi
    sleep 1
  done
  wait $pg; echo "$name: rc=$? after $(( $(date +%s) - t0 )) s"
  tail -c 1500 gpurun_out/two_gpu_$name.log
}
export NCCL_DEBUG=WARN
CASES=${1:-"single mlmc"}
for c in $CASES; do
  run $c ${2:-50} examples/run_multi_gpu.sh 2 examples/driver_qft /tmp/p_$c.in 64
done
