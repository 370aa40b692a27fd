#!/bin/bash
# multilevel method: one process x 64 / x 128 chains against two processes x 64 chains
cd /root/repo
g++ -std=c++17 -O2 -w -Iinclude examples/driver_qft.cc -Lmlmcpathintegral_b200 -lmlmcpi -lmlmcpi_comm -Wl,-rpath,$PWD/mlmcpathintegral_b200 -o /tmp/driver_qft
python - <<'PY'
import sys
sys.path.insert(0,'tests')
import test_drivers as t
open('/tmp/pm.in','w').write(t.QFT.format(**dict(t.QFT_DEFAULTS, method="multilevel", n_max_level=2, epsilon=0.05)))
open('/tmp/pm2.in','w').write(t.QFT.format(**dict(t.QFT_DEFAULTS, method="multilevel", n_max_level=2, epsilon=0.02)))
PY
F='level|E\[|Avg|analytical - numerical|tolerance|^ +[0-9]'
echo "== 1 x 64";  /tmp/driver_qft /tmp/pm.in 64 | grep -E "$F"
echo "== 1 x 128"; /tmp/driver_qft /tmp/pm.in 128 | grep -E "$F"
echo "== 2 x 64";  examples/run_multi_gpu.sh 2 /tmp/driver_qft /tmp/pm.in 64 | grep -E "$F"
echo "== 1 x 128 eps 0.02"; /tmp/driver_qft /tmp/pm2.in 128 | grep -E "$F"
echo "== 2 x 64 eps 0.02";  examples/run_multi_gpu.sh 2 /tmp/driver_qft /tmp/pm2.in 64 | grep -E "$F"
