#!/bin/bash
# single-level hierarchical Schwinger 64^2 through the C++ driver on 1, 2, 4, 8 GPUs
set -e
cd /root/repo
for d in qft; do g++ -std=c++17 -O2 -w -Iinclude examples/driver_$d.cc -Lmlmcpathintegral_b200 -lmlmcpi -lmlmcpi_comm -Wl,-rpath,$PWD/mlmcpathintegral_b200 -o /tmp/driver_$d; done
sed 's/n_samples = 100000 /n_samples = 4000000 /' examples/parameters_qft_schwinger.in > /tmp/p.in
for n in 1 2 4 8; do
  echo "== $n GPUs"
  examples/run_multi_gpu.sh $n /tmp/driver_qft /tmp/p.in 256 2>&1 | grep -E "Avg \+/- Err|# samples|throughput|timer SinglevelMC|analytical - numerical"
done
