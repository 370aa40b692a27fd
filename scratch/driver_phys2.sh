#!/bin/bash
cd /root/repo
g++ -std=c++17 -O2 -w -Iinclude examples/driver_qft.cc -Lmlmcpathintegral_b200 -lmlmcpi -lmlmcpi_comm -Wl,-rpath,$PWD/mlmcpathintegral_b200 -o /tmp/driver_qft
for L in 32 64; do
sed "s/n_samples = 100000 /n_samples = 4000000 /; s/Mt_lat = 64/Mt_lat = $L/; s/Mx_lat = 64/Mx_lat = $L/; s/sampler = 'hierarchical'      # HMC/sampler = 'cluster'      # HMC/; s/n_updates = 10/n_updates = 100/; s/n_autocorr_window = 20/n_autocorr_window = 200/; s/n_burnin = 100$/n_burnin = 1000/" examples/parameters_qft_schwinger.in > /tmp/p2.in
/tmp/driver_qft /tmp/p2.in 512 2>&1 | grep -E "Avg \+/- Err|tau_|# samples|analytical - numerical|E\[V|timer Single"
done
