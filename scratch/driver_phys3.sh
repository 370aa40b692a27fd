#!/bin/bash
cd /root/repo
g++ -std=c++17 -O2 -w -Iinclude examples/driver_qft.cc -Lmlmcpathintegral_b200 -lmlmcpi -lmlmcpi_comm -Wl,-rpath,$PWD/mlmcpathintegral_b200 -o /tmp/driver_qft
for cs in cluster HMC; do
sed "s/n_samples = 100000 /n_samples = 4000000 /; s/Mt_lat = 64/Mt_lat = 32/; s/Mx_lat = 64/Mx_lat = 32/; s/beta = 4.0/beta = 16.0/; s/n_max_level = 3/n_max_level = 2/; s/coarsesampler = 'HMC'/coarsesampler = '$cs'/; s/n_updates = 10/n_updates = 50/; s/n_autocorr_window = 20/n_autocorr_window = 200/; s/n_burnin = 100$/n_burnin = 2000/" examples/parameters_qft_schwinger.in > /tmp/p4.in
/tmp/driver_qft /tmp/p4.in 512 2>&1 | grep -E "coarsesampler|Avg \+/- Err|tau_|# samples|analytical - numerical|E\[V|level|timer Single"
done
