#!/bin/bash
cd /root/repo
g++ -std=c++17 -O2 -w -Iinclude examples/driver_qft.cc -Lmlmcpathintegral_b200 -lmlmcpi -lmlmcpi_comm -Wl,-rpath,$PWD/mlmcpathintegral_b200 -o /tmp/driver_qft
sed "s/n_samples = 100000 /n_samples = 2000000 /; s/sampler = 'hierarchical'      # HMC/sampler = 'heatbath'      # HMC/; s/n_burnin = 100$/n_burnin = 500/" examples/parameters_qft_schwinger.in > /tmp/p1.in
/tmp/driver_qft /tmp/p1.in 256 2>&1 | grep -E "sampler = |Avg \+/- Err|tau_|# samples|analytical - numerical|E\[V"
sed "s/n_samples = 100000 /n_samples = 2000000 /; s/sampler = 'hierarchical'      # HMC/sampler = 'cluster'      # HMC/" examples/parameters_qft_schwinger.in > /tmp/p2.in
/tmp/driver_qft /tmp/p2.in 256 2>&1 | grep -E "Avg \+/- Err|tau_|# samples|analytical - numerical"
# hierarchical, 2 levels, 16^2, long run
sed "s/n_samples = 100000 /n_samples = 4000000 /; s/Mt_lat = 64/Mt_lat = 16/; s/Mx_lat = 64/Mx_lat = 16/; s/n_max_level = 3/n_max_level = 2/; s/n_autocorr_window = 20/n_autocorr_window = 100/" examples/parameters_qft_schwinger.in > /tmp/p3.in
/tmp/driver_qft /tmp/p3.in 512 2>&1 | grep -E "Avg \+/- Err|tau_|# samples|analytical - numerical|level"
