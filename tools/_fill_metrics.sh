#!/bin/bash
# ncu counters that decide the fused fill-in kernel: instruction-cache misses that go to the GPC-level cache
# (gcc__*), issue utilisation, executed warp instructions.  Usage (on the GPU box): tools/_fill_metrics.sh OUT.csv
M="smsp__issue_active.avg.pct_of_peak_sustained_active,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warp_latency_per_inst_issued.ratio,sm__icc_requests.sum,sm__icc_requests_lookup_miss.sum,smsp__warps_active.avg.per_cycle_active"
MLMCPI_N=1 ncu --metrics $M --clock-control none -k regex:prolong_fill -c 1 --csv --log-file "$1" python tools/_fill_time.py > /dev/null 2>&1
python - "$1" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]
for r in rows[1:]:
    d = dict(zip(h, r))
    print(d['Metric Name'], d['Metric Value'])
PY
