"""Per-region view of an `ncu --page source --csv --print-source sass` export: executed warp instructions,
stall samples and the footprint (instructions executed by a given fraction of the warps).  Usage:
  ncu -i X.ncu-rep --page source --csv --print-source sass > x.csv; python tools/_sass_profile.py x.csv [chunk]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 128
h = rows[1]; ix = {n: i for i, n in enumerate(h)}; data = rows[2:]
def f(r, n):
    try: return float(r[ix[n]])
    except Exception: return 0.0
nw = max(f(r, 'Instructions Executed') for r in data[:40])
tot = sum(f(r, 'Instructions Executed') for r in data)
print('static', len(data), 'warps', nw, 'executed per warp', tot / nw)
for thr in (0.9, 0.5, 0.2, 0.05, 0.01, 0.0001):
    print(' executed by >= %g of warps: %d instructions' % (thr, sum(1 for r in data if f(r, 'Instructions Executed') >= thr * nw)))
for k in range(0, len(data), chunk):
    ch = data[k:k + chunk]
    ex = sum(f(r, 'Instructions Executed') for r in ch) / nw
    if ex == 0: continue
    ops = {}
    for r in ch:
        op = r[ix['Source']].split()[0] if r[ix['Source']].split() else ''
        if op.startswith('@'): op = r[ix['Source']].split()[1]
        ops[op.split('.')[0]] = ops.get(op.split('.')[0], 0) + f(r, 'Instructions Executed') / nw
    top = sorted(ops.items(), key=lambda t: -t[1])[:6]
    print('%5d exec/warp %7.1f samples %6d no_inst %6d wait %6d barrier %6d  %s' % (k, ex, sum(f(r, '# Samples') for r in ch),
          sum(f(r, 'stall_no_inst') for r in ch), sum(f(r, 'stall_wait') for r in ch), sum(f(r, 'stall_barrier') for r in ch),
          ' '.join('%s:%.0f' % t for t in top)))
