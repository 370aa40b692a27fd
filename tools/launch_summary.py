#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/launch_summary.py launches.csv [n_last_launches]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    if len(sys.argv) > 2:
        rows = rows[-int(sys.argv[2]):]
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r".*::", "", name)
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)  # -> us
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {v / 1e3:.3f} | {v / n:.1f} | {v / tot:.1%} |")
    print(f"| total | {len(rows)} | {tot / 1e3:.3f} | | |")


if __name__ == "__main__":
    main()
