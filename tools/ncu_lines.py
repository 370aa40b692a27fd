#!/usr/bin/env python
"""Attribute executed warp-instructions and stall samples to CUDA source lines.
Input: `ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:...`"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file, hdr = r[1], None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].strip():  # a CUDA source line (SASS rows have empty Line No)
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        out.append((cur_file, d))
key, skey = "Instructions Executed", "Warp Stall Sampling (All Samples)"
def num(d, k):
    try:
        return int(d.get(k, "0") or 0)
    except ValueError:
        return 0
tot = sum(num(d, key) for _, d in out) or 1
stot = sum(num(d, skey) for _, d in out) or 1
print(f"total warp-instructions {tot}, stall samples {stot}")
for f, d in sorted(out, key=lambda fd: -num(fd[1], key))[:top_n]:
    print(f"{num(d, key) / tot:6.2%} inst {num(d, skey) / stot:6.2%} stall  {f.split('/')[-1]}:{d['Line No']:>4} {d['Source'].strip()[:90]}")
