#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: top instructions by stall samples and the kernel-wide
stall-reason mix.  Usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_top.py [N]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr_i]
body = [r for r in rows[hdr_i + 1:] if len(r) == len(h)]
ci = {n: i for i, n in enumerate(h)}
samp = ci["# Samples"]
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
total = sum(int(r[samp]) for r in body) or 1
print(f"kernel: {rows[0][1] if rows and len(rows[0]) > 1 else '?'}  instructions: {len(body)}  samples: {total}")
mix = {s: sum(int(r[ci[s]]) for r in body) for s in stalls}
print("stall mix: " + ", ".join(f"{k[6:]} {100*v/total:.1f}%" for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:8]))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
order = sorted(range(len(body)), key=lambda i: -int(body[i][samp]))[:n]
for i in sorted(order):
    r = body[i]
    top = sorted(((int(r[ci[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {100*int(r[samp])/total:5.1f}%  {r[ci['Source']].strip()[:90]:90s} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}")
