#!/usr/bin/env python
"""Top stall locations of an ncu report (development tool): python tools/ncu_hot.py report.ncu-rep [topN]
Prints the SASS instructions with the most warp-stall samples, with their dominant stall reasons."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
S = ci["# Samples"]
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in body if r[S].isdigit()]
tot = sum(int(r[S] or 0) for r in body)
print(f"total samples {tot}, instructions {len(body)}")
agg = {}
for r in body:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("by reason:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
idx = sorted(range(len(body)), key=lambda k: -int(body[k][S] or 0))[:top]
for k in sorted(idx):
    r = body[k]
    reasons = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{k:5d} {int(r[S]):6d} {100*int(r[S])/tot:5.1f}%  {r[ci['Source']].strip():70s} {reasons}")
