#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/summarize_launches.py in.csv "command" """
import csv, re, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
n = 0
for r in rd:
    name = re.sub(r"^void ", "", r[ki])
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(.*", "", name)
    v = float(r[vi].replace(",", ""))
    us = v / 1e3 if r[ui] in ("ns", "nsecond") else v
    a = agg.setdefault(name, [0.0, 0])
    a[0] += us; a[1] += 1
    n += 1
tot = sum(a[0] for a in agg.values())
print(f"ncu --metrics gpu__time_duration.sum --clock-control none, command: {sys.argv[2] if len(sys.argv) > 2 else ''}")
print(f"total {tot / 1e3:.1f} ms over {n} launches (cold-cache, serialised)")
for name, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{us:10.1f} us  n={c:4d}  avg {us / c:8.1f} us  {100 * us / tot:5.1f}%  {name[:150]}")
