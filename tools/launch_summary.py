#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/launch_summary.py launches.csv [out.txt] [title]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "")
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v * 1e3 if u == "s" else v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
lines = [f"# {sys.argv[3] if len(sys.argv) > 3 else sys.argv[1]}",
         "# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)",
         f"# total {tot:.2f} ms over {sum(a[0] for a in agg.values())} launches"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{t:12.3f} ms {100 * t / tot:6.2f}%  x{c:4d}  {k}")
txt = "\n".join(lines)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
