"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel: launches, total / average time, share.

usage: python tools/summarize_launches.py gpurun_out/launches_TAG.csv > profiles/TAG_launches_summary.txt
"""
import csv, re, sys
from collections import OrderedDict

path = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, iv, ib, ig, iu = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Block Size", "Grid Size", "Metric Unit"))
agg = OrderedDict()
for r in rows:
    name = re.sub(r"^void ", "", r[ik])
    name = re.sub(r"\(.*$", "", name)
    us = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[r[iu]]
    a = agg.setdefault(name, [0, 0.0, r[ig], r[ib]])
    a[0] += 1; a[1] += us
total = sum(a[1] for a in agg.values())
print(f"# ncu --metrics gpu__time_duration.sum --clock-control none, command: python bench.py --steps 3 --warmup 3 --no-cpu-baseline")
print(f"# (per-launch times are cold-cache and serialised: compare shares, not absolutes; the list is cut at the first -c launches). Source: {path}")
print(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}  grid block")
for name, (n, us, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    short = name if len(name) <= 70 else name[:67] + "..."
    print(f"{short:70s} {n:8d} {us:12.1f} {us / n:10.2f} {100 * us / total:6.2f}%  {g} {b}")
