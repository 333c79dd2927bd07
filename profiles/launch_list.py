#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches and time per kernel.

    python profiles/launch_list.py gpurun_out/launches_r1k.csv "python bench.py" > profiles/r01k_launch_list_default_bench.txt
"""
import collections, csv, sys
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "us")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        rows.append((r["Kernel Name"], v))
agg = collections.OrderedDict()
for k, v in rows:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in rows)
cmd = sys.argv[2] if len(sys.argv) > 2 else "?"
print(f"# ncu --metrics gpu__time_duration.sum --clock-control none --csv : {cmd}")
print(f"# total {tot:.1f} us over {len(rows)} launches; per-launch times are cold-cache and serialised")
for k, (n, v) in sorted(agg.items(), key=lambda t: -t[1][1]):
    print(f"{n:6d} launches {v:12.1f} us total {v / n:10.2f} us avg {100 * v / tot:5.1f} %  {k[:110]}")
