"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
usage: launch_times.py launches.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
i_name, i_metric, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[i_metric] != "gpu__time_duration.sum":
        continue
    v = float(r[i_val].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[i_unit], 1.0)
    name = r[i_name].split("(")[0]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:40s} n={v[0]:5d} total={v[1] / 1e3:9.3f} ms  mean={v[1] / v[0]:9.1f} us  {100 * v[1] / tot:5.1f}%")
