"""Per-kernel totals of a launch list written by tools/launch_list.sh.  usage: ll_summary.py gpurun_out/<tag>_ll.csv [first_launch last_launch]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
L = collections.OrderedDict()
for r in rows:
    if not lo <= int(r[0]) <= hi:
        continue
    e = L.setdefault(int(r[0]), {"k": r[4].split("(")[0].replace("void ", "")})
    v = float(r[-1].replace(",", "")) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e3, "ms": 1e6}.get(r[-2], 1)
    e[r[-3].split(".")[0]] = v
agg = collections.defaultdict(lambda: collections.defaultdict(float))
for e in L.values():
    a = agg[e["k"]]
    t = e["gpu__time_duration"]
    a["n"] += 1; a["ns"] += t
    a["dram"] += e.get("dram__bytes_read", 0) + e.get("dram__bytes_write", 0)
    a["winst"] += e.get("smsp__inst_executed", 0)
    a["tinst"] += e.get("smsp__inst_executed", 0) * e.get("smsp__thread_inst_executed_per_inst_executed", 0)
    a["issue"] += e.get("smsp__issue_active", 0) * t
    a["l1"] += e.get("l1tex__t_sector_hit_rate", 0) * t
    a["l2"] += e.get("lts__t_sector_hit_rate", 0) * t
tot = sum(a["ns"] for a in agg.values())
print(f"{'kernel':28s} {'n':>5s} {'ms':>8s} {'share':>6s} {'us/launch':>9s} {'DRAM GB':>8s} {'Minst':>8s} {'lanes':>6s} {'issue%':>6s} {'L1%':>5s} {'L2%':>5s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    print(f"{k:28s} {int(a['n']):5d} {a['ns'] / 1e6:8.2f} {100 * a['ns'] / tot:5.1f}% {a['ns'] / a['n'] / 1e3:9.1f} {a['dram'] / 1e9:8.2f} {a['winst'] / 1e6:8.1f} "
          f"{a['tinst'] / max(1, a['winst']):6.1f} {a['issue'] / a['ns']:6.1f} {a['l1'] / a['ns']:5.1f} {a['l2'] / a['ns']:5.1f}")
print(f"total {tot / 1e6:.2f} ms")
