"""One steady-state iteration out of a launch list (tools/launch_list.sh): the launches between the
k-th and (k+1)-th trace kernel.  usage: ll_iter.py gpurun_out/<tag>_ll.csv [k=10]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
L = collections.OrderedDict()
for r in rows:
    e = L.setdefault(int(r[0]), {"k": r[4].split("(")[0].replace("void ", "")})
    e[r[-3].split(".")[0]] = float(r[-1].replace(",", "")) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e3, "ms": 1e6}.get(r[-2], 1)
ids = list(L)
tr = [i for i in ids if L[i]["k"].startswith("k_wf_trace")]
lo, hi = tr[k] + 1, tr[k + 1]
tot = 0.0
for i in ids:
    if lo <= i <= hi:
        e = L[i]
        tot += e["gpu__time_duration"]
        print(f"  {e['k']:22s} {e['gpu__time_duration'] / 1e3:8.1f} us  dram {(e['dram__bytes_read'] + e['dram__bytes_write']) / 1e6:7.1f} MB  Minst {e['smsp__inst_executed'] / 1e6:6.1f} "
              f"lanes {e['smsp__thread_inst_executed_per_inst_executed']:5.1f} issue {e['smsp__issue_active']:5.1f} L1 {e['l1tex__t_sector_hit_rate']:5.1f} L2 {e['lts__t_sector_hit_rate']:5.1f}")
print(f"  total {tot / 1e3:.1f} us (serialised under the profiler)")
