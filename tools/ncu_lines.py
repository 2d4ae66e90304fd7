"""Aggregates an ncu source page (`--page source --csv --print-source cuda,sass`) per source line.
usage: ncu_lines.py report.ncu-rep [top_n] [kernel-name-regex]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kern = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + kern, capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file, hdr = None, None
agg = defaultdict(lambda: [0, 0, 0, ""])      # (file,line) -> [inst, thread_inst, samples, src]
ftot = defaultdict(lambda: [0, 0, 0])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_inst, i_tinst, i_smp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or r[0] in ("Function Name", ""):
        continue
    try:
        line = int(r[0])
        inst, tinst, smp = int(r[i_inst]), int(r[i_tinst]), int(r[i_smp])
    except (ValueError, IndexError):
        continue
    a = agg[(cur_file, line)]
    a[0] += inst; a[1] += tinst; a[2] += smp; a[3] = r[1][:90]
    f = ftot[cur_file]
    f[0] += inst; f[1] += tinst; f[2] += smp
tot = sum(v[0] for v in ftot.values()) or 1
tots = sum(v[2] for v in ftot.values()) or 1
print("per file: inst%  lanes/inst  samples%")
for f, v in sorted(ftot.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:18s} {100*v[0]/tot:6.2f}%  {v[1]/max(1,v[0]):5.1f}  {100*v[2]/tots:6.2f}%")
print("top lines by issued instructions:")
for (f, line), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"  {f}:{line:<4d} {100*v[0]/tot:5.2f}% lanes {v[1]/max(1,v[0]):5.1f} smp {100*v[2]/tots:5.2f}%  {v[3]}")
print("top lines by stall samples:")
for (f, line), v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"  {f}:{line:<4d} smp {100*v[2]/tots:5.2f}% inst {100*v[0]/tot:5.2f}% lanes {v[1]/max(1,v[0]):5.1f}  {v[3]}")
