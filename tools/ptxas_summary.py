"""Registers / stack / spills per kernel from the `-Xptxas -v` logs the Makefile keeps (csrc/*.ptxas.log).
usage: ptxas_summary.py [log ...]"""
import glob
import os
import re
import subprocess
import sys

logs = sys.argv[1:] or sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "cpu-raytracing-rt_b200", "csrc", "*.ptxas.log")))
for log in logs:
    name, rows = None, []
    for line in open(log):
        m = re.search(r"Compiling entry function '(\w+)'", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void rt::", "")
            rows.append([name, "", "", ""])
            continue
        m = re.search(r"Function properties for (\w+)", line)
        if m and (not rows or rows[-1][2]):
            name = None
        if name is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and not rows[-1][2]:
            rows[-1][2] = f"stack {m.group(1)} spill st/ld {m.group(2)}/{m.group(3)}"
        m = re.search(r"Used (\d+) registers.*?(?:, (\d+) bytes smem)?$", line.strip())
        if m:
            rows[-1][1] = f"{m.group(1)} regs" + (f", {m.group(2)} B smem" if m.group(2) else "")
    print(os.path.basename(log))
    for r in rows:
        print(f"  {r[0]:55s} {r[1]:24s} {r[2]}")
