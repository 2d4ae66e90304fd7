"""Key metrics of an `ncu --set full` capture (tools/prof_kernels.sh) -> profiles/<tag>_ncu_full_summary.json.
usage: full_summary.py <report.ncu-rep> <tag>"""
import csv
import json
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active"]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(txt.splitlines()))
i_name = rr[0].index("Kernel Name")
summ = {}
for row in rr[2:]:
    m = {}
    for h, u, v in zip(rr[0], rr[1], row):
        if h in want or ("smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            try:
                fv = float(v.replace(",", ""))
            except ValueError:
                continue
            if "stalled" in h and fv < 0.2:
                continue
            m[h] = [fv, u]
    name = row[i_name].split("(")[0].replace("void ", "")
    k = name
    n = 2
    while k in summ:
        k = f"{name} #{n}"
        n += 1
    summ[k] = m
json.dump(summ, open(f"profiles/{tag}_ncu_full_summary.json", "w"), indent=1)
for k, m in summ.items():
    print(k, {h.split(".")[0].replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active", ""): round(v[0], 2) for h, v in m.items()})
