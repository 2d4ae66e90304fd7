"""Turns gpurun_out/r01_launches.csv and the --set full capture (r01_wf_all.ncu-rep: one launch each of
k_wf_confirm, k_wf_shade, k_wf_trace of the same iteration) into the tracked summaries
under profiles/ (launch summary, per-kernel key metrics, measured DRAM traffic)."""
import collections
import csv
import json
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rows = [r for r in csv.reader(open(f"gpurun_out/{tag}_launches.csv")) if len(r) > 10 and r[0].isdigit()]
d = collections.OrderedDict()
for r in rows:
    e = d.setdefault(int(r[0]), {"k": r[4].split("(")[0].replace("void ", "")})
    e[r[-3]] = float(r[-1].replace(",", ""))
    e["unit_" + r[-3]] = r[-2]
T = "gpu__time_duration.sum"


def tobytes(e, m):
    return e.get(m, 0.0) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(e.get("unit_" + m, "byte"), 1)


agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
for e in d.values():
    a = agg[e["k"]]
    a[0] += 1
    a[1] += e[T]
    a[2] += tobytes(e, "dram__bytes_read.sum") + tobytes(e, "dram__bytes_write.sum")
    a[3] += e["smsp__thread_inst_executed_per_inst_executed.ratio"] * e[T]
    a[4] += e["smsp__issue_active.avg.pct_of_peak_sustained_active"] * e[T]
tot = sum(a[1] for a in agg.values())
lines = []
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k} | {a[0]} | {a[1] / 1e6:.2f} | {100 * a[1] / tot:.1f}% | {a[2] / 1e9:.2f} | {a[3] / a[1]:.1f} | {a[4] / a[1]:.1f}")
open(f"profiles/{tag}_launch_summary.md", "w").write(
    f"# {tag} launch list summary (ncu --metrics gpu__time_duration.sum,dram__bytes_*.sum,... --clock-control none)\n\n"
    "command: `python bench.py --steps 2 --warmup 3 --spp-per-step 8 --no-cpu-baseline` (1 EXACT counter pass + 3 warm-up + 2 timed + 2+2 e2e renders)\n\n"
    "kernel | launches | total ms | share | DRAM GB (read+write) | time-weighted lanes/inst | issue active %\n---|---|---|---|---|---|---\n" + "\n".join(lines) +
    "\n\nper-launch times are cold-cache and serialised by the profiler: compare SHARES, not absolutes.\n")
print("\n".join(lines))
wf = sum(agg[k][2] for k in agg if k.startswith("k_wf_"))
renders = agg["k_wf_init"][0]
json.dump({"dram_bytes_per_spp_at_1080p": wf / renders / 8,
           "source": f"profiles/{tag}_launches.csv: sum of dram__bytes_read.sum + dram__bytes_write.sum over the k_wf_* kernels of {renders} wavefront renders of 8 spp (C3, 1920x1080), divided by {renders * 8} spp",
           "note": "one bench 'launch' is one render call of spp_per_step samples: traffic = this * spp_per_step"},
          open("profiles/traffic.json", "w"), indent=1)
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active"]
summ = {}
txt = subprocess.run(["ncu", "-i", f"gpurun_out/{tag}_wf_all.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(txt.splitlines()))
i_name = rr[0].index("Kernel Name")
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
    summ[row[i_name].split("(")[0].replace("void ", "").split("<")[0]] = m
json.dump(summ, open(f"profiles/{tag}_ncu_full_summary.json", "w"), indent=1)
for k, m in summ.items():
    print(k, {h.split(".")[0].replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active", ""): round(v[0], 2) for h, v in m.items()})
