"""Launch list (ncu --metrics ... --csv of the short bench command) -> profiles/<tag>_issue.json and
profiles/<tag>_launch_summary.md: per-kernel time share, DRAM bytes, lanes per instruction, issue activity,
and the per-SEGMENT instruction counts bench.py's issue roofline is computed from.
usage: issue_summary.py <tag> <spp_per_render> [segments_per_path] [csv] [command]
The launch list is gpurun_out/<tag>_launches.csv (the short bench command) or, with [csv], the list of one
render written by tools/launch_list.sh."""
import collections
import csv
import json
import sys

tag, spp = sys.argv[1], int(sys.argv[2])
seg_per_path = float(sys.argv[3]) if len(sys.argv) > 3 else 1.8330860533537687
W, H = 1920, 1080
src = sys.argv[4] if len(sys.argv) > 4 else f"gpurun_out/{tag}_launches.csv"
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
launch = collections.OrderedDict()
for r in rows:
    e = launch.setdefault(int(r[0]), {"k": r[4].split("(")[0].replace("void ", "")})
    v = float(r[-1].replace(",", ""))
    unit = r[-2]
    v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)    # bytes, ns
    e[r[-3]] = v
agg = collections.defaultdict(lambda: collections.defaultdict(float))
for e in launch.values():
    a = agg[e["k"]]
    a["n"] += 1
    t = e["gpu__time_duration.sum"]
    a["ns"] += t
    a["dram"] += e.get("dram__bytes_read.sum", 0) + e.get("dram__bytes_write.sum", 0)
    a["warp_inst"] += e.get("smsp__inst_executed.sum", 0)
    a["thread_inst"] += e.get("smsp__thread_inst_executed.sum", 0)
    a["fp64_warp_inst"] += e.get("smsp__inst_executed_pipe_fp64.sum", 0)
    a["issue_x_t"] += e.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0) * t
renders = int(agg["k_wf_init"]["n"])
segments = renders * spp * W * H * seg_per_path
tot_ns = sum(a["ns"] for a in agg.values())
wf = {k: a for k, a in agg.items() if k.startswith("k_wf_")}
lines, per_kernel = [], {}
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    lanes = a["thread_inst"] / max(1.0, a["warp_inst"])
    lines.append(f"{k} | {int(a['n'])} | {a['ns'] / 1e6:.2f} | {100 * a['ns'] / tot_ns:.1f}% | {a['dram'] / 1e9:.2f} | {lanes:.1f} | {a['issue_x_t'] / max(1.0, a['ns']):.1f}")
    if k in wf:
        per_kernel[k] = {"time_share_of_wavefront": a["ns"] / sum(x["ns"] for x in wf.values()), "lanes_per_instruction": lanes,
                         "issue_active_pct": a["issue_x_t"] / max(1.0, a["ns"]), "warp_inst_per_segment": a["warp_inst"] / segments,
                         "thread_inst_per_segment": a["thread_inst"] / segments, "dram_bytes_per_segment": a["dram"] / segments}
cmd = sys.argv[5] if len(sys.argv) > 5 else f"python bench.py --steps 2 --warmup 3 --spp-per-step {spp} --no-cpu-baseline --no-strong"
open(f"profiles/{tag}_launch_summary.md", "w").write(
    f"# {tag} launch list summary (ncu --metrics gpu__time_duration.sum,dram__bytes_*.sum,smsp__inst_executed.sum,... --clock-control none)\n\n"
    f"command: `{cmd}` ({renders} wavefront render(s) of {spp} spp)\n\n"
    "kernel | launches | total ms | share | DRAM GB (read+write) | lanes/inst | issue active % (time-weighted)\n---|---|---|---|---|---|---\n" + "\n".join(lines) +
    "\n\nper-launch times are cold-cache and serialised by the profiler: compare SHARES, not absolutes.\n")
out = {"source": f"profiles/{tag}_launches.csv ({cmd}): sums over every k_wf_* launch of {renders} render calls / {segments:.0f} path segments",
       "renders": renders, "spp_per_render": spp, "segments": segments,
       "warp_inst_per_segment": sum(a["warp_inst"] for a in wf.values()) / segments,
       "thread_inst_per_segment": sum(a["thread_inst"] for a in wf.values()) / segments,
       "fp64_warp_inst_per_segment": sum(a["fp64_warp_inst"] for a in wf.values()) / segments,
       "dram_bytes_per_segment": sum(a["dram"] for a in wf.values()) / segments,
       "sm_mhz": 1965.0, "per_kernel": per_kernel}
json.dump(out, open(f"profiles/{tag}_issue.json", "w"), indent=1)
json.dump({"dram_bytes_per_spp_at_1080p": sum(a["dram"] for a in wf.values()) / renders / spp,
           "source": f"profiles/{tag}_launches.csv: sum of dram__bytes_read.sum + dram__bytes_write.sum over the k_wf_* kernels of {renders} wavefront renders of {spp} spp (C3, 1920x1080), divided by {renders * spp} spp",
           "note": "one bench 'launch' is one render call of spp_per_step samples: traffic = this * spp_per_step"},
          open("profiles/traffic.json", "w"), indent=1)
print("\n".join(lines))
print(json.dumps({k: v for k, v in out.items() if k != "per_kernel"}, indent=1))
