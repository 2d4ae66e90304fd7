#!/usr/bin/env python
"""bench.py — headline benchmark: Mrays/s of path segments on BASELINE.json's glTF scenes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3|c4|c5]     our arm (one rank per GPU)
  python bench.py --impl reference [...]                                      the reference's CPU path

Configs (BASELINE.json `configs`):
  c3 (default)  1 010 304-triangle glTF, 1920x1080, 256 spp, depth 8     — the config `metric` is quoted on
  c4            same scene, 3840x2160, 1024 spp sample-sharded over the GPUs (strong scaling by definition)
  c5            20 M-triangle glTF (HBM-resident hierarchy), 1920x1080, 64 spp

A "step" is one pass of the hot path over one batch.  WEAK scaling (c3, c5; `scaling: "weak"`): every
rank renders the config's spp of every pixel on a disjoint global sample range.  STRONG scaling (c4; and
for every config the extra `strong` object of the line): the config's spp are split over the ranks
(rt_shard_sample_range), so the whole job is fixed as N grows; `time_to_1024spp_s` comes from that leg.
In both, the f64 accumulation buffers are sum-reduced to rank 0 INSIDE the library (rt_reduce_accum:
ncclReduce on the context's stream, communicator created by rt_comm_init) and rank 0 tonemaps.  torch is
only the launcher's plumbing: rendezvous, the broadcast of the 128-byte communicator id, barriers, the
max over ranks of the device times.

`value` counts path SEGMENTS (closest-hit scene queries, reference src/raytrace.rs:14) per second with the
scene resident in HBM; `e2e` is the same metric through the C ABI with host buffers: upload of the
flattened scene (H2D), render, reduce, tonemap and read-back of the image (D2H) inside the timed region.
The reference arm is the CPU oracle (oracle/, a C++ restatement of the Rust renderer — the Rust
toolchain is not available, see oracle/README.md) on all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mrays/s (path segments)"
UNIT = "Mrays/s"

CONFIGS = {
    # name: heightfield quads (nx x nz -> 2*nx*nz triangles + lights + spheres), image, spp of the job, default scaling
    "c3": dict(nx=500, nz=1000, width=1920, height=1080, spp=256, scaling="weak",
               what="configs[2]: full path tracing of the 1M-triangle glTF scene, 1920x1080, 256 spp, depth 8"),
    "c4": dict(nx=500, nz=1000, width=3840, height=2160, spp=1024, scaling="strong",
               what="configs[3]: same scene at 3840x2160, 1024 spp sample-sharded across the GPUs, NCCL reduce of the accumulation buffers"),
    "c5": dict(nx=2000, nz=5000, width=1920, height=1080, spp=64, scaling="weak",
               what="configs[4]: 20M-triangle glTF scene (HBM-resident hierarchy), 1920x1080, 64 spp, depth 8"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"], help="default: the config's (c3, c5 weak; c4 strong)")
    ap.add_argument("--spp-per-step", type=int, default=None, help="override the config's samples per pixel of one step")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--nx", type=int, default=None, help="heightfield quads in x")
    ap.add_argument("--nz", type=int, default=None)
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"])
    ap.add_argument("--sched", default="wavefront", choices=["megakernel", "wavefront"])
    ap.add_argument("--seed", type=lambda s: int(s, 0), default=0xC0FFEE)
    ap.add_argument("--shard", default="samples", choices=["samples", "rows"],
                    help="N>1: every rank renders its own sample range of the whole image (default), or all the "
                         "samples of its own band of rows (tile sharding; same total work, shows the load imbalance)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the extra strong-scaling leg")
    ap.add_argument("--strong-steps", type=int, default=3)
    ap.add_argument("--ref-spp", type=int, default=1, help="samples per step of the CPU arms (a bounded sample of the workload)")
    ap.add_argument("--cache-dir", default=os.environ.get("RT_BVH_CACHE") or os.path.join(tempfile.gettempdir(), "rtb200_bench_cache"))
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    for k in ("nx", "nz", "width", "height"):
        if getattr(args, k) is None:
            setattr(args, k, cfg[k])
    args.spp = args.spp_per_step if args.spp_per_step is not None else cfg["spp"]
    if args.scaling is None:
        args.scaling = cfg["scaling"]
    return args


def scene_path(args):
    return os.path.join(args.cache_dir, f"hf_{args.nx}x{args.nz}.gltf")


def make_scene_file(args):
    """Writes the synthetic glTF (idempotent; a marker file says it is complete)."""
    from rtb200 import scenes
    os.makedirs(args.cache_dir, exist_ok=True)
    path, done = scene_path(args), scene_path(args) + ".ok"
    if os.path.exists(done):
        return path, int(open(done).read())
    n = scenes.write_heightfield_gltf(path, nx=args.nx, nz=args.nz, extent_x=100.0 * args.nx / 500.0,
                                      extent_z=200.0 * args.nz / 1000.0, lights=True, spheres=True)
    with open(done, "w") as fh:
        fh.write(str(n))
    return path, n


def workload_name(args, n_tris):
    return (f"{args.config.upper()}: synthetic {n_tris}-triangle glTF (heightfield + 64 emissive triangles + metallic and dielectric "
            f"icospheres), {args.width}x{args.height}, depth 8, full path tracing with light/MIS sampling")


def config_block(args, n_tris, world):
    """The part of `config` both arms share (the driver compares them)."""
    return {"workload": workload_name(args, n_tris), "name": args.config, "baseline_config": CONFIGS[args.config]["what"],
            "spp": args.spp, "width": args.width, "height": args.height, "triangles": n_tris,
            "l2": "inputs larger than L2: scene records (>= 313 MB) + path state (>= 0.6 GB) per GPU against 126 MB of L2, no flush"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(counters):
    """bytes one launch has to touch by the reference's own algorithm.
    B_seg = N_aabb*64 + N_prim*80 + 112 (shading record) + 96 (ray state); light-PDF rays add
    N'_aabb*64 + N'_prim*80.  Counts are the deterministic EXACT-mode counters (== the oracle's)."""
    seg = counters["segments"]
    return (counters["aabb_tests"] * 64 + counters["prim_tests"] * 80 + seg * (112 + 96) +
            counters["light_aabb_tests"] * 64 + counters["light_prim_tests"] * 80)


def run_reference(args):
    """The reference's CPU implementation of the path (oracle port, all host threads) on a bounded
    sample of the same workload: --ref-spp samples of every pixel per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import loaders, orc
    orc.build()
    path, n_tris = make_scene_file(args)
    scene = orc.OracleScene(loaders.load_gltf(path, args.width, args.height, args.ref_spp))
    threads = orc.hardware_threads()
    for w in range(min(args.warmup, 2)):
        scene.render_accum(1000 + w, 1, args.seed, 0)
    seg = 0
    t0 = time.perf_counter()
    for k in range(args.steps):
        _, c = scene.render_accum(k * args.ref_spp, args.ref_spp, args.seed, 0)
        seg += c["segments"]
    dt = time.perf_counter() - t0
    value = seg / dt / 1e6
    sample = f"{args.ref_spp} spp of the {args.width}x{args.height} image per step ({args.steps} steps), same scene"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / max(1, args.steps) * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_block(args, n_tris, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def issue_roofline(kernel_ms, seg_per_launch, sm_mhz, n_sms=148):
    """The resource that binds this path is warp-instruction ISSUE (the FAST working set is L2-resident, DRAM
    runs at a fifth of its peak).  achieved = thread-instructions per second: the per-segment instruction
    counts of the tracked ncu launch list (profiles/r02_issue.json: smsp__thread_inst_executed.sum and friends,
    summed over every kernel of whole render calls, divided by the segments of those calls) times the
    segments per second measured live; peak = SMs x 4 schedulers x 32 lanes x SM clock."""
    path = os.path.join(ROOT, "profiles", "r02_issue.json")
    if not os.path.exists(path):
        return None
    p = json.load(open(path))
    clock_hz = (sm_mhz or p.get("sm_mhz") or 1965.0) * 1e6
    peak = n_sms * 4 * 32 * clock_hz
    seg_per_s = seg_per_launch / (kernel_ms * 1e-3)
    achieved = p["thread_inst_per_segment"] * seg_per_s
    out = {"bound": "issue", "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "T thread-inst/s", "frac": achieved / peak,
           "warp_inst_per_segment": p["warp_inst_per_segment"], "thread_inst_per_segment": p["thread_inst_per_segment"],
           "lanes_per_instruction": p["thread_inst_per_segment"] / p["warp_inst_per_segment"],
           "warp_issue_frac": p["warp_inst_per_segment"] * seg_per_s / (n_sms * 4 * clock_hz),
           "fp64_issue_frac": p["fp64_warp_inst_per_segment"] * seg_per_s * 32 / (n_sms * 64 * clock_hz) if "fp64_warp_inst_per_segment" in p else None,
           "per_kernel": p.get("per_kernel"), "source": p.get("source")}
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import rtb200
    import rtb200.ctypes_types as T
    from rtb200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(1, args.gpus) and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    mode = T.MODE_FAST if args.mode == "fast" else T.MODE_EXACT
    sched = T.SCHED_MEGAKERNEL if args.sched == "megakernel" else T.SCHED_WAVEFRONT
    S = args.spp

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- scene: rank 0 writes the glTF and fills the on-disk cache of the built hierarchies, the other ranks read it
    t_load = time.perf_counter()
    if rank == 0:
        path, n_tris = make_scene_file(args)
        scene = rtb200.Scene.load_gltf(path, args.width, args.height, S, cache_dir=args.cache_dir)
    barrier()
    if rank != 0:
        path, n_tris = make_scene_file(args)
        scene = rtb200.Scene.load_gltf(path, args.width, args.height, S, cache_dir=args.cache_dir)
    load_s = time.perf_counter() - t_load
    ctx = rtb200.Context(local_rank).upload(scene)
    if world > 1:
        # the library's own communicator: rank 0 makes the id, torch carries the 128 bytes
        box = [rtb200.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(box[0], rank, world)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=local_rank)
    scene_bytes = ctx.scene_device_bytes()

    def render_shard(k, weak=True):
        """this rank's share of pass k"""
        if args.shard == "rows" and world > 1:
            band = sharding.row_band(rank, world, args.width, args.height)
            total = S * world if weak else S
            if band is None:
                ctx.render_accum(k * total, 0, args.seed, mode, sched, clear=True)
            else:
                ctx.render_accum(k * total, total, args.seed, mode, sched, clear=True, region=band)
        elif weak:
            ctx.render_accum(*sharding.step_sample_range(rank, world, k, S), args.seed, mode, sched, clear=True)
        else:
            ctx.render_accum(*rtb200.shard_sample_range(rank, world, S, k * S), args.seed, mode, sched, clear=True)

    def step(k, weak=True):
        """one pass: render this rank's samples, sum-reduce inside the library, tonemap on rank 0"""
        render_shard(k, weak)
        ctx.reduce_accum(0)
        if rank == 0:
            ctx.postprocess_device(S * world if weak else S)

    def timed_loop(n_steps, weak, first_k=0):
        """device time of n_steps steps on the launching stream (max over ranks), segments of all ranks"""
        barrier()
        ctx.reset_counters()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        barrier()
        ev[0].record(stream)
        for k in range(n_steps):
            k_ev[k][0].record(stream)
            render_shard(first_k + k, weak)
            k_ev[k][1].record(stream)
            ctx.reduce_accum(0)
            if rank == 0:
                ctx.postprocess_device(S * world if weak else S)
        ev[1].record(stream)
        barrier()
        ms = ev[0].elapsed_time(ev[1])
        kernel_ms = sum(a.elapsed_time(b) for a, b in k_ev) / max(1, n_steps)
        c = ctx.counters()
        t = torch.tensor([ms, kernel_ms, float(c["segments"]), float(c["light_pdf_rays"]), float(c["paths"]), float(c.get("exact_fallbacks", 0)),
                          float(c["kernel_launches"])], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t[0], t[1] = tmax[0], tmax[1]
        return dict(ms=float(t[0]), kernel_ms=float(t[1]), segments=float(t[2]), lrays=float(t[3]), paths=float(t[4]),
                    xf=float(t[5]), launches_rank0=int(c["kernel_launches"]))

    # ---- deterministic per-segment counts for the algorithmic-bytes figure (EXACT mode == the oracle's counters)
    ctx.reset_counters()
    ctx.render_accum(0, 1, args.seed, T.MODE_EXACT, T.SCHED_MEGAKERNEL, clear=True)
    exact_c = ctx.counters()
    bytes_per_segment = algorithmic_bytes(exact_c) / max(1, exact_c["segments"])

    # ---- timed region (the config's scaling)
    weak = args.scaling == "weak"
    for w in range(args.warmup):
        step(10_000 + w, weak)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    main_leg = timed_loop(args.steps, weak)
    clocks = sampler.stop() if rank == 0 else None
    ms, kernel_ms, segments = main_leg["ms"], main_leg["kernel_ms"], main_leg["segments"]
    value = segments / (ms * 1e-3) / 1e6

    # ---- strong-scaling leg: the config's spp split over the ranks (the whole job fixed as N grows)
    strong = None
    if not args.no_strong:
        if weak:
            for w in range(2):
                step(30_000 + w, False)
            sl = timed_loop(args.strong_steps, False, first_k=40_000)
            n_steps = args.strong_steps
        else:
            sl, n_steps = main_leg, args.steps
        strong = {"value": sl["segments"] / (sl["ms"] * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": sl["ms"] / max(1, n_steps), "steps": n_steps,
                  "spp_total": S, "spp_per_gpu": [rtb200.shard_sample_range(r, world, S)[1] for r in range(world)],
                  "time_to_1024spp_s": sl["ms"] * 1e-3 / max(1, n_steps) / S * 1024,
                  "what": f"the config's {S} spp split over {world} GPU(s) by rt_shard_sample_range, reduce + tonemap included"}

    # ---- e2e through the C ABI with host buffers: upload (H2D) + render + reduce + tonemap + image D2H
    img = np.zeros((args.height, args.width, 3))
    e2e_steps = args.steps if scene_bytes < (1 << 30) else min(args.steps, 3)
    for w in range(2):
        ctx.upload(scene)
        render_shard(20_000 + w, weak)
        ctx.reduce_accum(0)
        ctx.postprocess_into(S, img)
    barrier()
    ctx.reset_counters()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        ctx.upload(scene)                                                   # H2D: the flattened scene
        render_shard(k, weak)
        ctx.reduce_accum(0)
        if rank == 0:
            ctx.postprocess_into(S * world if weak else S, img)             # D2H: the image
        else:
            ctx.sync()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_c = ctx.counters()
    te = torch.tensor([e2e_s, float(e2e_c["segments"])], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        tm = te.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(te, op=dist.ReduceOp.SUM)
        e2e_s = float(tm[0])
    e2e_value = float(te[1]) / e2e_s / 1e6
    comm = ctx.comm_info()

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
        seg_per_launch = segments / world / max(1, args.steps)
        achieved = seg_per_launch * bytes_per_segment / (kernel_ms * 1e-3) / 1e9
        prof = {}
        pj = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(pj):
            prof = json.load(open(pj))
        spp_per_launch = seg_per_launch / max(1.0, segments / max(1.0, main_leg["paths"])) / (args.width * args.height)
        traffic = None
        if args.sched == "wavefront" and args.config in ("c3", "c4") and "dram_bytes_per_spp_at_1080p" in prof:
            traffic = prof["dram_bytes_per_spp_at_1080p"] * spp_per_launch * (args.width * args.height / 2073600.0)
        kernel_name = ("k_wf_confirm + k_wf_shade_hit + k_wf_regen + k_wf_trace per iteration (one wavefront render call)" if args.sched == "wavefront" else "k_render_mega<FastPolicy>") if args.mode == "fast" else "k_render_mega<ExactPolicy>"
        hbm = {"achieved": achieved, "peak": peak, "unit": "GB/s", "algorithmic_hbm_frac": achieved / peak, "traffic": traffic,
               "measured_dram_frac": (traffic / (kernel_ms * 1e-3) / 1e9 / peak) if traffic else None,
               "traffic_source": prof.get("source"), "peak_source": peak_src, "algorithmic_bytes_per_segment": bytes_per_segment,
               "aabb_tests_per_segment": exact_c["aabb_tests"] / max(1, exact_c["segments"]),
               "prim_tests_per_segment": exact_c["prim_tests"] / max(1, exact_c["segments"]),
               "note": "SURVEY §8(d) figure: the REFERENCE algorithm's bytes per segment / device time; the real DRAM traffic "
                       "(`traffic`) is a fraction of it because the fp32 culling hierarchy is L2-resident"}
        issue = issue_roofline(kernel_ms, seg_per_launch, clocks.get("sm_mhz") if clocks else None) if args.sched == "wavefront" and args.mode == "fast" else None
        if issue is not None:
            roofline = dict(issue, kernel=kernel_name, kernel_ms_per_launch=kernel_ms, traffic=traffic, hbm=hbm, algorithmic_hbm_frac=hbm["algorithmic_hbm_frac"])
        else:       # no tracked instruction counts for this path: the contract's HBM form
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                        "kernel": kernel_name, "kernel_ms_per_launch": kernel_ms, "hbm": hbm, "algorithmic_hbm_frac": hbm["algorithmic_hbm_frac"]}
        cfg = config_block(args, n_tris, world)       # the same object in both arms (the driver compares them)
        details = {"spp_per_step_per_gpu": S if weak else [rtb200.shard_sample_range(r, world, S)[1] for r in range(world)],
                   "mode": args.mode, "scheduler": args.sched, "shard": args.shard if world > 1 else "none",
                   "reduce": f"rt_reduce_accum (in-library ncclReduce, f64, NCCL {comm['nccl_version']}, {comm['world']} ranks)" if world > 1 else "none (1 GPU)",
                   "scene_device_bytes": int(scene_bytes), "scene_load_s": load_s, "scene_cache_hit": bool(scene.cache_hit),
                   "segments_per_path": segments / max(1.0, main_leg["paths"]), "light_pdf_rays_per_segment": main_leg["lrays"] / max(1.0, segments),
                   "exact_fallbacks_per_million_segments": 1e6 * main_leg["xf"] / max(1.0, segments)}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / max(1, args.steps), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg, "details": details,
            "time_to_1024spp_s": strong["time_to_1024spp_s"] if strong else (ms * 1e-3) / max(1, args.steps) / (S * world if weak else S) * 1024,
            "strong": strong,
            "gpu_launches": main_leg["launches_rank0"],
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(scene_bytes), "d2h_bytes_per_step": int(img.nbytes),
                    "ms_per_step": e2e_s / max(1, e2e_steps) * 1e3, "steps": e2e_steps},
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import loaders, orc
            orc.build()
            osc = orc.OracleScene(loaders.load_gltf(path, args.width, args.height, 1))
            t0 = time.perf_counter()
            _, oc = osc.render_accum(0, 1, args.seed, 0)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": oc["segments"] / dt / 1e6, "unit": UNIT, "cores": orc.hardware_threads(), "kind": "port",
                                   "sample": f"1 spp of the {args.width}x{args.height} image ({oc['paths']} paths, {oc['segments']} segments), same scene and seed",
                                   "counters_equal_gpu_exact": all(oc[k] == exact_c[k] for k in ("segments", "aabb_tests", "prim_tests", "light_aabb_tests", "light_prim_tests"))}
        print(json.dumps(out), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
