#!/usr/bin/env python
"""bench.py — headline benchmark: Mrays/s of path segments on the 1M-triangle 1080p glTF scene.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one rank per GPU)
  python bench.py --impl reference [...]                          the reference's CPU path

A "step" is one pass of the hot path over one batch: every rank renders --spp-per-step samples
of every pixel of the 1920x1080 image (disjoint global sample ranges: weak scaling), the f64
accumulation buffers are sum-reduced to rank 0 over NCCL, and rank 0 tonemaps.  `value` counts
path SEGMENTS (closest-hit scene queries, reference src/raytrace.rs:14) per second with the scene
resident in HBM; `e2e` is the same metric through the C ABI with host buffers: upload of the
flattened scene (H2D), render, tonemap and read-back of the image (D2H) inside the timed region.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp-per-step", type=int, default=256, help="samples per pixel of one step (one render call); 256 = the whole C3 job")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--nx", type=int, default=500, help="heightfield quads in x (500 x 1000 = 1M triangles)")
    ap.add_argument("--nz", type=int, default=1000)
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"])
    ap.add_argument("--sched", default="wavefront", choices=["megakernel", "wavefront"])
    ap.add_argument("--seed", type=lambda s: int(s, 0), default=0xC0FFEE)
    ap.add_argument("--shard", default="samples", choices=["samples", "rows"],
                    help="N>1: every rank renders its own sample range of the whole image (default), or all the "
                         "samples of its own band of rows (tile sharding; same total work, shows the load imbalance)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-spp", type=int, default=1, help="samples per step of the CPU reference arm")
    return ap.parse_args()


def make_scene_file(args, tmp):
    from rtb200 import scenes
    path = os.path.join(tmp, "c3.gltf")
    n = scenes.write_heightfield_gltf(path, nx=args.nx, nz=args.nz, extent_x=100.0 * args.nx / 500.0,
                                      extent_z=200.0 * args.nz / 1000.0, lights=True, spheres=True)
    return path, n


def workload_name(args, n_tris):
    return (f"C3: synthetic {n_tris}-triangle glTF (heightfield + 64 emissive triangles + metallic and dielectric "
            f"icospheres), {args.width}x{args.height}, depth 8, full path tracing with light/MIS sampling")


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
    """The reference's CPU implementation of the path (oracle port, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import loaders, orc
    orc.build()
    with tempfile.TemporaryDirectory() as tmp:
        path, n_tris = make_scene_file(args, tmp)
        scene = orc.OracleScene(loaders.load_gltf(path, args.width, args.height, args.ref_spp))
    threads = orc.hardware_threads()
    for w in range(args.warmup):
        scene.render_accum(1000 + w, 1 if args.ref_spp > 1 else 1, args.seed, 0)
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
        "warmup": args.warmup, "ms_per_step": dt / max(1, args.steps) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, n_tris), "spp_per_step": args.ref_spp},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


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
    S = args.spp_per_step

    tmp = tempfile.TemporaryDirectory()
    path, n_tris = make_scene_file(args, tmp.name)
    scene = rtb200.Scene.load_gltf(path, args.width, args.height, S)
    ctx = rtb200.Context(local_rank).upload(scene)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=local_rank)
    dptr, n_doubles = ctx.accum_device_ptr()

    class _Accum:       # the library's accumulation buffer as a torch tensor (no copy)
        __cuda_array_interface__ = {"shape": (n_doubles,), "typestr": "<f8", "data": (dptr, False), "version": 2}
    accum_t = torch.as_tensor(_Accum(), device=f"cuda:{local_rank}")
    scene_bytes = ctx.scene_device_bytes()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def render_shard(k):
        """this rank's share of pass k: S samples of every pixel, or S*world samples of its rows"""
        if args.shard == "rows" and world > 1:
            band = sharding.row_band(rank, world, args.width, args.height)
            if band is None:
                ctx.render_accum(k * world * S, 0, args.seed, mode, sched, clear=True)
            else:
                ctx.render_accum(k * world * S, S * world, args.seed, mode, sched, clear=True, region=band)
        else:
            ctx.render_accum(*sharding.step_sample_range(rank, world, k, S), args.seed, mode, sched, clear=True)

    def step(k):
        """one pass: S spp per GPU on a disjoint global sample range, reduce, tonemap"""
        render_shard(k)
        if world > 1:
            with torch.cuda.stream(stream):
                dist.reduce(accum_t, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            ctx.postprocess_device(S * world)

    # ---- deterministic per-segment counts for the roofline (EXACT mode == the oracle's counters)
    ctx.reset_counters()
    ctx.render_accum(0, 1, args.seed, T.MODE_EXACT, T.SCHED_MEGAKERNEL, clear=True)
    exact_c = ctx.counters()
    bytes_per_segment = algorithmic_bytes(exact_c) / max(1, exact_c["segments"])

    # ---- timed region: device time on the launching stream, max over ranks
    for w in range(args.warmup):
        step(10_000 + w)
    barrier()
    ctx.reset_counters()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ev[0].record(stream)
    for k in range(args.steps):
        k_ev[k][0].record(stream)
        render_shard(k)
        k_ev[k][1].record(stream)
        if world > 1:
            with torch.cuda.stream(stream):
                dist.reduce(accum_t, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            ctx.postprocess_device(S * world)
    ev[1].record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev[0].elapsed_time(ev[1])
    kernel_ms = sum(a.elapsed_time(b) for a, b in k_ev) / max(1, args.steps)
    counters = ctx.counters()
    t = torch.tensor([ms, kernel_ms, float(counters["segments"]), float(counters["light_pdf_rays"]), float(counters["paths"])],
                     dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, kernel_ms = float(tmax[0]), float(tmax[1])
    segments, lrays, paths = float(t[2]), float(t[3]), float(t[4])
    value = segments / (ms * 1e-3) / 1e6

    # ---- e2e through the C ABI with host buffers: upload (H2D) + render + tonemap + image D2H
    img = np.zeros((args.height, args.width, 3))
    for w in range(2):
        ctx.upload(scene)
        render_shard(20_000 + w)
        ctx.postprocess_into(S, img)
    barrier()
    ctx.reset_counters()
    t0 = time.perf_counter()
    for k in range(args.steps):
        ctx.upload(scene)                                                   # H2D: the flattened scene
        render_shard(k)
        if world > 1:
            dptr2, _ = ctx.accum_device_ptr()
            assert dptr2 == dptr, "accumulation buffer moved"
            with torch.cuda.stream(stream):
                dist.reduce(accum_t, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            ctx.postprocess_into(S * world, img)                            # D2H: the image
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
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, n_tris), "spp_per_step_per_gpu": S, "mode": args.mode,
                       "scheduler": args.sched, "shard": args.shard if world > 1 else "none", "l2": f"inputs larger than L2: {scene_bytes / 1e6:.0f} MB of scene records per GPU, no flush",
                       "segments_per_path": segments / max(1.0, paths), "light_pdf_rays_per_segment": lrays / max(1.0, segments),
                       "exact_fallbacks_per_million_segments": 1e6 * counters.get("exact_fallbacks", 0) / max(1.0, segments / world)},
            "time_to_1024spp_s": (ms * 1e-3) / max(1, args.steps) / (S * world) * 1024,
            "gpu_launches": int(counters["kernel_launches"]),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(scene_bytes), "d2h_bytes_per_step": int(img.nbytes),
                    "ms_per_step": e2e_s / max(1, args.steps) * 1e3},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (prof["dram_bytes_per_spp_at_1080p"] * S * (args.width * args.height / 2073600.0)) if (args.sched == "wavefront" and "dram_bytes_per_spp_at_1080p" in prof) else None,
                         "traffic_source": prof.get("source"), "peak_source": peak_src, "kernel": ("k_wf_confirm + k_wf_shade + k_wf_trace per iteration (one wavefront render call)" if args.sched == "wavefront" else "k_render_mega<FastPolicy>") if args.mode == "fast" else "k_render_mega<ExactPolicy>",
                         "kernel_ms_per_launch": kernel_ms, "algorithmic_bytes_per_segment": bytes_per_segment,
                         "aabb_tests_per_segment": exact_c["aabb_tests"] / max(1, exact_c["segments"]),
                         "prim_tests_per_segment": exact_c["prim_tests"] / max(1, exact_c["segments"])},
        }
        # issue-side view of the same kernels (static, from the tracked ncu summaries): the path is
        # latency / issue bound, not HBM bound, so lanes per instruction and issue activity say more
        # about kernel quality than the HBM fraction
        ncu_path = os.path.join(ROOT, "profiles", "r01_ncu_full_summary.json")
        if args.sched == "wavefront" and os.path.exists(ncu_path):
            ns = json.load(open(ncu_path))
            out["issue"] = {k: {"lanes_per_instruction": v["smsp__thread_inst_executed_per_inst_executed.ratio"][0],
                                "issue_active_pct": v["smsp__issue_active.avg.pct_of_peak_sustained_active"][0],
                                "registers": v["launch__registers_per_thread"][0],
                                "source": "profiles/r01_ncu_full_summary.json (ncu --set full, one launch)"} for k, v in ns.items()}
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
    tmp.cleanup()


if __name__ == "__main__":
    main()
