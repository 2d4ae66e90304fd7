"""Soak: the wavefront scheduler against the megakernel, bit for bit, over seeds, image sizes, sample counts and
regions on the C3 scene — looks for rare ordering / race effects in the list-driven pipeline (development aid)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
import rtb200.ctypes_types as T
from rtb200 import scenes

os.makedirs("/tmp/sc", exist_ok=True)
path = "/tmp/sc/soak.gltf"
scenes.write_heightfield_gltf(path, lights=True, spheres=True)
rng = np.random.default_rng(int(os.environ.get("SEED", 1)))
bad = 0
cases = [(1920, 1080, 3), (1920, 1080, 17), (640, 360, 40), (333, 217, 9), (64, 48, 33), (3840, 2160, 2)]
for w, h, spp in cases:
    scene = rtb200.Scene.load_gltf(path, w, h, spp)
    ctx = rtb200.Context(0).upload(scene)
    for rep in range(int(os.environ.get("REPS", 3))):
        seed = int(rng.integers(1, 2 ** 62))
        region = None
        if rep == 2:
            x0, y0 = int(rng.integers(0, w // 2)), int(rng.integers(0, h // 2))
            region = (x0, y0, int(rng.integers(1, w - x0 + 1)), int(rng.integers(1, h - y0 + 1)))
        ctx.reset_counters()
        ctx.render_accum(0, spp, seed, T.MODE_FAST, T.SCHED_MEGAKERNEL, clear=True, region=region)
        want, wc = ctx.accum(), ctx.counters()
        for again in range(2):
            ctx.reset_counters()
            ctx.render_accum(0, spp, seed, T.MODE_FAST, T.SCHED_WAVEFRONT, clear=True, region=region)
            got, gc = ctx.accum(), ctx.counters()
            ok = got.tobytes() == want.tobytes() and all(gc[k] == wc[k] for k in ("segments", "light_pdf_rays", "paths"))
            bad += not ok
            print(f"{w}x{h} spp {spp} seed {seed} region {region} run {again}: {'ok' if ok else 'MISMATCH'} segments {gc['segments']}", flush=True)
        # split over sample ranges == one call
        ctx.render_accum(0, spp // 2, seed, T.MODE_FAST, T.SCHED_WAVEFRONT, clear=True, region=region)
        ctx.render_accum(spp // 2, spp - spp // 2, seed, T.MODE_FAST, T.SCHED_WAVEFRONT, region=region)
        ok = ctx.accum().tobytes() == want.tobytes()
        bad += not ok
        print(f"   split at {spp // 2}: {'ok' if ok else 'MISMATCH'}", flush=True)
    ctx.close()
print("mismatches:", bad)
sys.exit(1 if bad else 0)
