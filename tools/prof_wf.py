"""Renders SPP samples of the C3 scene with the chosen scheduler (profiling target)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
import rtb200.ctypes_types as T
from rtb200 import scenes

spp = int(os.environ.get("SPP", 16))
sched = T.SCHED_WAVEFRONT if os.environ.get("SCHED", "wavefront") == "wavefront" else T.SCHED_MEGAKERNEL
reps = int(os.environ.get("REPS", 2))
os.makedirs("/tmp/sc", exist_ok=True)
path = "/tmp/sc/p.gltf"
scenes.write_heightfield_gltf(path, lights=True, spheres=True)
W, H = int(os.environ.get("W", 1920)), int(os.environ.get("H", 1080))
scene = rtb200.Scene.load_gltf(path, W, H, spp)
ctx = rtb200.Context(0).upload(scene)
for rep in range(reps):
    ctx.reset_counters()
    ctx.sync()
    t = time.time()
    ctx.render_accum(0, spp, 0xC0FFEE, T.MODE_FAST, sched, clear=True)
    ctx.sync()
    dt = time.time() - t
    c = ctx.counters()
    print("render %d spp: %.1f ms, %.1f Mseg/s launches %d%s" % (spp, dt * 1e3, c["segments"] / dt / 1e6, c["kernel_launches"],
          " exact fallbacks %d" % c["exact_fallbacks"]), flush=True)
