"""Quick device timing of the hot path on the C3 scene (development aid, not the bench)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
import rtb200.ctypes_types as T
from rtb200 import scenes

nx = int(os.environ.get("NX", 500))
nz = int(os.environ.get("NZ", 1000))
spp = int(os.environ.get("SPP", 1))
os.makedirs("/tmp/sc", exist_ok=True)
path = "/tmp/sc/q.gltf"
t = time.time()
n = scenes.write_heightfield_gltf(path, nx=nx, nz=nz, extent_x=100.0 * nx / 500, extent_z=200.0 * nz / 1000, lights=True, spheres=True)
print("gen", n, round(time.time() - t, 2), flush=True)
t = time.time()
scene = rtb200.Scene.load_gltf(path, 1920, 1080, spp)
print("host build", round(time.time() - t, 2), "depth", scene.info.bvh_depth[2], flush=True)
ctx = rtb200.Context(0)
t = time.time()
ctx.upload(scene)
print("upload", round(time.time() - t, 2), flush=True)
for mode, name in ((T.MODE_EXACT, "exact"), (T.MODE_FAST, "fast")):
    for rep in range(2):
        ctx.reset_counters()
        ctx.sync()
        t = time.time()
        ctx.trace_primary(0.5, 0.5, mode, want_hits=False)
        dt = time.time() - t
    c = ctx.counters()
    print(name, "primary: %.1f ms, %.1f Mrays/s" % (dt * 1e3, c["segments"] / dt / 1e6), c, flush=True)
    for rep in range(2):
        ctx.reset_counters()
        ctx.sync()
        t = time.time()
        ctx.render_accum(0, spp, 0xC0FFEE, mode, clear=True)
        ctx.sync()
        dt = time.time() - t
    c = ctx.counters()
    print(name, "render %d spp: %.1f ms, %.1f Mseg/s" % (spp, dt * 1e3, c["segments"] / dt / 1e6), c, flush=True)
for rep in range(2):
    ctx.reset_counters()
    ctx.sync()
    t = time.time()
    ctx.render_accum(0, spp, 0xC0FFEE, T.MODE_FAST, T.SCHED_WAVEFRONT, clear=True)
    ctx.sync()
    dt = time.time() - t
c = ctx.counters()
print("wavefront render %d spp: %.1f ms, %.1f Mseg/s" % (spp, dt * 1e3, c["segments"] / dt / 1e6), c, flush=True)
acc = ctx.accum()
print("accum mean", acc.mean(axis=(0, 1)))
