"""Per-call fixed cost of the render path: the rate of short render calls against
the 256-spp call, and rt_render_progressive in 20 chunks against rt_render, on the C3 scene at 1080p.
Development / evidence script (prints one JSON line), not the bench."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
import rtb200.ctypes_types as T
from rtb200 import scenes

os.makedirs("/tmp/sc", exist_ok=True)
path = "/tmp/sc/c3.gltf"
if not os.path.exists(path):
    scenes.write_heightfield_gltf(path, lights=True, spheres=True)
spp_total = int(os.environ.get("SPP", 256))
scene = rtb200.Scene.load_gltf(path, 1920, 1080, spp_total)
ctx = rtb200.Context(0).upload(scene)
out = {}


def rate(spp, reps):
    ctx.render_accum(0, spp, 1, clear=True)
    ctx.sync()
    ctx.reset_counters()
    t = time.perf_counter()
    for k in range(reps):
        ctx.render_accum(k * spp, spp, 1, clear=True)
    ctx.sync()
    dt = time.perf_counter() - t
    return ctx.counters()["segments"] / dt / 1e6, dt / reps * 1e3


for spp, reps in ((256, 3), (64, 6), (32, 10), (8, 30), (2, 60), (1, 60)):
    if spp > spp_total:
        continue
    r, ms = rate(spp, reps)
    out[f"spp{spp}"] = {"Mseg_per_s": round(r, 1), "ms_per_call": round(ms, 3)}
base = out[f"spp{min(256, spp_total)}"]["Mseg_per_s"]
for k in list(out):
    out[k]["vs_longest"] = round(out[k]["Mseg_per_s"] / base, 4)
for name, fn in (("render", lambda: ctx.render(seed=1)), ("progressive_20", lambda: ctx.render_progressive((spp_total + 19) // 20, None, seed=1))):
    fn()
    t = time.perf_counter()
    for _ in range(2):
        fn()
    out[name + "_ms"] = round((time.perf_counter() - t) / 2 * 1e3, 2)
out["progressive_over_render"] = round(out["progressive_20_ms"] / out["render_ms"], 4)
print(json.dumps(out))
