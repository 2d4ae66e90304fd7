"""C5: synthetic 20M-triangle glTF scene (2000 x 5000 quads), 1080p — host build time, HBM
footprint and render throughput (development / evidence script, not the bench)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
import rtb200.ctypes_types as T
from rtb200 import scenes

nx, nz = int(os.environ.get("NX", 2000)), int(os.environ.get("NZ", 5000))
spp = int(os.environ.get("SPP", 8))
out = {}
os.makedirs("/tmp/sc", exist_ok=True)
t = time.time()
n = scenes.write_heightfield_gltf("/tmp/sc/c5.gltf", nx=nx, nz=nz, extent_x=100.0 * nx / 500, extent_z=200.0 * nz / 1000, lights=True, spheres=True)
out["triangles"], out["generate_s"] = n, round(time.time() - t, 2)
t = time.time()
scene = rtb200.Scene.load_gltf("/tmp/sc/c5.gltf", 1920, 1080, spp)
out["load_and_bvh_build_s"] = round(time.time() - t, 2)
if os.environ.get("CACHE", "1") != "0":      # on-disk cache of the built hierarchies (rt_scene_load_gltf_cached)
    del scene
    os.makedirs("/tmp/sc/cache", exist_ok=True)
    t = time.time()
    scene = rtb200.Scene.load_gltf("/tmp/sc/c5.gltf", 1920, 1080, spp, cache_dir="/tmp/sc/cache")
    out["load_build_and_write_cache_s"], hit0 = round(time.time() - t, 2), scene.cache_hit
    del scene
    t = time.time()
    scene = rtb200.Scene.load_gltf("/tmp/sc/c5.gltf", 1920, 1080, spp, cache_dir="/tmp/sc/cache")
    out["load_from_cache_s"], out["cache_hit"] = round(time.time() - t, 2), [hit0, scene.cache_hit]
    out["cache_file_GB"] = round(sum(os.path.getsize(os.path.join("/tmp/sc/cache", f)) for f in os.listdir("/tmp/sc/cache")) / 1e9, 2)
out["bvh_nodes"], out["bvh_depth"] = int(scene.info.bvh_nodes[2]), int(scene.info.bvh_depth[2])
ctx = rtb200.Context(0)
t = time.time()
ctx.upload(scene)
out["first_upload_s"] = round(time.time() - t, 2)
out["scene_device_GB"] = round(ctx.scene_device_bytes() / 1e9, 2)
for name, sched in (("wavefront", T.SCHED_WAVEFRONT), ("megakernel", T.SCHED_MEGAKERNEL)):
    for rep in range(2):
        ctx.reset_counters()
        ctx.sync()
        t = time.time()
        ctx.render_accum(0, spp, 0xC0FFEE, T.MODE_FAST, sched, clear=True)
        ctx.sync()
        dt = time.time() - t
    c = ctx.counters()
    out[name] = {"ms": round(dt * 1e3, 1), "Mseg_per_s": round(c["segments"] / dt / 1e6, 1), "segments": c["segments"], "paths": c["paths"]}
    out[name + "_accum_sum"] = float(ctx.accum().sum())
# FAST == EXACT on a sample of primary rays at this size
import numpy as np
e, _ = ctx.trace_primary(0.5, 0.5, T.MODE_EXACT)
f, _ = ctx.trace_primary(0.5, 0.5, T.MODE_FAST)
out["primary_fast_equals_exact"] = bool(np.array_equal(e["prim"], f["prim"]) and e["t"].tobytes() == f["t"].tobytes())
print(json.dumps(out))
