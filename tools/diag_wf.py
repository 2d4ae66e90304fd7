import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
import rtb200.ctypes_types as T
from rtb200 import scenes
os.makedirs("/tmp/sc", exist_ok=True)
scenes.write_heightfield_gltf("/tmp/sc/d.gltf", lights=True, spheres=True)
scene = rtb200.Scene.load_gltf("/tmp/sc/d.gltf", 1920, 1080, 4)
ctx = rtb200.Context(0).upload(scene)
ctx.render_accum(0, 4, 0xC0FFEE, T.MODE_FAST, T.SCHED_WAVEFRONT, clear=True)
nc, fl, rays = ctx.debug_wavefront()
ov = nc >> 16
idx = np.nonzero(ov == 3)[0]
print("overflow slots", len(idx))
np.set_printoptions(precision=6, linewidth=200)
print(rays[idx[:12]])
print("flags", [hex(v) for v in fl[idx[:12]]])
r = rays[idx]
print("origin y: min/mean/max", r[:, 1].min(), r[:, 1].mean(), r[:, 1].max(), " |d|:", np.linalg.norm(r[:, 3:], axis=1)[:8])
print("dir y hist", np.histogram(r[:, 4], bins=8, range=(-1, 1))[0])
hits = ctx.trace_rays(r[:2000, :3], r[:2000, 3:], T.MODE_EXACT)
print("exact hit kinds", np.bincount(hits["kind"] + 1), "t min/max", hits["t"].min(), hits["t"].max())
