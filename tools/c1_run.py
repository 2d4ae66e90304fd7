"""C1: Cornell box (text format), 512x512, 64 spp, depth 6 — GPU both schedulers + CPU oracle."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
import rtb200.ctypes_types as T
from rtb200 import scenes
from oracle import loaders, orc
os.makedirs("/tmp/sc", exist_ok=True)
path = "/tmp/sc/c1.txt"
open(path, "w").write(scenes.cornell_text(512, 512, 64, 6))
scene = rtb200.Scene.load_text(path)
ctx = rtb200.Context(0).upload(scene)
out = {}
for name, mode, sched in (("wavefront_fast", T.MODE_FAST, T.SCHED_WAVEFRONT), ("megakernel_fast", T.MODE_FAST, T.SCHED_MEGAKERNEL), ("megakernel_exact", T.MODE_EXACT, T.SCHED_MEGAKERNEL)):
    for rep in range(2):
        ctx.reset_counters(); ctx.sync(); t = time.time()
        ctx.render_accum(0, 64, 0xC0FFEE, mode, sched, clear=True); ctx.sync(); dt = time.time() - t
    c = ctx.counters()
    out[name] = {"ms": round(dt * 1e3, 1), "Mseg_per_s": round(c["segments"] / dt / 1e6, 1), "segments": c["segments"], "launches": c["kernel_launches"]}
    out[name + "_sum"] = float(ctx.accum().sum())
osc = orc.OracleScene(loaders.parse_text_scene(open(path).read()))
t = time.time(); acc, oc = osc.render_accum(0, 8, 0xC0FFEE); dt = time.time() - t
out["oracle_8spp"] = {"ms": round(dt * 1e3, 1), "Mseg_per_s": round(oc["segments"] / dt / 1e6, 2), "threads": orc.hardware_threads()}
print(json.dumps(out))
