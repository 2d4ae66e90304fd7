"""Diagnostics for tests/test_gpu_fuzz.py: prints how FAST and EXACT differ, per ray kind."""
import importlib.util
import sys

import numpy as np

sys.path.insert(0, ".")
import rtb200
import rtb200.ctypes_types as T
from rtb200.scene_desc import RawScene

spec = importlib.util.spec_from_file_location("fz", "tests/test_gpu_fuzz.py")
fz = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fz)
ctx = rtb200.Context(0)
for flavour in sys.argv[1:] or ["far", "scales", "slivers"]:
    for seed in (1, 2, 3):
        rng = np.random.default_rng(1000 * seed + len(flavour))
        n = int(rng.integers(300, 3000))
        tri = fz._soup(rng, n, flavour)
        raw = RawScene(width=8, height=8)
        raw.materials = [{"kind": T.MAT_DIFFUSE, "color": (0.5, 0.5, 0.5)}, {"kind": T.MAT_DIFFUSE, "emission": (2.0, 2.0, 2.0)}]
        raw.tri_pos = tri
        raw.tri_mat = (rng.uniform(size=len(tri)) < 0.15).astype(np.uint32)
        ctx.upload(rtb200.Scene.from_raw(raw))
        o, d, kind = fz._rays(rng, tri, 40000)
        e = ctx.trace_rays(o, d, T.MODE_EXACT)
        f = ctx.trace_rays(o, d, T.MODE_FAST)
        bad = np.nonzero((e["kind"] != f["kind"]) | (e["prim"] != f["prim"]) | (e["t"].view(np.uint64) != f["t"].view(np.uint64)))[0]
        print(f"{flavour} seed {seed}: n={n} mismatches={len(bad)} by ray kind {np.bincount(kind[bad], minlength=6)}")
        for i in bad[:6]:
            print(f"   ray {i} kind {kind[i]} o={o[i]} d={d[i]}\n      exact: k={e['kind'][i]} prim={e['prim'][i]} t={e['t'][i]!r}   fast: k={f['kind'][i]} prim={f['prim'][i]} t={f['t'][i]!r}")
        dn = d / np.linalg.norm(d, axis=1, keepdims=True)
        ok = np.isfinite(dn).all(1)
        pe = ctx.light_pdf(o[ok], dn[ok], T.MODE_EXACT)
        pf = ctx.light_pdf(o[ok], dn[ok], T.MODE_FAST)
        badp = np.nonzero(pe.view(np.uint64) != pf.view(np.uint64))[0]
        print(f"   light pdf mismatches {len(badp)}", [(int(i), int(kind[ok][i]), float(pe[i]), float(pf[i])) for i in badp[:5]])
