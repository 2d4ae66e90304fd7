"""Writes tests/golden/oracle_v1.npz — outputs of the CPU oracle on two small scenes.

The reference is Rust and cannot be built in this environment, so these are NOT outputs of the
reference itself (the only vectors it holds are its 13 inline tests, tests/test_oracle_kat.py).
They freeze what the oracle — and, bit for bit, the CUDA path — produced when the parity suite was
green, so that a change to either that moves a result is caught even on a machine where only one of
the two can run.  Regenerate (only deliberately) with:  python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loaders, orc  # noqa: E402
from rtb200 import scenes  # noqa: E402

SEED = 0xC0FFEE


def build_cases(tmp):
    out = {}
    # Cornell box (text format): planes, boxes, ellipsoids, a model-space triangle, three materials
    text = scenes.cornell_text(32, 24, 2, 6, scenes.CORNELL_EXTRA)
    cornell = orc.OracleScene(loaders.parse_text_scene(text))
    hits, rays, _ = cornell.trace_primary(0.5, 0.5)
    out["cornell_kind"], out["cornell_prim"], out["cornell_t"] = hits["kind"], hits["prim"], hits["t"]
    out["cornell_ns"] = hits["shading_normal"]
    acc, counters = cornell.render_accum(0, 2, SEED)
    out["cornell_accum"] = acc
    out["cornell_counters"] = np.array([counters[k] for k in ("segments", "light_pdf_rays", "paths", "aabb_tests", "prim_tests")], np.uint64)
    hit = hits["kind"] >= 0
    pos = rays[hit, :3] + rays[hit, 3:] * hits["t"][hit, None]
    ld = cornell.light_sample(7, pos)
    out["cornell_light_dirs"] = ld
    out["cornell_light_pdf"] = cornell.light_pdf(pos, ld)[0]
    # heightfield glTF (the C2 / C3 generator at 12 x 16 quads) with lights and spheres
    path = os.path.join(tmp, "hf.gltf")
    scenes.write_heightfield_gltf(path, nx=12, nz=16, lights=True, spheres=True, sphere_subdiv=1)
    hf = orc.OracleScene(loaders.load_gltf(path, 40, 24, 2))
    b = hf.bvh(2)
    for k in ("bounds", "left", "right", "first", "count", "prim_order"):
        out["hf_bvh_" + k] = b[k]
    hits, _, _ = hf.trace_primary(0.5, 0.5)
    out["hf_kind"], out["hf_prim"], out["hf_t"] = hits["kind"], hits["prim"], hits["t"]
    acc, _ = hf.render_accum(0, 2, SEED)
    out["hf_accum"] = acc
    # counter RNG: first draws of one stream and one value of every distribution on the path
    out["rng_u64"] = orc.rng_u64(SEED, 1234, 5, 16)
    out["rng_probe"] = orc.rng_probe(SEED, 1234, 5)
    return out


if __name__ == "__main__":
    orc.build()
    with tempfile.TemporaryDirectory() as tmp:
        cases = build_cases(tmp)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_v1.npz")
    np.savez_compressed(dst, **cases)
    print(dst, os.path.getsize(dst), "bytes;", len(cases), "arrays")
