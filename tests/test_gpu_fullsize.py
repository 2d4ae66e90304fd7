"""Full-size checks at BASELINE.json's sizes (C2: 1 010 304 triangles... here the C3 mesh, 1920x1080).

The oracle needs ~15 s to build its hierarchy and a few seconds per million rays, so the full
primary-ray comparison is one test; the rest are size-independent properties on the GPU alone:
FAST == EXACT, wavefront == megakernel, sample-range additivity."""
import os

import numpy as np
import pytest

import rtb200
import rtb200.ctypes_types as T
from rtb200 import scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big(tmp_path_factory):
    d = tmp_path_factory.mktemp("c3")
    path = os.path.join(str(d), "c3.gltf")
    n = scenes.write_heightfield_gltf(path, lights=True, spheres=True)
    assert n == 1010304
    scene = rtb200.Scene.load_gltf(path, 1920, 1080, 2)
    ctx = rtb200.Context(0).upload(scene)
    yield path, scene, ctx
    ctx.close()


def test_c2_primary_ids_bit_exact_against_oracle(orc, big):
    """configs[1]: primary-ray closest hit on the 1M-triangle mesh at 1920x1080, pixel-centre
    jitter: hierarchy bit-exact, prim ids bit-exact (t and normals too), in EXACT and FAST mode."""
    from oracle import loaders
    path, scene, ctx = big
    oracle = orc.OracleScene(loaders.load_gltf(path, 1920, 1080, 2))
    ob, pb = oracle.bvh(T.BVH_TRIANGLES), scene.bvh(T.BVH_TRIANGLES)
    for k in ("left", "right", "first", "count", "prim_order"):
        assert np.array_equal(ob[k], pb[k]), k
    assert ob["bounds"].tobytes() == pb["bounds"].tobytes()
    want, _, oc = oracle.trace_primary(0.5, 0.5, want_rays=False)
    for mode in (T.MODE_EXACT, T.MODE_FAST):
        ctx.reset_counters()
        got, _ = ctx.trace_primary(0.5, 0.5, mode)
        assert np.array_equal(got["kind"], want["kind"]) and np.array_equal(got["prim"], want["prim"])
        hit = want["kind"] >= 0
        assert hit.sum() > 1_000_000
        assert got["t"][hit].tobytes() == want["t"][hit].tobytes()
        assert got["geometry_normal"][hit].tobytes() == want["geometry_normal"][hit].tobytes()
        assert got["shading_normal"][hit].tobytes() == want["shading_normal"][hit].tobytes()
        if mode == T.MODE_EXACT:
            gc = ctx.counters()
            assert gc["aabb_tests"] == oc["aabb_tests"] and gc["prim_tests"] == oc["prim_tests"]


def test_full_size_scheduler_and_mode_properties(big):
    _, scene, ctx = big
    seed = 0xC0FFEE
    ctx.reset_counters()
    ctx.render_accum(0, 2, seed, T.MODE_EXACT, T.SCHED_MEGAKERNEL, clear=True)
    exact, ec = ctx.accum(), ctx.counters()
    ctx.reset_counters()
    ctx.render_accum(0, 2, seed, T.MODE_FAST, T.SCHED_MEGAKERNEL, clear=True)
    fast, fc = ctx.accum(), ctx.counters()
    ctx.reset_counters()
    ctx.render_accum(0, 2, seed, T.MODE_FAST, T.SCHED_WAVEFRONT, clear=True)
    wave, wc = ctx.accum(), ctx.counters()
    assert exact.sum() > 0 and np.isfinite(exact).all()
    assert fast.tobytes() == exact.tobytes()              # 7.6 M segments, zero mismatches
    assert wave.tobytes() == exact.tobytes()
    for k in ("segments", "light_pdf_rays", "paths"):
        assert ec[k] == fc[k] == wc[k], k
    ctx.render_accum(0, 1, seed, T.MODE_FAST, T.SCHED_WAVEFRONT, clear=True)
    ctx.render_accum(1, 1, seed, T.MODE_FAST, T.SCHED_MEGAKERNEL)
    assert ctx.accum().tobytes() == exact.tobytes()       # additivity across launches and schedulers
