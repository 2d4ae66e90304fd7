"""Golden fixtures (tests/golden/oracle_v1.npz, written by tests/golden/make_golden.py).

CPU: the oracle still produces them bit for bit.  GPU: so does the CUDA path through the C ABI.
(They are oracle outputs, not outputs of a Rust build — see the generating script.)"""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "oracle_v1.npz"))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(spec)
spec.loader.exec_module(make_golden)


def _same(a, b, name):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, name
    assert a.tobytes() == b.tobytes(), name


def test_oracle_reproduces_the_golden_fixtures(orc, tmp_path):
    now = make_golden.build_cases(str(tmp_path))
    assert set(now) == set(GOLD.files)
    for k in GOLD.files:
        _same(now[k], GOLD[k], k)


@pytest.mark.gpu
def test_device_reproduces_the_golden_fixtures(tmp_path):
    import rtb200
    import rtb200.ctypes_types as T
    from rtb200 import scenes
    ctx = rtb200.Context(0)
    text = os.path.join(str(tmp_path), "c.txt")
    open(text, "w").write(scenes.cornell_text(32, 24, 2, 6, scenes.CORNELL_EXTRA))
    gltf = os.path.join(str(tmp_path), "hf.gltf")
    scenes.write_heightfield_gltf(gltf, nx=12, nz=16, lights=True, spheres=True, sphere_subdiv=1)
    for name, scene in (("cornell", rtb200.Scene.load_text(text)), ("hf", rtb200.Scene.load_gltf(gltf, 40, 24, 2))):
        ctx.upload(scene)
        for mode in (T.MODE_EXACT, T.MODE_FAST):
            hits, rays = ctx.trace_primary(0.5, 0.5, mode, want_rays=True)
            _same(hits["kind"], GOLD[name + "_kind"], name)
            _same(hits["prim"], GOLD[name + "_prim"], name)
            hit = hits["kind"] >= 0
            assert hits["t"][hit].tobytes() == GOLD[name + "_t"][hit].tobytes()
        # radiance: the loop carries throughput instead of unwinding the reference's recursion
        for mode, sched in ((T.MODE_EXACT, T.SCHED_MEGAKERNEL), (T.MODE_FAST, T.SCHED_WAVEFRONT)):
            ctx.render_accum(0, 2, make_golden.SEED, mode, sched, clear=True)
            np.testing.assert_allclose(ctx.accum().reshape(GOLD[name + "_accum"].shape), GOLD[name + "_accum"], rtol=1e-9, atol=1e-12)
        if name == "cornell":
            _same(hits["shading_normal"][hit], GOLD["cornell_ns"][hit], "ns")
            pos = rays[hit, :3] + rays[hit, 3:] * hits["t"][hit, None]
            for mode in (T.MODE_EXACT, T.MODE_FAST):
                _same(ctx.light_pdf(pos, GOLD["cornell_light_dirs"], mode), GOLD["cornell_light_pdf"], "light pdf")
        else:
            b = scene.bvh(T.BVH_TRIANGLES)
            for k in ("bounds", "left", "right", "first", "count", "prim_order"):
                _same(b[k], GOLD["hf_bvh_" + k], k)
    u, probe = ctx.debug_rng(make_golden.SEED, 1234, 5, 16)
    _same(u, GOLD["rng_u64"], "rng")
    _same(probe, GOLD["rng_probe"], "rng probe")
    ctx.close()
