"""Edge cases of the hot path on the GPU, each against the oracle: empty scenes, no lights,
depth 0/1, odd image sizes, background colour, planes hit edge-on (t = inf / NaN), a leaf far
longer than 4 primitives (SameNode, src/bvh.rs:93-96), degenerate triangles."""
import numpy as np
import pytest

import rtb200
import rtb200.ctypes_types as T
from rtb200.scene_desc import RawScene
from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = rtb200.Context(0)
    yield c
    c.close()


def _check(orc, ctx, raw, spp, wavefront=True):
    prod, oracle = rtb200.Scene.from_raw(raw), orc.OracleScene(raw)
    ctx.upload(prod)
    want, oc = oracle.render_accum(0, spp, 99)
    ctx.reset_counters()
    ctx.render_accum(0, spp, 99, T.MODE_EXACT, clear=True)
    exact, ec = ctx.accum(), ctx.counters()
    for k in ("segments", "light_pdf_rays", "paths", "aabb_tests", "prim_tests"):
        assert ec[k] == oc[k], k
    ok = np.isfinite(want)
    assert np.array_equal(np.isfinite(exact), ok)
    err = np.abs(exact[ok] - want[ok]) / np.maximum(np.abs(want[ok]), 1e-300)
    err[exact[ok] == want[ok]] = 0
    assert err.size == 0 or err.max() <= 1e-9
    ctx.render_accum(0, spp, 99, T.MODE_FAST, clear=True)
    assert ctx.accum().tobytes() == exact.tobytes()
    if wavefront:
        ctx.render_accum(0, spp, 99, T.MODE_FAST, T.SCHED_WAVEFRONT, clear=True)
        assert ctx.accum().tobytes() == exact.tobytes()
    g_hits, _ = ctx.trace_primary(0.5, 0.5, T.MODE_FAST)
    c_hits, _, _ = oracle.trace_primary(0.5, 0.5)
    assert np.array_equal(g_hits["kind"], c_hits["kind"]) and np.array_equal(g_hits["prim"], c_hits["prim"])
    return want


def test_empty_scene_returns_background(orc, ctx):
    raw = RawScene(width=13, height=7, samples=2, ray_depth=3, bg_color=(0.25, 0.5, 0.75))
    want = _check(orc, ctx, raw, 2)
    assert np.allclose(want, np.array([0.5, 1.0, 1.5]))


def test_depth_zero_and_one(orc, ctx):
    for depth in (0, 1):
        raw = RawScene(width=9, height=5, samples=3, ray_depth=depth, bg_color=(1.0, 1.0, 1.0), tri_mode=T.TRI_MODEL)
        raw.materials = [{"kind": T.MAT_DIFFUSE, "color": (0.5, 0.5, 0.5), "emission": (2.0, 0.0, 0.0)}]
        raw.shapes = [{"type": T.SHAPE_ELLIPSOID, "material": 0, "param": (0.5, 0.4, 0.3), "position": (0.0, 0.0, 2.0)}]
        want = _check(orc, ctx, raw, 3)
        if depth == 0:
            assert want.sum() == 0.0


def test_no_lights_uses_cosine_sampler_only(orc, ctx):
    rng = np.random.default_rng(3)
    raw = RawScene(width=24, height=16, samples=3, ray_depth=5, bg_color=(0.6, 0.7, 0.9), cam_position=(0.0, 0.5, -3.0))
    raw.materials = [{"kind": T.MAT_DIFFUSE, "color": (0.8, 0.4, 0.2)}, {"kind": T.MAT_METALLIC, "color": (0.9, 0.9, 0.9)},
                     {"kind": T.MAT_DIELECTRIC, "ior": 1.33, "color": (0.9, 1.0, 0.9)}]
    c = rng.uniform(-1.5, 1.5, (300, 1, 3)) + np.array([0.0, 0.0, 2.0])
    raw.tri_pos = (c + rng.normal(scale=0.35, size=(300, 3, 3))).reshape(300, 9)
    raw.tri_mat = rng.integers(0, 3, 300).astype(np.uint32)
    raw.shapes = [{"type": T.SHAPE_PLANE, "material": 0, "param": (0.0, 2.0, 0.0), "position": (0.0, -1.0, 0.0)}]
    _check(orc, ctx, raw, 3)


def test_plane_seen_edge_on_and_axis_parallel_rays(orc, ctx):
    """Camera in the plane of a PLANE looking along it: n.d = 0 gives t = NaN / +-inf, which the
    reference reports as a hit candidate and then drops (plane.rs:12-14, intersections.rs:56)."""
    raw = RawScene(width=8, height=8, samples=2, ray_depth=4, bg_color=(0.1, 0.2, 0.3), cam_position=(0.0, 0.0, 0.0),
                   fov=0.5, tri_mode=T.TRI_MODEL)
    raw.materials = [{"kind": T.MAT_DIFFUSE, "color": (0.7, 0.7, 0.7)}, {"kind": T.MAT_DIFFUSE, "emission": (3.0, 3.0, 3.0), "color": (0.5, 0.5, 0.5)}]
    raw.shapes = [{"type": T.SHAPE_PLANE, "material": 0, "param": (0.0, 1.0, 0.0), "position": (0.0, 0.0, 0.0)},
                  {"type": T.SHAPE_BOX, "material": 1, "param": (0.5, 0.5, 0.5), "position": (0.0, 0.0, 4.0)},
                  {"type": T.SHAPE_BOX, "material": 0, "param": (0.3, 0.3, 0.3), "position": (1.5, 0.3, 4.0), "rotation": (0.0, 0.3826834323650898, 0.0, 0.9238795325112867)}]
    prod, oracle = rtb200.Scene.from_raw(raw), orc.OracleScene(raw)
    ctx.upload(prod)
    o = np.zeros((6, 3))
    d = np.array([[0, 0, 1.0], [1.0, 0, 0], [0, 1.0, 0], [0, -1.0, 0], [0, 0, -1.0], [1.0, 0, 1.0]])
    for mode in (T.MODE_EXACT, T.MODE_FAST):
        g = ctx.trace_rays(o, d, mode)
        c, _ = oracle.trace_rays(o, d)
        assert np.array_equal(g["kind"], c["kind"]) and np.array_equal(g["prim"], c["prim"])
        hit = c["kind"] >= 0
        assert g["t"][hit].tobytes() == c["t"][hit].tobytes()
    _check(orc, ctx, raw, 2)


def test_long_leaf_and_degenerate_triangles(orc, ctx):
    """40 coincident triangles cannot be split by SAH -> one reference leaf of 40 primitives
    (src/bvh.rs:93-96), which the fp32 hierarchy must split into <= 8-triangle runs; plus zero-area
    and needle triangles (|det| < 1e-11 rejects, triangle.rs:51)."""
    base = np.array([-1.0, -1.0, 3.0, 1.0, -1.0, 3.0, 0.0, 1.0, 3.0])
    tris = [base] * 40
    tris.append(np.array([0.0, 0.0, 2.0, 0.0, 0.0, 2.0, 0.0, 0.0, 2.0]))            # a point
    tris.append(np.array([-2.0, 0.0, 2.5, 2.0, 0.0, 2.5, 0.0, 1e-13, 2.5]))          # a needle
    rng = np.random.default_rng(5)
    c = rng.uniform(-2, 2, (30, 1, 3)) + np.array([0.0, 0.0, 5.0])
    tris += list((c + rng.normal(scale=0.5, size=(30, 3, 3))).reshape(30, 9))
    raw = RawScene(width=20, height=12, samples=2, ray_depth=4, bg_color=(0.2, 0.2, 0.2))
    raw.materials = [{"kind": T.MAT_DIFFUSE, "color": (0.6, 0.6, 0.6)}, {"kind": T.MAT_DIFFUSE, "color": (0.5, 0.5, 0.5), "emission": (4.0, 4.0, 4.0)}]
    raw.tri_pos = np.array(tris)
    raw.tri_mat = np.array([0] * 42 + [1] * 5 + [0] * 25, np.uint32)
    prod = rtb200.Scene.from_raw(raw)
    counts = prod.bvh(T.BVH_TRIANGLES)["count"]
    assert counts.max() >= 40
    _check(orc, ctx, raw, 2)


def test_prebuilt_scene_renders_identically(orc, ctx, tmp_path):
    """rt_scene_from_built (the caller's own hierarchies as flat arrays) renders bit for bit like
    the scene it was exported from."""
    from tests import helpers
    path = helpers.cornell(tmp_path, 40, 32, 3, 5)
    prod = rtb200.Scene.load_text(path)
    ctx.upload(prod)
    ctx.render_accum(0, 3, 17, T.MODE_FAST, T.SCHED_WAVEFRONT, clear=True)
    want = ctx.accum()
    built, keep = prod.export_built()
    ctx.upload(rtb200.Scene.from_built(built))
    ctx.render_accum(0, 3, 17, T.MODE_FAST, T.SCHED_WAVEFRONT, clear=True)
    assert ctx.accum().tobytes() == want.tobytes() and want.sum() > 0


def test_cached_scene_renders_identically(orc, ctx, tmp_path):
    """A scene read back from the on-disk hierarchy cache renders bit for bit like a fresh build."""
    from tests import helpers
    path, _ = helpers.heightfield(tmp_path, 24, 32, lights=True, spheres=True)
    cache = tmp_path / "cache"
    cache.mkdir()
    images = []
    for expect_hit in (False, True):
        sc = rtb200.Scene.load_gltf(path, 48, 32, 3, cache_dir=str(cache))
        assert sc.cache_hit == expect_hit
        ctx.upload(sc)
        ctx.render_accum(0, 3, 23, T.MODE_FAST, T.SCHED_WAVEFRONT, clear=True)
        images.append(ctx.accum())
    ctx.upload(rtb200.Scene.load_gltf(path, 48, 32, 3))
    ctx.render_accum(0, 3, 23, T.MODE_FAST, T.SCHED_WAVEFRONT, clear=True)
    want = ctx.accum()
    assert images[0].tobytes() == want.tobytes() == images[1].tobytes() and want.sum() > 0


@pytest.mark.parametrize("sched", [T.SCHED_MEGAKERNEL, T.SCHED_WAVEFRONT])
def test_region_render_equals_the_full_render_there(ctx, tmp_path, sched):
    """Tile sharding: a render restricted to a rectangle writes bit for bit what the full render
    has in that rectangle and leaves the rest of the buffer alone; disjoint row bands (what
    sharding.row_band hands to the ranks) assemble the whole image."""
    from rtb200 import sharding
    from tests import helpers
    path = helpers.cornell(tmp_path, 45, 38, 3, 5)          # odd sizes: partial 8x4 tiles, w != W
    sc = rtb200.Scene.load_text(path)
    ctx.upload(sc)
    ctx.render_accum(0, 3, 99, T.MODE_FAST, sched, clear=True)
    full = ctx.accum()
    x0, y0, w, h = 7, 5, 21, 19
    ctx.render_accum(0, 3, 99, T.MODE_FAST, sched, clear=True, region=(x0, y0, w, h))
    part = ctx.accum()
    assert part[y0:y0 + h, x0:x0 + w].tobytes() == full[y0:y0 + h, x0:x0 + w].tobytes()
    outside = part.copy()
    outside[y0:y0 + h, x0:x0 + w] = 0
    assert not outside.any() and part.any()
    ctx.render_accum(0, 0, 99, T.MODE_FAST, sched, clear=True)
    for rank in range(3):
        ctx.render_accum(0, 3, 99, T.MODE_FAST, sched, region=sharding.row_band(rank, 3, 45, 38))
    assert ctx.accum().tobytes() == full.tobytes()
    with pytest.raises(rtb200.RtError):
        ctx.render_accum(0, 1, 99, T.MODE_FAST, sched, region=(40, 0, 10, 10))


def test_progressive_render(ctx, tmp_path):
    """rt_render_progressive: the callback sees the image of the samples done so far (== a render
    of that many samples), the final image is render()'s, and a truthy return stops early."""
    from tests import helpers
    path = helpers.cornell(tmp_path, 32, 24, 7, 4)
    ctx.upload(rtb200.Scene.load_text(path))
    want = ctx.render(seed=5)
    seen = []
    rgb, done = ctx.render_progressive(3, lambda d, t, img: seen.append((d, t, img.copy())) and False, seed=5)
    assert done == 7 and [s[:2] for s in seen] == [(3, 7), (6, 7), (7, 7)]
    assert rgb.tobytes() == want.tobytes() == seen[-1][2].tobytes()
    ctx.render_accum(0, 3, 5, clear=True)
    assert ctx.postprocess(3).tobytes() == seen[0][2].tobytes()
    rgb2, done2 = ctx.render_progressive(3, lambda d, t, img: d >= 6, seed=5)
    assert done2 == 6 and rgb2.tobytes() == seen[1][2].tobytes()
    with pytest.raises(ZeroDivisionError):                   # exceptions come back out, after the C call returned
        ctx.render_progressive(3, lambda d, t, img: 1 // 0, seed=5)
    with pytest.raises(rtb200.RtError):
        ctx.render_progressive(0, None, seed=5)


def test_auto_mode_picks_by_scene_and_changes_nothing(ctx, tmp_path):
    """RT_MODE_AUTO / RT_SCHED_AUTO: wavefront FAST for scenes with triangles, megakernel EXACT for
    purely analytic ones (one launch); the image is the explicit modes' image bit for bit."""
    from tests import helpers
    text = helpers.cornell(tmp_path, 40, 32, 3, 5, extra=False)          # planes, boxes, ellipsoids only
    ctx.upload(rtb200.Scene.load_text(text))
    ctx.reset_counters()
    ctx.render_accum(0, 3, 9, T.MODE_AUTO, T.SCHED_AUTO, clear=True)
    auto = ctx.accum()
    assert ctx.counters()["kernel_launches"] == 1                         # the megakernel
    ctx.render_accum(0, 3, 9, T.MODE_FAST, T.SCHED_WAVEFRONT, clear=True)
    assert ctx.accum().tobytes() == auto.tobytes() and auto.sum() > 0
    path, _ = helpers.heightfield(tmp_path, 24, 32)
    ctx.upload(rtb200.Scene.load_gltf(path, 48, 32, 3))
    ctx.reset_counters()
    ctx.render_accum(0, 3, 9, T.MODE_AUTO, T.SCHED_AUTO, clear=True)
    auto = ctx.accum()
    assert ctx.counters()["kernel_launches"] > 4                          # the wavefront pipeline
    ctx.render_accum(0, 3, 9, T.MODE_EXACT, T.SCHED_MEGAKERNEL, clear=True)
    assert ctx.accum().tobytes() == auto.tobytes()


@pytest.mark.parametrize("w,h", [(1, 1), (2, 3), (7, 5), (33, 1)])
def test_tiny_images_and_regions(ctx, tmp_path, w, h):
    """Images far smaller than a warp (the item queue then has fewer items per plane than a warp has
    lanes, and most of the ring is reserve): wavefront == megakernel bit for bit, also for a
    one-pixel region and for more samples than the ring has planes."""
    from tests import helpers
    ctx.upload(rtb200.Scene.load_text(helpers.cornell(tmp_path, w, h, 40, 5)))
    ctx.render_accum(0, 40, 3, T.MODE_FAST, T.SCHED_MEGAKERNEL, clear=True)
    want = ctx.accum()
    ctx.render_accum(0, 40, 3, T.MODE_FAST, T.SCHED_WAVEFRONT, clear=True)
    assert ctx.accum().tobytes() == want.tobytes() and want.sum() > 0
    ctx.render_accum(0, 40, 3, T.MODE_FAST, T.SCHED_WAVEFRONT, clear=True, region=(w - 1, h - 1, 1, 1))
    one = ctx.accum()
    assert one[h - 1, w - 1].tobytes() == want[h - 1, w - 1].tobytes()
    one[h - 1, w - 1] = 0
    assert not one.any()


def test_shared_memory_staging_of_the_top_levels_is_bitwise_the_same(tmp_path, monkeypatch):
    """north_star's "shared-memory staging of the top BVH levels": the trace kernel instantiation that
    copies the first levels of the fp32 hierarchy (breadth-first, 85 nodes) to shared memory is selected
    with RT_WF_SMEM_TOP=1, read at rt_create.  It is off by default (measured slower on B200, DESIGN.md
    §4) — and must give the same image bit for bit, with the same number of segments."""
    path, _ = helpers.heightfield(tmp_path, 60, 90)
    scene = rtb200.Scene.load_gltf(path, 160, 96, 6)
    plain = rtb200.Context(0).upload(scene)
    plain.reset_counters()
    plain.render_accum(0, 6, 0x5EED, T.MODE_FAST, T.SCHED_WAVEFRONT, clear=True)
    want, wc = plain.accum(), plain.counters()
    plain.close()
    monkeypatch.setenv("RT_WF_SMEM_TOP", "1")
    staged = rtb200.Context(0).upload(scene)
    monkeypatch.delenv("RT_WF_SMEM_TOP")
    staged.reset_counters()
    staged.render_accum(0, 6, 0x5EED, T.MODE_FAST, T.SCHED_WAVEFRONT, clear=True)
    got, gc = staged.accum(), staged.counters()
    staged.close()
    assert want.sum() > 0 and got.tobytes() == want.tobytes()
    assert gc["segments"] == wc["segments"] and gc["light_pdf_rays"] == wc["light_pdf_rays"]
