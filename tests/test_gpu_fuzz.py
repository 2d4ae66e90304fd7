"""Randomised parity (run with -m gpu): FAST mode's fp32 culling must never lose or invent a hit.

The conservative bounds of FAST mode (fast.cuh: outward-rounded boxes, folded slab error constants,
the 8u bound of the ray-relative triangle pre-test) are proofs on paper; this file attacks them with
triangle soups built to be hard for fp32 — coordinates far from the origin, sliver and needle
triangles, many orders of magnitude between the largest and the smallest triangle, coplanar
duplicates, rays that graze edges, start on surfaces, run along axes or have tiny components — and
requires FAST == EXACT == oracle, bit for bit, on ids, t, normals and light PDFs.
"""
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


def _soup(rng, n, flavour):
    """n triangles (n, 9) of one adversarial flavour."""
    if flavour == "far":                    # a small cluster 1e5 units from the origin: |o| >> size
        c = rng.uniform(-1, 1, (n, 1, 3)) * 3.0 + np.array([1.0e5, -7.0e4, 3.0e4])
        return (c + rng.normal(scale=0.05, size=(n, 3, 3))).reshape(n, 9)
    if flavour == "scales":                 # sizes from 1e-4 to 1e2 in one scene
        c = rng.uniform(-50, 50, (n, 1, 3))
        s = 10.0 ** rng.uniform(-4, 2, (n, 1, 1))
        return (c + s * rng.normal(size=(n, 3, 3))).reshape(n, 9)
    if flavour == "slivers":                # needles: two vertices 1e-6 apart, the third far away
        a = rng.uniform(-5, 5, (n, 3))
        b = a + rng.normal(scale=1e-6, size=(n, 3))
        c = a + rng.normal(scale=3.0, size=(n, 3))
        return np.concatenate([a, b, c], 1)
    if flavour == "coplanar":               # overlapping triangles in one plane, many exact duplicates
        base = rng.uniform(-4, 4, (max(1, n // 4), 3, 3))
        base[:, :, 1] = 0.25                # all in the plane y = 0.25
        tri = base[rng.integers(0, len(base), n)]
        return tri.reshape(n, 9)
    if flavour == "grid":                   # axis-aligned unit quads: box faces coincide with triangle edges
        k = int(np.ceil(np.sqrt(n / 2)))
        xs, zs = np.meshgrid(np.arange(k, dtype=float), np.arange(k, dtype=float), indexing="ij")
        a = np.stack([xs, np.zeros_like(xs), zs], -1).reshape(-1, 3)
        t1 = np.concatenate([a, a + [1, 0, 0], a + [0, 0, 1]], 1)
        t2 = np.concatenate([a + [1, 0, 0], a + [1, 0, 1], a + [0, 0, 1]], 1)
        return np.stack([t1, t2], 1).reshape(-1, 9)[:n]
    raise ValueError(flavour)


def _rays(rng, tri, n):
    """Rays aimed AT the geometry in ways that stress the bounds."""
    m = len(tri)
    v = tri.reshape(m, 3, 3)
    pick = rng.integers(0, m, n)
    w = rng.dirichlet([1, 1, 1], n)
    target = np.einsum("nk,nkc->nc", w, v[pick])                     # a point inside a random triangle
    edge = v[pick, 0] + rng.uniform(0, 1, (n, 1)) * (v[pick, 1] - v[pick, 0])
    vert = v[pick, rng.integers(0, 3, n)]
    lo, hi = tri.reshape(-1, 3).min(0), tri.reshape(-1, 3).max(0)
    span = np.maximum(hi - lo, 1e-3)
    o = lo + rng.uniform(-0.5, 1.5, (n, 3)) * span
    kind = rng.integers(0, 6, n)
    aim = np.where((kind == 0)[:, None], target, np.where((kind == 1)[:, None], edge, np.where((kind == 2)[:, None], vert, target)))
    d = aim - o
    # 3: start ON a triangle (self-intersection distance 0 / EPSILON cases), random direction
    on = kind == 3
    o[on] = target[on]
    d[on] = rng.normal(size=(on.sum(), 3))
    # 4: axis-parallel rays through a target (zero direction components -> the slab special cases)
    ax = kind == 4
    axis = rng.integers(0, 3, n)
    dd = np.zeros((n, 3))
    dd[np.arange(n), axis] = rng.choice([-1.0, 1.0], n)
    o[ax] = target[ax] - dd[ax] * span[axis[ax]][:, None] * 2.0
    d[ax] = dd[ax]
    # 5: nearly axis-parallel: one or two components 1e-12 .. 1e-300 of the main one (huge reciprocals)
    tiny = kind == 5
    d[tiny] = d[tiny] * (10.0 ** rng.uniform(-300, -12, (tiny.sum(), 3))) ** (rng.uniform(size=(tiny.sum(), 3)) < 0.5)
    # unnormalised on purpose, over many magnitudes (refracted / camera rays are not unit length)
    d *= 10.0 ** rng.uniform(-3, 3, (n, 1))
    return np.ascontiguousarray(o), np.ascontiguousarray(d), kind


def _assert_equal(a, b, what):
    assert helpers.hits_equal(a, b) == 0, what
    hit = b["kind"] >= 0
    for f in ("t", "geometry_normal", "shading_normal"):
        assert a[f][hit].tobytes() == b[f][hit].tobytes(), (what, f)


@pytest.mark.parametrize("flavour", ["far", "scales", "slivers", "coplanar", "grid"])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fast_equals_exact_equals_oracle_on_adversarial_soups(orc, ctx, flavour, seed):
    rng = np.random.default_rng(1000 * seed + len(flavour))
    n = int(rng.integers(300, 3000))
    tri = _soup(rng, n, flavour)
    raw = RawScene(width=8, height=8)
    raw.materials = [{"kind": T.MAT_DIFFUSE, "color": (0.5, 0.5, 0.5)}, {"kind": T.MAT_DIFFUSE, "emission": (2.0, 2.0, 2.0)}]
    raw.tri_pos = tri
    raw.tri_mat = (rng.uniform(size=len(tri)) < 0.15).astype(np.uint32)          # 15 % of them are lights
    prod, oracle = rtb200.Scene.from_raw(raw), orc.OracleScene(raw)
    helpers.assert_scene_equal(prod, oracle)
    ctx.upload(prod)
    o, d, _ = _rays(rng, tri, 40000)
    exact = ctx.trace_rays(o, d, T.MODE_EXACT)
    ctx.reset_counters()
    fast = ctx.trace_rays(o, d, T.MODE_FAST)
    _assert_equal(fast, exact, "FAST vs EXACT")
    # these inputs are built so that the proof fails often: the exact traversal must have been used
    assert ctx.counters()["exact_fallbacks"] > 0
    sub = slice(0, 6000)
    _assert_equal(exact[sub], oracle.trace_rays(o[sub], d[sub])[0], "EXACT vs oracle")
    assert (exact["kind"] >= 0).mean() > 0.2
    # Light::pdf: all-hits search over the emissive subset, sum formed in the reference's order
    with np.errstate(divide="ignore", invalid="ignore"):           # directions that underflowed to zero are filtered out below
        dn = d / np.linalg.norm(d, axis=1, keepdims=True)
    ok = np.isfinite(dn).all(1)
    pe = ctx.light_pdf(o[ok], dn[ok], T.MODE_EXACT)
    pf = ctx.light_pdf(o[ok], dn[ok], T.MODE_FAST)
    assert pf.tobytes() == pe.tobytes()
    pc, _ = oracle.light_pdf(o[ok][sub], dn[ok][sub])
    assert pe[sub].tobytes() == pc.tobytes()
    assert (pe > 0).mean() > 0.02


@pytest.mark.parametrize("flavour", ["scales", "coplanar"])
def test_renders_agree_on_adversarial_soups(ctx, flavour):
    """Whole paths through such a soup (emissive triangles, depth 6): EXACT megakernel == FAST
    megakernel == FAST wavefront, bit for bit."""
    rng = np.random.default_rng(77)
    tri = _soup(rng, 1500, flavour)
    raw = RawScene(width=48, height=40, samples=3, ray_depth=6)
    lo, hi = tri.reshape(-1, 3).min(0), tri.reshape(-1, 3).max(0)
    raw.cam_position = tuple((lo + hi) / 2 + np.array([0.0, 0.3 * (hi - lo)[1] + 1.0, -0.9 * (hi - lo)[2] - 1.0]))
    raw.bg_color = (0.2, 0.3, 0.4)
    raw.materials = [{"kind": T.MAT_DIFFUSE, "color": (0.7, 0.6, 0.5)}, {"kind": T.MAT_DIFFUSE, "emission": (3.0, 3.0, 3.0)},
                     {"kind": T.MAT_METALLIC, "color": (0.9, 0.9, 0.9)}, {"kind": T.MAT_DIELECTRIC, "ior": 1.5, "color": (1.0, 1.0, 1.0)}]
    raw.tri_pos = tri
    raw.tri_mat = rng.choice([0, 0, 0, 1, 2, 3], len(tri)).astype(np.uint32)
    ctx.upload(rtb200.Scene.from_raw(raw))
    images = []
    for mode, sched in ((T.MODE_EXACT, T.SCHED_MEGAKERNEL), (T.MODE_FAST, T.SCHED_MEGAKERNEL), (T.MODE_FAST, T.SCHED_WAVEFRONT)):
        ctx.render_accum(0, 3, 31, mode, sched, clear=True)
        images.append(ctx.accum())
    assert images[0].tobytes() == images[1].tobytes() == images[2].tobytes()
    assert np.isfinite(images[0]).all() and images[0].sum() > 0


def _random_text_scene(rng, n_shapes, n_tris):
    """A random scene of the TEXT format's kinds: planes, rotated boxes and ellipsoids, and
    model-space triangles with their own position + rotation (src/scene.rs:139-165), a third of the
    primitives emissive so that all three light hierarchies are populated."""
    raw = RawScene(width=40, height=32, samples=2, ray_depth=5)
    raw.tri_mode = T.TRI_MODEL
    raw.cam_position = (0.0, 0.5, -9.0)
    raw.bg_color = (0.1, 0.1, 0.2)
    raw.materials = [{"kind": T.MAT_DIFFUSE, "color": (0.7, 0.7, 0.7)}, {"kind": T.MAT_DIFFUSE, "emission": (4.0, 3.0, 2.0)},
                     {"kind": T.MAT_METALLIC, "color": (0.9, 0.8, 0.7)}, {"kind": T.MAT_DIELECTRIC, "ior": 1.4, "color": (1.0, 1.0, 1.0)}]

    def quat():
        q = rng.normal(size=4)
        return tuple(q / np.linalg.norm(q))

    shapes = [{"type": T.SHAPE_PLANE, "material": 0, "param": (0.0, 1.0, 0.0), "position": (0.0, -3.0, 0.0)},
              {"type": T.SHAPE_PLANE, "material": 0, "param": tuple(rng.normal(size=3)), "position": (0.0, 0.0, 6.0), "rotation": quat()}]
    for _ in range(n_shapes):
        shapes.append({"type": int(rng.choice([T.SHAPE_BOX, T.SHAPE_ELLIPSOID])), "material": int(rng.choice([0, 0, 1, 2, 3])),
                       "param": tuple(10.0 ** rng.uniform(-1.5, 0.3, 3)), "position": tuple(rng.uniform(-3, 3, 3)), "rotation": quat()})
    raw.shapes = shapes
    c = rng.uniform(-3, 3, (n_tris, 1, 3))
    raw.tri_pos = (c + rng.normal(scale=0.6, size=(n_tris, 3, 3))).reshape(n_tris, 9)
    raw.tri_mat = rng.choice([0, 0, 1, 2, 3], n_tris).astype(np.uint32)
    xf = np.zeros((n_tris, 7))
    xf[:, :3] = rng.uniform(-1, 1, (n_tris, 3))
    q = rng.normal(size=(n_tris, 4))
    xf[:, 3:] = q / np.linalg.norm(q, axis=1, keepdims=True)
    raw.tri_xform = xf
    return raw


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_text_format_scenes(orc, ctx, seed):
    """All four primitive kinds with random rotations, all three materials, all three light
    hierarchies: hits and light PDFs bitwise against the oracle in EXACT mode, FAST == EXACT,
    and the three render paths agree bit for bit."""
    rng = np.random.default_rng(seed)
    raw = _random_text_scene(rng, int(rng.integers(3, 40)), int(rng.integers(0, 400)))
    prod, oracle = rtb200.Scene.from_raw(raw), orc.OracleScene(raw)
    helpers.assert_scene_equal(prod, oracle)
    ctx.upload(prod)
    n = 20000
    o = rng.uniform(-5, 5, (n, 3))
    d = rng.normal(size=(n, 3)) * 10.0 ** rng.uniform(-2, 2, (n, 1))
    axis = rng.uniform(size=n) < 0.15                           # zero direction components: the slab special cases
    d[axis] = np.eye(3)[rng.integers(0, 3, axis.sum())] * rng.choice([-1.0, 1.0], (axis.sum(), 1))
    inside = rng.uniform(size=n) < 0.2                          # origins inside boxes / ellipsoids
    pick = rng.integers(2, len(raw.shapes), inside.sum())
    o[inside] = np.array([raw.shapes[i]["position"] for i in pick]) + rng.normal(scale=0.01, size=(inside.sum(), 3))
    exact = ctx.trace_rays(o, d, T.MODE_EXACT)
    _assert_equal(exact, oracle.trace_rays(o, d)[0], "EXACT vs oracle")
    _assert_equal(ctx.trace_rays(o, d, T.MODE_FAST), exact, "FAST vs EXACT")
    assert len(set(exact["kind"].tolist())) >= 4                # misses, planes, boxes or ellipsoids, triangles all occur
    dn = d / np.linalg.norm(d, axis=1, keepdims=True)
    pe = ctx.light_pdf(o, dn, T.MODE_EXACT)
    assert pe.tobytes() == oracle.light_pdf(o, dn)[0].tobytes()
    assert ctx.light_pdf(o, dn, T.MODE_FAST).tobytes() == pe.tobytes()
    images = []
    for mode, sched in ((T.MODE_EXACT, T.SCHED_MEGAKERNEL), (T.MODE_FAST, T.SCHED_MEGAKERNEL), (T.MODE_FAST, T.SCHED_WAVEFRONT)):
        ctx.render_accum(0, 2, 5, mode, sched, clear=True)
        images.append(ctx.accum())
    assert images[0].tobytes() == images[1].tobytes() == images[2].tobytes()
    acc, _ = oracle.render_accum(0, 2, 5)
    np.testing.assert_allclose(images[0], acc.reshape(images[0].shape), rtol=1e-9, atol=1e-12)
