"""Pins the CPU oracle against the reference's own 13 inline tests:
src/aabb.rs:118-151, src/primitives/box.rs:129-165, src/primitives/triangle.rs:98-144,
src/gltf/scene_builder.rs:408-426.  These are the only golden vectors the reference holds."""
import ctypes as C
import math

import numpy as np
import pytest

import rtb200.ctypes_types as T
from rtb200.scene_desc import RawScene


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _aabb(orc, o, d):
    mn, mx = np.array([-1.0, -2.0, -1.0]), np.array([1.0, 2.0, 1.0])
    t = np.zeros(1)
    hit = orc.lib().orc_kat_aabb(_p(mn), _p(mx), _p(np.array(o, float)), _p(np.array(d, float)), _p(t))
    return (t[0] if hit else None)


def _normalize(orc, v):
    out = np.zeros(3)
    orc.lib().orc_kat_normalize(_p(np.array(v, float)), _p(out))
    return out


def test_aabb_a(orc):
    assert _aabb(orc, (0, 0, 2), _normalize(orc, (0, 0, 1))) is None


def test_aabb_b(orc):
    assert _aabb(orc, (0, 0, -2), _normalize(orc, (0, 0, 1))) == 1.0


def test_aabb_c(orc):
    assert _aabb(orc, (2, 0, -2), _normalize(orc, (0, 0, 1))) is None


def test_aabb_d(orc):
    # bit-exact: assert!(aabb.intersects(&ray) == Some(2f64.sqrt()))
    assert _aabb(orc, (-2, 0, -2), _normalize(orc, (1, 0, 1))) == math.sqrt(2.0)


def test_aabb_e(orc):
    assert _aabb(orc, (-1, 0, -2), _normalize(orc, (0, 0, 1))) == 1.0


def _box(orc, o, d):
    t, n, inside = np.zeros(1), np.zeros(3), C.c_int()
    hit = orc.lib().orc_kat_box(_p(np.array([1.0, 2.0, 1.0])), _p(np.array(o, float)), _p(np.array(d, float)),
                                _p(t), _p(n), C.byref(inside))
    return (t[0], tuple(n), bool(inside.value)) if hit else None


def test_box_a(orc):
    assert _box(orc, (0, 0, 2), _normalize(orc, (0, 0, 1))) is None


def test_box_b(orc):
    assert _box(orc, (0, 0, -2), _normalize(orc, (0, 0, 1))) == (1.0, (0.0, 0.0, -1.0), False)


def test_box_c(orc):
    assert _box(orc, (2, 0, -2), _normalize(orc, (0, 0, 1))) is None


def test_box_d(orc):
    # pins the tie rule: equal near-t on x and z -> the later axis wins (box.rs:57-59)
    assert _box(orc, (-2, 0, -2), _normalize(orc, (1, 0, 1))) == (math.sqrt(2.0), (0.0, 0.0, -1.0), False)


def test_box_e(orc):
    assert _box(orc, (-1, 0, -2), _normalize(orc, (0, 0, 1))) == (1.0, (0.0, 0.0, -1.0), False)


def test_triangle_aaa(orc):
    """intersect_lights over a 1-triangle light BVH fires its callback (all-hits path).
    The callback firing is observed as a non-zero Light::pdf contribution."""
    a = np.array([-4.0, -2.0, 10.0])
    ba, ca = np.array([1.0, 6.0, 0.0]), np.array([3.0, 0.0, 0.0])
    rs = RawScene(width=1, height=1, tri_mode=T.TRI_MODEL)
    rs.materials = [{"kind": T.MAT_DIFFUSE, "emission": (1.0, 1.0, 1.0)}]
    rs.tri_pos = np.concatenate([a, a + ba, a + ca]).reshape(1, 9)
    rs.tri_mat = np.zeros(1, np.uint32)
    rs.tri_xform = np.array([[0.0, 0.0, -6.0, 0.0, 0.0, 0.0, 1.0]])
    scene = orc.OracleScene(rs)
    tri, _ = scene.triangle(T.BVH_LIGHT_TRIANGLES, 0)
    wa, wba, wca = tri[0:3], tri[3:6], tri[6:9]
    assert tuple(wa) == (-4.0, -2.0, 4.0) and tuple(wba) == (1.0, 6.0, 0.0) and tuple(wca) == (3.0, 0.0, 0.0)
    assert tri[21] == 0.05555555555555555 * 2.0           # inverted_area of the reference fixture is
    # 1/18 = 1/(2*area); the reference's own constructor gives 1/area (triangle.rs:45-46)
    u, v = 0.6, 0.3
    world = wba * u + wca * v + wa
    pos = np.array([-3.0, 2.0, 4.0])
    d = _normalize(orc, world + pos)
    # the reference shoots `world + pos` from `pos`, which lies inside the triangle itself (t = 0)
    calls = orc.lib().orc_kat_intersect_lights(scene._h, _p(pos), _p(d))
    assert calls == 1


def test_triangle_bbb(orc):
    """ray parallel to the triangle's plane: |det| < 1e-11 -> None (triangle.rs:51)."""
    t = np.zeros(3)
    hit = orc.lib().orc_kat_triangle(_p(np.array([0.0, 0.0, 2.0])), _p(np.array([1.0, 0.0, 2.0])), _p(np.array([0.0, 1.0, 0.0])),
                                     _p(np.array([0.1541891385674881, 0.7047585918803002, 0.5904828162393995])),
                                     _p(np.array([-0.0759650747603601, -0.4459213624433466, 0.8918427248866934])), _p(t))
    assert hit == 0


def test_gltf_cof(orc):
    def rot(axis, deg):
        c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
        if axis == 0:
            return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
        if axis == 1:
            return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    m = rot(0, 10.0) @ rot(1, 20.0) @ rot(2, 30.0) @ np.diag([2.0, 3.0, 4.0])
    cof_cm = np.zeros(9)
    orc.lib().orc_kat_cof(_p(np.ascontiguousarray(m.T).reshape(-1)), _p(cof_cm))     # column-major in/out
    cof = cof_cm.reshape(3, 3).T
    inv_t = np.linalg.inv(m.T)
    for n in ((1.0, 2.0, 3.0), (-1.0, 2.0, 3.0), (-1.0, -2.0, 1.0)):
        n = np.array(n) / np.linalg.norm(n)
        a, b = cof @ n, inv_t @ n
        np.testing.assert_allclose(a / np.linalg.norm(a), b / np.linalg.norm(b), atol=1e-12)
