"""Shared scene fixtures for the tests (small versions of BASELINE.json's configs)."""
import os

import numpy as np

import rtb200
import rtb200.ctypes_types as T
from rtb200 import scenes


def write_text(tmp_path, text, name="scene.txt"):
    p = os.path.join(str(tmp_path), name)
    with open(p, "w") as fh:
        fh.write(text)
    return p


def cornell(tmp_path, w=64, h=64, spp=4, depth=6, extra=True):
    return write_text(tmp_path, scenes.cornell_text(w, h, spp, depth, scenes.CORNELL_EXTRA if extra else ""))


def heightfield(tmp_path, nx=40, nz=80, lights=True, spheres=True, name="hf.gltf", **kw):
    p = os.path.join(str(tmp_path), name)
    n = scenes.write_heightfield_gltf(p, nx=nx, nz=nz, lights=lights, spheres=spheres, sphere_subdiv=kw.pop("sphere_subdiv", 2), **kw)
    return p, n


def assert_bvh_equal(a, b, prim_order=True):
    for k in ("left", "right", "first", "count") + (("prim_order",) if prim_order else ()):
        assert np.array_equal(a[k], b[k]), k
    assert a["bounds"].tobytes() == b["bounds"].tobytes(), "node bounds differ bitwise"


def assert_scene_equal(prod, oracle, prim_order=True):
    """Product host scene (rtb200.Scene) vs oracle scene: bit-exact hierarchies and primitives.
    prim_order=False: a scene imported with its hierarchies already built does not know the order
    its primitives had before the build."""
    for which in range(6):
        pa, ob = prod.bvh(which), oracle.bvh(which)
        assert len(pa["left"]) == len(ob["left"]), (which, len(pa["left"]), len(ob["left"]))
        assert_bvh_equal(pa, ob, prim_order)
    for which in (T.BVH_TRIANGLES, T.BVH_LIGHT_TRIANGLES):
        n = int(prod.info.bvh_prims[which])
        step = max(1, n // 2000)
        for i in list(range(0, n, step)) + ([n - 1] if n else []):
            ta, ma = prod.triangle(which, i)
            tb, mb = oracle.triangle(which, i)
            assert ta.tobytes() == tb.tobytes() and ma == mb, (which, i)
    for which in (-1, T.BVH_BOXES, T.BVH_ELLIPSOIDS, T.BVH_LIGHT_BOXES, T.BVH_LIGHT_ELLIPSOIDS):
        n = int(prod.info.n_planes) if which < 0 else int(prod.info.bvh_prims[which])
        for i in range(n):
            sa, ba = prod.shape(which, i)
            sb, bb = oracle.shape(which, i)
            assert bytes(sa) == bytes(sb), (which, i)
            if which >= 0:
                assert ba.tobytes() == bb.tobytes(), (which, i)
    th = oracle.tan_half_fov()
    assert th[0] == prod.info.tan_half_fov_x and th[1] == prod.info.tan_half_fov_y


def hits_equal(a, b, rtol=0.0):
    """Bit-exact (rtol=0) or tolerance comparison of two rt_hit arrays; returns mismatch count of ids."""
    ids = (a["kind"] != b["kind"]) | (a["prim"] != b["prim"]) | (a["inside"] != b["inside"]) | (a["material"] != b["material"])
    return int(ids.sum())
