"""ctypes front-end of liboracle.so — TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  It shares struct LAYOUTS with the product (include/rt_b200.h mirrored in
rtb200.ctypes_types) and nothing else."""
import ctypes as C
import os
import subprocess

import numpy as np

import rtb200.ctypes_types as T

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "oracle_math.hpp", "Makefile")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_scene_build.restype = C.c_void_p
        L.orc_scene_build.argtypes = [C.POINTER(T.RawScene)]
        L.orc_scene_free.argtypes = [C.c_void_p]
        L.orc_bvh_counts.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_bvh_get.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
        L.orc_get_triangle.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint32)]
        L.orc_get_shape.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.POINTER(T.Shape), C.c_void_p]
        L.orc_camera.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_trace_rays.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(T.Counters)]
        L.orc_trace_primary.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(T.Counters)]
        L.orc_light_pdf.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(T.Counters)]
        L.orc_light_sample.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_cosine_sample.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_render_accum.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p, C.POINTER(T.Counters)]
        L.orc_postprocess.argtypes = [C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_to_bytes.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_kat_aabb.argtypes = [C.c_void_p] * 5
        L.orc_kat_box.argtypes = [C.c_void_p] * 5 + [C.POINTER(C.c_int)]
        L.orc_kat_triangle.argtypes = [C.c_void_p] * 6
        L.orc_kat_normalize.argtypes = [C.c_void_p] * 2
        L.orc_kat_cof.argtypes = [C.c_void_p] * 2
        L.orc_kat_intersect_lights.argtypes = [C.c_void_p] * 3
        L.orc_rng_u64.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p]
        L.orc_rng_probe.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_hardware_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class OracleScene:
    """Scene::new / build_scene + the hot path, restated on the CPU."""

    def __init__(self, raw_scene):
        raw, keep = raw_scene.as_ctypes()
        self._h = lib().orc_scene_build(C.byref(raw))
        del keep
        self.width, self.height = raw_scene.width, raw_scene.height
        self.samples, self.ray_depth = raw_scene.samples, raw_scene.ray_depth

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_scene_free(self._h)
            self._h = None

    def bvh(self, which):
        nn, npr = C.c_uint64(), C.c_uint64()
        assert lib().orc_bvh_counts(self._h, which, C.byref(nn), C.byref(npr)) == 0
        n, p = nn.value, npr.value
        out = dict(bounds=np.zeros((n, 6)), left=np.zeros(n, np.uint32), right=np.zeros(n, np.uint32),
                   first=np.zeros(n, np.uint32), count=np.zeros(n, np.uint32), prim_order=np.zeros(p, np.uint32))
        lib().orc_bvh_get(self._h, which, *[_p(out[k]) for k in ("bounds", "left", "right", "first", "count", "prim_order")])
        return out

    def triangle(self, which, i):
        o = np.zeros(28)
        m = C.c_uint32()
        assert lib().orc_get_triangle(self._h, which, i, _p(o), C.byref(m)) == 0
        return o, m.value

    def shape(self, which, i):
        s = T.Shape()
        bb = np.zeros(6)
        assert lib().orc_get_shape(self._h, which, i, C.byref(s), _p(bb)) == 0
        return s, bb

    def tan_half_fov(self):
        o = np.zeros(2)
        lib().orc_camera(self._h, _p(o))
        return o

    def trace_rays(self, origins, dirs, nthreads=0):
        origins, dirs = _f64(origins, (-1, 3)), _f64(dirs, (-1, 3))
        hits = np.zeros(origins.shape[0], T.HIT_DTYPE)
        c = T.Counters()
        lib().orc_trace_rays(self._h, origins.shape[0], _p(origins), _p(dirs), _p(hits), nthreads, C.byref(c))
        return hits, c.as_dict()

    def trace_primary(self, jx=0.5, jy=0.5, nthreads=0, want_rays=True):
        n = self.width * self.height
        hits = np.zeros(n, T.HIT_DTYPE)
        rays = np.zeros((n, 6)) if want_rays else None
        c = T.Counters()
        lib().orc_trace_primary(self._h, jx, jy, _p(hits), _p(rays) if want_rays else None, nthreads, C.byref(c))
        return hits, rays, c.as_dict()

    def light_pdf(self, pos, dirs):
        pos, dirs = _f64(pos, (-1, 3)), _f64(dirs, (-1, 3))
        out = np.zeros(pos.shape[0])
        c = T.Counters()
        lib().orc_light_pdf(self._h, pos.shape[0], _p(pos), _p(dirs), _p(out), C.byref(c))
        return out, c.as_dict()

    def light_sample(self, seed, pos):
        pos = _f64(pos, (-1, 3))
        out = np.zeros_like(pos)
        assert lib().orc_light_sample(self._h, seed, pos.shape[0], _p(pos), _p(out)) == 0
        return out

    def render_accum(self, sample_begin, sample_count, seed, nthreads=0, accum=None):
        if accum is None:
            accum = np.zeros((self.height, self.width, 3))
        accum = _f64(accum)
        c = T.Counters()
        lib().orc_render_accum(self._h, sample_begin, sample_count, seed, nthreads, _p(accum), C.byref(c))
        return accum, c.as_dict()


def postprocess(accum, total_samples):
    accum = _f64(accum)
    out = np.zeros_like(accum)
    lib().orc_postprocess(accum.size // 3, _p(accum), total_samples, _p(out))
    return out


def to_bytes(rgb):
    rgb = _f64(rgb)
    out = np.zeros(rgb.shape, np.uint8)
    lib().orc_to_bytes(rgb.size, _p(rgb), _p(out))
    return out


def cosine_sample(seed, n, normal):
    normal = _f64(normal)
    out = np.zeros((n, 3))
    lib().orc_cosine_sample(seed, n, _p(normal), _p(out))
    return out


def rng_u64(seed, pixel, sample, n):
    out = np.zeros(n, np.uint64)
    lib().orc_rng_u64(seed, pixel, sample, n, _p(out))
    return out


def rng_probe(seed, pixel, sample):
    out = np.zeros(8)
    lib().orc_rng_probe(seed, pixel, sample, _p(out))
    return out


def hardware_threads():
    return lib().orc_hardware_threads()
