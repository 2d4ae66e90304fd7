"""ctypes binding of librt_b200.so (the C ABI in include/rt_b200.h) — what a Python caller, the
tests and bench.py use.  Mirrors the reference's own seams: `Scene` is reference `scene::Scene`
(src/scene.rs:82-90) after `Scene::new` / `gltf::build_scene`; `Context.render()` is
`generate_image(&scene)` (src/main.rs:85).  There is no CPU fallback: `Context()` raises when no
CUDA device is present, and importing never silently substitutes another implementation."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import ctypes_types as T

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("RT_B200_LIB") or os.path.join(_HERE, "librt_b200.so")     # env override: tuning builds only
_LIB = None

# every symbol include/rt_b200.h declares (tests check that the library exports each of them)
SYMBOLS = [
    "rt_scene_build", "rt_scene_build_cached", "rt_scene_load_gltf_cached", "rt_scene_from_built", "rt_scene_free", "rt_scene_load_text", "rt_scene_load_gltf", "rt_scene_get_info",
    "rt_scene_get_bvh", "rt_scene_get_triangle", "rt_scene_get_shape", "rt_scene_get_material", "rt_save_ppm",
    "rt_device_count", "rt_create", "rt_destroy", "rt_last_error", "rt_upload_scene", "rt_scene_device_bytes", "rt_trace_rays",
    "rt_trace_primary", "rt_light_pdf", "rt_render_accum", "rt_accum_device_ptr", "rt_get_accum", "rt_set_accum",
    "rt_postprocess", "rt_render", "rt_render_progressive", "rt_sync", "rt_stream", "rt_get_counters", "rt_reset_counters", "rt_debug_rng",
    "rt_shard_sample_range", "rt_comm_unique_id", "rt_comm_init", "rt_comm_info", "rt_reduce_accum",
    "rt_group_create", "rt_group_destroy", "rt_group_size", "rt_group_ctx", "rt_group_last_error", "rt_group_upload_scene",
    "rt_group_render_accum", "rt_group_reduce_accum", "rt_group_render", "rt_group_render_progressive",
]
COMM_ID_BYTES = 128


class RtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rt_b200 error {code}: {msg}")
        self.code = code


def build(force=False):
    """Compiles librt_b200.so + the CLI in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    args = ["make", "-C", os.path.join(_HERE, "csrc"), "-s", "-j8"]
    if force:
        args.append("-B")
    subprocess.check_call(args)
    return _LIB_PATH


def _pick_nccl():
    """The library dlopen()s NCCL by its soname, which finds the SYSTEM copy.  In a Python process that also
    imports PyTorch that is the wrong one: torch links against its own bundled (newer) libnccl.so.2, the
    loader reuses whatever object of that soname is already in the process, and `import torch` then fails
    with an undefined NCCL symbol.  So in Python the library is pointed at the copy torch would load
    (RT_NCCL_LIB, read once at the first multi-GPU call); a C or Rust host keeps the system copy."""
    if os.environ.get("RT_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for root in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(root, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["RT_NCCL_LIB"] = cand
                return
    except (ImportError, ValueError, AttributeError):
        pass


def lib():
    global _LIB
    if _LIB is None:
        _pick_nccl()
        if not os.path.exists(_LIB_PATH):
            raise ImportError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no fallback implementation)")
        L = C.CDLL(_LIB_PATH)
        vp, u32, u64, i32, dbl = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_double
        L.rt_scene_build.argtypes = [C.POINTER(T.RawScene), C.POINTER(vp)]
        L.rt_scene_from_built.argtypes = [C.POINTER(T.BuiltScene), C.POINTER(vp)]
        L.rt_scene_free.argtypes = [vp]
        L.rt_scene_free.restype = None
        L.rt_scene_load_text.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.rt_scene_load_gltf.argtypes = [C.c_char_p, u32, u32, u32, C.POINTER(vp)]
        L.rt_scene_build_cached.argtypes = [C.POINTER(T.RawScene), C.c_char_p, C.POINTER(vp), C.POINTER(i32)]
        L.rt_scene_load_gltf_cached.argtypes = [C.c_char_p, u32, u32, u32, C.c_char_p, C.POINTER(vp), C.POINTER(i32)]
        L.rt_scene_get_info.argtypes = [vp, C.POINTER(T.SceneInfo)]
        L.rt_scene_get_bvh.argtypes = [vp, i32] + [vp] * 6
        L.rt_scene_get_triangle.argtypes = [vp, i32, u64, vp, C.POINTER(u32)]
        L.rt_scene_get_shape.argtypes = [vp, i32, u64, C.POINTER(T.Shape), vp]
        L.rt_scene_get_material.argtypes = [vp, u32, C.POINTER(T.Material)]
        L.rt_save_ppm.argtypes = [C.c_char_p, u32, u32, vp]
        L.rt_device_count.argtypes = [C.POINTER(i32)]
        L.rt_create.argtypes = [i32, C.POINTER(vp)]
        L.rt_destroy.argtypes = [vp]
        L.rt_destroy.restype = None
        L.rt_last_error.argtypes = [vp]
        L.rt_last_error.restype = C.c_char_p
        L.rt_upload_scene.argtypes = [vp, vp]
        L.rt_scene_device_bytes.argtypes = [vp, C.POINTER(u64)]
        L.rt_trace_rays.argtypes = [vp, i32, u64, vp, vp, vp]
        L.rt_trace_primary.argtypes = [vp, i32, dbl, dbl, vp, vp]
        L.rt_light_pdf.argtypes = [vp, i32, u64, vp, vp, vp]
        L.rt_render_accum.argtypes = [vp, C.POINTER(T.RenderParams)]
        L.rt_accum_device_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
        L.rt_get_accum.argtypes = [vp, vp]
        L.rt_set_accum.argtypes = [vp, vp]
        L.rt_postprocess.argtypes = [vp, u32, vp]
        L.rt_render.argtypes = [vp, u64, i32, i32, vp]
        L.rt_render_progressive.argtypes = [vp, u64, i32, i32, u32, T.ProgressFn, vp, vp, C.POINTER(u32)]
        L.rt_sync.argtypes = [vp]
        L.rt_stream.argtypes = [vp, C.POINTER(vp)]
        L.rt_get_counters.argtypes = [vp, C.POINTER(T.Counters)]
        L.rt_reset_counters.argtypes = [vp]
        L.rt_debug_rng.argtypes = [vp, u64, u32, u32, u64, vp, vp]
        L.rt_shard_sample_range.argtypes = [i32, i32, u32, u32, C.POINTER(u32), C.POINTER(u32)]
        L.rt_comm_unique_id.argtypes = [vp]
        L.rt_comm_init.argtypes = [vp, vp, i32, i32]
        L.rt_comm_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
        L.rt_reduce_accum.argtypes = [vp, i32]
        L.rt_group_create.argtypes = [i32, C.POINTER(i32), C.POINTER(vp)]
        L.rt_group_destroy.argtypes = [vp]
        L.rt_group_destroy.restype = None
        L.rt_group_size.argtypes = [vp, C.POINTER(i32)]
        L.rt_group_ctx.argtypes = [vp, i32]
        L.rt_group_ctx.restype = vp
        L.rt_group_last_error.argtypes = [vp]
        L.rt_group_last_error.restype = C.c_char_p
        L.rt_group_upload_scene.argtypes = [vp, vp]
        L.rt_group_render_accum.argtypes = [vp, C.POINTER(T.RenderParams)]
        L.rt_group_reduce_accum.argtypes = [vp]
        L.rt_group_render.argtypes = [vp, u64, i32, i32, vp]
        L.rt_group_render_progressive.argtypes = [vp, u64, i32, i32, u32, T.ProgressFn, vp, vp, C.POINTER(u32)]
        _LIB = L
    return _LIB


def _check(code, ctx=None):
    if code != 0:
        raise RtError(code, (lib().rt_last_error(ctx) or b"").decode("utf-8", "replace"))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.reshape(shape) if shape is not None else a


def device_count():
    n = C.c_int()
    lib().rt_device_count(C.byref(n))
    return n.value


class Scene:
    """Host scene: world-space primitives + the six SAH hierarchies (built on construction)."""

    def __init__(self, handle):
        self._h = handle
        info = T.SceneInfo()
        _check(lib().rt_scene_get_info(self._h, C.byref(info)))
        self.info = info
        self.width, self.height, self.samples, self.ray_depth = info.width, info.height, info.samples, info.ray_depth

    @classmethod
    def from_raw(cls, raw_scene, cache_dir=None):
        """Scene::new (src/scene.rs:180) / build_scene on an in-memory description.

        cache_dir: keep the built hierarchies on disk, keyed by a hash of the geometry; the next
        call with the same mesh reads them back instead of rebuilding (`scene.cache_hit`)."""
        raw, keep = raw_scene.as_ctypes()
        h = C.c_void_p()
        hit = C.c_int32(0)
        if cache_dir is None:
            _check(lib().rt_scene_build(C.byref(raw), C.byref(h)))
        else:
            _check(lib().rt_scene_build_cached(C.byref(raw), os.fsencode(cache_dir), C.byref(h), C.byref(hit)))
        del keep
        sc = cls(h)
        sc.cache_hit = bool(hit.value)
        return sc

    @classmethod
    def load_text(cls, path):
        """scene_parser::parse_scene + Scene::new on a text-format file."""
        h = C.c_void_p()
        _check(lib().rt_scene_load_text(os.fsencode(path), C.byref(h)))
        return cls(h)

    @classmethod
    def load_gltf(cls, path, width, height, samples, cache_dir=None):
        """gltf::parse + gltf::build_scene (src/main.rs:47-66); cache_dir as in from_raw."""
        h = C.c_void_p()
        hit = C.c_int32(0)
        if cache_dir is None:
            _check(lib().rt_scene_load_gltf(os.fsencode(path), width, height, samples, C.byref(h)))
        else:
            _check(lib().rt_scene_load_gltf_cached(os.fsencode(path), width, height, samples, os.fsencode(cache_dir),
                                                   C.byref(h), C.byref(hit)))
        sc = cls(h)
        sc.cache_hit = bool(hit.value)
        return sc

    def export_built(self):
        """This scene as an rt_built_scene (already-built hierarchies as flat arrays) + keepalive."""
        keep = {"bvh": [], "other": []}     # owns every array the struct points to; keep["bvh"][w] = node arrays
        b = T.BuiltScene()
        i = self.info
        b.width, b.height, b.samples, b.ray_depth = i.width, i.height, i.samples, i.ray_depth
        for name in ("bg_color", "cam_position", "cam_right", "cam_up", "cam_forward"):
            getattr(b, name)[:] = list(getattr(i, name))
        b.tan_half_fov_x, b.tan_half_fov_y = i.tan_half_fov_x, i.tan_half_fov_y
        mats = (T.Material * max(1, i.n_materials))()
        for k in range(i.n_materials):
            mats[k] = self.material(k)
        planes = (T.Shape * max(1, i.n_planes))()
        for k in range(i.n_planes):
            planes[k] = self.shape(-1, k)[0]
        keep["other"] += [mats, planes]
        b.n_materials, b.n_planes, b.materials, b.planes = i.n_materials, i.n_planes, mats, planes
        for w in range(6):
            d = self.bvh(w)
            keep["bvh"].append(d)
            bb = b.bvh[w]
            bb.n_nodes, bb.n_prims = len(d["left"]), len(d["prim_order"])
            bb.bounds = d["bounds"].ctypes.data_as(T.c_double_p)
            for name in ("left", "right", "first", "count"):
                setattr(bb, name, d[name].ctypes.data_as(T.c_u32_p))
            if w in (T.BVH_TRIANGLES, T.BVH_LIGHT_TRIANGLES):
                tri = np.zeros((max(1, bb.n_prims), 28))
                mat = np.zeros(max(1, bb.n_prims), np.uint32)
                for k in range(bb.n_prims):
                    tri[k], mat[k] = self.triangle(w, k)
                keep["other"] += [tri, mat]
                bb.tri28, bb.tri_material = tri.ctypes.data_as(T.c_double_p), mat.ctypes.data_as(T.c_u32_p)
            else:
                sh = (T.Shape * max(1, bb.n_prims))()
                boxes = np.zeros((max(1, bb.n_prims), 6))
                for k in range(bb.n_prims):
                    sh[k], boxes[k] = self.shape(w, k)
                keep["other"] += [sh, boxes]
                bb.shapes, bb.shape_aabb = sh, boxes.ctypes.data_as(T.c_double_p)
        return b, keep

    @classmethod
    def from_built(cls, built):
        """A scene whose hierarchies the caller built (rt_scene_from_built)."""
        h = C.c_void_p()
        _check(lib().rt_scene_from_built(C.byref(built), C.byref(h)))
        return cls(h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().rt_scene_free(self._h)
            self._h = None

    def bvh(self, which):
        n, p = int(self.info.bvh_nodes[which]), int(self.info.bvh_prims[which])
        out = dict(bounds=np.zeros((n, 6)), left=np.zeros(n, np.uint32), right=np.zeros(n, np.uint32),
                   first=np.zeros(n, np.uint32), count=np.zeros(n, np.uint32), prim_order=np.zeros(p, np.uint32))
        _check(lib().rt_scene_get_bvh(self._h, which, *[_p(out[k]) for k in ("bounds", "left", "right", "first", "count", "prim_order")]))
        return out

    def triangle(self, which, i):
        o = np.zeros(28)
        m = C.c_uint32()
        _check(lib().rt_scene_get_triangle(self._h, which, i, _p(o), C.byref(m)))
        return o, m.value

    def shape(self, which, i):
        s = T.Shape()
        bb = np.zeros(6)
        _check(lib().rt_scene_get_shape(self._h, which, i, C.byref(s), _p(bb)))
        return s, bb

    def material(self, i):
        m = T.Material()
        _check(lib().rt_scene_get_material(self._h, i, C.byref(m)))
        return m


def shard_sample_range(rank, world, total_samples, first_sample=0):
    """The library's partition of a sample range over `world` devices -> (begin, count)."""
    b, c = C.c_uint32(), C.c_uint32()
    _check(lib().rt_shard_sample_range(rank, world, total_samples, first_sample, C.byref(b), C.byref(c)))
    return b.value, c.value


def comm_unique_id():
    """128 bytes rank 0 creates and the launcher broadcasts (ncclGetUniqueId inside the library)."""
    buf = (C.c_ubyte * COMM_ID_BYTES)()
    _check(lib().rt_comm_unique_id(buf))
    return bytes(buf)


class Context:
    """One CUDA device.  render() == generate_image (src/main.rs:85-114)."""

    def __init__(self, device=0, _borrowed=None):
        if _borrowed is not None:           # a context owned by a Group
            self._h, self._owned = C.c_void_p(_borrowed), False
        else:
            h = C.c_void_p()
            _check(lib().rt_create(device, C.byref(h)))
            self._h, self._owned = h, True
        self.device = device
        self.scene = None

    def close(self):
        if getattr(self, "_h", None):
            if self._owned:
                lib().rt_destroy(self._h)
            self._h = None

    def comm_init(self, unique_id, rank, world):
        """Join the in-library NCCL communicator (collective: every rank calls it)."""
        assert len(unique_id) == COMM_ID_BYTES
        buf = (C.c_ubyte * COMM_ID_BYTES).from_buffer_copy(unique_id)
        _check(lib().rt_comm_init(self._h, buf, rank, world), self._h)
        return self

    def comm_info(self):
        r, w, v = C.c_int(), C.c_int(), C.c_int()
        _check(lib().rt_comm_info(self._h, C.byref(r), C.byref(w), C.byref(v)), self._h)
        return {"rank": r.value, "world": w.value, "nccl_version": v.value}

    def reduce_accum(self, root=0):
        """Sum-reduce of the accumulation buffers onto `root` (ncclReduce on the context's stream)."""
        _check(lib().rt_reduce_accum(self._h, root), self._h)

    __del__ = close

    def upload(self, scene):
        _check(lib().rt_upload_scene(self._h, scene._h), self._h)
        self.scene = scene
        return self

    def trace_rays(self, origins, dirs, mode=T.MODE_EXACT):
        origins, dirs = _f64(origins, (-1, 3)), _f64(dirs, (-1, 3))
        hits = np.zeros(origins.shape[0], T.HIT_DTYPE)
        _check(lib().rt_trace_rays(self._h, mode, origins.shape[0], _p(origins), _p(dirs), _p(hits)), self._h)
        return hits

    def trace_primary(self, jx=0.5, jy=0.5, mode=T.MODE_EXACT, want_rays=False, want_hits=True):
        n = self.scene.width * self.scene.height
        hits = np.zeros(n, T.HIT_DTYPE) if want_hits else None
        rays = np.zeros((n, 6)) if want_rays else None
        _check(lib().rt_trace_primary(self._h, mode, jx, jy, _p(hits), _p(rays)), self._h)
        return hits, rays

    def light_pdf(self, pos, dirs, mode=T.MODE_EXACT):
        pos, dirs = _f64(pos, (-1, 3)), _f64(dirs, (-1, 3))
        out = np.zeros(pos.shape[0])
        _check(lib().rt_light_pdf(self._h, mode, pos.shape[0], _p(pos), _p(dirs), _p(out)), self._h)
        return out

    def render_accum(self, sample_begin, sample_count, seed, mode=T.MODE_FAST, sched=None, clear=False, region=None):
        """accum += sum of `sample_count` samples per pixel (src/main.rs:94-103 without :104).
        region = (x0, y0, w, h): only those pixels (same values as the full render has there)."""
        if sched is None:       # FAST renders with the wavefront scheduler, EXACT is the megakernel parity anchor
            sched = T.SCHED_WAVEFRONT if mode == T.MODE_FAST else T.SCHED_MEGAKERNEL
        p = T.RenderParams(sample_begin, sample_count, seed, mode, sched, 1 if clear else 0, 0)
        if region is not None:
            p.region[:] = [int(v) for v in region]
        _check(lib().rt_render_accum(self._h, C.byref(p)), self._h)

    def accum(self):
        out = np.zeros((self.scene.height, self.scene.width, 3))
        _check(lib().rt_get_accum(self._h, _p(out)), self._h)
        return out

    def set_accum(self, accum):
        accum = _f64(accum, (self.scene.height, self.scene.width, 3))
        _check(lib().rt_set_accum(self._h, _p(accum)), self._h)

    def accum_device_ptr(self):
        p, n = C.c_void_p(), C.c_uint64()
        _check(lib().rt_accum_device_ptr(self._h, C.byref(p), C.byref(n)), self._h)
        return p.value, n.value

    def postprocess(self, total_samples):
        out = np.zeros((self.scene.height, self.scene.width, 3))
        _check(lib().rt_postprocess(self._h, total_samples, _p(out)), self._h)
        return out

    def postprocess_device(self, total_samples):
        """tonemap + gamma on the device only (asynchronous; the image stays in HBM)"""
        _check(lib().rt_postprocess(self._h, total_samples, None), self._h)

    def postprocess_into(self, total_samples, out):
        assert out.dtype == np.float64 and out.flags["C_CONTIGUOUS"] and out.size == self.scene.width * self.scene.height * 3
        _check(lib().rt_postprocess(self._h, total_samples, _p(out)), self._h)
        return out

    def scene_device_bytes(self):
        n = C.c_uint64()
        _check(lib().rt_scene_device_bytes(self._h, C.byref(n)), self._h)
        return n.value

    def render(self, seed=0xC0FFEE, mode=T.MODE_FAST, sched=None):
        if sched is None:
            sched = T.SCHED_WAVEFRONT if mode == T.MODE_FAST else T.SCHED_MEGAKERNEL
        out = np.zeros((self.scene.height, self.scene.width, 3))
        _check(lib().rt_render(self._h, seed, mode, sched, _p(out)), self._h)
        return out

    def render_progressive(self, chunk_samples, callback=None, seed=0xC0FFEE, mode=T.MODE_FAST, sched=None):
        """generate_image in chunks with the image so far after each chunk (rt_render_progressive).
        callback(samples_done, samples_total, rgb) -> truthy stops the render; rgb is a view that
        is only valid during the call.  Returns (rgb, samples_done); run to the end, rgb is bit
        for bit what render() returns."""
        if sched is None:
            sched = T.SCHED_WAVEFRONT if mode == T.MODE_FAST else T.SCHED_MEGAKERNEL
        h, w = self.scene.height, self.scene.width
        out = np.zeros((h, w, 3))
        raised = []

        @T.ProgressFn
        def trampoline(_user, done, total, rgb):
            if callback is None:
                return 0
            try:
                view = np.ctypeslib.as_array(rgb, shape=(h, w, 3))
                return 1 if callback(int(done), int(total), view) else 0
            except BaseException as e:        # an exception must not unwind through the C frames
                raised.append(e)
                return 1

        done = C.c_uint32(0)
        _check(lib().rt_render_progressive(self._h, seed, mode, sched, chunk_samples, trampoline, None, _p(out), C.byref(done)), self._h)
        if raised:
            raise raised[0]
        return out, done.value

    def sync(self):
        _check(lib().rt_sync(self._h), self._h)

    def stream(self):
        s = C.c_void_p()
        _check(lib().rt_stream(self._h, C.byref(s)), self._h)
        return s.value or 0

    def counters(self):
        c = T.Counters()
        _check(lib().rt_get_counters(self._h, C.byref(c)), self._h)
        return c.as_dict()

    def reset_counters(self):
        _check(lib().rt_reset_counters(self._h), self._h)

    def debug_rng(self, seed, pixel, sample, n):
        u = np.zeros(n, np.uint64)
        probe = np.zeros(8)
        _check(lib().rt_debug_rng(self._h, seed, pixel, sample, n, _p(u), _p(probe)), self._h)
        return u, probe


class Group:
    """Several CUDA devices of this process behind one handle (rt_group_*): the scene is replicated,
    the sample range is split, the accumulation buffers are sum-reduced to device 0 over NCCL."""

    def __init__(self, devices=None):
        h = C.c_void_p()
        if devices is None:
            _check(lib().rt_group_create(0, None, C.byref(h)))
        else:
            ids = (C.c_int * len(devices))(*devices)
            _check(lib().rt_group_create(len(devices), ids, C.byref(h)))
        self._h = h
        n = C.c_int()
        lib().rt_group_size(self._h, C.byref(n))
        self.size = n.value
        self.scene = None

    def _check(self, code):
        if code != 0:
            raise RtError(code, (lib().rt_group_last_error(self._h) or b"").decode("utf-8", "replace"))

    def close(self):
        if getattr(self, "_h", None):
            lib().rt_group_destroy(self._h)
            self._h = None

    __del__ = close

    def ctx(self, i):
        """Borrowed Context of device i (0 = root); valid while the group lives."""
        h = lib().rt_group_ctx(self._h, i)
        if not h:
            raise IndexError(i)
        c = Context(_borrowed=h)
        c.scene = self.scene
        return c

    def upload(self, scene):
        self._check(lib().rt_group_upload_scene(self._h, scene._h))
        self.scene = scene
        return self

    def render_accum(self, sample_begin, sample_count, seed, mode=T.MODE_FAST, sched=None, clear=False):
        if sched is None:
            sched = T.SCHED_WAVEFRONT if mode == T.MODE_FAST else T.SCHED_MEGAKERNEL
        p = T.RenderParams(sample_begin, sample_count, seed, mode, sched, 1 if clear else 0, 0)
        self._check(lib().rt_group_render_accum(self._h, C.byref(p)))

    def reduce_accum(self):
        self._check(lib().rt_group_reduce_accum(self._h))

    def render(self, seed=0xC0FFEE, mode=T.MODE_FAST, sched=None):
        if sched is None:
            sched = T.SCHED_WAVEFRONT if mode == T.MODE_FAST else T.SCHED_MEGAKERNEL
        out = np.zeros((self.scene.height, self.scene.width, 3))
        self._check(lib().rt_group_render(self._h, seed, mode, sched, _p(out)))
        return out

    def render_progressive(self, chunk_samples, seed=0xC0FFEE, mode=T.MODE_FAST, sched=None):
        if sched is None:
            sched = T.SCHED_WAVEFRONT if mode == T.MODE_FAST else T.SCHED_MEGAKERNEL
        out = np.zeros((self.scene.height, self.scene.width, 3))
        done = C.c_uint32(0)
        self._check(lib().rt_group_render_progressive(self._h, seed, mode, sched, chunk_samples, C.cast(None, T.ProgressFn), None, _p(out), C.byref(done)))
        return out, done.value


def save_ppm(path, rgb):
    rgb = _f64(rgb)
    h, w = rgb.shape[0], rgb.shape[1]
    _check(lib().rt_save_ppm(os.fsencode(path), w, h, _p(rgb)))
