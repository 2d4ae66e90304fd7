"""Host-side Python mirror of `rt_raw_scene` (include/rt_b200.h): the pre-BVH scene that the
reference holds as `parsed_scene::Scene` (src/parsed_scene.rs:43-51) or as the glTF triangle
lists (src/gltf/scene_builder.rs:177).  Pure data + marshalling to ctypes; no arithmetic."""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import ctypes_types as T


@dataclass
class RawScene:
    width: int
    height: int
    samples: int = 64                 # Scene::new default, src/scene.rs:187
    ray_depth: int = 16               # src/scene.rs:185
    bg_color: tuple = (0.0, 0.0, 0.0)
    cam_position: tuple = (0.0, 0.0, 0.0)
    cam_right: tuple = (1.0, 0.0, 0.0)
    cam_up: tuple = (0.0, 1.0, 0.0)
    cam_forward: tuple = (0.0, 0.0, 1.0)
    fov_kind: int = T.FOV_X
    fov: float = float(np.pi / 2.0)   # src/scene.rs:174
    tri_mode: int = T.TRI_BAKED
    materials: list = field(default_factory=list)   # dicts: kind, ior, color, emission
    shapes: list = field(default_factory=list)      # dicts: type, material, param, position, rotation(xyzw)
    tri_pos: np.ndarray = None        # (n, 9) f64
    tri_nrm: np.ndarray = None        # (n, 9) f64 or None
    tri_mat: np.ndarray = None        # (n,) u32
    tri_xform: np.ndarray = None      # (n, 7) f64 or None

    def n_triangles(self):
        return 0 if self.tri_pos is None else int(self.tri_pos.shape[0])

    def as_ctypes(self):
        """Returns (rt_raw_scene, keepalive) — keepalive owns every buffer the struct points to."""
        keep = []
        raw = T.RawScene()
        raw.width, raw.height, raw.samples, raw.ray_depth = self.width, self.height, self.samples, self.ray_depth
        raw.bg_color[:] = [float(v) for v in self.bg_color]
        raw.cam_position[:] = [float(v) for v in self.cam_position]
        raw.cam_right[:] = [float(v) for v in self.cam_right]
        raw.cam_up[:] = [float(v) for v in self.cam_up]
        raw.cam_forward[:] = [float(v) for v in self.cam_forward]
        raw.fov_kind, raw.tri_mode, raw.fov = self.fov_kind, self.tri_mode, float(self.fov)
        mats = (T.Material * max(1, len(self.materials)))()
        for i, m in enumerate(self.materials):
            mats[i].kind = m.get("kind", T.MAT_DIFFUSE)
            mats[i].ior = float(m.get("ior", 1.0))
            mats[i].color[:] = [float(v) for v in m.get("color", (0.0, 0.0, 0.0))]
            mats[i].emission[:] = [float(v) for v in m.get("emission", (0.0, 0.0, 0.0))]
        shapes = (T.Shape * max(1, len(self.shapes)))()
        for i, s in enumerate(self.shapes):
            shapes[i].type = s["type"]
            shapes[i].material = s["material"]
            shapes[i].param[:] = [float(v) for v in s["param"]]
            shapes[i].position[:] = [float(v) for v in s.get("position", (0.0, 0.0, 0.0))]
            shapes[i].rotation[:] = [float(v) for v in s.get("rotation", (0.0, 0.0, 0.0, 1.0))]
        keep += [mats, shapes]
        raw.n_materials, raw.n_shapes = len(self.materials), len(self.shapes)
        raw.materials, raw.shapes = mats, shapes
        n = self.n_triangles()
        raw.n_triangles = n

        def ptr(arr, dtype, ctype, shape):
            if arr is None:
                return None
            a = np.ascontiguousarray(arr, dtype=dtype)
            assert a.shape == shape, (a.shape, shape)
            keep.append(a)
            return a.ctypes.data_as(C.POINTER(ctype))

        if n:
            raw.tri_pos = ptr(self.tri_pos, np.float64, C.c_double, (n, 9))
            raw.tri_nrm = ptr(self.tri_nrm, np.float64, C.c_double, (n, 9))
            tri_mat = self.tri_mat if self.tri_mat is not None else np.zeros(n, np.uint32)
            raw.tri_mat = ptr(tri_mat, np.uint32, C.c_uint32, (n,))
            raw.tri_xform = ptr(self.tri_xform, np.float64, C.c_double, (n, 7))
        return raw, keep
