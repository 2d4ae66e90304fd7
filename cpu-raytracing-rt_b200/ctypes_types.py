"""ctypes mirrors of the structs in include/rt_b200.h (data layout only, no behaviour)."""
import ctypes as C

c_double_p = C.POINTER(C.c_double)
c_u32_p = C.POINTER(C.c_uint32)


class Material(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("ior", C.c_double),
                ("color", C.c_double * 3), ("emission", C.c_double * 3)]


class Shape(C.Structure):
    _fields_ = [("type", C.c_int32), ("material", C.c_uint32), ("param", C.c_double * 3),
                ("position", C.c_double * 3), ("rotation", C.c_double * 4)]


class RawScene(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("samples", C.c_uint32),
                ("ray_depth", C.c_uint32), ("bg_color", C.c_double * 3),
                ("cam_position", C.c_double * 3), ("cam_right", C.c_double * 3),
                ("cam_up", C.c_double * 3), ("cam_forward", C.c_double * 3),
                ("fov_kind", C.c_int32), ("tri_mode", C.c_int32), ("fov", C.c_double),
                ("n_materials", C.c_uint32), ("n_shapes", C.c_uint32),
                ("materials", C.POINTER(Material)), ("shapes", C.POINTER(Shape)),
                ("n_triangles", C.c_uint64), ("tri_pos", c_double_p), ("tri_nrm", c_double_p),
                ("tri_mat", c_u32_p), ("tri_xform", c_double_p)]


class SceneInfo(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("samples", C.c_uint32),
                ("ray_depth", C.c_uint32), ("bg_color", C.c_double * 3),
                ("cam_position", C.c_double * 3), ("cam_right", C.c_double * 3),
                ("cam_up", C.c_double * 3), ("cam_forward", C.c_double * 3),
                ("tan_half_fov_x", C.c_double), ("tan_half_fov_y", C.c_double),
                ("n_planes", C.c_uint32), ("n_materials", C.c_uint32),
                ("bvh_nodes", C.c_uint64 * 6), ("bvh_prims", C.c_uint64 * 6),
                ("bvh_depth", C.c_uint32 * 6)]


class BuiltBvh(C.Structure):
    _fields_ = [("n_nodes", C.c_uint64), ("n_prims", C.c_uint64), ("bounds", c_double_p), ("left", c_u32_p),
                ("right", c_u32_p), ("first", c_u32_p), ("count", c_u32_p), ("shapes", C.POINTER(Shape)),
                ("shape_aabb", c_double_p), ("tri28", c_double_p), ("tri_material", c_u32_p)]


class BuiltScene(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("samples", C.c_uint32), ("ray_depth", C.c_uint32),
                ("bg_color", C.c_double * 3), ("cam_position", C.c_double * 3), ("cam_right", C.c_double * 3),
                ("cam_up", C.c_double * 3), ("cam_forward", C.c_double * 3),
                ("tan_half_fov_x", C.c_double), ("tan_half_fov_y", C.c_double),
                ("n_materials", C.c_uint32), ("n_planes", C.c_uint32),
                ("materials", C.POINTER(Material)), ("planes", C.POINTER(Shape)), ("bvh", BuiltBvh * 6)]


class Hit(C.Structure):
    _fields_ = [("t", C.c_double), ("geometry_normal", C.c_double * 3),
                ("shading_normal", C.c_double * 3), ("kind", C.c_int32), ("prim", C.c_uint32),
                ("inside", C.c_int32), ("material", C.c_uint32)]


class RenderParams(C.Structure):
    _fields_ = [("sample_begin", C.c_uint32), ("sample_count", C.c_uint32), ("seed", C.c_uint64),
                ("mode", C.c_int32), ("sched", C.c_int32), ("clear", C.c_int32),
                ("_pad", C.c_int32), ("region", C.c_uint32 * 4)]


ProgressFn = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_double))


class Counters(C.Structure):
    _fields_ = [("segments", C.c_uint64), ("light_pdf_rays", C.c_uint64), ("paths", C.c_uint64),
                ("aabb_tests", C.c_uint64), ("prim_tests", C.c_uint64),
                ("light_aabb_tests", C.c_uint64), ("light_prim_tests", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("exact_fallbacks", C.c_uint64)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


# numpy structured dtype with the same layout as Hit (sizeof = 72)
import numpy as _np

HIT_DTYPE = _np.dtype([("t", "<f8"), ("geometry_normal", "<f8", (3,)), ("shading_normal", "<f8", (3,)),
                       ("kind", "<i4"), ("prim", "<u4"), ("inside", "<i4"), ("material", "<u4")])
assert HIT_DTYPE.itemsize == C.sizeof(Hit) == 72

MAT_DIFFUSE, MAT_METALLIC, MAT_DIELECTRIC = 0, 1, 2
SHAPE_PLANE, SHAPE_BOX, SHAPE_ELLIPSOID = 0, 1, 2
FOV_X, FOV_Y = 0, 1
TRI_BAKED, TRI_MODEL = 0, 1
BVH_BOXES, BVH_ELLIPSOIDS, BVH_TRIANGLES, BVH_LIGHT_BOXES, BVH_LIGHT_ELLIPSOIDS, BVH_LIGHT_TRIANGLES = range(6)
NO_CHILD = 0xFFFFFFFF
MODE_EXACT, MODE_FAST, MODE_AUTO = 0, 1, 2
SCHED_MEGAKERNEL, SCHED_WAVEFRONT, SCHED_AUTO = 0, 1, 2
