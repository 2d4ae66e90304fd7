"""Oracle-side scene loaders — TEST INFRASTRUCTURE ONLY.

Python/numpy restatement of the reference's two ingestion paths, producing the pre-BVH RawScene:
  * text:  src/scene_parser.rs:5-85 + Metadata::new / CameraParams::new (src/scene.rs:92-106, :167-177)
  * glTF:  src/gltf/parser.rs (serde field names + defaults) + src/gltf/scene_builder.rs:9-388
Arithmetic follows cgmath 0.18's operation order (see oracle_math.hpp); Python floats and numpy
element-wise ops are IEEE f64 without contraction, so results are bit-comparable with the C++
host loader of the product, which restates the same code independently.
"""
import json
import math

import numpy as np

import rtb200.ctypes_types as T
from rtb200.scene_desc import RawScene


# ---- cgmath helpers on python floats -------------------------------------------------------------
def _normalize(v):
    m = math.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])
    r = 1.0 / m
    return (v[0] * r, v[1] * r, v[2] * r)


def _mat4_mul(a, b):
    """cgmath Matrix4 * Matrix4; matrices are lists of 4 columns of 4."""
    out = []
    for j in range(4):
        col = []
        for r in range(4):
            col.append(((a[0][r] * b[j][0] + a[1][r] * b[j][1]) + a[2][r] * b[j][2]) + a[3][r] * b[j][3])
        out.append(col)
    return out


def _mat4_identity():
    return [[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]]


def _mat4_from_quat(w, x, y, z):
    """impl From<Quaternion> for Matrix4 (cgmath 0.18 quaternion.rs)."""
    x2, y2, z2 = x + x, y + y, z + z
    xx2, xy2, xz2 = x2 * x, x2 * y, x2 * z
    yy2, yz2, zz2 = y2 * y, y2 * z, z2 * z
    sy2, sz2, sx2 = y2 * w, z2 * w, x2 * w
    return [[1.0 - yy2 - zz2, xy2 + sz2, xz2 - sy2, 0.0],
            [xy2 - sz2, 1.0 - xx2 - zz2, yz2 + sx2, 0.0],
            [xz2 + sy2, yz2 - sx2, 1.0 - xx2 - yy2, 0.0],
            [0.0, 0.0, 0.0, 1.0]]


# ---- text format -----------------------------------------------------------------------------------
def parse_text_scene(text):
    """scene_parser::parse_scene + Scene::new's defaulting (src/scene.rs:180-191)."""
    prims = []          # dicts: type, param / tri, props
    cam = {}
    scene = {}

    def f(parts, n):
        vals = [float(parts.pop(0)) for _ in range(n)]
        return vals

    for line in text.splitlines():
        parts = line.split()
        if not parts:
            continue
        key = parts.pop(0)
        if key == "NEW_PRIMITIVE":
            prims.append({"type": None, "props": {}})
        elif key in ("BOX", "PLANE", "ELLIPSOID"):
            prims[-1]["type"] = key
            prims[-1]["param"] = f(parts, 3)
        elif key == "TRIANGLE":
            prims[-1]["type"] = key
            prims[-1]["tri"] = f(parts, 9)
        elif key == "POSITION":
            prims[-1]["props"]["position"] = f(parts, 3)
        elif key == "ROTATION":
            prims[-1]["props"]["rotation"] = f(parts, 4)     # x y z w, scene_parser.rs:51-57
        elif key == "COLOR":
            prims[-1]["props"]["color"] = f(parts, 3)
        elif key == "EMISSION":
            prims[-1]["props"]["emission"] = f(parts, 3)
        elif key == "METALLIC":
            prims[-1]["props"]["material"] = T.MAT_METALLIC
        elif key == "DIELECTRIC":
            prims[-1]["props"]["material"] = T.MAT_DIELECTRIC
        elif key == "IOR":
            prims[-1]["props"]["ior"] = f(parts, 1)[0]
        elif key == "CAMERA_POSITION":
            cam["position"] = f(parts, 3)
        elif key == "CAMERA_RIGHT":
            cam["right"] = f(parts, 3)
        elif key == "CAMERA_UP":
            cam["up"] = f(parts, 3)
        elif key == "CAMERA_FORWARD":
            cam["forward"] = f(parts, 3)
        elif key == "CAMERA_FOV_X":
            cam["fov_x"] = f(parts, 1)[0]
        elif key == "DIMENSIONS":
            scene["dimensions"] = (int(parts[0]), int(parts[1]))
        elif key == "RAY_DEPTH":
            scene["ray_depth"] = int(parts[0])
            if not 0 <= scene["ray_depth"] <= 255:
                raise ValueError("RAY_DEPTH is a u8")
        elif key == "BG_COLOR":
            scene["bg_color"] = f(parts, 3)
        elif key == "SAMPLES":
            scene["samples"] = int(parts[0])
        # anything else: skipped (scene_parser.rs:37-38)

    if "dimensions" not in scene:
        raise ValueError("DIMENSIONS missing (the reference unwraps it, src/scene.rs:188)")
    rs = RawScene(width=scene["dimensions"][0], height=scene["dimensions"][1],
                  samples=scene.get("samples", 64), ray_depth=scene.get("ray_depth", 16),
                  bg_color=tuple(scene.get("bg_color", (0.0, 0.0, 0.0))),
                  cam_position=tuple(cam.get("position", (0.0, 0.0, 0.0))),
                  cam_right=_normalize(cam.get("right", (1.0, 0.0, 0.0))),
                  cam_up=_normalize(cam.get("up", (0.0, 1.0, 0.0))),
                  cam_forward=_normalize(cam.get("forward", (0.0, 0.0, 1.0))),
                  fov_kind=T.FOV_X, fov=cam.get("fov_x", math.pi / 2.0), tri_mode=T.TRI_MODEL)
    tri_pos, tri_mat, tri_xform = [], [], []
    for p in prims:
        if p["type"] is None:
            raise ValueError("primitive without a type (prim_type.unwrap(), src/scene.rs:201)")
        pr = p["props"]
        kind = pr.get("material", T.MAT_DIFFUSE)
        if kind == T.MAT_DIELECTRIC and "ior" not in pr:
            raise ValueError("DIELECTRIC without IOR (ior.unwrap(), src/scene.rs:97)")
        rs.materials.append({"kind": kind, "ior": pr.get("ior", 1.0),
                             "color": pr.get("color", (0.0, 0.0, 0.0)),
                             "emission": pr.get("emission", (0.0, 0.0, 0.0))})
        mat = len(rs.materials) - 1
        pos = pr.get("position", (0.0, 0.0, 0.0))
        rot = pr.get("rotation", (0.0, 0.0, 0.0, 1.0))
        if p["type"] == "TRIANGLE":
            tri_pos.append(p["tri"])
            tri_mat.append(mat)
            tri_xform.append(list(pos) + list(rot))
        else:
            ty = {"PLANE": T.SHAPE_PLANE, "BOX": T.SHAPE_BOX, "ELLIPSOID": T.SHAPE_ELLIPSOID}[p["type"]]
            rs.shapes.append({"type": ty, "material": mat, "param": p["param"], "position": pos, "rotation": rot})
    if tri_pos:
        rs.tri_pos = np.array(tri_pos, np.float64).reshape(-1, 9)
        rs.tri_mat = np.array(tri_mat, np.uint32)
        rs.tri_xform = np.array(tri_xform, np.float64).reshape(-1, 7)
    return rs


# ---- glTF --------------------------------------------------------------------------------------------
_UNSIGNED_SHORT, _UNSIGNED_INT, _FLOAT = 5123, 5125, 5126


def _extract_trs(node):
    """extract_trs, gltf/scene_builder.rs:108-123."""
    if node.get("matrix") is not None:
        m = [float(v) for v in node["matrix"]]
        assert len(m) == 16
        return [m[0:4], m[4:8], m[8:12], m[12:16]]
    t = [float(v) for v in node.get("translation") or (0.0, 0.0, 0.0)]
    r = [float(v) for v in node.get("rotation") or (0.0, 0.0, 0.0, 1.0)]
    s = [float(v) for v in node.get("scale") or (1.0, 1.0, 1.0)]
    assert len(t) == 3 and len(r) == 4 and len(s) == 3
    mt = _mat4_identity()
    mt[3] = [t[0], t[1], t[2], 1.0]                          # from_translation
    mr = _mat4_from_quat(r[3], r[0], r[1], r[2])             # extract_quat: Quat::new(w, x, y, z)
    ms = [[s[0], 0.0, 0.0, 0.0], [0.0, s[1], 0.0, 0.0], [0.0, 0.0, s[2], 0.0], [0.0, 0.0, 0.0, 1.0]]
    return _mat4_mul(_mat4_mul(mt, mr), ms)


def _accessor_bytes(model, buffers, accessor):
    view = model["bufferViews"][accessor["bufferView"]]
    buf = model["buffers"][view["buffer"]]
    data = buffers[buf["uri"]]
    offset = view.get("byteOffset", 0) + accessor.get("byteOffset", 0)
    return data, offset, view.get("byteStride")


def _read_vec3(model, buffers, accessor):
    assert accessor["componentType"] == _FLOAT and accessor["type"] == "VEC3"
    if accessor.get("bufferView") is None:
        return np.zeros((0, 3))
    data, offset, stride = _accessor_bytes(model, buffers, accessor)
    stride = stride or 12
    n = accessor["count"]
    raw = np.ndarray((n, 3), dtype="<f4", buffer=data, offset=offset, strides=(stride, 4))
    return raw.astype(np.float64)                            # `as Float`


def _read_indices(model, buffers, accessor):
    """read_indices, :237-267."""
    if accessor.get("bufferView") is None:
        return np.zeros(0, np.uint32)
    assert accessor["componentType"] in (_UNSIGNED_SHORT, _UNSIGNED_INT) and accessor["type"] == "SCALAR"
    data, offset, stride = _accessor_bytes(model, buffers, accessor)
    dt, size = ("<u2", 2) if accessor["componentType"] == _UNSIGNED_SHORT else ("<u4", 4)
    stride = stride or size
    return np.ndarray((accessor["count"],), dtype=dt, buffer=data, offset=offset, strides=(stride,)).astype(np.uint32)


def _cof(c):
    """cof, :367-388; c = 3 columns of 3; returns 3 columns of 3."""
    def others(i):
        return {0: (1, 2), 1: (0, 2), 2: (0, 1)}[i]
    out = [[0.0] * 3 for _ in range(3)]
    for col in range(3):
        for row in range(3):
            lc, rc = others(col)
            tr, br = others(row)
            d = c[lc][tr] * c[rc][br] - c[rc][tr] * c[lc][br]
            out[col][row] = -d if (col + row) & 1 else d
    return out


def load_gltf(path, width, height, samples):
    """gltf::parse + gltf::build_scene (src/main.rs:47-66)."""
    prefix = path[: path.rfind("/") + 1] if "/" in path else ""          # bin_file_prefix, main.rs:77-83
    with open(path, "r") as fh:
        model = json.load(fh)
    for key in ("scenes", "nodes", "buffers", "bufferViews", "cameras", "materials", "meshes", "accessors"):
        model.setdefault(key, [])
    nodes = [{"trs": _extract_trs(n), "children": n.get("children", []), "camera": n.get("camera"),
              "mesh": n.get("mesh")} for n in model["nodes"]]

    def propagate(index, parent):                            # propagate_trs_nodes, :163-169
        nodes[index]["trs"] = _mat4_mul(parent, nodes[index]["trs"])
        for ch in nodes[index]["children"]:
            propagate(ch, nodes[index]["trs"])

    for sc in model["scenes"]:                               # every scene's roots, :155-161
        for n in sc["nodes"]:
            propagate(n, _mat4_identity())

    buffers = {}
    for b in model["buffers"]:                               # load_buffers, :80-87
        with open(prefix + b["uri"], "rb") as fh:
            buffers[b["uri"]] = fh.read()

    tri_pos, tri_nrm, tri_mat, materials = [], [], [], []

    def make_metadata(mat):                                  # :227-235
        pbr = mat.get("pbrMetallicRoughness", {})
        color = [float(v) for v in pbr.get("baseColorFactor") or (1.0, 1.0, 1.0, 1.0)]
        assert len(color) == 4
        emission = [float(v) for v in mat.get("emissiveFactor") or (0.0, 0.0, 0.0)]
        strength = float(mat.get("extensions", {}).get("KHR_materials_emissive_strength", {}).get("emissive_strength", 1.0))
        metallic = float(pbr.get("metallicFactor", 1.0))
        if color[3] < 1.0:
            kind, ior = T.MAT_DIELECTRIC, 1.5
        elif metallic > 0.0:
            kind, ior = T.MAT_METALLIC, 1.0
        else:
            kind, ior = T.MAT_DIFFUSE, 1.0
        return {"kind": kind, "ior": ior, "color": color[:3], "emission": [e * strength for e in emission]}

    def convert_primitive(prim, trs):                        # :209-225
        assert prim.get("mode", 4) == 4, "supported only triangles for primitive.mode"
        attrs = prim["attributes"]
        assert attrs.get("NORMAL") is not None, "empty normals"
        v = _read_vec3(model, buffers, model["accessors"][attrs["POSITION"]])
        c = [np.array(col[:3]) for col in trs]
        w = ((trs[0][3] * v[:, 0] + trs[1][3] * v[:, 1]) + trs[2][3] * v[:, 2]) + trs[3][3] * 1.0
        assert np.all(w == 1.0)
        pos = ((c[0][None, :] * v[:, 0:1] + c[1][None, :] * v[:, 1:2]) + c[2][None, :] * v[:, 2:3]) + c[3][None, :] * 1.0
        n = _read_vec3(model, buffers, model["accessors"][attrs["NORMAL"]])
        rs = [np.array(col) for col in _cof([trs[0][:3], trs[1][:3], trs[2][:3]])]
        nn = (rs[0][None, :] * n[:, 0:1] + rs[1][None, :] * n[:, 1:2]) + rs[2][None, :] * n[:, 2:3]
        mag = np.sqrt((nn[:, 0] * nn[:, 0] + nn[:, 1] * nn[:, 1]) + nn[:, 2] * nn[:, 2])
        with np.errstate(divide="ignore", invalid="ignore"):
            nn = nn * (1.0 / mag)[:, None]
        assert len(pos) == len(nn)
        if prim.get("indices") is not None:
            idx = _read_indices(model, buffers, model["accessors"][prim["indices"]])
            assert len(idx) % 3 == 0
        else:
            assert len(pos) % 3 == 0
            idx = np.arange(len(pos), dtype=np.uint32)
        idx = idx.astype(np.int64).reshape(-1, 3)
        tri_pos.append(pos[idx].reshape(-1, 9))
        tri_nrm.append(nn[idx].reshape(-1, 9))
        mat = model["materials"][prim["material"]] if prim.get("material") is not None else {}
        materials.append(make_metadata(mat))
        tri_mat.append(np.full(len(idx), len(materials) - 1, np.uint32))

    def convert_node(index):                                 # :190-200 (mesh first, then children)
        node = nodes[index]
        if node["mesh"] is not None:
            for prim in model["meshes"][node["mesh"]]["primitives"]:
                convert_primitive(prim, node["trs"])
        for ch in node["children"]:
            convert_node(ch)

    for root in model["scenes"][model.get("scene", 0)]["nodes"]:          # :183-185
        convert_node(root)

    cams = model["cameras"]                                  # extract_camera_params, :57-78
    assert len(cams) == 1 and cams[0]["type"] == "perspective" and cams[0].get("perspective") is not None, \
        "Supported only single perspective camera"
    cam_nodes = [n for n in nodes if n["camera"] is not None]
    assert len(cam_nodes) == 1, "You must specify only one a node with the camera"
    trs = cam_nodes[0]["trs"]
    rs = RawScene(width=width, height=height, samples=samples, ray_depth=8, bg_color=(0.0, 0.0, 0.0),
                  cam_position=tuple(trs[3][:3]), cam_right=tuple(trs[0][:3]), cam_up=tuple(trs[1][:3]),
                  cam_forward=tuple(-v for v in trs[2][:3]), fov_kind=T.FOV_Y,
                  fov=float(cams[0]["perspective"]["yfov"]), tri_mode=T.TRI_BAKED)
    rs.materials = materials
    if tri_pos:
        rs.tri_pos = np.concatenate(tri_pos)
        rs.tri_nrm = np.concatenate(tri_nrm)
        rs.tri_mat = np.concatenate(tri_mat)
    return rs
