"""Synthetic inputs of BASELINE.json's configs, written in the reference's own scene formats so
that they go through the same loaders a user's files would.

  cornell_text()            C1: text format, 512x512, 64 spp, depth 6
  write_heightfield_gltf()  C2/C3/C5: heightfield glTF + external .bin, optional lights / spheres
"""
import json
import os

import numpy as np

CORNELL_TEXT = """\
DIMENSIONS {w} {h}
SAMPLES {spp}
RAY_DEPTH {depth}
BG_COLOR 0 0 0
CAMERA_POSITION 0 0 -3.6
CAMERA_RIGHT 1 0 0
CAMERA_UP 0 1 0
CAMERA_FORWARD 0 0 1
CAMERA_FOV_X 0.69
NEW_PRIMITIVE
PLANE 0 1 0
POSITION 0 -1 0
COLOR 0.73 0.73 0.73
NEW_PRIMITIVE
PLANE 0 -1 0
POSITION 0 1 0
COLOR 0.73 0.73 0.73
NEW_PRIMITIVE
PLANE 0 0 -1
POSITION 0 0 1
COLOR 0.73 0.73 0.73
NEW_PRIMITIVE
PLANE 0 0 1
POSITION 0 0 -4
COLOR 0.73 0.73 0.73
NEW_PRIMITIVE
PLANE 1 0 0
POSITION -1 0 0
COLOR 0.65 0.05 0.05
NEW_PRIMITIVE
PLANE -1 0 0
POSITION 1 0 0
COLOR 0.12 0.45 0.15
NEW_PRIMITIVE
BOX 0.35 0.01 0.3
POSITION 0 0.985 0
EMISSION 15 15 15
NEW_PRIMITIVE
BOX 0.3 0.6 0.3
POSITION -0.4 -0.4 0.3
ROTATION 0 0.258819 0 0.965926
METALLIC
COLOR 0.9 0.9 0.9
NEW_PRIMITIVE
ELLIPSOID 0.35 0.35 0.35
POSITION 0.45 -0.65 -0.3
DIELECTRIC
IOR 1.5
COLOR 1 1 1
NEW_PRIMITIVE
ELLIPSOID 0.25 0.15 0.25
POSITION 0.1 -0.85 -0.65
COLOR 0.2 0.3 0.8
"""


def cornell_text(width=512, height=512, spp=64, depth=6, extra=""):
    """Config C1.  `extra` appends more primitives (tests use it for
    emissive ellipsoids / text triangles so that every light kind is exercised)."""
    return CORNELL_TEXT.format(w=width, h=height, spp=spp, depth=depth) + extra


# extra primitives that exercise what the plain Cornell box does not: an emissive ellipsoid, an
# emissive rotated text TRIANGLE, a rotated diffuse box and a metallic triangle.
CORNELL_EXTRA = """\
NEW_PRIMITIVE
ELLIPSOID 0.12 0.08 0.1
POSITION -0.55 0.55 -0.5
ROTATION 0.1 0.2 0.3 0.9273618495495704
EMISSION 6 5 4
COLOR 0.5 0.5 0.5
NEW_PRIMITIVE
TRIANGLE -0.2 0 0 0.2 0 0 0 0.3 0
POSITION 0.55 0.3 0.6
ROTATION 0 0.3826834323650898 0 0.9238795325112867
EMISSION 3 6 9
COLOR 0.4 0.4 0.4
NEW_PRIMITIVE
BOX 0.15 0.15 0.15
POSITION 0.55 -0.85 0.55
ROTATION 0.2 0.4 0.1 0.8888194417315589
COLOR 0.7 0.6 0.2
NEW_PRIMITIVE
TRIANGLE -0.9 -0.95 0.9 0.9 -0.95 0.9 0 -0.2 0.95
METALLIC
COLOR 0.8 0.8 0.9
"""


def _icosphere(subdiv):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int64)
    verts = [tuple(p) for p in v]
    for _ in range(subdiv):
        cache = {}

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = (np.array(verts[a]) + np.array(verts[b])) / 2.0
                m /= np.linalg.norm(m)
                verts.append(tuple(m))
                cache[key] = len(verts) - 1
            return cache[key]

        nf = []
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        f = np.array(nf, np.int64)
    return np.array(verts, np.float64), f.astype(np.uint32)


def _look_at_quat(pos, target, up=(0.0, 1.0, 0.0)):
    pos, target, up = (np.asarray(a, np.float64) for a in (pos, target, up))
    f = target - pos
    f /= np.linalg.norm(f)
    z = -f                                       # glTF cameras look along local -Z
    x = np.cross(up, z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    m = np.stack([x, y, z], axis=1)              # columns = local axes in world space
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = [(m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, 0.25 * s]
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = [0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s, (m[2, 1] - m[1, 2]) / s]
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = [(m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s, (m[0, 2] - m[2, 0]) / s]
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = [(m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s, (m[1, 0] - m[0, 1]) / s]
    return [float(v) for v in q]


def heightfield_arrays(nx, nz, extent_x, extent_z, seed=12345):
    """(nx x nz quads) -> positions f32 (V,3), normals f32 (V,3), indices u32 (2*nx*nz*3,)."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(-extent_x / 2, extent_x / 2, nx + 1)
    zs = np.linspace(-extent_z / 2, extent_z / 2, nz + 1)
    x, z = np.meshgrid(xs, zs, indexing="ij")
    x = x + rng.uniform(-0.02, 0.02, x.shape) * (extent_x / nx) / 0.2
    z = z + rng.uniform(-0.02, 0.02, z.shape) * (extent_z / nz) / 0.2
    y = 4 * np.sin(0.13 * x) * np.cos(0.11 * z) + 1.5 * np.sin(0.71 * x + 0.5) * np.sin(0.53 * z)
    dydx = 4 * 0.13 * np.cos(0.13 * x) * np.cos(0.11 * z) + 1.5 * 0.71 * np.cos(0.71 * x + 0.5) * np.sin(0.53 * z)
    dydz = -4 * 0.11 * np.sin(0.13 * x) * np.sin(0.11 * z) + 1.5 * 0.53 * np.sin(0.71 * x + 0.5) * np.cos(0.53 * z)
    y = y + rng.uniform(-0.05, 0.05, y.shape) * min(1.0, (extent_x / nx) / 0.2)
    pos = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)
    nrm = np.stack([-dydx, np.ones_like(dydx), -dydz], axis=-1).reshape(-1, 3)
    nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    i, j = np.meshgrid(np.arange(nx), np.arange(nz), indexing="ij")
    v00 = (i * (nz + 1) + j).reshape(-1)
    v10, v01, v11 = v00 + (nz + 1), v00 + 1, v00 + (nz + 1) + 1
    idx = np.stack([v00, v01, v10, v10, v01, v11], axis=1).reshape(-1).astype(np.uint32)   # +Y facing
    return pos, nrm, idx


def write_heightfield_gltf(path, nx=500, nz=1000, extent_x=100.0, extent_z=200.0, lights=False,
                           spheres=False, seed=12345, light_strength=40.0, sphere_subdiv=4):
    """Writes <path> (.gltf) and <path minus ext>.bin.  nx*nz quads = 2*nx*nz triangles.
    C2: defaults (1 000 000 triangles).  C3: lights=True, spheres=True.  C5: nx=2000, nz=5000,
    extent 400 x 1000.  Returns the triangle count."""
    scale = extent_x / 100.0
    prims = []          # (positions f32, normals f32, indices u32, material dict)
    pos, nrm, idx = heightfield_arrays(nx, nz, extent_x, extent_z, seed)
    prims.append((pos, nrm, idx, {"pbrMetallicRoughness": {"baseColorFactor": [0.7, 0.7, 0.7, 1.0], "metallicFactor": 0.0}}))
    if lights:
        lp, li = [], []
        for a in range(4):
            for b in range(8):
                cx = (-30.0 + 20.0 * a) * scale
                cz = (-70.0 + 20.0 * b) * scale
                h = 1.5 * scale
                base = len(lp)
                lp += [[cx - h, 25.0, cz - h], [cx + h, 25.0, cz - h], [cx - h, 25.0, cz + h], [cx + h, 25.0, cz + h]]
                li += [base, base + 1, base + 2, base + 2, base + 1, base + 3]                 # facing -Y
        lp = np.array(lp, np.float32)
        ln = np.tile(np.array([[0.0, -1.0, 0.0]], np.float32), (len(lp), 1))
        prims.append((lp, ln, np.array(li, np.uint32),
                      {"pbrMetallicRoughness": {"baseColorFactor": [0.8, 0.8, 0.8, 1.0], "metallicFactor": 0.0},
                       "emissiveFactor": [1.0, 1.0, 1.0],
                       "extensions": {"KHR_materials_emissive_strength": {"emissive_strength": light_strength}}}))
    if spheres:
        sv, sf = _icosphere(sphere_subdiv)
        for centre, radius, mat in (((-15.0 * scale, 12.0, 0.0), 8.0, {"pbrMetallicRoughness": {"baseColorFactor": [0.9, 0.85, 0.7, 1.0], "metallicFactor": 1.0}}),
                                    ((15.0 * scale, 12.0, -20.0 * scale), 8.0, {"pbrMetallicRoughness": {"baseColorFactor": [0.95, 0.97, 1.0, 0.5], "metallicFactor": 0.0}})):
            p = (sv * radius + np.array(centre)).astype(np.float32)
            prims.append((p, sv.astype(np.float32), sf.reshape(-1), mat))

    bin_path = os.path.splitext(path)[0] + ".bin"
    views, accessors, gl_prims, materials = [], [], [], []
    offset = 0
    with open(bin_path, "wb") as fh:
        def put(arr, comp, typ):
            nonlocal offset
            data = np.ascontiguousarray(arr).tobytes()
            fh.write(data)
            views.append({"buffer": 0, "byteOffset": offset, "byteLength": len(data)})
            acc = {"bufferView": len(views) - 1, "componentType": comp, "count": int(arr.shape[0]), "type": typ}
            if typ == "VEC3" and comp == 5126 and len(accessors) % 3 == 0:
                acc["min"] = [float(v) for v in arr.min(axis=0)]
                acc["max"] = [float(v) for v in arr.max(axis=0)]
            accessors.append(acc)
            offset += len(data)
            pad = (-offset) % 4
            if pad:
                fh.write(b"\0" * pad)
                offset += pad
            return len(accessors) - 1

        n_tris = 0
        for p, n, i, mat in prims:
            a_pos, a_nrm, a_idx = put(p, 5126, "VEC3"), put(n, 5126, "VEC3"), put(i, 5125, "SCALAR")
            materials.append(mat)
            gl_prims.append({"attributes": {"POSITION": a_pos, "NORMAL": a_nrm}, "indices": a_idx,
                             "material": len(materials) - 1, "mode": 4})
            n_tris += len(i) // 3
    cam_pos = [0.0, 30.0 * max(1.0, scale ** 0.5), -110.0 * scale]
    doc = {
        "asset": {"version": "2.0", "generator": "rtb200.scenes"},
        "extensionsUsed": ["KHR_materials_emissive_strength"],
        "scene": 0,
        "scenes": [{"nodes": [0, 1]}],
        "nodes": [{"mesh": 0}, {"camera": 0, "translation": cam_pos, "rotation": _look_at_quat(cam_pos, (0.0, 0.0, 0.0))}],
        "cameras": [{"type": "perspective", "perspective": {"yfov": 0.8, "znear": 0.1, "aspectRatio": 16.0 / 9.0}}],
        "meshes": [{"primitives": gl_prims}],
        "materials": materials,
        "buffers": [{"uri": os.path.basename(bin_path), "byteLength": offset}],
        "bufferViews": views,
        "accessors": accessors,
    }
    with open(path, "w") as fh:
        json.dump(doc, fh)
    return n_tris


def deep_chain_text(n=130, growth=13.0, width=32, height=24, spp=2, depth=4, lights=True):
    """A text scene whose triangle hierarchy is a CHAIN: triangle k is `growth` times larger than
    triangle k-1, so the surface-area cost of the reference's sweep SAH (src/bvh.rs:120-135) is
    minimised by peeling the largest triangle off at every level (area ratio growth^2 > n) and the
    tree is about n levels deep — deeper than any fixed traversal stack (SURVEY.md §7 hard part 5).
    Triangles face the camera at increasing distance; every fourth one emits."""
    lines = [f"DIMENSIONS {width} {height}", f"SAMPLES {spp}", f"RAY_DEPTH {depth}", "BG_COLOR 0.1 0.2 0.3",
             "CAMERA_POSITION 0 0 -3", "CAMERA_FOV_X 1.2"]
    s = 0.5
    for k in range(n):
        # distance and size grow together, so every triangle has the same angular size; a 6 x 4 grid of
        # directions keeps two dozen of them visible side by side instead of one hiding the rest
        z = 2.0 * s
        x0, y0 = (-1.1 + 0.36 * (k % 6)) * s, (-0.75 + 0.36 * ((k // 6) % 4)) * s
        lines += ["NEW_PRIMITIVE", f"TRIANGLE {x0!r} {y0!r} {z!r} {x0 + 0.33 * s!r} {y0!r} {z!r} {x0!r} {y0 + 0.33 * s!r} {z * 1.05!r}",
                  f"COLOR {0.3 + 0.5 * (k % 3 == 0)} {0.3 + 0.5 * (k % 3 == 1)} {0.3 + 0.5 * (k % 3 == 2)}"]
        if lights and k % 4 == 1:
            lines.append("EMISSION 3 3 3")
        if k % 5 == 2:
            lines.append("METALLIC")
        s *= growth
    return "\n".join(lines) + "\n"
