"""CPU tests of the host side (no GPU): the product's scene loaders and bit-exact SAH build
against the oracle's restatement of reference src/scene.rs, src/bvh.rs:12-17,:75-144,
src/scene_parser.rs and src/gltf/scene_builder.rs; PPM bytes against src/ppm.rs; the C ABI."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import rtb200
import rtb200.ctypes_types as T
from rtb200.scene_desc import RawScene
from tests import helpers


def test_abi_exports_every_declared_symbol():
    L = rtb200.lib()
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "rt_b200.h")).read()
    declared = set(re.findall(r"RT_API\s+[\w\s\*]+?\b(rt_\w+)\s*\(", hdr))
    assert declared == set(rtb200.SYMBOLS), declared ^ set(rtb200.SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name


def test_no_device_is_an_error_not_a_fallback():
    if rtb200.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(rtb200.RtError) as e:
        rtb200.Context(0)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_text_scene_matches_oracle(orc, tmp_path):
    from oracle import loaders
    path = helpers.cornell(tmp_path, 96, 64, 8, 6)
    prod = rtb200.Scene.load_text(path)
    raw = loaders.parse_text_scene(open(path).read())
    oracle = orc.OracleScene(raw)
    assert (prod.width, prod.height, prod.samples, prod.ray_depth) == (96, 64, 8, 6)
    assert prod.info.n_planes == 6 and prod.info.bvh_prims[T.BVH_BOXES] == 3 and prod.info.bvh_prims[T.BVH_ELLIPSOIDS] == 3
    assert prod.info.bvh_prims[T.BVH_LIGHT_BOXES] == 1 and prod.info.bvh_prims[T.BVH_LIGHT_ELLIPSOIDS] == 1
    assert prod.info.bvh_prims[T.BVH_TRIANGLES] == 2 and prod.info.bvh_prims[T.BVH_LIGHT_TRIANGLES] == 1
    helpers.assert_scene_equal(prod, oracle)
    # the same description through rt_scene_build
    helpers.assert_scene_equal(rtb200.Scene.from_raw(raw), oracle)


def test_text_defaults_and_errors(tmp_path):
    p = helpers.write_text(tmp_path, "DIMENSIONS 3 2\nNEW_PRIMITIVE\nPLANE 0 1 0\nFOO bar\n\n")
    s = rtb200.Scene.load_text(p)
    assert (s.samples, s.ray_depth) == (64, 16)                     # src/scene.rs:185-187
    assert s.info.tan_half_fov_x == np.tan(np.pi / 2.0 / 2.0)       # CAMERA_FOV_X default pi/2
    for bad in ("NEW_PRIMITIVE\nPLANE 0 1 0\n",                     # DIMENSIONS missing -> unwrap panic
                "DIMENSIONS 2 2\nCOLOR 1 1 1\n",                    # property before NEW_PRIMITIVE
                "DIMENSIONS 2 2\nNEW_PRIMITIVE\n",                  # primitive without a shape
                "DIMENSIONS 2 2\nNEW_PRIMITIVE\nBOX 1 1\n",         # missing value
                "DIMENSIONS 2 2\nNEW_PRIMITIVE\nBOX 1 1 1\nDIELECTRIC\n",   # ior.unwrap()
                "DIMENSIONS 2 2\nRAY_DEPTH 300\n"):                 # u8 overflow
        with pytest.raises(rtb200.RtError) as e:
            rtb200.Scene.load_text(helpers.write_text(tmp_path, bad, "bad.txt"))
        assert e.value.code == 5
    with pytest.raises(rtb200.RtError) as e:
        rtb200.Scene.load_text(os.path.join(str(tmp_path), "missing.txt"))
    assert e.value.code == 4


def test_gltf_scene_matches_oracle(orc, tmp_path):
    from oracle import loaders
    path, n = helpers.heightfield(tmp_path, 40, 80)
    prod = rtb200.Scene.load_gltf(path, 160, 90, 4)
    raw = loaders.load_gltf(path, 160, 90, 4)
    assert raw.n_triangles() == n == prod.info.bvh_prims[T.BVH_TRIANGLES]
    assert prod.info.bvh_prims[T.BVH_LIGHT_TRIANGLES] == 64 and prod.ray_depth == 8
    assert prod.info.bvh_nodes[T.BVH_BOXES] == 0                    # BVH::new_empty for glTF scenes
    helpers.assert_scene_equal(prod, orc.OracleScene(raw))
    kinds = sorted(prod.material(i).kind for i in range(prod.info.n_materials))
    assert kinds == [T.MAT_DIFFUSE, T.MAT_DIFFUSE, T.MAT_METALLIC, T.MAT_DIELECTRIC]


def test_gltf_node_transforms_u16_indices_and_stride(orc, tmp_path):
    """TRS propagation through a node tree, matrix nodes, UNSIGNED_SHORT indices, byteStride,
    non-indexed primitives and the default (metallic) material: src/gltf/scene_builder.rs."""
    import json
    from oracle import loaders
    rng = np.random.default_rng(7)
    v = rng.uniform(-1, 1, (12, 3)).astype(np.float32)
    nrm = rng.normal(size=(12, 3)).astype(np.float32)
    inter = np.concatenate([v, nrm], axis=1)                        # interleaved, stride 24
    idx = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 2, 1, 9], np.uint16)
    blob = inter.tobytes() + idx.tobytes()
    open(os.path.join(str(tmp_path), "t.bin"), "wb").write(blob)
    doc = {
        "scenes": [{"nodes": [0, 3]}],
        "nodes": [{"children": [1], "translation": [1, 2, 3], "rotation": [0.1, 0.2, 0.3, 0.9273618495495704], "scale": [2, 3, 4]},
                  {"mesh": 0, "children": [2], "matrix": [0, 0, 1, 0, 0, 1, 0, 0, -1, 0, 0, 0, 0.5, 0.25, 2, 1]},
                  {"mesh": 1, "scale": [1, -1, 1]},
                  {"camera": 0, "translation": [0, 0, 5]}],
        "cameras": [{"type": "perspective", "perspective": {"yfov": 0.7}}],
        "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "NORMAL": 1}, "indices": 2, "material": 0}]},
                   {"primitives": [{"attributes": {"POSITION": 0, "NORMAL": 1}}]}],
        "materials": [{"emissiveFactor": [1, 0.5, 0.25]}],
        "buffers": [{"uri": "t.bin", "byteLength": len(blob)}],
        "bufferViews": [{"buffer": 0, "byteLength": 288, "byteStride": 24}, {"buffer": 0, "byteOffset": 288, "byteLength": 24}],
        "accessors": [{"bufferView": 0, "componentType": 5126, "count": 12, "type": "VEC3"},
                      {"bufferView": 0, "byteOffset": 12, "componentType": 5126, "count": 12, "type": "VEC3"},
                      {"bufferView": 1, "componentType": 5123, "count": 12, "type": "SCALAR"}],
    }
    path = os.path.join(str(tmp_path), "t.gltf")
    json.dump(doc, open(path, "w"))
    prod = rtb200.Scene.load_gltf(path, 32, 16, 2)
    raw = loaders.load_gltf(path, 32, 16, 2)
    assert raw.n_triangles() == 8
    helpers.assert_scene_equal(prod, orc.OracleScene(raw))
    assert prod.material(0).kind == T.MAT_METALLIC and tuple(prod.material(0).emission) == (1.0, 0.5, 0.25)
    assert prod.material(1).kind == T.MAT_METALLIC                  # default material: metallicFactor 1.0
    assert tuple(prod.info.cam_position) == (0.0, 0.0, 5.0) and tuple(prod.info.cam_forward) == (-0.0, -0.0, -1.0)


def test_gltf_errors(tmp_path):
    import json
    base = {"scenes": [{"nodes": [0]}], "nodes": [{"camera": 0}],
            "cameras": [{"type": "perspective", "perspective": {"yfov": 0.7}}]}
    ok = os.path.join(str(tmp_path), "ok.gltf")
    json.dump(base, open(ok, "w"))
    assert rtb200.Scene.load_gltf(ok, 4, 4, 1).info.bvh_prims[T.BVH_TRIANGLES] == 0
    for mutate in (lambda d: d.update(cameras=[]),                                  # single perspective camera
                   lambda d: d.update(nodes=[{"camera": 0}, {"camera": 0}], scenes=[{"nodes": [0, 1]}]),
                   lambda d: d.update(nodes=[{}]),                                  # no camera node
                   lambda d: d.update(nodes=[{"camera": 0, "mesh": 0}], meshes=[{"primitives": [{"attributes": {"POSITION": 0}}]}])):
        d = json.loads(json.dumps(base))
        mutate(d)
        bad = os.path.join(str(tmp_path), "bad.gltf")
        json.dump(d, open(bad, "w"))
        with pytest.raises(rtb200.RtError):
            rtb200.Scene.load_gltf(bad, 4, 4, 1)
    open(os.path.join(str(tmp_path), "junk.gltf"), "w").write("{ not json")
    with pytest.raises(rtb200.RtError) as e:
        rtb200.Scene.load_gltf(os.path.join(str(tmp_path), "junk.gltf"), 4, 4, 1)
    assert e.value.code == 5


@pytest.mark.parametrize("n,seed", [(0, 1), (1, 2), (4, 3), (5, 4), (33, 5), (1000, 6), (20000, 7)])
def test_bvh_build_bit_exact_random_soups(orc, n, seed):
    """Random triangle soups incl. the edge sizes around the leaf rule (<= 4) and the parallel
    build threshold; node array + primitive order bit-exact against the oracle."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-10, 10, (n, 1, 3))
    tri = (c + rng.normal(scale=0.3, size=(n, 3, 3))).reshape(n, 9)
    raw = RawScene(width=4, height=4)
    raw.materials = [{"kind": T.MAT_DIFFUSE, "color": (0.5, 0.5, 0.5)}, {"kind": T.MAT_DIFFUSE, "emission": (1.0, 1.0, 1.0)}]
    raw.tri_pos = tri if n else None
    raw.tri_mat = (rng.uniform(size=n) < 0.1).astype(np.uint32) if n else None
    helpers.assert_scene_equal(rtb200.Scene.from_raw(raw), orc.OracleScene(raw))


def test_bvh_build_with_midpoint_ties(orc):
    """A regular grid of identical quads: every axis sort is full of equal midpoints, and many
    splits tie on score.  Stable tie order + first-best split must match the oracle exactly."""
    xs, zs = np.meshgrid(np.arange(24.0), np.arange(16.0), indexing="ij")
    a = np.stack([xs, np.zeros_like(xs), zs], -1).reshape(-1, 3)
    t1 = np.concatenate([a, a + [1, 0, 0], a + [0, 0, 1]], 1)
    t2 = np.concatenate([a + [1, 0, 0], a + [1, 0, 1], a + [0, 0, 1]], 1)
    raw = RawScene(width=4, height=4)
    raw.materials = [{"kind": T.MAT_DIFFUSE}]
    raw.tri_pos = np.stack([t1, t2], 1).reshape(-1, 9)
    raw.tri_mat = np.zeros(len(raw.tri_pos), np.uint32)
    prod, oracle = rtb200.Scene.from_raw(raw), orc.OracleScene(raw)
    helpers.assert_scene_equal(prod, oracle)
    assert prod.info.bvh_depth[T.BVH_TRIANGLES] >= 7


def test_bvh_build_larger_mesh_parallel_path(orc, tmp_path):
    """120k triangles: exercises the radix sort and the task-parallel subtree stitching."""
    from oracle import loaders
    path, n = helpers.heightfield(tmp_path, 200, 300, lights=True, spheres=False)
    prod = rtb200.Scene.load_gltf(path, 64, 36, 1)
    oracle = orc.OracleScene(loaders.load_gltf(path, 64, 36, 1))
    assert n == 120064
    helpers.assert_scene_equal(prod, oracle)


def test_ppm_bytes(orc, tmp_path):
    rng = np.random.default_rng(3)
    img = rng.uniform(-0.2, 1.2, (5, 7, 3))
    img[0, 0] = [0.5 / 255.0, 1.5 / 255.0, 2.5 / 255.0]             # round-half-away-from-zero cases
    img[0, 1] = [np.nan, np.inf, -np.inf]
    out = os.path.join(str(tmp_path), "o.ppm")
    rtb200.save_ppm(out, img)
    data = open(out, "rb").read()
    assert data.startswith(b"P6\n7 5\n255\n")
    body = np.frombuffer(data[len(b"P6\n7 5\n255\n"):], np.uint8).reshape(5, 7, 3)
    assert np.array_equal(body, orc.to_bytes(img))
    assert tuple(body[0, 0]) == (1, 2, 3) and tuple(body[0, 1]) == (0, 255, 0)


def test_prebuilt_scene_round_trip(orc, tmp_path):
    """rt_scene_from_built: a host that keeps its own BVH::new hands the hierarchies over as flat
    arrays.  Exporting a built scene and importing it again must give the same scene bit for bit;
    malformed hierarchies are rejected."""
    from oracle import loaders
    path = helpers.cornell(tmp_path, 32, 24, 2, 5)
    prod = rtb200.Scene.load_text(path)
    built, keep = prod.export_built()
    again = rtb200.Scene.from_built(built)
    helpers.assert_scene_equal(again, orc.OracleScene(loaders.parse_text_scene(open(path).read())), prim_order=False)
    assert list(again.info.bvh_depth) == list(prod.info.bvh_depth)
    gltf, _ = helpers.heightfield(tmp_path, 12, 20)
    prod2 = rtb200.Scene.load_gltf(gltf, 16, 9, 1)
    built2, keep2 = prod2.export_built()
    helpers.assert_scene_equal(rtb200.Scene.from_built(built2), orc.OracleScene(loaders.load_gltf(gltf, 16, 9, 1)), prim_order=False)
    # not in pre-order: swap the children of the root
    d = keep2["bvh"][T.BVH_TRIANGLES]
    d["left"][0], d["right"][0] = d["right"][0], d["left"][0]
    with pytest.raises(rtb200.RtError):
        rtb200.Scene.from_built(built2)
    d["left"][0], d["right"][0] = d["right"][0], d["left"][0]
    d["count"][-1] += 5                                              # leaf range past the primitive array
    with pytest.raises(rtb200.RtError):
        rtb200.Scene.from_built(built2)
    d["count"][-1] -= 5
    # a child shared by two parents (a DAG): passes every per-node check, but the real path length
    # is no longer what the depth pass computes and the traversal stacks are sized from that depth
    assert d["left"][1] != 0xFFFFFFFF and d["right"][1] > 2
    keep_right = d["right"][0]
    d["right"][0] = d["right"][1]
    with pytest.raises(rtb200.RtError) as e:
        rtb200.Scene.from_built(built2)
    assert "two parents" in str(e.value)
    d["right"][0] = keep_right
    rtb200.Scene.from_built(built2)                                  # restored: accepted again


def test_scene_cache_round_trip(orc, tmp_path):
    """the built hierarchies are stored on disk keyed by a hash of the geometry.
    A hit must give the bits a fresh build gives (== the oracle's), with the VIEW of the current
    call (image size, samples, camera), and another mesh must not hit."""
    from oracle import loaders
    cache = tmp_path / "cache"
    cache.mkdir()
    path, _ = helpers.heightfield(tmp_path, 30, 40, lights=True, spheres=True)
    first = rtb200.Scene.load_gltf(path, 64, 36, 2, cache_dir=str(cache))
    assert not first.cache_hit and len(list(cache.glob("*.rtbvh"))) == 1
    again = rtb200.Scene.load_gltf(path, 96, 54, 7, cache_dir=str(cache))      # same mesh, other view
    assert again.cache_hit and len(list(cache.glob("*"))) == 1                 # no temporary left behind
    assert (again.info.width, again.info.height, again.info.samples) == (96, 54, 7)
    helpers.assert_scene_equal(again, orc.OracleScene(loaders.load_gltf(path, 96, 54, 7)))
    helpers.assert_scene_equal(first, orc.OracleScene(loaders.load_gltf(path, 64, 36, 2)))
    # text-format scene through the in-memory entry point: shapes + planes + model-space triangles
    raw = loaders.parse_text_scene(open(helpers.cornell(tmp_path)).read())
    a = rtb200.Scene.from_raw(raw, cache_dir=str(cache))
    b = rtb200.Scene.from_raw(raw, cache_dir=str(cache))
    assert not a.cache_hit and b.cache_hit and len(list(cache.glob("*.rtbvh"))) == 2
    helpers.assert_scene_equal(b, orc.OracleScene(raw))
    # one changed coordinate is another mesh
    raw.shapes[0]["position"] = tuple(np.add(raw.shapes[0]["position"], (0.0, 1e-9, 0.0)))
    c = rtb200.Scene.from_raw(raw, cache_dir=str(cache))
    assert not c.cache_hit and len(list(cache.glob("*.rtbvh"))) == 3


def test_scene_cache_rejects_damaged_files(tmp_path):
    """A truncated or corrupted cache file is an error (RT_ERR_IO = 4), not a silent rebuild and
    not an out-of-range index later on the device."""
    cache = tmp_path / "cache"
    cache.mkdir()
    path, _ = helpers.heightfield(tmp_path, 12, 12, lights=True, spheres=False)
    rtb200.Scene.load_gltf(path, 16, 16, 1, cache_dir=str(cache))
    (f,) = list(cache.glob("*.rtbvh"))
    blob = f.read_bytes()
    f.write_bytes(blob[:-40])
    with pytest.raises(rtb200.RtError) as e:
        rtb200.Scene.load_gltf(path, 16, 16, 1, cache_dir=str(cache))
    assert e.value.code == 4 and "size" in str(e.value)
    # same size, a child link pointing outside the node array
    import struct
    n_mat = struct.unpack_from("<I", blob, 28)[0]
    n_planes = struct.unpack_from("<Q", blob, 40)[0]
    assert n_planes == 0
    first_node = 48 + 8 * 12 + n_mat * C.sizeof(T.Material)        # header, materials, (no planes / shapes)
    bad = bytearray(blob)
    n_nodes = struct.unpack_from("<6Q", blob, 48)
    assert n_nodes[T.BVH_TRIANGLES] > 1 and n_nodes[T.BVH_BOXES] == 0
    struct.pack_into("<I", bad, first_node + 48 + 4, 0x7FFFFFFF)    # root.right of the triangle hierarchy
    f.write_bytes(bytes(bad))
    with pytest.raises(rtb200.RtError) as e:
        rtb200.Scene.load_gltf(path, 16, 16, 1, cache_dir=str(cache))
    assert e.value.code == 4 and "pre-order" in str(e.value)
    # same size, a child shared by two parents: root.right := node1.right
    bad = bytearray(blob)
    left1, right1 = struct.unpack_from("<2I", blob, first_node + 64 + 48)
    assert left1 == 2 and right1 > 2
    struct.pack_into("<I", bad, first_node + 48 + 4, right1)
    f.write_bytes(bytes(bad))
    with pytest.raises(rtb200.RtError) as e:
        rtb200.Scene.load_gltf(path, 16, 16, 1, cache_dir=str(cache))
    assert e.value.code == 4 and "two parents" in str(e.value)
    # a file of another mesh that happens to carry this mesh's hash (collision / stale file): the
    # counts in the header differ from the caller's scene -> a miss, rebuilt and rewritten, no error
    other, _ = helpers.heightfield(tmp_path, 10, 12, lights=True, spheres=False, name="other.gltf")
    cache2 = tmp_path / "cache2"
    cache2.mkdir()
    rtb200.Scene.load_gltf(other, 16, 16, 1, cache_dir=str(cache2))
    (g,) = list(cache2.glob("*.rtbvh"))
    forged = bytearray(g.read_bytes())
    forged[32:40] = blob[32:40]                                      # the hash field of the first mesh
    f.write_bytes(bytes(forged))
    sc = rtb200.Scene.load_gltf(path, 16, 16, 1, cache_dir=str(cache))
    assert not sc.cache_hit and f.read_bytes() == blob
    with pytest.raises(rtb200.RtError):                             # unwritable cache directory
        rtb200.Scene.load_gltf(path, 16, 16, 1, cache_dir=str(tmp_path / "missing" / "dir"))


def test_bvh_build_team_path_is_bit_exact():
    """The top-level nodes of a large build split their own sorts and sweeps over idle threads
    (scene_build.cpp: sort_team / sweep_team).  RT_TEAM_MIN=64 forces that path on every node of
    the small test hierarchies (read once per process, hence the child process): the soups, the
    tie-heavy grid and the 120k-triangle mesh must still match the oracle bit for bit."""
    import subprocess
    import sys
    if os.environ.get("RT_TEAM_MIN"):
        pytest.skip("already inside the forced-team run")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_host_scene.py", "-x", "-q", "-k", "soups or ties or larger_mesh"],
                       cwd=root, env=dict(os.environ, RT_TEAM_MIN="64"), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def _random_scene_text(rng):
    """A valid scene file with the keywords in random order inside each block, random optional
    properties, mixed number formats, blank lines and unknown keywords (ignored by the parser,
    src/scene_parser.rs:40-42)."""
    def num(v):
        return rng.choice([repr(float(v)), "%.6f" % v, "%.3e" % v, str(int(v)) if float(v).is_integer() else repr(float(v))])

    def vec(n, lo=-3.0, hi=3.0):
        return " ".join(num(v) for v in rng.uniform(lo, hi, n))

    head = [f"DIMENSIONS {int(rng.integers(1, 40))} {int(rng.integers(1, 40))}"]
    opt = [f"SAMPLES {int(rng.integers(1, 9))}", f"RAY_DEPTH {int(rng.integers(0, 20))}", "BG_COLOR " + vec(3, 0, 1),
           "CAMERA_POSITION " + vec(3), "CAMERA_RIGHT " + vec(3), "CAMERA_UP " + vec(3), "CAMERA_FORWARD " + vec(3),
           f"CAMERA_FOV_X {num(rng.uniform(0.2, 2.5))}", "SOMETHING_ELSE 1 2 3", ""]
    head += [o for o in opt if rng.uniform() < 0.6]
    rng.shuffle(head)
    lines = list(head)
    for _ in range(int(rng.integers(0, 25))):
        kind = rng.choice(["BOX", "PLANE", "ELLIPSOID", "TRIANGLE"])
        block = [kind + " " + (vec(9) if kind == "TRIANGLE" else vec(3, 0.05, 2.0))]
        if rng.uniform() < 0.7:
            block.append("POSITION " + vec(3))
        if rng.uniform() < 0.5:
            q = rng.normal(size=4)
            block.append("ROTATION " + " ".join(num(v) for v in q / np.linalg.norm(q)))
        if rng.uniform() < 0.7:
            block.append("COLOR " + vec(3, 0, 1))
        if rng.uniform() < 0.3:
            block.append("EMISSION " + vec(3, 0, 9))
        m = rng.uniform()
        if m < 0.2:
            block.append("METALLIC")
        elif m < 0.4:
            block += ["DIELECTRIC", f"IOR {num(rng.uniform(1.1, 2.4))}"]
        rng.shuffle(block)
        lines += ["NEW_PRIMITIVE"] + block + ([""] if rng.uniform() < 0.2 else [])
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("seed", range(12))
def test_random_text_scenes_parse_and_build_like_the_oracle(orc, tmp_path, seed):
    """Parser + Scene::new defaults + baking + the six hierarchies on random files: the C++ host
    (scene_parser.rs / scene.rs restated in loaders.cpp, scene_build.cpp) against the oracle's own
    Python parser and C++ build, bit for bit."""
    from oracle import loaders
    rng = np.random.default_rng(100 + seed)
    text = _random_scene_text(rng)
    prod = rtb200.Scene.load_text(helpers.write_text(tmp_path, text))
    raw = loaders.parse_text_scene(text)
    helpers.assert_scene_equal(prod, orc.OracleScene(raw))
    i = prod.info
    assert (i.width, i.height, i.samples, i.ray_depth) == (raw.width, raw.height, raw.samples, raw.ray_depth)
    for name in ("bg_color", "cam_position", "cam_right", "cam_up", "cam_forward"):
        assert list(getattr(i, name)) == [float(v) for v in getattr(raw, name)], name


def _random_gltf(rng, tmp_path, name):
    """A random glTF inside the subset the reference accepts: several meshes and
    primitives, interleaved or separate buffer views, u16 / u32 / no indices, a node tree mixing TRS
    and matrix nodes, random materials (alpha -> dielectric, metallicFactor, emissive + strength)."""
    import json
    blob = bytearray()
    views, accessors, meshes, materials = [], [], [], []

    def add_view(data, stride=None):
        while len(blob) % 4:
            blob.append(0)
        v = {"buffer": 0, "byteOffset": len(blob), "byteLength": len(data)}
        if stride:
            v["byteStride"] = stride
        blob.extend(data)
        views.append(v)
        return len(views) - 1

    for _ in range(int(rng.integers(1, 4))):
        m = {}
        if rng.uniform() < 0.5:
            m["pbrMetallicRoughness"] = {}
            if rng.uniform() < 0.6:
                m["pbrMetallicRoughness"]["metallicFactor"] = float(rng.choice([0.0, 0.0, 0.4, 1.0]))
            if rng.uniform() < 0.6:
                m["pbrMetallicRoughness"]["baseColorFactor"] = [float(v) for v in rng.uniform(0, 1, 3)] + [float(rng.choice([1.0, 0.5]))]
        if rng.uniform() < 0.4:
            m["emissiveFactor"] = [float(v) for v in rng.uniform(0, 1, 3)]
            if rng.uniform() < 0.5:
                m["extensions"] = {"KHR_materials_emissive_strength": {"emissive_strength": float(rng.uniform(1, 30))}}   # the reference reads this snake_case name
        materials.append(m)
    for _ in range(int(rng.integers(1, 4))):
        prims = []
        for _ in range(int(rng.integers(1, 3))):
            nv = int(rng.integers(3, 40))
            pos = rng.uniform(-2, 2, (nv, 3)).astype(np.float32)
            nrm = rng.normal(size=(nv, 3)).astype(np.float32)
            if rng.uniform() < 0.5:                                  # interleaved, stride 24 (+ optional padding)
                pad = int(rng.choice([0, 8]))
                inter = np.concatenate([pos, nrm, np.zeros((nv, pad // 4), np.float32)], axis=1)
                bv = add_view(inter.tobytes(), 24 + pad)
                accessors.append({"bufferView": bv, "componentType": 5126, "count": nv, "type": "VEC3"})
                accessors.append({"bufferView": bv, "byteOffset": 12, "componentType": 5126, "count": nv, "type": "VEC3"})
            else:
                accessors.append({"bufferView": add_view(pos.tobytes()), "componentType": 5126, "count": nv, "type": "VEC3"})
                accessors.append({"bufferView": add_view(nrm.tobytes()), "componentType": 5126, "count": nv, "type": "VEC3"})
            p = {"attributes": {"POSITION": len(accessors) - 2, "NORMAL": len(accessors) - 1}}
            kind = rng.integers(0, 3)
            if kind:                                                 # indexed: u16 or u32
                ni = 3 * int(rng.integers(1, 30))
                dt, ct = (np.uint16, 5123) if kind == 1 else (np.uint32, 5125)
                accessors.append({"bufferView": add_view(rng.integers(0, nv, ni).astype(dt).tobytes()), "componentType": ct, "count": ni, "type": "SCALAR"})
                p["indices"] = len(accessors) - 1
            else:
                nv3 = nv - nv % 3                                     # non-indexed: whole triangles only
                accessors[-2]["count"] = accessors[-1]["count"] = nv3
            if rng.uniform() < 0.8:
                p["material"] = int(rng.integers(0, len(materials)))
            if rng.uniform() < 0.3:
                p["mode"] = 4
            prims.append(p)
        meshes.append({"primitives": prims})
    nodes = []

    def make_node(depth):
        n = {}
        if rng.uniform() < 0.3:
            q = rng.normal(size=4)
            q /= np.linalg.norm(q)
            x, y, z, w = q
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            M = np.eye(4)
            M[:3, :3] = R * rng.uniform(0.5, 2.0, 3)
            M[:3, 3] = rng.uniform(-3, 3, 3)
            n["matrix"] = [float(v) for v in M.T.reshape(-1)]         # column-major
        else:
            if rng.uniform() < 0.6:
                n["translation"] = [float(v) for v in rng.uniform(-3, 3, 3)]
            if rng.uniform() < 0.6:
                q = rng.normal(size=4)
                n["rotation"] = [float(v) for v in q / np.linalg.norm(q)]
            if rng.uniform() < 0.6:
                n["scale"] = [float(v) for v in rng.uniform(0.3, 2.5, 3) * rng.choice([-1.0, 1.0], 3)]
        if rng.uniform() < 0.7:
            n["mesh"] = int(rng.integers(0, len(meshes)))
        nodes.append(n)
        me = len(nodes) - 1
        if depth < 3:
            kids = [make_node(depth + 1) for _ in range(int(rng.integers(0, 3)))]
            if kids:
                nodes[me]["children"] = kids
        return me

    roots = [make_node(0) for _ in range(int(rng.integers(1, 4)))]
    cam = {"camera": 0, "translation": [float(v) for v in rng.uniform(-5, 5, 3)]}
    q = rng.normal(size=4)
    cam["rotation"] = [float(v) for v in q / np.linalg.norm(q)]
    nodes.append(cam)
    host = int(rng.integers(0, len(nodes) - 1))                       # the camera hangs somewhere in the tree
    if rng.uniform() < 0.5:
        nodes[host].setdefault("children", []).append(len(nodes) - 1)
    else:
        roots.append(len(nodes) - 1)
    doc = {"scenes": [{"nodes": roots}], "scene": 0, "nodes": nodes, "meshes": meshes, "materials": materials,
           "cameras": [{"type": "perspective", "perspective": {"yfov": float(rng.uniform(0.3, 1.5)), "znear": 0.1}}],
           "buffers": [{"uri": name + ".bin", "byteLength": len(blob)}], "bufferViews": views, "accessors": accessors}
    open(os.path.join(str(tmp_path), name + ".bin"), "wb").write(bytes(blob))
    path = os.path.join(str(tmp_path), name + ".gltf")
    json.dump(doc, open(path, "w"))
    return path


@pytest.mark.parametrize("seed", range(10))
def test_random_gltf_files_load_and_build_like_the_oracle(orc, tmp_path, seed):
    """gltf::parse + build_scene (src/gltf/*.rs) restated twice — loaders.cpp for the product,
    oracle/loaders.py for the oracle — must agree bit for bit on random files: triangle order,
    world-space vertices and normals, materials, camera axes, then the hierarchies."""
    from oracle import loaders
    rng = np.random.default_rng(500 + seed)
    path = _random_gltf(rng, tmp_path, f"r{seed}")
    w, h, spp = int(rng.integers(4, 60)), int(rng.integers(4, 60)), int(rng.integers(1, 9))
    prod = rtb200.Scene.load_gltf(path, w, h, spp)
    raw = loaders.load_gltf(path, w, h, spp)
    helpers.assert_scene_equal(prod, orc.OracleScene(raw))
    assert prod.info.n_materials == len(raw.materials)
    for k in range(len(raw.materials)):
        m = prod.material(k)
        assert m.kind == raw.materials[k]["kind"] and list(m.emission) == [float(v) for v in raw.materials[k].get("emission", (0, 0, 0))], k
    for name in ("cam_position", "cam_right", "cam_up", "cam_forward"):
        assert list(getattr(prod.info, name)) == [float(v) for v in getattr(raw, name)], name


def test_scene_cache_survives_random_corruption(tmp_path):
    """Random byte flips anywhere in a cache file: the load either fails with RT_ERR_IO or yields a
    scene whose hierarchies pass the structural checks (flipped float payload is not detectable and
    not dangerous) — it never crashes and never hands out-of-range indices to the device code."""
    cache = tmp_path / "cache"
    cache.mkdir()
    path, _ = helpers.heightfield(tmp_path, 10, 12, lights=True, spheres=False)
    rtb200.Scene.load_gltf(path, 16, 16, 1, cache_dir=str(cache))
    (f,) = list(cache.glob("*.rtbvh"))
    blob = f.read_bytes()
    rng = np.random.default_rng(3)
    outcomes = {"error": 0, "loaded": 0}
    for trial in range(150):
        bad = bytearray(blob)
        region = len(blob) if trial % 3 else 4096                      # a third of the trials hit the header / first nodes
        for _ in range(int(rng.integers(1, 6))):
            pos = int(rng.integers(16, region))                        # keep the hash field region reachable, magic intact sometimes
            bad[pos] ^= 1 << int(rng.integers(0, 8))
        f.write_bytes(bytes(bad))
        try:
            sc = rtb200.Scene.load_gltf(path, 16, 16, 1, cache_dir=str(cache))
        except rtb200.RtError as e:
            assert e.code == 4, e
            outcomes["error"] += 1
            f.write_bytes(blob)                                        # (a hash mismatch rebuilt and rewrote the file: restore)
            continue
        outcomes["loaded"] += 1
        for which in (T.BVH_TRIANGLES, T.BVH_LIGHT_TRIANGLES):
            b = sc.bvh(which)
            n, n_prims = len(b["left"]), int(sc.info.bvh_prims[which])
            leaf = b["left"] == 0xFFFFFFFF
            assert np.all(b["right"][~leaf] < n) and np.all(b["left"][~leaf] == np.nonzero(~leaf)[0] + 1)
            assert np.all(b["first"][leaf].astype(np.int64) + b["count"][leaf] <= n_prims)
        f.write_bytes(blob)
    assert outcomes["error"] > 0 and outcomes["loaded"] > 0, outcomes


def test_loaders_survive_mutated_files(tmp_path):
    """Truncated / byte-flipped scene files (text, glTF JSON, glTF .bin): the loaders either reject
    them (RT_ERR_PARSE / RT_ERR_IO, where the reference would panic) or load something — they never
    crash or read outside the buffers they were given."""
    rng = np.random.default_rng(9)
    text = _random_scene_text(np.random.default_rng(1))
    gltf = _random_gltf(np.random.default_rng(2), tmp_path, "base")
    gltf_json = open(gltf, "rb").read()
    gltf_bin = open(gltf[:-5] + ".bin", "rb").read()
    seen = {"ok": 0, "rejected": 0}

    def mutate(data):
        data = bytearray(data)
        kind = rng.integers(0, 3)
        if kind == 0 and len(data) > 8:
            del data[int(rng.integers(1, len(data))):]                  # truncate
        elif kind == 1:
            for _ in range(int(rng.integers(1, 8))):
                data[int(rng.integers(0, len(data)))] = int(rng.integers(32, 127))
        else:
            a = int(rng.integers(0, len(data)))
            data[a:a] = bytes(rng.integers(32, 127, int(rng.integers(1, 12)), dtype=np.uint8))
        return bytes(data)

    def attempt(load):
        try:
            load()
            seen["ok"] += 1
        except rtb200.RtError as e:
            assert e.code in (4, 5), e
            seen["rejected"] += 1

    for trial in range(int(os.environ.get("RT_FUZZ_TRIALS", 120))):
        p = helpers.write_text(tmp_path, mutate(text.encode()).decode("latin-1"), "m.txt")
        attempt(lambda: rtb200.Scene.load_text(p))
        which = trial % 2
        open(os.path.join(str(tmp_path), "m.gltf"), "wb").write(mutate(gltf_json.replace(b"base.bin", b"m.bin")) if which == 0 else gltf_json.replace(b"base.bin", b"m.bin"))
        open(os.path.join(str(tmp_path), "m.bin"), "wb").write(gltf_bin if which == 0 else mutate(gltf_bin))
        attempt(lambda: rtb200.Scene.load_gltf(os.path.join(str(tmp_path), "m.gltf"), 8, 8, 1))
    assert seen["ok"] > 0 and seen["rejected"] > 0, seen
