"""The drop-in command line (reference src/main.rs:30-83, run.sh) on the GPU box."""
import os
import subprocess

import numpy as np
import pytest

import rtb200
import rtb200.ctypes_types as T
from tests import helpers

pytestmark = pytest.mark.gpu
BIN = os.path.join(os.path.dirname(rtb200.__file__), "cpu-raytracing-rt")


def _read_ppm(path):
    data = open(path, "rb").read()
    assert data.startswith(b"P6\n")
    hdr_end = 0
    for _ in range(3):
        hdr_end = data.index(b"\n", hdr_end) + 1
    w, h = (int(v) for v in data.split(b"\n")[1].split())
    return np.frombuffer(data[hdr_end:], np.uint8).reshape(h, w, 3)


def test_custom_and_gltf_cli(orc, tmp_path):
    scene_txt = helpers.cornell(tmp_path, 48, 40, 3, 5)
    out = os.path.join(str(tmp_path), "o.ppm")
    env = dict(os.environ, RT_SEED="12345", RT_MODE="fast")
    r = subprocess.run([BIN, "custom", out], stdin=open(scene_txt), capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.split("\n")
    assert lines[0] == "Started scene parsing" and lines[1] == "Started raytracing" and "100%" in lines
    img = _read_ppm(out)
    assert img.shape == (40, 48, 3) and img.max() > 50
    # the same render through the library: identical bytes
    ctx = rtb200.Context(0).upload(rtb200.Scene.load_text(scene_txt))
    ctx.render_accum(0, 3, 12345, T.MODE_FAST, clear=True)
    assert np.array_equal(orc.to_bytes(ctx.postprocess(3)), img)
    ctx.close()
    gltf, _ = helpers.heightfield(tmp_path, 30, 60)
    out2 = os.path.join(str(tmp_path), "g.ppm")
    r = subprocess.run([BIN, "glTF", gltf, "64", "36", "2", out2], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    assert _read_ppm(out2).shape == (36, 64, 3)
    # usage errors print a message and exit 0 like the reference (src/main.rs:31-43)
    r = subprocess.run([BIN], capture_output=True, text=True)
    assert r.returncode == 0 and "You must specify the input format" in r.stdout
    r = subprocess.run([BIN, "obj", "x"], capture_output=True, text=True)
    assert r.returncode == 0 and "Input format must be 'custom' or 'glTF'" in r.stdout
    r = subprocess.run([BIN, "custom"], capture_output=True, text=True)
    assert r.returncode == 0 and "You must specify the output file" in r.stdout
    # a panic in the reference is a non-zero exit here
    r = subprocess.run([BIN, "glTF", os.path.join(str(tmp_path), "missing.gltf"), "4", "4", "1", out2], capture_output=True, text=True)
    assert r.returncode == 101


def test_cli_cache_and_preview(tmp_path):
    """RT_BVH_CACHE (on-disk hierarchy cache) and RT_PREVIEW (output rewritten after every chunk)
    change how the image is produced, never the image."""
    gltf, _ = helpers.heightfield(tmp_path, 30, 60)
    cache = os.path.join(str(tmp_path), "cache")
    os.makedirs(cache)
    env = dict(os.environ, RT_SEED="777")
    images = []
    for k, extra in enumerate(({}, {"RT_BVH_CACHE": cache}, {"RT_BVH_CACHE": cache, "RT_PREVIEW": "1"})):
        out = os.path.join(str(tmp_path), f"c{k}.ppm")
        r = subprocess.run([BIN, "glTF", gltf, "48", "32", "45", out], capture_output=True, text=True, env=dict(env, **extra), timeout=300)
        assert r.returncode == 0, r.stderr
        percents = [ln for ln in r.stdout.split("\n") if ln.endswith("%")]
        assert percents[-1] == "100%" and len(percents) == len(set(percents)) >= 10       # 45 spp in chunks of 3
        images.append(_read_ppm(out))
    assert len(os.listdir(cache)) == 1                      # written by the second run, read by the third
    assert np.array_equal(images[0], images[1]) and np.array_equal(images[0], images[2])
