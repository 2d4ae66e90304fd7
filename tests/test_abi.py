"""The C ABI seen from plain C and from the Rust `-sys` crate.

No Rust toolchain exists in the build image, so the `#[repr(C)]` mirrors in rust/rt_b200-sys/src/lib.rs
cannot be compiled here.  Instead every struct there carries an `@layout` line, and tests/abi/abi_smoke.c —
compiled by gcc against include/rt_b200.h ONLY — prints the same table from `sizeof` / `offsetof`; the
two must be equal.  The same C program then walks the flow a host performs in place of the reference's
generate_image (src/main.rs:73): build -> create -> upload -> render -> destroy."""
import ctypes
import os
import re
import subprocess

import pytest

import rtb200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.path.dirname(rtb200.__file__)
LIB_RS = os.path.join(ROOT, "rust", "rt_b200-sys", "src", "lib.rs")
HEADER = os.path.join(ROOT, "include", "rt_b200.h")


@pytest.fixture(scope="module")
def abi_smoke(tmp_path_factory):
    rtb200.lib()                                    # builds / loads librt_b200.so first
    exe = str(tmp_path_factory.mktemp("abi") / "abi_smoke")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "abi", "abi_smoke.c"), "-L" + LIB_DIR, "-lrt_b200",
                           "-Wl,-rpath," + LIB_DIR, "-lm", "-o", exe])
    return exe


def _layout_lines(text):
    out = {}
    for line in text.splitlines():
        m = re.search(r"@layout (\w+) (.*)$", line)
        if m:
            out[m.group(1)] = dict(kv.split("=") for kv in m.group(2).split())
    return out


def test_rust_layout_tables_equal_the_header(abi_smoke):
    c_text = subprocess.run([abi_smoke, "--layout"], capture_output=True, text=True, check=True).stdout
    c_tab, rs_text = _layout_lines(c_text), open(LIB_RS).read()
    rs_tab = _layout_lines(rs_text)
    assert c_tab and set(c_tab) == set(rs_tab)
    for name in c_tab:
        assert c_tab[name] == rs_tab[name], name
    # every struct of the header is in the table, and every table has a #[repr(C)] struct with those fields in that order
    hdr_structs = set(re.findall(r"typedef struct (rt_\w+) \{", open(HEADER).read()))
    assert hdr_structs == set(c_tab)
    for name, tab in rs_tab.items():
        m = re.search(r"#\[repr\(C\)\]\n(?:#\[derive[^\n]*\]\n)?pub struct %s \{(.*?)\n\}" % name, rs_text, re.S)
        assert m, name
        fields = re.findall(r"pub (\w+):", m.group(1))
        want = [("ty" if k == "type" else k) for k in tab if k != "size"]
        assert fields == want, name
    consts = dict(kv.split("=") for kv in re.search(r"@const (.*)", c_text).group(1).split())
    assert re.search(r"@const (.*)", rs_text).group(1).split() == re.search(r"@const (.*)", c_text).group(1).split()
    assert int(consts["RT_NO_CHILD"]) == 0xFFFFFFFF and int(consts["RT_COMM_ID_BYTES"]) == 128


def test_rust_declares_exactly_the_exported_functions():
    declared_h = set(re.findall(r"RT_API\s+[\w\s\*]+?\b(rt_\w+)\s*\(", open(HEADER).read()))
    declared_rs = set(re.findall(r"pub fn (rt_\w+)\(", open(LIB_RS).read()))
    assert declared_h == declared_rs, declared_h ^ declared_rs
    lib = ctypes.CDLL(os.path.join(LIB_DIR, "librt_b200.so"))
    for name in declared_rs:
        assert hasattr(lib, name), name


def test_flow_refuses_without_a_device_or_renders(abi_smoke):
    """On a box without a GPU the flow must stop at rt_create with RT_ERR_NO_DEVICE (the program checks
    that itself); with one it renders — the same binary serves the CPU and the GPU suites."""
    r = subprocess.run([abi_smoke, "--flow"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("flow ok")


@pytest.mark.gpu
def test_flow_renders_on_the_gpu(abi_smoke):
    r = subprocess.run([abi_smoke, "--flow"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "device(s)" in r.stdout and "segments" in r.stdout, r.stdout
