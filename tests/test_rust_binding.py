"""The Rust half of the drop-in (integration/rust/) cannot be compiled here (no cargo / rustc): these checks keep it
honest against the C header it binds.

  * every #[repr(C)] struct in rtgpu-sys/src/lib.rs == the header's struct: same fields, same order, matching types;
  * the ctypes binding (abi.py) agrees with the header on the same structs (names, order, size of each field);
  * the extern "C" block declares exactly the functions of the header, with the same number of parameters;
  * camera_gpu.rs initialises every field of the structs it builds;
  * the patch applies to the reference tree (only where /root/reference exists).
"""
import ctypes as C
import os
import re
import shutil
import subprocess
import tempfile

import pytest

from ray_tracer_challenge_rs_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUST = os.path.join(ROOT, "integration", "rust")
HEADER = open(os.path.join(ROOT, "include", "rtgpu.h")).read()
LIB_RS = open(os.path.join(RUST, "rtgpu-sys", "src", "lib.rs")).read()
CAMERA_GPU = open(os.path.join(RUST, "ray-tracer", "src", "composites", "camera_gpu.rs")).read()

C_TO_RUST = {
    "uint32_t": "u32", "int32_t": "i32", "uint64_t": "u64", "double": "f64", "uint8_t": "u8", "size_t": "usize", "int": "c_int",
    "const uint8_t *": "*const u8", "const double *": "*const f64", "const int32_t *": "*const i32", "const uint32_t *": "*const u32",
}
RUST_SIZE = {"u32": 4, "i32": 4, "u64": 8, "f64": 8, "u8": 1}


def strip_comments(text):
    return re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def header_structs():
    out = {}
    for body, name in re.findall(r"typedef struct \w+ \{(.*?)\} (\w+);", strip_comments(HEADER), flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            m = re.match(r"(.*?)(\w+)(\[(\d+)\])?$", decl)
            ctype, fname, _, count = m.group(1).strip(), m.group(2), m.group(3), m.group(4)
            ctype = ctype.replace(" *", " *").strip()
            fields.append((fname, ctype, int(count) if count else None))
        out[name] = fields
    return out


def rust_structs():
    out = {}
    for name, body in re.findall(r"#\[repr\(C\)\](?:\s*#\[derive\([^\]]*\)\])?\s*pub struct (\w+) \{(.*?)\n\}", LIB_RS, flags=re.S):
        fields = []
        for fname, rtype in re.findall(r"pub (\w+): ([^,\n]+),", body):
            fields.append((fname, rtype.strip()))
        out[name] = fields
    return out


def expected_rust_type(ctype, count):
    base = C_TO_RUST[ctype if ctype in C_TO_RUST else ctype.replace("*", " *").replace("  ", " ")]
    return f"[{base}; {count}]" if count else base


@pytest.mark.parametrize("name", ["rtgpu_scene", "rtgpu_camera", "rtgpu_rows", "rtgpu_opts", "rtgpu_stats"])
def test_repr_c_structs_mirror_the_header(name):
    c_fields, r_fields = header_structs()[name], rust_structs()[name]
    assert [f[0] for f in c_fields] == [f[0] for f in r_fields], "field names / order"
    for (fname, ctype, count), (_, rtype) in zip(c_fields, r_fields):
        assert expected_rust_type(ctype, count) == rtype, (name, fname, ctype, rtype)


@pytest.mark.parametrize("name,ctypes_struct", [("rtgpu_scene", abi.RtgpuScene), ("rtgpu_camera", abi.RtgpuCamera), ("rtgpu_rows", abi.RtgpuRows),
                                                 ("rtgpu_opts", abi.RtgpuOpts), ("rtgpu_stats", abi.RtgpuStats)])
def test_ctypes_binding_mirrors_the_header(name, ctypes_struct):
    c_fields = header_structs()[name]
    assert [f[0] for f in c_fields] == [f[0] for f in ctypes_struct._fields_]
    # C layout rules applied to the header's declarations == what ctypes computed
    offset = 0
    for (fname, ctype, count), (_, ct) in zip(c_fields, ctypes_struct._fields_):
        size = 8 if "*" in ctype else RUST_SIZE[C_TO_RUST[ctype]]
        align = size
        offset = (offset + align - 1) // align * align
        assert getattr(ctypes_struct, fname).offset == offset, (name, fname)
        assert C.sizeof(ct) == size * (count or 1), (name, fname)
        offset += size * (count or 1)


def test_extern_block_declares_the_header_functions():
    declared = dict((n, len([a for a in args.split(",") if a.strip() and a.strip() != "void"]))
                    for n, args in re.findall(r"\b(rtgpu_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", strip_comments(HEADER), flags=re.S))
    block = re.sub(r"//[^\n]*", "", re.search(r'unsafe extern "C" \{(.*)\n\}', LIB_RS, flags=re.S).group(1))
    rust = dict((n, len([a for a in args.split(",") if a.strip()])) for n, args in re.findall(r"pub fn (rtgpu_\w+)\(([^)]*)\)", block, flags=re.S))
    assert set(rust) == set(declared) == set(abi.EXPORTED_SYMBOLS)
    assert rust == declared, {n: (rust[n], declared[n]) for n in rust if rust[n] != declared[n]}


def test_constants_agree():
    for name, value in re.findall(r"#define (RTGPU_\w+) (\d+)u?\b", HEADER):
        m = re.search(rf"pub const {name}: \w+ = (\d+);", LIB_RS)
        if m:
            assert int(m.group(1)) == int(value), name
    assert re.search(r"pub const RTGPU_ABI_VERSION: u32 = (\d+);", LIB_RS).group(1) == str(abi.ABI_VERSION)
    assert "pub const RTGPU_MAT_PARAM_COUNT: usize = 7;" in LIB_RS and abi.MAT_PARAM_COUNT == 7


@pytest.mark.parametrize("name", ["rtgpu_scene", "rtgpu_camera", "rtgpu_opts"])
def test_camera_gpu_initialises_every_field(name):
    literal = re.search(rf"sys::{name} \{{(.*?)\n        \}};", CAMERA_GPU, flags=re.S).group(1)
    used = set(re.findall(r"^\s*(\w+)(?::|,)", literal, flags=re.M))
    assert used == {f[0] for f in header_structs()[name]}, (used ^ {f[0] for f in header_structs()[name]})


def test_patch_applies_to_the_reference_tree():
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "ray-tracer-cli")) or shutil.which("patch") is None:
        pytest.skip("reference tree (or patch) not available here")
    with tempfile.TemporaryDirectory() as tmp:
        for sub in ("Cargo.toml", "README.md", "ray-tracer", "ray-tracer-cli"):
            src = os.path.join(ref, sub)
            (shutil.copytree if os.path.isdir(src) else shutil.copy)(src, os.path.join(tmp, sub))
        proc = subprocess.run(["patch", "-p1", "--forward", "-i", os.path.join(RUST, "gpu-rendering-mode.patch")], cwd=tmp, capture_output=True, text=True)
        assert proc.returncode == 0, proc.stdout + proc.stderr
        assert "RenderingMode::Gpu => camera.render_gpu(&world)," in open(os.path.join(tmp, "ray-tracer-cli/src/main.rs")).read()
        assert "mod camera_gpu;" in open(os.path.join(tmp, "ray-tracer/src/composites.rs")).read()
        assert "pub(crate) const fn colors" in open(os.path.join(tmp, "ray-tracer/src/patterns/ring_pattern.rs")).read()
