"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/rtgpu.h declares,
its host-only helpers work, and without a GPU the render entry points fail LOUDLY (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ray_tracer_challenge_rs_b200 import abi, build
from ray_tracer_challenge_rs_b200.fixtures import load_scene_fixture
from ray_tracer_challenge_rs_b200.flatten import camera_to_c

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return abi.load_library()


def test_header_and_binding_agree(lib):
    header = open(os.path.join(ROOT, "include", "rtgpu.h")).read()
    declared = set(re.findall(r"\b(rtgpu_[a-z0-9_]+)\s*\(", header))
    assert declared == set(abi.EXPORTED_SYMBOLS)
    for name in abi.EXPORTED_SYMBOLS:
        assert hasattr(lib, name), name
    assert lib.rtgpu_abi_version() == abi.ABI_VERSION


def test_struct_sizes_match_c_layout():
    # 8-byte aligned C layout of include/rtgpu.h
    assert C.sizeof(abi.RtgpuCamera) == 8 + 3 * 8 + 12 * 8 + 3 * 8
    assert C.sizeof(abi.RtgpuStats) == 8 * 8
    assert C.sizeof(abi.RtgpuRows) == 12 and C.sizeof(abi.RtgpuOpts) == 20


@pytest.mark.parametrize("vsize", [0, 1, 7, 16, 100, 1080])
@pytest.mark.parametrize("band,count", [(0, 1), (4, 2), (16, 8), (3, 5), (16, 3)])
def test_rows_helpers_partition_the_frame(lib, vsize, band, count):
    seen = []
    for index in range(count):
        r = abi.RtgpuRows(band, index, count)
        n = lib.rtgpu_rows_count(C.byref(r), vsize)
        buf = (C.c_uint32 * max(1, n))()
        assert lib.rtgpu_rows_list(C.byref(r), vsize, buf, n) == n
        rows = list(buf[:n])
        assert rows == sorted(rows)
        b = band or max(vsize, 1)
        assert all((y // b) % count == index for y in rows)
        seen += rows
    assert sorted(seen) == list(range(vsize))


def test_rows_arithmetic_cannot_wrap(lib):
    """band_rows * shard_count is computed in 64 bits: a product that overflows 32 bits is rejected (count 0), a band
    taller than the image is the whole image."""
    assert lib.rtgpu_rows_count(C.byref(abi.RtgpuRows(65536, 0, 65536)), 1080) == 1080  # band clamped to the image: one band, shard 0
    assert lib.rtgpu_rows_count(C.byref(abi.RtgpuRows(65536, 1, 65536)), 1080) == 0
    assert lib.rtgpu_rows_count(C.byref(abi.RtgpuRows(4096, 0, 0x00200000)), 0xFFFFFFFF) == 0  # 2^12 * 2^21 wraps: rejected
    assert lib.rtgpu_rows_count(C.byref(abi.RtgpuRows(0xFFFFFFFF, 0, 1)), 100) == 100
    assert lib.rtgpu_rows_count(C.byref(abi.RtgpuRows(16, 3, 2)), 100) == 0  # shard_index >= shard_count


def test_no_device_is_a_loud_error(lib):
    if lib.rtgpu_device_count() > 0:
        pytest.skip("a GPU is visible")
    flat, camera = load_scene_fixture("three_sphere_scene")
    camera = camera.resized(16, 8)
    cs, cc = flat.as_c(), camera_to_c(camera)
    out = np.zeros((16 * 8, 3))
    st = lib.rtgpu_render(C.byref(cs), C.byref(cc), None, out.ctypes.data, None, None)
    assert st == abi.ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.rtgpu_last_error()
    ctx = C.c_void_p()
    assert lib.rtgpu_context_create(C.byref(cs), 0, C.byref(ctx)) == abi.ERR_NO_DEVICE
    with pytest.raises(abi.RtgpuError):
        camera.render_gpu(flat)


def test_invalid_scenes_are_rejected(lib):
    flat, camera = load_scene_fixture("cover")
    cc = camera_to_c(camera.resized(8, 8))
    out = np.zeros((64, 3))
    cs = flat.as_c()
    cs.abi_version = 99
    assert lib.rtgpu_render(C.byref(cs), C.byref(cc), None, out.ctypes.data, None, None) == abi.ERR_INVALID_ARGUMENT
    bad = flat.shape_material.copy()
    bad[0] = 1000
    cs = flat.as_c()
    cs.shape_material = bad.ctypes.data_as(C.POINTER(C.c_uint32))
    assert lib.rtgpu_render(C.byref(cs), C.byref(cc), None, out.ctypes.data, None, None) == abi.ERR_INVALID_ARGUMENT
    assert b"material" in lib.rtgpu_last_error()
    cs = flat.as_c()
    assert lib.rtgpu_render(C.byref(cs), C.byref(cc), None, None, None, None) == abi.ERR_INVALID_ARGUMENT
    opts = abi.RtgpuOpts(7, 6, 1, 0, 0)
    st = lib.rtgpu_render(C.byref(cs), C.byref(cc), C.byref(opts), out.ctypes.data, None, None)
    assert st in (abi.ERR_INVALID_ARGUMENT, abi.ERR_NO_DEVICE)


def test_product_never_touches_the_oracle():
    """The package (the product) must not import, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "ray_tracer_challenge_rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "rt_oracle" not in text and "librtoracle" not in text and "from oracle" not in text and "import oracle" not in text, f
