"""ctypes view of ``include/rtgpu.h`` — struct layouts, enums and the library loader.

Pure declarations: nothing here computes.  The product library is ``librtgpu.so`` next to this
file (built by ``__graft_entry__.build()`` / :mod:`.build`); loading fails loudly when it is
missing — there is no Python or CPU fallback for the render path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

ABI_VERSION = 1

# rtgpu_status
OK = 0
ERR_INVALID_ARGUMENT = -1
ERR_UNSUPPORTED = -2
ERR_NO_DEVICE = -3
ERR_CUDA = -4
ERR_OUT_OF_MEMORY = -5

# rtgpu_shape_type
SPHERE, PLANE, CUBE, CYLINDER, CONE, TRIANGLE = range(6)
SHAPE_NAMES = ("sphere", "plane", "cube", "cylinder", "cone", "triangle")
# rtgpu_pattern_type
PATTERN_STRIPE, PATTERN_GRADIENT, PATTERN_RING, PATTERN_CHECKER, PATTERN_COMPLEX, PATTERN_TEST = range(6)
# material scalar block
MAT_AMBIENT, MAT_DIFFUSE, MAT_SPECULAR, MAT_SHININESS, MAT_REFLECTIVENESS, MAT_TRANSPARENCY, MAT_REFRACTIVE_INDEX = range(7)
MAT_PARAM_COUNT = 7
# rtgpu_precision
PRECISION_F64 = 0
PRECISION_F32 = 1
# rtgpu_opts.flags
FLAG_WAVEFRONT = 1
FLAG_PERSISTENT = 2

_pd = C.POINTER(C.c_double)
_pu8 = C.POINTER(C.c_uint8)
_pu32 = C.POINTER(C.c_uint32)
_pi32 = C.POINTER(C.c_int32)
_pu64 = C.POINTER(C.c_uint64)


class RtgpuScene(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("n_shapes", C.c_uint32),
        ("shape_type", _pu8),
        ("shape_inv", _pd),
        ("shape_min", _pd),
        ("shape_max", _pd),
        ("shape_closed", _pu8),
        ("shape_triangle", _pi32),
        ("shape_material", _pu32),
        ("shape_eq_class", _pu32),
        ("n_triangles", C.c_uint32),
        ("tri_vertex_1", _pd),
        ("tri_edge_1", _pd),
        ("tri_edge_2", _pd),
        ("tri_normal", _pd),
        ("n_materials", C.c_uint32),
        ("mat_color", _pd),
        ("mat_params", _pd),
        ("mat_casts_shadow", _pu8),
        ("mat_pattern", _pi32),
        ("n_patterns", C.c_uint32),
        ("pat_type", _pu8),
        ("pat_color_a", _pd),
        ("pat_color_b", _pd),
        ("pat_inv", _pd),
        ("pat_child_a", _pi32),
        ("pat_child_b", _pi32),
        ("n_lights", C.c_uint32),
        ("light_position", _pd),
        ("light_intensity", _pd),
    ]


class RtgpuCamera(C.Structure):
    _fields_ = [
        ("hsize", C.c_uint32),
        ("vsize", C.c_uint32),
        ("half_width", C.c_double),
        ("half_height", C.c_double),
        ("pixel_size", C.c_double),
        ("inv", C.c_double * 12),
        ("origin", C.c_double * 3),
    ]


class RtgpuRows(C.Structure):
    _fields_ = [("band_rows", C.c_uint32), ("shard_index", C.c_uint32), ("shard_count", C.c_uint32)]


class RtgpuOpts(C.Structure):
    _fields_ = [
        ("precision", C.c_uint32),
        ("max_depth", C.c_uint32),
        ("n_gpus", C.c_int32),
        ("band_rows", C.c_uint32),
        ("flags", C.c_uint32),
    ]


class RtgpuStats(C.Structure):
    _fields_ = [
        ("rays_primary", C.c_uint64),
        ("rays_shadow", C.c_uint64),
        ("rays_reflect", C.c_uint64),
        ("rays_refract", C.c_uint64),
        ("hit_nodes", C.c_uint64),
        ("pixels", C.c_uint64),
        ("kernel_ms", C.c_double),
        ("total_ms", C.c_double),
    ]

    def as_dict(self) -> dict:
        d = {name: getattr(self, name) for name, _ in self._fields_}
        d["rays"] = self.rays_primary + self.rays_shadow + self.rays_reflect + self.rays_refract
        return d


#: every symbol include/rtgpu.h declares (tests assert the library exports each of them)
EXPORTED_SYMBOLS = (
    "rtgpu_abi_version",
    "rtgpu_last_error",
    "rtgpu_device_count",
    "rtgpu_rows_count",
    "rtgpu_rows_list",
    "rtgpu_render",
    "rtgpu_context_create",
    "rtgpu_context_destroy",
    "rtgpu_context_render_device",
    "rtgpu_context_render",
    "rtgpu_last_family",
    "rtgpu_context_frame_records",
    "rtgpu_context_launch_count",
    "rtgpu_host_alloc",
    "rtgpu_host_free",
    "rtgpu_measure_fma_peak",
    "rtgpu_selftest_arith",
    "rtgpu_debug_probe",
    "rtgpu_debug_color_at",
)

# rtgpu_probe_kind
(PROBE_RAY_FOR_PIXEL, PROBE_INTERSECT, PROBE_LOCAL_NORMAL, PROBE_NORMAL, PROBE_PATTERN, PROBE_LIGHTING, PROBE_IN_SHADOW,
 PROBE_PREPARE, PROBE_QUANTISE, PROBE_COLLECT) = range(10)

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
LIBRARY_PATH = os.path.join(PACKAGE_DIR, "librtgpu.so")

_lib: Optional[C.CDLL] = None


class RtgpuError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"rtgpu status {status}: {message}")
        self.status = status


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Load ``librtgpu.so`` and declare the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("RTGPU_LIBRARY") or LIBRARY_PATH  # RTGPU_LIBRARY: A/B builds of the same ABI
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} is missing: the CUDA library has not been built (run `python -c 'import "
            "__graft_entry__ as g; g.build()'`).  There is no CPU fallback for the render path."
        )
    lib = C.CDLL(p)
    lib.rtgpu_abi_version.restype = C.c_uint32
    lib.rtgpu_abi_version.argtypes = []
    lib.rtgpu_last_error.restype = C.c_char_p
    lib.rtgpu_last_error.argtypes = []
    lib.rtgpu_device_count.restype = C.c_int
    lib.rtgpu_device_count.argtypes = []
    lib.rtgpu_rows_count.restype = C.c_uint32
    lib.rtgpu_rows_count.argtypes = [C.POINTER(RtgpuRows), C.c_uint32]
    lib.rtgpu_rows_list.restype = C.c_uint32
    lib.rtgpu_rows_list.argtypes = [C.POINTER(RtgpuRows), C.c_uint32, _pu32, C.c_uint32]
    lib.rtgpu_render.restype = C.c_int
    lib.rtgpu_render.argtypes = [
        C.POINTER(RtgpuScene),
        C.POINTER(RtgpuCamera),
        C.POINTER(RtgpuOpts),
        C.c_void_p,
        C.c_void_p,
        C.POINTER(RtgpuStats),
    ]
    lib.rtgpu_context_create.restype = C.c_int
    lib.rtgpu_context_create.argtypes = [C.POINTER(RtgpuScene), C.c_int, C.POINTER(C.c_void_p)]
    lib.rtgpu_context_destroy.restype = None
    lib.rtgpu_context_destroy.argtypes = [C.c_void_p]
    lib.rtgpu_context_render_device.restype = C.c_int
    lib.rtgpu_context_render_device.argtypes = [
        C.c_void_p,
        C.POINTER(RtgpuCamera),
        C.POINTER(RtgpuOpts),
        C.POINTER(RtgpuRows),
        C.c_void_p,
        C.c_void_p,
        C.c_void_p,
        C.c_void_p,
    ]
    lib.rtgpu_context_render.restype = C.c_int
    lib.rtgpu_context_render.argtypes = [
        C.c_void_p,
        C.POINTER(RtgpuCamera),
        C.POINTER(RtgpuOpts),
        C.POINTER(RtgpuRows),
        C.c_void_p,
        C.c_void_p,
        C.POINTER(RtgpuStats),
    ]
    lib.rtgpu_measure_fma_peak.restype = C.c_int
    lib.rtgpu_measure_fma_peak.argtypes = [C.c_int, C.c_uint32, _pd, _pd]
    lib.rtgpu_context_frame_records.restype = C.c_int
    lib.rtgpu_context_frame_records.argtypes = [C.c_void_p, _pu64]
    lib.rtgpu_context_launch_count.restype = C.c_uint64
    lib.rtgpu_context_launch_count.argtypes = [C.c_void_p]
    lib.rtgpu_last_family.restype = C.c_int
    lib.rtgpu_last_family.argtypes = []
    lib.rtgpu_host_alloc.restype = C.c_void_p
    lib.rtgpu_host_alloc.argtypes = [C.c_size_t]
    lib.rtgpu_host_free.restype = None
    lib.rtgpu_host_free.argtypes = [C.c_void_p]
    lib.rtgpu_selftest_arith.restype = C.c_int
    lib.rtgpu_selftest_arith.argtypes = [C.c_int, _pd, _pd, C.c_size_t, _pu64, _pu64, _pu64, _pu64]
    lib.rtgpu_debug_probe.restype = C.c_int
    lib.rtgpu_debug_probe.argtypes = [C.c_void_p, C.POINTER(RtgpuCamera), C.c_uint32, _pd, C.c_size_t, _pd, C.c_size_t]
    lib.rtgpu_debug_color_at.restype = C.c_int
    lib.rtgpu_debug_color_at.argtypes = [C.c_void_p, _pd, _pd, C.POINTER(RtgpuOpts), _pd]
    if lib.rtgpu_abi_version() != ABI_VERSION:
        raise RuntimeError(f"{p}: ABI version {lib.rtgpu_abi_version()} != {ABI_VERSION}")
    if path is None:
        _lib = lib
    return lib


def check(lib: C.CDLL, status: int) -> None:
    if status != OK:
        msg = lib.rtgpu_last_error()
        raise RtgpuError(status, msg.decode("utf-8", "replace") if msg else "")
