"""ctypes binding of the CPU oracle (oracle/rt_oracle.c).  TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs import this module; nothing under ``ray_tracer_challenge_rs_b200/`` does.
It borrows the *struct declarations* of include/rtgpu.h from the package (types only).
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

from ray_tracer_challenge_rs_b200 import abi
from ray_tracer_challenge_rs_b200.flatten import FlatScene, camera_to_c

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
# RTORACLE_LIBRARY: another build of the same source (bench.py times the -march=native variant through it)
LIBRARY_PATH = os.environ.get("RTORACLE_LIBRARY") or os.path.join(ORACLE_DIR, "librtoracle.so")
NATIVE_LIBRARY_PATH = os.path.join(ORACLE_DIR, "_native", "librtoracle_native.so")

_pd = C.POINTER(C.c_double)


class ComputedHit(C.Structure):
    _fields_ = [
        ("distance", C.c_double),
        ("shape", C.c_uint32),
        ("is_inside", C.c_uint32),
        ("point", C.c_double * 3),
        ("over_point", C.c_double * 3),
        ("under_point", C.c_double * 3),
        ("camera_direction", C.c_double * 3),
        ("normal", C.c_double * 3),
        ("reflect_direction", C.c_double * 3),
        ("refractive_index_1", C.c_double),
        ("refractive_index_2", C.c_double),
        ("schlick", C.c_double),
    ]


def build(force: bool = False) -> str:
    """Compile librtoracle.so with the committed Makefile (gcc from PATH, not $CC)."""
    src = [os.path.join(ORACLE_DIR, n) for n in ("rt_oracle.c", "rt_oracle.h")]
    src.append(os.path.join(ORACLE_DIR, "..", "include", "rtgpu.h"))
    if (
        not force
        and os.path.exists(LIBRARY_PATH)
        and all(os.path.getmtime(LIBRARY_PATH) >= os.path.getmtime(s) for s in src)
    ):
        return LIBRARY_PATH
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CFLAGS", None)
    make = shutil.which("make")
    if make is None:
        raise RuntimeError("make not found")
    subprocess.run([make, "-C", ORACLE_DIR, "-B" if force else "-s"], check=True, env=env, capture_output=True)
    return LIBRARY_PATH


def build_native() -> str:
    """Compile the -march=native variant for THIS host (always rebuilt: the flags depend on the CPU it runs on)."""
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CFLAGS", None)
    subprocess.run([shutil.which("make") or "make", "-C", ORACLE_DIR, "-B", "native"], check=True, env=env, capture_output=True)
    return NATIVE_LIBRARY_PATH


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBRARY_PATH):
        build()
    L = C.CDLL(LIBRARY_PATH)
    S, Cam, Rows, Stats = abi.RtgpuScene, abi.RtgpuCamera, abi.RtgpuRows, abi.RtgpuStats
    L.rto_render.restype = C.c_int
    L.rto_render.argtypes = [C.POINTER(S), C.POINTER(Cam), C.c_uint32, C.POINTER(Rows), C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Stats)]
    L.rto_render_pixels.restype = C.c_int
    L.rto_render_pixels.argtypes = [C.POINTER(S), C.POINTER(Cam), C.c_uint32, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Stats)]
    L.rto_max_threads.restype = C.c_int
    L.rto_ray_for_pixel.restype = None
    L.rto_ray_for_pixel.argtypes = [C.POINTER(Cam), C.c_uint32, C.c_uint32, _pd]
    L.rto_intersect_shape.restype = C.c_int
    L.rto_intersect_shape.argtypes = [C.POINTER(S), C.c_uint32, _pd, _pd]
    L.rto_collect_intersections.restype = C.c_size_t
    L.rto_collect_intersections.argtypes = [C.POINTER(S), _pd, _pd, C.POINTER(C.c_uint32), C.c_size_t]
    L.rto_normal_at.restype = None
    L.rto_normal_at.argtypes = [C.POINTER(S), C.c_uint32, _pd, _pd]
    L.rto_local_normal_at.restype = None
    L.rto_local_normal_at.argtypes = [C.POINTER(S), C.c_uint32, _pd, _pd]
    L.rto_pattern_at_shape.restype = None
    L.rto_pattern_at_shape.argtypes = [C.POINTER(S), C.c_uint32, C.c_uint32, _pd, _pd]
    L.rto_lighting.restype = None
    L.rto_lighting.argtypes = [C.POINTER(S), C.c_uint32, C.c_uint32, _pd, _pd, _pd, _pd, _pd, C.c_int, _pd]
    L.rto_is_in_shadow.restype = C.c_int
    L.rto_is_in_shadow.argtypes = [C.POINTER(S), C.c_uint32, _pd]
    L.rto_prepare_computations.restype = C.c_int
    L.rto_prepare_computations.argtypes = [C.POINTER(S), _pd, _pd, C.POINTER(C.c_uint32), C.c_size_t, C.c_int, C.POINTER(ComputedHit)]
    L.rto_color_at.restype = None
    L.rto_color_at.argtypes = [C.POINTER(S), _pd, C.c_uint32, _pd]
    L.rto_shade_entry.restype = C.c_int
    L.rto_shade_entry.argtypes = [C.POINTER(S), _pd, _pd, C.POINTER(C.c_uint32), C.c_size_t, C.c_int, C.c_uint32, C.c_int, _pd]
    L.rto_quantise.restype = C.c_uint8
    L.rto_quantise.argtypes = [C.c_double]
    _lib = L
    return L


def _vec(values: Sequence[float]):
    return (C.c_double * len(values))(*[float(v) for v in values])


def _ray(origin, direction):
    return _vec(list(origin) + list(direction))


class Oracle:
    """The reference's World + Camera entry points, evaluated by the C restatement."""

    def __init__(self, scene: FlatScene):
        self.scene = scene
        self._c = scene.as_c()
        self._L = lib()

    # ---- Camera::render / render_parallel ------------------------------------------------
    def render(self, camera, max_depth: int = 6, threads: int = 0, rows: Optional[Tuple[int, int, int]] = None,
               want_rgb: bool = True, want_rgb8: bool = True):
        cam = camera_to_c(camera) if not isinstance(camera, abi.RtgpuCamera) else camera
        n = cam.hsize * cam.vsize
        rgb = np.zeros((n, 3), np.float64) if want_rgb else None
        rgb8 = np.zeros((n, 3), np.uint8) if want_rgb8 else None
        stats = abi.RtgpuStats()
        r = abi.RtgpuRows(*rows) if rows else None
        st = self._L.rto_render(
            C.byref(self._c), C.byref(cam), max_depth, C.byref(r) if r else None, threads,
            rgb.ctypes.data if rgb is not None else None, rgb8.ctypes.data if rgb8 is not None else None,
            C.byref(stats))
        if st != 0:
            raise RuntimeError(f"rto_render failed: {st}")
        return rgb, rgb8, stats.as_dict()

    def render_pixels(self, camera, pixels: np.ndarray, max_depth: int = 6, threads: int = 0):
        cam = camera_to_c(camera) if not isinstance(camera, abi.RtgpuCamera) else camera
        px = np.ascontiguousarray(pixels, dtype=np.uint64)
        rgb = np.zeros((px.size, 3), np.float64)
        rgb8 = np.zeros((px.size, 3), np.uint8)
        stats = abi.RtgpuStats()
        st = self._L.rto_render_pixels(C.byref(self._c), C.byref(cam), max_depth, px.ctypes.data, px.size, threads,
                                       rgb.ctypes.data, rgb8.ctypes.data, C.byref(stats))
        if st != 0:
            raise RuntimeError(f"rto_render_pixels failed: {st}")
        return rgb, rgb8, stats.as_dict()

    # ---- micro entry points ----------------------------------------------------------------
    def ray_for_pixel(self, camera, px: int, py: int):
        cam = camera_to_c(camera)
        out = (C.c_double * 6)()
        self._L.rto_ray_for_pixel(C.byref(cam), px, py, out)
        return tuple(out[:3]), tuple(out[3:])

    def intersect_shape(self, shape: int, origin, direction):
        out = (C.c_double * 4)()
        n = self._L.rto_intersect_shape(C.byref(self._c), shape, _ray(origin, direction), out)
        return [out[i] for i in range(n)]

    def collect_intersections(self, origin, direction):
        cap = 4 * max(1, self.scene.n_shapes)
        t = (C.c_double * cap)()
        sh = (C.c_uint32 * cap)()
        n = self._L.rto_collect_intersections(C.byref(self._c), _ray(origin, direction), t, sh, cap)
        return [(t[i], sh[i]) for i in range(n)]

    def normal_at(self, shape: int, point):
        out = (C.c_double * 3)()
        self._L.rto_normal_at(C.byref(self._c), shape, _vec(point), out)
        return tuple(out)

    def local_normal_at(self, shape: int, point):
        out = (C.c_double * 3)()
        self._L.rto_local_normal_at(C.byref(self._c), shape, _vec(point), out)
        return tuple(out)

    def pattern_at_shape(self, pattern: int, shape: int, point):
        out = (C.c_double * 3)()
        self._L.rto_pattern_at_shape(C.byref(self._c), pattern, shape, _vec(point), out)
        return tuple(out)

    def lighting(self, material: int, shape: int, light_position, light_intensity, point, eye, normal, in_shadow: bool):
        out = (C.c_double * 3)()
        self._L.rto_lighting(C.byref(self._c), material, shape, _vec(light_position), _vec(light_intensity),
                             _vec(point), _vec(eye), _vec(normal), 1 if in_shadow else 0, out)
        return tuple(out)

    def is_in_shadow(self, light: int, point) -> bool:
        return bool(self._L.rto_is_in_shadow(C.byref(self._c), light, _vec(point)))

    @staticmethod
    def _list(xs):
        """xs: [(distance, shape index), ...] hand-built Intersections, or None = the world's own."""
        if xs is None:
            return None, None, 0
        t = (C.c_double * len(xs))(*[float(x[0]) for x in xs])
        sh = (C.c_uint32 * len(xs))(*[int(x[1]) for x in xs])
        return t, sh, len(xs)

    def prepare_computations(self, origin, direction, k: int = -1, xs=None) -> Optional[ComputedHit]:
        h = ComputedHit()
        t, sh, n = self._list(xs)
        ok = self._L.rto_prepare_computations(C.byref(self._c), _ray(origin, direction), t, sh, n, k, C.byref(h))
        return h if ok else None

    def color_at(self, origin, direction, remaining: int = 6):
        out = (C.c_double * 3)()
        self._L.rto_color_at(C.byref(self._c), _ray(origin, direction), remaining, out)
        return tuple(out)

    def shade_entry(self, origin, direction, k: int, remaining: int, which: str, xs=None):
        out = (C.c_double * 3)()
        code = {"shade_hit": 0, "reflected": 1, "refracted": 2}[which]
        t, sh, n = self._list(xs)
        ok = self._L.rto_shade_entry(C.byref(self._c), _ray(origin, direction), t, sh, n, k, remaining, code, out)
        return tuple(out) if ok else None

    def quantise(self, v: float) -> int:
        return int(self._L.rto_quantise(v))


def max_threads() -> int:
    return int(lib().rto_max_threads())
