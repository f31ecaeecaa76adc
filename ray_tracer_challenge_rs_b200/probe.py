"""Known-answer probes of the device code (``rtgpu_debug_probe`` / ``rtgpu_debug_color_at``, include/rtgpu.h).

:class:`DeviceProbe` has the method names of the CPU oracle's micro entry points, so the replay of the reference's
unit tests (tests/test_oracle_kat.py) runs unchanged against the GPU: each call evaluates, on
one device thread, the same device functions the render kernels are built from.  Test infrastructure — nothing on
the render path imports this module.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Optional, Sequence

import numpy as np

from . import abi
from .flatten import FlatScene, camera_to_c
from .render import _opts, render_gpu


class ProbeUnsupported(NotImplementedError):
    """The reference entry point has no device-side equivalent that could be probed in isolation."""


class DeviceProbe:
    def __init__(self, scene: FlatScene, device: int = 0, family: str = "persistent"):
        self.scene = scene
        self.family = family
        self._lib = abi.load_library()
        self._cscene = scene.as_c()
        self._ctx = C.c_void_p()
        abi.check(self._lib, self._lib.rtgpu_context_create(C.byref(self._cscene), device, C.byref(self._ctx)))

    def close(self):
        if self._ctx:
            self._lib.rtgpu_context_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _probe(self, kind: int, values: Sequence[float], n_out: int, camera=None) -> np.ndarray:
        vin = np.ascontiguousarray(values, dtype=np.float64)
        out = np.zeros(n_out, np.float64)
        cam = camera_to_c(camera) if camera is not None else None
        abi.check(self._lib, self._lib.rtgpu_debug_probe(
            self._ctx, C.byref(cam) if cam is not None else None, kind, vin.ctypes.data_as(C.POINTER(C.c_double)), vin.size,
            out.ctypes.data_as(C.POINTER(C.c_double)), out.size))
        return out

    # ---- the oracle's micro entry points, on the device ------------------------------------------
    def ray_for_pixel(self, camera, px: int, py: int):
        o = self._probe(abi.PROBE_RAY_FOR_PIXEL, [px, py], 6, camera)
        return tuple(o[:3].tolist()), tuple(o[3:].tolist())

    def intersect_shape(self, shape: int, origin, direction):
        o = self._probe(abi.PROBE_INTERSECT, [shape, *origin, *direction], 5)
        return o[1:1 + int(o[0])].tolist()

    def collect_intersections(self, origin, direction):
        cap = 4 * max(1, self.scene.n_shapes)
        o = self._probe(abi.PROBE_COLLECT, [*origin, *direction], 1 + 2 * cap)
        n = int(o[0])
        pushed = [(float(o[1 + 2 * k]), int(o[2 + 2 * k])) for k in range(n)]
        return sorted(pushed, key=lambda x: x[0])  # stable, like world.rs:34 (NaN-free inputs in the tests)

    def normal_at(self, shape: int, point):
        return tuple(self._probe(abi.PROBE_NORMAL, [shape, *point], 3).tolist())

    def local_normal_at(self, shape: int, point):
        return tuple(self._probe(abi.PROBE_LOCAL_NORMAL, [shape, *point], 3).tolist())

    def pattern_at_shape(self, pattern: int, shape: int, point):
        return tuple(self._probe(abi.PROBE_PATTERN, [pattern, shape, *point], 3).tolist())

    def lighting(self, material: int, shape: int, light_position, light_intensity, point, eye, normal, in_shadow: bool):
        v = [material, shape, *light_position, *light_intensity, *point, *eye, *normal, 1.0 if in_shadow else 0.0]
        return tuple(self._probe(abi.PROBE_LIGHTING, v, 3).tolist())

    def is_in_shadow(self, light: int, point) -> bool:
        return bool(self._probe(abi.PROBE_IN_SHADOW, [light, *point], 1)[0])

    def prepare_computations(self, origin, direction, k: int = -1, xs=None):
        if xs is None:
            xs = self.collect_intersections(origin, direction)
        else:
            xs = sorted(((float(t), int(s)) for t, s in xs), key=lambda x: x[0])  # Intersections::new sorts (stable)
        flat = [v for x in xs for v in x]
        o = self._probe(abi.PROBE_PREPARE, [*origin, *direction, k, len(xs), *flat], 25)
        if not o[0]:
            return None
        return SimpleNamespace(
            distance=float(o[1]), shape=int(o[2]), is_inside=bool(o[3]), point=o[4:7].tolist(), over_point=o[7:10].tolist(),
            under_point=o[10:13].tolist(), camera_direction=o[13:16].tolist(), normal=o[16:19].tolist(),
            reflect_direction=o[19:22].tolist(), refractive_index_1=float(o[22]), refractive_index_2=float(o[23]), schlick=float(o[24]))

    def color_at(self, origin, direction, remaining: int = 6):
        o = (C.c_double * 3)(*[float(v) for v in origin])
        d = (C.c_double * 3)(*[float(v) for v in direction])
        out = (C.c_double * 3)()
        opts = _opts("f64", remaining, family=self.family)
        abi.check(self._lib, self._lib.rtgpu_debug_color_at(self._ctx, o, d, C.byref(opts), out))
        return tuple(out)

    def shade_entry(self, origin, direction, k: int, remaining: int, which: str, xs=None):
        """World::shade_hit of the k-th entry of `xs`.  On the device a node is only ever shaded as part of
        World::color_at, so this is color_at of the same ray — valid when xs[k] IS the ray's hit, which is checked
        with the device's own hit rule; reflected_color / refracted_color alone are not device entry points."""
        if which != "shade_hit":
            raise ProbeUnsupported(f"{which}_color is not evaluated in isolation on the device")
        want = sorted(((float(t), int(s)) for t, s in xs), key=lambda x: x[0])[k] if xs is not None else None
        hit = self.prepare_computations(origin, direction, k=-1)
        if hit is None or (want is not None and (hit.distance, hit.shape) != want):
            raise ProbeUnsupported("the hand-built intersection is not the ray's own hit")
        return self.color_at(origin, direction, remaining)

    def quantise(self, v: float) -> int:
        return int(self._probe(abi.PROBE_QUANTISE, [v], 1)[0])

    def render(self, camera, max_depth: int = 6, threads: int = 0, rows=None, want_rgb: bool = True, want_rgb8: bool = True):
        canvas, stats = render_gpu(camera, self.scene, max_depth=max_depth, return_stats=True, family=self.family)
        return canvas.pixels, canvas.to_rgb8().reshape(-1, 3), stats
