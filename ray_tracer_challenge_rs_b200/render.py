"""``Camera.render_gpu`` and the resident-scene renderer: thin callers of the C ABI (include/rtgpu.h).

Nothing here computes pixels.  If ``librtgpu.so`` is missing, or there is no CUDA device, every
entry point raises — the gpu rendering mode has no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple, Union

import numpy as np

from . import abi
from .flatten import FlatScene, camera_to_c, flatten_world
from .scene import Camera, Canvas, World

SceneLike = Union[World, FlatScene]


def _flat(scene: SceneLike) -> FlatScene:
    return scene if isinstance(scene, FlatScene) else flatten_world(scene)


def _opts(precision: str = "f64", max_depth: int = World.MAX_REFLECTION_ITERATIONS, n_gpus: int = 1, band_rows: int = 16,
          family: Optional[str] = None) -> abi.RtgpuOpts:
    """family: None (auto: RTGPU_FAMILY, else the library times both and keeps the faster), "wavefront" or
    "persistent" (include/rtgpu.h RTGPU_FLAG_*)."""
    prec = {"f64": abi.PRECISION_F64, "f32": abi.PRECISION_F32}[precision]
    flags = {None: 0, "wavefront": abi.FLAG_WAVEFRONT, "persistent": abi.FLAG_PERSISTENT}[family]
    return abi.RtgpuOpts(prec, int(max_depth), int(n_gpus), int(band_rows), flags)


FAMILY_NAMES = ("persistent", "wavefront")


def _check_out(arr, dtype, size: int, name: str) -> None:
    """The library writes `size` elements of `dtype` through the raw pointer: anything else would overflow the
    caller's buffer or be filled with reinterpreted values."""
    if arr is None:
        return
    if not isinstance(arr, np.ndarray) or arr.dtype != np.dtype(dtype):
        raise ValueError(f"{name} must be a numpy array of dtype {np.dtype(dtype).name} (got {getattr(arr, 'dtype', type(arr))})")
    if not arr.flags["C_CONTIGUOUS"] or not arr.flags["WRITEABLE"]:
        raise ValueError(f"{name} must be C-contiguous and writeable")
    if arr.size != size:
        raise ValueError(f"{name} must hold exactly {size} elements (full frame x 3), got {arr.size}")


def last_family() -> str:
    """The kernel family this thread's most recent render ran (``rtgpu_last_family``)."""
    return FAMILY_NAMES[int(abi.load_library().rtgpu_last_family())]


def device_count() -> int:
    return int(abi.load_library().rtgpu_device_count())


def render_gpu(
    camera: Camera,
    world: SceneLike,
    precision: str = "f64",
    max_depth: int = World.MAX_REFLECTION_ITERATIONS,
    n_gpus: int = 1,
    band_rows: int = 16,
    want_rgb: bool = True,
    want_rgb8: bool = True,
    return_stats: bool = False,
    family: Optional[str] = None,
):
    """One-shot render through ``rtgpu_render`` (the call a ``RenderingMode::Gpu`` arm makes).

    Returns a :class:`Canvas` whose ``pixels`` are the linear colours (f64; f32 in fast mode) and
    whose ``to_rgb8()`` are the bytes the device quantised (canvas.rs:117-123)."""
    lib = abi.load_library()
    flat = _flat(world)
    cscene = flat.as_c()
    ccam = camera_to_c(camera)
    n = camera.horizontal_size * camera.vertical_size
    dtype = np.float64 if precision == "f64" else np.float32
    rgb = np.zeros((n, 3), dtype) if want_rgb else None
    rgb8 = np.zeros((n, 3), np.uint8) if want_rgb8 else None
    stats = abi.RtgpuStats()
    opts = _opts(precision, max_depth, n_gpus, band_rows, family)
    st = lib.rtgpu_render(
        C.byref(cscene),
        C.byref(ccam),
        C.byref(opts),
        rgb.ctypes.data if rgb is not None else None,
        rgb8.ctypes.data if rgb8 is not None else None,
        C.byref(stats),
    )
    abi.check(lib, st)
    canvas = Canvas(camera.horizontal_size, camera.vertical_size, rgb if rgb is not None else np.zeros((n, 3)), rgb8)
    if return_stats:
        return canvas, dict(stats.as_dict(), family=last_family())
    return canvas


class Renderer:
    """A scene resident on one device (``rtgpu_context``): what a caller rendering many frames, or
    one rank of a row-band-sharded job, uses."""

    def __init__(self, world: SceneLike, device: int = 0):
        self._lib = abi.load_library()
        self.flat = _flat(world)
        self._cscene = self.flat.as_c()
        self._ctx = C.c_void_p()
        self.device = device
        abi.check(self._lib, self._lib.rtgpu_context_create(C.byref(self._cscene), device, C.byref(self._ctx)))

    def close(self) -> None:
        if self._ctx:
            self._lib.rtgpu_context_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def rows_count(self, camera: Camera, rows: Optional[Tuple[int, int, int]]) -> int:
        r = abi.RtgpuRows(*(rows or (0, 0, 1)))
        return int(self._lib.rtgpu_rows_count(C.byref(r), camera.vertical_size))

    def rows_list(self, camera: Camera, rows: Optional[Tuple[int, int, int]]):
        """Image rows of the selection, in the order the compact output of a shard holds them (``rtgpu_rows_list``)."""
        r = abi.RtgpuRows(*(rows or (0, 0, 1)))
        n = self.rows_count(camera, rows)
        out = (C.c_uint32 * max(n, 1))()
        self._lib.rtgpu_rows_list(C.byref(r), camera.vertical_size, out, n)
        return [int(out[i]) for i in range(n)]

    def render(
        self,
        camera: Camera,
        precision: str = "f64",
        max_depth: int = World.MAX_REFLECTION_ITERATIONS,
        rows: Optional[Tuple[int, int, int]] = None,
        out_rgb: Optional[np.ndarray] = None,
        out_rgb8: Optional[np.ndarray] = None,
        want_rgb: bool = True,
        want_rgb8: bool = True,
        family: Optional[str] = None,
    ):
        """Host buffers in, host buffers out (``rtgpu_context_render``): full-frame arrays; only the
        rows that ``rows = (band_rows, shard_index, shard_count)`` selects are written."""
        n = camera.horizontal_size * camera.vertical_size
        dtype = np.float64 if precision == "f64" else np.float32
        if out_rgb is None and want_rgb:
            out_rgb = np.zeros((n, 3), dtype)
        if out_rgb8 is None and want_rgb8:
            out_rgb8 = np.zeros((n, 3), np.uint8)
        _check_out(out_rgb, dtype, n * 3, "out_rgb")
        _check_out(out_rgb8, np.uint8, n * 3, "out_rgb8")
        ccam = camera_to_c(camera)
        opts = _opts(precision, max_depth, family=family)
        r = abi.RtgpuRows(*(rows or (0, 0, 1)))
        stats = abi.RtgpuStats()
        st = self._lib.rtgpu_context_render(
            self._ctx,
            C.byref(ccam),
            C.byref(opts),
            C.byref(r),
            out_rgb.ctypes.data if out_rgb is not None else None,
            out_rgb8.ctypes.data if out_rgb8 is not None else None,
            C.byref(stats),
        )
        abi.check(self._lib, st)
        return out_rgb, out_rgb8, dict(stats.as_dict(), family=last_family())

    def launch_count(self) -> int:
        """Render kernels launched for this context so far (``rtgpu_context_launch_count``)."""
        return int(self._lib.rtgpu_context_launch_count(self._ctx))

    def frame_records(self) -> dict:
        """What the wavefront family moved through HBM for the most recent host-buffer frame (``rtgpu_context_frame_records``)."""
        out = (C.c_uint64 * 4)()
        abi.check(self._lib, self._lib.rtgpu_context_frame_records(self._ctx, out))
        return {"queued_rays": int(out[0]), "node_records": int(out[1]), "ray_record_bytes": int(out[2]), "node_record_bytes": int(out[3])}

    def render_device(
        self,
        camera: Camera,
        d_out_rgb: int,
        d_out_rgb8: int,
        d_counters: int,
        stream: int,
        precision: str = "f64",
        max_depth: int = World.MAX_REFLECTION_ITERATIONS,
        rows: Optional[Tuple[int, int, int]] = None,
        family: Optional[str] = None,
    ) -> None:
        """Asynchronous launch on device pointers and a caller-owned stream
        (``rtgpu_context_render_device``); outputs are compact over the selected rows."""
        ccam = camera_to_c(camera)
        opts = _opts(precision, max_depth, family=family)
        r = abi.RtgpuRows(*(rows or (0, 0, 1)))
        st = self._lib.rtgpu_context_render_device(
            self._ctx, C.byref(ccam), C.byref(opts), C.byref(r), d_out_rgb or None, d_out_rgb8 or None, d_counters or None, stream or None
        )
        abi.check(self._lib, st)


class PinnedArray:
    """A numpy array over ``rtgpu_host_alloc`` memory (page-locked, device-mapped): host-buffer renders
    into it are zero-copy."""

    def __init__(self, shape, dtype=np.float64):
        self._lib = abi.load_library()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._ptr = self._lib.rtgpu_host_alloc(self.nbytes)
        if not self._ptr:
            raise abi.RtgpuError(abi.ERR_OUT_OF_MEMORY, (self._lib.rtgpu_last_error() or b"").decode())
        buf = (C.c_char * self.nbytes).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def close(self):
        if self._ptr:
            self.array = None
            self._lib.rtgpu_host_free(self._ptr)
            self._ptr = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def measure_fma_peak(precision: str = "f64", device: int = 0) -> Tuple[float, float]:
    """(TFLOP/s, ms) of a dependent-free FMA chain on every SM — the FP-pipe roofline denominator."""
    lib = abi.load_library()
    tf, ms = C.c_double(), C.c_double()
    prec = {"f64": abi.PRECISION_F64, "f32": abi.PRECISION_F32}[precision]
    abi.check(lib, lib.rtgpu_measure_fma_peak(device, prec, C.byref(tf), C.byref(ms)))
    return tf.value, ms.value
