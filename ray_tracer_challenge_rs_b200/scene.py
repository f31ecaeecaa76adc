"""Host-side mirror of the reference's scene types — the caller side of the drop-in boundary.

Same names, fields, defaults and construction rules as the reference (paths under
``ray-tracer/src/``):

  * ``Material``                      -> composites/material.rs:9-20, 157-161
  * ``Light``                         -> primitives/light.rs:6-9
  * ``Sphere/Plane/Cube/Cylinder/Cone/Triangle`` (each stores only ``transformation_inverse``)
                                      -> shapes/sphere.rs:8-19, plane.rs, cube.rs:9-20, cylinder.rs:9-32,
                                         cone.rs:9-32, triangle.rs:9-35
  * ``StripePattern/GradientPattern/RingPattern/CheckerPattern/ComplexPattern/TestPattern``
                                      -> patterns/*.rs
  * ``World``                         -> composites/world.rs:9-22, 160-169 (default world)
  * ``Camera``                        -> composites/camera.rs:10-49, 114-127
  * ``Canvas``                        -> composites/canvas.rs:13-17, 53-55, 75-97, 117-137

These objects only *describe* a scene.  The per-pixel work of ``Camera::render`` /
``render_parallel`` (camera.rs:79-112) is replaced by :meth:`Camera.render_gpu`, which flattens the
World (see :mod:`.flatten`) and calls the C ABI in ``include/rtgpu.h``.  There is deliberately no
CPU render method here: gpu mode has no fallback.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import primitives as P

Color = Tuple[float, float, float]
BLACK: Color = (0.0, 0.0, 0.0)
WHITE: Color = (1.0, 1.0, 1.0)


def _f3(v: Sequence[float]) -> Tuple[float, float, float]:
    return (float(v[0]), float(v[1]), float(v[2]))


# --------------------------------------------------------------------------------------------
# patterns


class Pattern:
    """patterns/pattern.rs:6-15.  Stores only the inverse, like every reference pattern."""

    TYPE = -1

    def __init__(self) -> None:
        self.transformation_inverse: P.Matrix = P.identity()

    def set_transformation(self, transformation: P.Matrix) -> None:
        self.transformation_inverse = P.inverse(transformation)

    def set_transformation_inverse(self, transformation: P.Matrix) -> None:
        self.transformation_inverse = P.matrix(transformation)

    def transformation(self) -> P.Matrix:
        return P.inverse(self.transformation_inverse)

    def value_key(self):
        raise NotImplementedError


class _TwoColorPattern(Pattern):
    def __init__(self, color_a: Sequence[float], color_b: Sequence[float]) -> None:
        super().__init__()
        self.color_a: Color = _f3(color_a)
        self.color_b: Color = _f3(color_b)

    def value_key(self):
        return (self.TYPE, self.color_a, self.color_b, _mkey(self.transformation_inverse))


class StripePattern(_TwoColorPattern):
    """patterns/stripe_pattern.rs:7-31"""

    TYPE = 0


class GradientPattern(_TwoColorPattern):
    """patterns/gradient_pattern.rs:7-31"""

    TYPE = 1


class RingPattern(_TwoColorPattern):
    """patterns/ring_pattern.rs:8-32"""

    TYPE = 2


class CheckerPattern(_TwoColorPattern):
    """patterns/checker_pattern.rs:7-31"""

    TYPE = 3


class ComplexPattern(Pattern):
    """patterns/complex_pattern.rs:8-33"""

    TYPE = 4

    def __init__(self, pattern_a: Pattern, pattern_b: Pattern) -> None:
        super().__init__()
        self.pattern_a = pattern_a
        self.pattern_b = pattern_b

    def value_key(self):
        return (self.TYPE, self.pattern_a.value_key(), self.pattern_b.value_key(), _mkey(self.transformation_inverse))


class TestPattern(Pattern):
    """patterns/pattern.rs:28-66 (crate-private in the reference; colour = pattern-space point)"""

    __test__ = False  # not a pytest class
    TYPE = 5

    def value_key(self):
        return (self.TYPE, _mkey(self.transformation_inverse))


def _mkey(m: P.Matrix):
    return tuple(tuple(row) for row in m)


# --------------------------------------------------------------------------------------------
# material, light


@dataclass
class Material:
    """composites/material.rs:9-20; defaults from material.rs:157-161."""

    color: Color = WHITE
    pattern: Optional[Pattern] = None
    ambient: float = 0.1
    diffuse: float = 0.9
    specular: float = 0.9
    shininess: float = 200.0
    reflectiveness: float = 0.0
    transparency: float = 0.0
    refractive_index: float = 1.0
    casts_shadow: bool = True

    DEFAULT_REFRACTIVE_INDEX = 1.0  # material.rs:24

    @staticmethod
    def glass() -> "Material":
        """material.rs:148-154"""
        return Material(transparency=1.0, refractive_index=1.5)

    def clone(self) -> "Material":
        return Material(
            self.color,
            self.pattern,
            self.ambient,
            self.diffuse,
            self.specular,
            self.shininess,
            self.reflectiveness,
            self.transparency,
            self.refractive_index,
            self.casts_shadow,
        )

    def value_key(self):
        return (
            _f3(self.color),
            None if self.pattern is None else self.pattern.value_key(),
            float(self.ambient),
            float(self.diffuse),
            float(self.specular),
            float(self.shininess),
            float(self.reflectiveness),
            float(self.transparency),
            float(self.refractive_index),
            bool(self.casts_shadow),
        )


@dataclass
class Light:
    """primitives/light.rs:6-9; default = light.rs:44-48 (-10, 10, -10), white."""

    position: Tuple[float, float, float] = (-10.0, 10.0, -10.0)
    intensity: Color = WHITE


# --------------------------------------------------------------------------------------------
# shapes

SPHERE, PLANE, CUBE, CYLINDER, CONE, TRIANGLE = range(6)


class Shape:
    """shapes/shape.rs:6-32.  ``transformation`` arguments are FORWARD transforms; the shape keeps
    ``transformation.inverse()`` (e.g. sphere.rs:14-19)."""

    TYPE = -1

    def __init__(self, material: Optional[Material] = None, transformation: Optional[P.Matrix] = None) -> None:
        self.material: Material = material if material is not None else Material()
        self.transformation_inverse: P.Matrix = (
            P.identity() if transformation is None else P.inverse(transformation)
        )

    def set_transformation(self, transformation: P.Matrix) -> None:
        self.transformation_inverse = P.inverse(transformation)

    def set_transformation_inverse(self, transformation: P.Matrix) -> None:
        self.transformation_inverse = P.matrix(transformation)

    def transformation(self) -> P.Matrix:
        return P.inverse(self.transformation_inverse)

    def _extra_key(self):
        return ()

    def value_key(self):
        """What the derived ``PartialEq`` compares (type + every field), used for
        ``dyn Shape == dyn Shape`` (shape.rs:34-38, dyn_partial_eq.rs:9-16)."""
        return (self.TYPE, self.material.value_key(), _mkey(self.transformation_inverse), self._extra_key())


class Sphere(Shape):
    TYPE = SPHERE


class Plane(Shape):
    TYPE = PLANE


class Cube(Shape):
    TYPE = CUBE


class Cylinder(Shape):
    """shapes/cylinder.rs:9-32; Default (cylinder.rs:133-143): min = f64::MIN, max = f64::MAX, open."""

    TYPE = CYLINDER

    def __init__(
        self,
        material: Optional[Material] = None,
        transformation: Optional[P.Matrix] = None,
        min: float = P.F64_MIN,
        max: float = P.F64_MAX,
        closed: bool = False,
    ) -> None:
        super().__init__(material, transformation)
        self.min = float(min)
        self.max = float(max)
        self.closed = bool(closed)

    def _extra_key(self):
        return (self.min, self.max, self.closed)


class Cone(Cylinder):
    """shapes/cone.rs:9-32; same defaults (cone.rs:140-150)."""

    TYPE = CONE


class Triangle(Shape):
    """shapes/triangle.rs:9-35"""

    TYPE = TRIANGLE

    def __init__(self, vertex_1: Sequence[float], vertex_2: Sequence[float], vertex_3: Sequence[float]) -> None:
        super().__init__()
        self.vertex_1 = _f3(vertex_1)
        self.vertex_2 = _f3(vertex_2)
        self.vertex_3 = _f3(vertex_3)
        self.edge_1 = P.sub(self.vertex_2, self.vertex_1)
        self.edge_2 = P.sub(self.vertex_3, self.vertex_1)
        self.normal = P.normalized(P.cross(self.edge_2, self.edge_1))

    def _extra_key(self):
        return (self.vertex_1, self.vertex_2, self.vertex_3, self.edge_1, self.edge_2, self.normal)


# --------------------------------------------------------------------------------------------
# world


@dataclass
class World:
    """composites/world.rs:9-22"""

    lights: List[Light] = field(default_factory=list)
    shapes: List[Shape] = field(default_factory=list)

    MAX_REFLECTION_ITERATIONS = 6  # world.rs:15

    @staticmethod
    def default() -> "World":
        """world.rs:160-169 + utils.rs:59-71"""
        s1 = Sphere()
        s1.material.color = (0.8, 1.0, 0.6)
        s1.material.diffuse = 0.7
        s1.material.specular = 0.2
        s2 = Sphere()
        s2.set_transformation(P.scaling(0.5, 0.5, 0.5))
        return World([Light()], [s1, s2])

    def flatten(self):
        from .flatten import flatten_world

        return flatten_world(self)


# --------------------------------------------------------------------------------------------
# canvas


class Canvas:
    """composites/canvas.rs:13-17: ``pixels`` is row-major, index = x + y*width (canvas.rs:44-55),
    3 f64 per pixel."""

    def __init__(self, width: int, height: int, pixels: Optional[np.ndarray] = None, rgb8: Optional[np.ndarray] = None):
        self.width = int(width)
        self.height = int(height)
        if pixels is None:
            pixels = np.zeros((self.width * self.height, 3), dtype=np.float64)  # Canvas::DEFAULT_COLOR = BLACK
        self.pixels = pixels.reshape(self.width * self.height, 3)
        self._rgb8 = rgb8

    def get_pixel(self, x: int, y: int) -> Color:
        r, g, b = self.pixels[x + y * self.width]
        return (float(r), float(g), float(b))

    def to_rgb8(self) -> np.ndarray:
        """canvas.rs:117-123: clamp to [0,1], * 255, round half away from zero, ``as u8``.
        Uses the bytes the device produced when the render asked for them."""
        if self._rgb8 is not None:
            return self._rgb8.reshape(self.height, self.width, 3)
        return quantise_rgb8(self.pixels).reshape(self.height, self.width, 3)

    def to_png_file(self, path: str) -> None:
        """canvas.rs:114-137 (RGB8 PNG)."""
        from PIL import Image

        Image.fromarray(self.to_rgb8(), mode="RGB").save(path, format="PNG")

    def to_ppm(self) -> str:
        """canvas.rs:68-97: header ``P3`` / ``w h`` / ``255``; then floor(70/12) = 5 pixels per line,
        every channel right-aligned to width 3, joined by single spaces; no trailing newline."""
        rgb = self.to_rgb8().reshape(-1, 3)
        lines = ["P3", f"{self.width} {self.height}", "255"]
        pixels_per_line = int(math.floor(70.0 / (3.0 * 4.0)))
        for start in range(0, rgb.shape[0], pixels_per_line):
            chunk = rgb[start : start + pixels_per_line].reshape(-1)
            lines.append(" ".join(f"{int(v):>3d}" for v in chunk))
        return "\n".join(lines)

    def to_ppm_file(self, path: str) -> None:
        """canvas.rs:107-112"""
        with open(path, "w") as f:
            f.write(self.to_ppm())


def quantise_rgb8(pixels: np.ndarray) -> np.ndarray:
    """canvas.rs:117-123 on an array: NaN -> 0 (``as u8``)."""
    v = np.clip(pixels, 0.0, 1.0) * 255.0
    # f64::round = half away from zero; values are >= 0 here (or NaN)
    v = np.floor(v + 0.5)
    # floor(x + 0.5) differs from round-half-away only when x + 0.5 rounds up across an integer;
    # repair that single case exactly
    v = np.where((v - 0.5) > (np.clip(pixels, 0.0, 1.0) * 255.0), v - 1.0, v)
    v = np.nan_to_num(v, nan=0.0)
    return v.astype(np.uint8)


# --------------------------------------------------------------------------------------------
# camera


class Camera:
    """composites/camera.rs:10-19."""

    def __init__(self, horizontal_size: int, vertical_size: int, field_of_view: float) -> None:
        # camera.rs:25-49
        self.horizontal_size = int(horizontal_size)
        self.vertical_size = int(vertical_size)
        self.field_of_view = float(field_of_view)
        half_view = math.tan(self.field_of_view / 2.0)
        if self.vertical_size == 0:
            aspect = math.nan if self.horizontal_size == 0 else math.inf
        else:
            aspect = float(self.horizontal_size) / float(self.vertical_size)
        if aspect >= 1.0:
            self.half_width = half_view
            self.half_height = half_view / aspect
        else:
            self.half_width = half_view * aspect
            self.half_height = half_view
        self.pixel_size = (
            (self.half_width * 2.0) / float(self.horizontal_size) if self.horizontal_size else math.nan
        )
        self.transformation_inverse: P.Matrix = P.identity()
        self.origin = (0.0, 0.0, 0.0)

    def set_transformation(self, transformation: P.Matrix) -> None:
        """camera.rs:124-127"""
        self.transformation_inverse = P.inverse(transformation)
        self._update_origin()

    def set_transformation_inverse(self, transformation: P.Matrix) -> None:
        """camera.rs:133-136"""
        self.transformation_inverse = P.matrix(transformation)
        self._update_origin()

    def transformation(self) -> P.Matrix:
        return P.inverse(self.transformation_inverse)

    def _update_origin(self) -> None:
        """camera.rs:114-116"""
        self.origin = P.mat_point(self.transformation_inverse, (0.0, 0.0, 0.0))

    def resized(self, horizontal_size: int, vertical_size: int) -> "Camera":
        """A camera with the same field of view and placement at another image size (the
        reference has no size flag; benches override the YAML ``width``/``height``)."""
        cam = Camera(horizontal_size, vertical_size, self.field_of_view)
        cam.set_transformation_inverse(self.transformation_inverse)
        return cam

    def render_gpu(self, world: World, **kwargs) -> Canvas:
        """The new ``RenderingMode::Gpu`` arm (ray-tracer-cli/src/main.rs:18-21): same contract as
        ``Camera::render_parallel`` (camera.rs:97-112) — returns the full Canvas — computed by the
        CUDA path.  Raises if the CUDA library or a device is missing (no CPU fallback)."""
        from .render import render_gpu

        return render_gpu(self, world, **kwargs)
