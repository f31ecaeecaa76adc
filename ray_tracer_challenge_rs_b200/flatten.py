"""Scene flattener: ``World`` (a list of trait objects in the reference, ``world.rs:9-12``) ->
structure-of-arrays buffers in the layout of ``rtgpu_scene`` (``include/rtgpu.h``).

This is subsystem (1) of the north star: no virtual dispatch survives past this point.  The same
``FlatScene`` is what the synthetic generator emits directly and what the committed scene
fixtures (``scenes/*.npz``) store.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, fields
from typing import Dict, List

import numpy as np

from . import abi
from .scene import Camera, ComplexPattern, Pattern, World


def _rows012(m) -> List[float]:
    return [m[r][c] for r in range(3) for c in range(4)]


@dataclass
class FlatScene:
    """Numpy twin of ``rtgpu_scene``.  Array names = the C field names."""

    shape_type: np.ndarray  # u8  [S]
    shape_inv: np.ndarray  # f64 [S,12]
    shape_min: np.ndarray  # f64 [S]
    shape_max: np.ndarray  # f64 [S]
    shape_closed: np.ndarray  # u8  [S]
    shape_triangle: np.ndarray  # i32 [S]
    shape_material: np.ndarray  # u32 [S]
    shape_eq_class: np.ndarray  # u32 [S]
    tri_vertex_1: np.ndarray  # f64 [T,3]
    tri_edge_1: np.ndarray
    tri_edge_2: np.ndarray
    tri_normal: np.ndarray
    mat_color: np.ndarray  # f64 [M,3]
    mat_params: np.ndarray  # f64 [M,7]
    mat_casts_shadow: np.ndarray  # u8 [M]
    mat_pattern: np.ndarray  # i32 [M]
    pat_type: np.ndarray  # u8 [Q]
    pat_color_a: np.ndarray  # f64 [Q,3]
    pat_color_b: np.ndarray
    pat_inv: np.ndarray  # f64 [Q,12]
    pat_child_a: np.ndarray  # i32 [Q]
    pat_child_b: np.ndarray
    light_position: np.ndarray  # f64 [L,3]
    light_intensity: np.ndarray

    _DTYPES = {
        "shape_type": np.uint8,
        "shape_closed": np.uint8,
        "shape_triangle": np.int32,
        "shape_material": np.uint32,
        "shape_eq_class": np.uint32,
        "mat_casts_shadow": np.uint8,
        "mat_pattern": np.int32,
        "pat_type": np.uint8,
        "pat_child_a": np.int32,
        "pat_child_b": np.int32,
    }
    _WIDTH = {
        "shape_inv": 12,
        "tri_vertex_1": 3,
        "tri_edge_1": 3,
        "tri_edge_2": 3,
        "tri_normal": 3,
        "mat_color": 3,
        "mat_params": abi.MAT_PARAM_COUNT,
        "pat_color_a": 3,
        "pat_color_b": 3,
        "pat_inv": 12,
        "light_position": 3,
        "light_intensity": 3,
    }

    def __post_init__(self) -> None:
        for f in fields(self):
            dt = self._DTYPES.get(f.name, np.float64)
            a = np.ascontiguousarray(np.asarray(getattr(self, f.name), dtype=dt))
            w = self._WIDTH.get(f.name)
            if w is not None:
                a = a.reshape(-1, w)
            setattr(self, f.name, a)
        self.validate()

    # ---- sizes -------------------------------------------------------------------------
    @property
    def n_shapes(self) -> int:
        return int(self.shape_type.shape[0])

    @property
    def n_triangles(self) -> int:
        return int(self.tri_vertex_1.shape[0])

    @property
    def n_materials(self) -> int:
        return int(self.mat_color.shape[0])

    @property
    def n_patterns(self) -> int:
        return int(self.pat_type.shape[0])

    @property
    def n_lights(self) -> int:
        return int(self.light_position.shape[0])

    def validate(self) -> None:
        S, M, Q, T = self.n_shapes, self.n_materials, self.n_patterns, self.n_triangles
        for name in ("shape_inv", "shape_min", "shape_max", "shape_closed", "shape_triangle", "shape_material", "shape_eq_class"):
            if getattr(self, name).shape[0] != S:
                raise ValueError(f"{name}: expected {S} entries")
        if S and (self.shape_type.max() >= 6):
            raise ValueError("shape_type out of range")
        if S and M == 0:
            raise ValueError("shapes without materials")
        if S and self.shape_material.max() >= M:
            raise ValueError("shape_material out of range")
        if S and self.shape_eq_class.max() >= S:
            raise ValueError("shape_eq_class out of range")
        tri = self.shape_type == abi.TRIANGLE
        if tri.any() and (self.shape_triangle[tri].min() < 0 or self.shape_triangle[tri].max() >= T):
            raise ValueError("shape_triangle out of range")
        for name in ("mat_params", "mat_casts_shadow", "mat_pattern"):
            if getattr(self, name).shape[0] != M:
                raise ValueError(f"{name}: expected {M} entries")
        if M and self.mat_pattern.max() >= Q:
            raise ValueError("mat_pattern out of range")
        for name in ("pat_color_a", "pat_color_b", "pat_inv", "pat_child_a", "pat_child_b"):
            if getattr(self, name).shape[0] != Q:
                raise ValueError(f"{name}: expected {Q} entries")
        if self.light_intensity.shape[0] != self.n_lights:
            raise ValueError("light_intensity: size mismatch")

    # ---- C view ------------------------------------------------------------------------
    def as_c(self) -> abi.RtgpuScene:
        """A ``rtgpu_scene`` whose pointers alias this object's arrays (keep ``self`` alive)."""

        def ptr(a: np.ndarray, ctype):
            return a.ctypes.data_as(C.POINTER(ctype)) if a.size else C.cast(None, C.POINTER(ctype))

        s = abi.RtgpuScene()
        s.abi_version = abi.ABI_VERSION
        s.n_shapes = self.n_shapes
        s.shape_type = ptr(self.shape_type, C.c_uint8)
        s.shape_inv = ptr(self.shape_inv, C.c_double)
        s.shape_min = ptr(self.shape_min, C.c_double)
        s.shape_max = ptr(self.shape_max, C.c_double)
        s.shape_closed = ptr(self.shape_closed, C.c_uint8)
        s.shape_triangle = ptr(self.shape_triangle, C.c_int32)
        s.shape_material = ptr(self.shape_material, C.c_uint32)
        s.shape_eq_class = ptr(self.shape_eq_class, C.c_uint32)
        s.n_triangles = self.n_triangles
        s.tri_vertex_1 = ptr(self.tri_vertex_1, C.c_double)
        s.tri_edge_1 = ptr(self.tri_edge_1, C.c_double)
        s.tri_edge_2 = ptr(self.tri_edge_2, C.c_double)
        s.tri_normal = ptr(self.tri_normal, C.c_double)
        s.n_materials = self.n_materials
        s.mat_color = ptr(self.mat_color, C.c_double)
        s.mat_params = ptr(self.mat_params, C.c_double)
        s.mat_casts_shadow = ptr(self.mat_casts_shadow, C.c_uint8)
        s.mat_pattern = ptr(self.mat_pattern, C.c_int32)
        s.n_patterns = self.n_patterns
        s.pat_type = ptr(self.pat_type, C.c_uint8)
        s.pat_color_a = ptr(self.pat_color_a, C.c_double)
        s.pat_color_b = ptr(self.pat_color_b, C.c_double)
        s.pat_inv = ptr(self.pat_inv, C.c_double)
        s.pat_child_a = ptr(self.pat_child_a, C.c_int32)
        s.pat_child_b = ptr(self.pat_child_b, C.c_int32)
        s.n_lights = self.n_lights
        s.light_position = ptr(self.light_position, C.c_double)
        s.light_intensity = ptr(self.light_intensity, C.c_double)
        s._owner = self  # keep the arrays alive as long as the struct
        return s

    # ---- (de)serialisation: bit-exact fixtures ------------------------------------------
    def to_arrays(self) -> Dict[str, np.ndarray]:
        return {f.name: getattr(self, f.name) for f in fields(self)}

    @staticmethod
    def from_arrays(arrays) -> "FlatScene":
        return FlatScene(**{f.name: arrays[f.name] for f in fields(FlatScene)})

    def shape_counts(self) -> Dict[str, int]:
        return {abi.SHAPE_NAMES[t]: int((self.shape_type == t).sum()) for t in range(6) if (self.shape_type == t).any()}


def flatten_world(world: World) -> FlatScene:
    """``World`` -> ``FlatScene``.  Shapes keep ``world.shapes`` order (tie-breaks depend on it);
    materials and patterns are de-duplicated by value; ``shape_eq_class`` encodes
    ``dyn Shape == dyn Shape`` (shapes/shape.rs:34-38)."""
    mat_index: Dict[object, int] = {}
    pat_index: Dict[object, int] = {}
    pat_rows: List[dict] = []
    mat_rows: List[dict] = []

    def add_pattern(p: Pattern) -> int:
        key = p.value_key()
        if key in pat_index:
            return pat_index[key]
        child_a = child_b = -1
        if isinstance(p, ComplexPattern):
            child_a = add_pattern(p.pattern_a)
            child_b = add_pattern(p.pattern_b)
        idx = len(pat_rows)
        pat_index[key] = idx
        pat_rows.append(
            dict(
                type=p.TYPE,
                a=getattr(p, "color_a", (0.0, 0.0, 0.0)),
                b=getattr(p, "color_b", (0.0, 0.0, 0.0)),
                inv=_rows012(p.transformation_inverse),
                child_a=child_a,
                child_b=child_b,
            )
        )
        return idx

    def add_material(m) -> int:
        key = m.value_key()
        if key in mat_index:
            return mat_index[key]
        pat = -1 if m.pattern is None else add_pattern(m.pattern)
        idx = len(mat_rows)
        mat_index[key] = idx
        mat_rows.append(
            dict(
                color=tuple(float(c) for c in m.color),
                params=(
                    float(m.ambient),
                    float(m.diffuse),
                    float(m.specular),
                    float(m.shininess),
                    float(m.reflectiveness),
                    float(m.transparency),
                    float(m.refractive_index),
                ),
                casts_shadow=1 if m.casts_shadow else 0,
                pattern=pat,
            )
        )
        return idx

    S = len(world.shapes)
    shape_type = np.zeros(S, np.uint8)
    shape_inv = np.zeros((S, 12), np.float64)
    shape_min = np.zeros(S, np.float64)
    shape_max = np.zeros(S, np.float64)
    shape_closed = np.zeros(S, np.uint8)
    shape_triangle = np.full(S, -1, np.int32)
    shape_material = np.zeros(S, np.uint32)
    shape_eq_class = np.zeros(S, np.uint32)
    tri_v1, tri_e1, tri_e2, tri_n = [], [], [], []
    classes: Dict[object, int] = {}
    for i, sh in enumerate(world.shapes):
        row3 = sh.transformation_inverse[3]
        if not (row3[0] == 0.0 and row3[1] == 0.0 and row3[2] == 0.0):
            raise ValueError(
                f"shape {i}: transformation_inverse is not affine (row 3 = {row3}); the flattened scene "
                "carries rows 0..2 only"
            )
        shape_type[i] = sh.TYPE
        shape_inv[i] = _rows012(sh.transformation_inverse)
        if sh.TYPE in (abi.CYLINDER, abi.CONE):
            shape_min[i], shape_max[i], shape_closed[i] = sh.min, sh.max, 1 if sh.closed else 0
        if sh.TYPE == abi.TRIANGLE:
            shape_triangle[i] = len(tri_v1)
            tri_v1.append(sh.vertex_1)
            tri_e1.append(sh.edge_1)
            tri_e2.append(sh.edge_2)
            tri_n.append(sh.normal)
        shape_material[i] = add_material(sh.material)
        shape_eq_class[i] = classes.setdefault(sh.value_key(), i)

    def arr(rows, key, width, dtype=np.float64):
        if not rows:
            return np.zeros((0, width) if width else (0,), dtype)
        return np.asarray([r[key] for r in rows], dtype=dtype)

    return FlatScene(
        shape_type=shape_type,
        shape_inv=shape_inv,
        shape_min=shape_min,
        shape_max=shape_max,
        shape_closed=shape_closed,
        shape_triangle=shape_triangle,
        shape_material=shape_material,
        shape_eq_class=shape_eq_class,
        tri_vertex_1=np.asarray(tri_v1, np.float64).reshape(-1, 3),
        tri_edge_1=np.asarray(tri_e1, np.float64).reshape(-1, 3),
        tri_edge_2=np.asarray(tri_e2, np.float64).reshape(-1, 3),
        tri_normal=np.asarray(tri_n, np.float64).reshape(-1, 3),
        mat_color=arr(mat_rows, "color", 3),
        mat_params=arr(mat_rows, "params", abi.MAT_PARAM_COUNT),
        mat_casts_shadow=arr(mat_rows, "casts_shadow", 0, np.uint8),
        mat_pattern=arr(mat_rows, "pattern", 0, np.int32),
        pat_type=arr(pat_rows, "type", 0, np.uint8),
        pat_color_a=arr(pat_rows, "a", 3),
        pat_color_b=arr(pat_rows, "b", 3),
        pat_inv=arr(pat_rows, "inv", 12),
        pat_child_a=arr(pat_rows, "child_a", 0, np.int32),
        pat_child_b=arr(pat_rows, "child_b", 0, np.int32),
        light_position=np.asarray([l.position for l in world.lights], np.float64).reshape(-1, 3),
        light_intensity=np.asarray([l.intensity for l in world.lights], np.float64).reshape(-1, 3),
    )


def camera_to_c(camera: Camera) -> abi.RtgpuCamera:
    """``Camera`` (camera.rs:10-19) -> ``rtgpu_camera``."""
    c = abi.RtgpuCamera()
    c.hsize = camera.horizontal_size
    c.vsize = camera.vertical_size
    c.half_width = camera.half_width
    c.half_height = camera.half_height
    c.pixel_size = camera.pixel_size
    inv = _rows012(camera.transformation_inverse)
    for i in range(12):
        c.inv[i] = inv[i]
    for i in range(3):
        c.origin[i] = camera.origin[i]
    return c


def camera_to_dict(camera: Camera) -> dict:
    return dict(
        horizontal_size=camera.horizontal_size,
        vertical_size=camera.vertical_size,
        field_of_view=camera.field_of_view,
        transformation_inverse=np.asarray(camera.transformation_inverse, np.float64),
    )


def camera_from_dict(d) -> Camera:
    cam = Camera(int(d["horizontal_size"]), int(d["vertical_size"]), float(d["field_of_view"]))
    cam.set_transformation_inverse([[float(v) for v in row] for row in np.asarray(d["transformation_inverse"])])
    return cam
