"""YAML scene loader — restates ``ray-tracer-cli/src/scene_loader.rs`` (all line numbers below are
in that file) so a scene file yields bit-identical ``(World, Camera)`` inputs to the render pass.

Quirks kept on purpose:
  * definitions are collected in a first pass in document order, dispatched by NAME SUFFIX
    (``-color`` / ``-material`` / ``-transform`` | ``-object``)                         (:46-62)
  * in a transform list a *named* transform is right-multiplied (``t = t * named``), an inline
    op is left-multiplied (``t = op * t``)                                               (:204-233)
  * unknown ops, unknown ``add:`` kinds and unknown keys are silently ignored            (:232, :330)
  * cylinders / cones start from ``Default`` (min = f64::MIN, max = f64::MAX, open)      (:296-329)
  * numbers: YAML integers via ``as f64``, reals via ``str::parse::<f64>``               (:338-344)
  * camera ``width``/``height`` are parsed as f64 and cast ``as u32``                    (:256-257)
  * triangles cannot be loaded (no ``triangle`` arm)                                     (:254-331)
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Tuple

import yaml

from . import primitives as P
from .scene import (
    Camera,
    CheckerPattern,
    Cone,
    Cube,
    Cylinder,
    GradientPattern,
    Light,
    Material,
    Pattern,
    Plane,
    RingPattern,
    Sphere,
    StripePattern,
    World,
)

_BAD = object()  # yaml_rust::Yaml::BadValue


def _get(node: Any, key: str) -> Any:
    """``yaml[key]`` of yaml-rust: BadValue unless ``yaml`` is a hash that has ``key``."""
    if isinstance(node, dict) and key in node:
        return node[key]
    return _BAD


def _iter(node: Any) -> List[Any]:
    """``yaml.into_iter()``: the elements of an array, nothing otherwise."""
    return list(node) if isinstance(node, list) else []


def parse_f64(node: Any) -> float:
    """:338-344"""
    if isinstance(node, bool):
        raise ValueError("cannot parse float from empty string")  # Yaml::Boolean -> "".parse()
    if isinstance(node, int):
        return float(node)
    if isinstance(node, float):
        return node
    if isinstance(node, str):
        # yaml-rust types any scalar that parses as f64 as Real (e.g. `1e3`, which YAML 1.1 /
        # PyYAML leaves a string)
        try:
            return float(node)
        except ValueError:
            pass
    raise ValueError("cannot parse float from empty string")


def parse_array_of_3(values: List[Any]) -> Tuple[float, float, float]:
    """:346-352"""
    vals = [parse_f64(v) for v in values[:3]]
    return (vals[0], vals[1], vals[2])


def _as_u32(v: float) -> int:
    """Rust ``f64 as u32``: truncate, saturate, NaN -> 0."""
    if v != v:
        return 0
    if v <= 0.0:
        return 0
    if v >= 4294967295.0:
        return 4294967295
    return int(math.trunc(v))


class SceneParser:
    def __init__(self) -> None:
        self.colors: Dict[str, Tuple[float, float, float]] = {}
        self.materials: Dict[str, Material] = {}
        self.transformations: Dict[str, P.Matrix] = {}

    # :46-62
    def process_definitions(self, doc: Any) -> None:
        for entry in _iter(doc):
            name = _get(entry, "define")
            if isinstance(name, str):
                if name.endswith("-color"):
                    self.colors[name] = self.parse_color(entry)
                elif name.endswith("-material"):
                    self.materials[name] = self.parse_material(entry)
                elif name.endswith("-transform") or name.endswith("-object"):
                    self.transformations[name] = self.parse_transformation(entry)

    # :64-90
    def parse_color(self, node: Any) -> Tuple[float, float, float]:
        if isinstance(node, dict):
            keyword = "color" if _get(node, "color") is not _BAD else "value"
            return self.parse_color(_get(node, keyword))
        if isinstance(node, list):
            return parse_array_of_3(node)
        if isinstance(node, str):
            return self.colors[node]
        raise ValueError("Incorrect color value")

    # :92-149
    def parse_material(self, node: Any) -> Material:
        if isinstance(node, str):
            return self.materials[node].clone()
        extend = _get(node, "extend")
        material = Material() if extend is _BAD else self.materials[extend].clone()
        if _get(node, "value") is not _BAD:
            node = _get(node, "value")
        if _get(node, "color") is not _BAD:
            material.color = self.parse_color(_get(node, "color"))
        if _get(node, "pattern") is not _BAD:
            material.pattern = self.parse_pattern(_get(node, "pattern"))
        for key, attr in (
            ("ambient", "ambient"),
            ("diffuse", "diffuse"),
            ("specular", "specular"),
            ("shininess", "shininess"),
            ("reflective", "reflectiveness"),
            ("transparency", "transparency"),
            ("refractive-index", "refractive_index"),
        ):
            if _get(node, key) is not _BAD:
                setattr(material, attr, parse_f64(_get(node, key)))
        casts = _get(node, "casts-shadow")
        if isinstance(casts, bool):
            material.casts_shadow = casts
        return material

    # :151-193
    def parse_pattern(self, node: Any) -> Pattern:
        colors = _get(node, "colors")
        color_a = self.parse_color(colors[0])
        color_b = self.parse_color(colors[1])
        transformation = None
        if _get(node, "transform") is not _BAD:
            transformation = self.parse_transformation(_get(node, "transform"))
        kind = _get(node, "type")
        ctor = {
            "stripes": StripePattern,
            "gradient": GradientPattern,
            "rings": RingPattern,
            "checkers": CheckerPattern,
        }.get(kind if isinstance(kind, str) else None)
        if ctor is None:
            raise ValueError("Incorrect pattern type")
        pattern = ctor(color_a, color_b)
        if transformation is not None:
            pattern.set_transformation(transformation)
        return pattern

    # :195-238
    def parse_transformation(self, node: Any) -> P.Matrix:
        transformation = P.identity()
        if _get(node, "value") is not _BAD:
            node = _get(node, "value")
        for transform in _iter(node):
            if isinstance(transform, str):
                transformation = P.mat_mul(transformation, self.transformations[transform])
            elif isinstance(transform, list):
                op = transform[0]
                if op == "scale":
                    v = parse_array_of_3(transform[1:])
                    transformation = P.mat_mul(P.scaling(*v), transformation)
                elif op == "translate":
                    v = parse_array_of_3(transform[1:])
                    transformation = P.mat_mul(P.translation(*v), transformation)
                elif op == "rotate-x":
                    transformation = P.mat_mul(P.rotation_x(parse_f64(transform[1])), transformation)
                elif op == "rotate-y":
                    transformation = P.mat_mul(P.rotation_y(parse_f64(transform[1])), transformation)
                elif op == "rotate-z":
                    transformation = P.mat_mul(P.rotation_z(parse_f64(transform[1])), transformation)
        return transformation

    # :240-247
    def parse_material_and_transformation(self, entry: Any) -> Tuple[Material, P.Matrix]:
        material = self.parse_material(_get(entry, "material")) if _get(entry, "material") is not _BAD else Material()
        transformation = self.parse_transformation(_get(entry, "transform"))
        return material, transformation

    # :249-335
    def parse_scene(self, doc: Any) -> Tuple[World, Camera]:
        world = World([], [])
        camera = Camera(0, 0, 0.0)
        for entry in _iter(doc):
            name = _get(entry, "add")
            if not isinstance(name, str):
                continue
            if name == "camera":
                horizontal_size = _as_u32(parse_f64(_get(entry, "width")))
                vertical_size = _as_u32(parse_f64(_get(entry, "height")))
                fov = parse_f64(_get(entry, "field-of-view"))
                from_ = parse_array_of_3(_get(entry, "from"))
                to = parse_array_of_3(_get(entry, "to"))
                up = parse_array_of_3(_get(entry, "up"))
                camera = Camera(horizontal_size, vertical_size, fov)
                camera.set_transformation(P.view_transform(from_, to, up))
            elif name == "light":
                position = parse_array_of_3(_get(entry, "at"))
                intensity = parse_array_of_3(_get(entry, "intensity"))
                world.lights.append(Light(position, intensity))
            elif name in ("plane", "sphere", "cube"):
                material, transformation = self.parse_material_and_transformation(entry)
                ctor = {"plane": Plane, "sphere": Sphere, "cube": Cube}[name]
                world.shapes.append(ctor(material, transformation))
            elif name in ("cone", "cylinder"):
                material, transformation = self.parse_material_and_transformation(entry)
                shape = Cone() if name == "cone" else Cylinder()
                shape.material = material
                shape.set_transformation(transformation)
                closed = _get(entry, "closed")
                if isinstance(closed, bool):
                    shape.closed = closed
                if _get(entry, "max") is not _BAD:
                    shape.max = parse_f64(_get(entry, "max"))
                if _get(entry, "min") is not _BAD:
                    shape.min = parse_f64(_get(entry, "min"))
                world.shapes.append(shape)
        return world, camera


def load_scene_from_str(text: str) -> Tuple[World, Camera]:
    docs = list(yaml.safe_load_all(text))
    if len(docs) != 1:
        raise ValueError("Incorrect yaml format")  # :357
    parser = SceneParser()
    parser.process_definitions(docs[0])
    return parser.parse_scene(docs[0])


def load_scene_description(path: str) -> Tuple[World, Camera]:
    """:361-367"""
    with open(path, "r") as f:
        return load_scene_from_str(f.read())
