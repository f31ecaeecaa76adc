"""Seeded random worlds for fuzzing the CUDA path against the oracle: every shape kind under random
affine transforms (rotations, anisotropic scales, shears), every pattern kind (nested complex patterns
included), random Phong / reflective / transparent materials, 0-3 lights, value-equal duplicates."""
import math

import numpy as np

from ray_tracer_challenge_rs_b200 import (
    Camera, CheckerPattern, ComplexPattern, Cone, Cube, Cylinder, GradientPattern, Light, Material, Plane, RingPattern,
    Sphere, StripePattern, Triangle, World,
)
from ray_tracer_challenge_rs_b200 import primitives as P


def _transform(rng, spread):
    t = P.identity()
    for _ in range(int(rng.integers(1, 4))):
        kind = int(rng.integers(0, 6))
        if kind == 0:
            t = P.mat_mul(P.scaling(*rng.uniform(0.3, 1.6, 3)), t)
        elif kind == 1:
            t = P.mat_mul(P.rotation_x(float(rng.uniform(-math.pi, math.pi))), t)
        elif kind == 2:
            t = P.mat_mul(P.rotation_y(float(rng.uniform(-math.pi, math.pi))), t)
        elif kind == 3:
            t = P.mat_mul(P.rotation_z(float(rng.uniform(-math.pi, math.pi))), t)
        elif kind == 4:
            t = P.mat_mul(P.shearing(*rng.uniform(-0.3, 0.3, 6)), t)
        else:
            s = float(rng.uniform(0.4, 1.5))
            t = P.mat_mul(P.scaling(s, s, s), t)
    return P.mat_mul(P.translation(*rng.uniform(-spread, spread, 3)), t)


def _pattern(rng, depth=0):
    kind = int(rng.integers(0, 5 if depth < 2 else 4))
    a, b = tuple(rng.uniform(0, 1, 3)), tuple(rng.uniform(0, 1, 3))
    if kind == 4:
        p = ComplexPattern(_pattern(rng, depth + 1), _pattern(rng, depth + 1))
    else:
        p = [StripePattern, GradientPattern, RingPattern, CheckerPattern][kind](a, b)
    if rng.random() < 0.7:
        p.set_transformation(_transform(rng, 1.0))
    return p


def _material(rng):
    m = Material(color=tuple(rng.uniform(0, 1, 3)), ambient=float(rng.uniform(0, 0.3)), diffuse=float(rng.uniform(0.2, 0.9)),
                 specular=float(rng.choice([0.0, rng.uniform(0, 1)])), shininess=float(rng.uniform(5, 300)))
    if rng.random() < 0.35:
        m.pattern = _pattern(rng)
    r = rng.random()
    if r < 0.25:
        m.reflectiveness = float(rng.uniform(0.1, 1.0))
    elif r < 0.45:
        m.transparency = float(rng.uniform(0.2, 1.0))
        m.refractive_index = float(rng.choice([1.0, 1.0000034, 1.33, 1.5, 2.4]))
        if rng.random() < 0.6:
            m.reflectiveness = float(rng.uniform(0.1, 1.0))
    if rng.random() < 0.15:
        m.casts_shadow = False
    return m


def random_world(seed, n_shapes=14, width=72, height=48):
    rng = np.random.default_rng(seed)
    shapes = []
    spread = 2.5
    for _ in range(n_shapes):
        kind = int(rng.integers(0, 6))
        m = _material(rng)
        if kind == 0:
            s = Sphere(m, _transform(rng, spread))
        elif kind == 1:
            s = Plane(m, P.mat_mul(P.translation(0, float(rng.uniform(-4, -2)), 0), P.rotation_z(float(rng.uniform(-0.2, 0.2)))))
        elif kind == 2:
            s = Cube(m, _transform(rng, spread))
        elif kind in (3, 4):
            lo, hi = sorted(rng.uniform(-1.5, 1.5, 2))
            cls = Cylinder if kind == 3 else Cone
            if rng.random() < 0.15:
                s = cls(m, _transform(rng, spread))  # untruncated: unbounded
            else:
                s = cls(m, _transform(rng, spread), min=float(lo), max=float(hi), closed=bool(rng.random() < 0.6))
        else:
            c = rng.uniform(-spread, spread, 3)
            s = Triangle(*(tuple(c + rng.uniform(-1.2, 1.2, 3)) for _ in range(3)))
            s.material = m
            if rng.random() < 0.5:
                s.set_transformation(_transform(rng, 0.5))
        shapes.append(s)
        if rng.random() < 0.12:  # a value-equal duplicate (shape_eq_class)
            dup = type(s).__new__(type(s))
            dup.__dict__.update(s.__dict__)
            shapes.append(dup)
    lights = [Light(tuple(rng.uniform(-8, 8, 3) + np.array([0, 6, -6])), tuple(rng.uniform(0.2, 1.0, 3))) for _ in range(0 if rng.random() < 0.08 else int(rng.integers(1, 4)))]
    cam = Camera(width, height, float(rng.uniform(0.6, 1.4)))
    frm = tuple(rng.uniform(-2, 2, 3) + np.array([0, 1.5, -8]))
    cam.set_transformation(P.view_transform(frm, tuple(rng.uniform(-1, 1, 3)), (0, 1, 0)))
    return World(lights, shapes), cam
