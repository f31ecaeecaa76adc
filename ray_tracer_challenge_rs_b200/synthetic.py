"""Synthetic scaling scenes (BASELINE.json configs[4], SURVEY.md 8d "C5"): N shapes, half spheres and
half triangles, random materials and patterns, two lights, emitted directly as a :class:`FlatScene`
(no per-shape Python objects, so 10^6 shapes take seconds).

Deterministic for a given (n_shapes, seed): numpy's PCG64 stream.  There is no reference render of
these scenes — parity is against the CPU oracle on the same flattened arrays.
"""
from __future__ import annotations

import numpy as np

from . import abi
from . import primitives as P
from .flatten import FlatScene
from .scene import Camera


def synthetic_scene(n_shapes: int, seed: int = 0xB200, extent: float = 50.0, sphere_fraction: float = 0.5) -> FlatScene:
    rng = np.random.default_rng(seed)
    n = int(n_shapes)
    is_sphere = rng.random(n) < sphere_fraction
    centre = rng.uniform(-extent, extent, (n, 3))
    radius = rng.uniform(0.05, 0.5, n)

    # spheres: translation(c) * scaling(r)  ->  inverse = scaling(1/r) * translation(-c)
    inv = np.zeros((n, 12))
    inv[:, 0] = inv[:, 5] = inv[:, 10] = 1.0  # triangles keep the identity
    s = is_sphere
    inv[s, 0] = inv[s, 5] = inv[s, 10] = 1.0 / radius[s]
    inv[s, 3] = -centre[s, 0] / radius[s]
    inv[s, 7] = -centre[s, 1] / radius[s]
    inv[s, 11] = -centre[s, 2] / radius[s]

    # triangles: vertices = centre + U[-0.5, 0.5]^3 (shapes/triangle.rs:21-35 for the derived fields)
    t = ~is_sphere
    nt = int(t.sum())
    v1 = centre[t] + rng.uniform(-0.5, 0.5, (nt, 3))
    v2 = centre[t] + rng.uniform(-0.5, 0.5, (nt, 3))
    v3 = centre[t] + rng.uniform(-0.5, 0.5, (nt, 3))
    e1, e2 = v2 - v1, v3 - v1
    normal = np.cross(e2, e1)
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    shape_triangle = np.full(n, -1, np.int32)
    shape_triangle[t] = np.arange(nt, dtype=np.int32)

    # one material per shape
    colour = rng.uniform(0.0, 1.0, (n, 3))
    params = np.zeros((n, abi.MAT_PARAM_COUNT))
    params[:, abi.MAT_AMBIENT] = 0.1
    params[:, abi.MAT_DIFFUSE] = rng.uniform(0.5, 0.9, n)
    params[:, abi.MAT_SPECULAR] = rng.uniform(0.0, 0.9, n)
    params[:, abi.MAT_SHININESS] = rng.uniform(10.0, 300.0, n)
    reflective = rng.random(n) < 0.10
    params[reflective, abi.MAT_REFLECTIVENESS] = rng.uniform(0.1, 0.9, int(reflective.sum()))
    transparent = rng.random(n) < 0.05
    params[transparent, abi.MAT_TRANSPARENCY] = rng.uniform(0.3, 0.9, int(transparent.sum()))
    params[:, abi.MAT_REFRACTIVE_INDEX] = 1.0
    params[transparent, abi.MAT_REFRACTIVE_INDEX] = 1.5
    patterned = rng.random(n) < 0.25
    nq = int(patterned.sum())
    mat_pattern = np.full(n, -1, np.int32)
    mat_pattern[patterned] = np.arange(nq, dtype=np.int32)
    pat_type = rng.integers(0, 4, nq).astype(np.uint8)  # stripe, gradient, ring, checker
    pat_scale = rng.uniform(0.1, 1.0, nq)
    pat_inv = np.zeros((nq, 12))
    pat_inv[:, 0] = pat_inv[:, 5] = pat_inv[:, 10] = 1.0 / pat_scale

    lights_pos = np.array([[-100.0, 120.0, -150.0], [150.0, 60.0, -100.0]])
    lights_int = np.array([[0.9, 0.9, 0.9], [0.35, 0.35, 0.4]])

    return FlatScene(
        shape_type=np.where(is_sphere, abi.SPHERE, abi.TRIANGLE).astype(np.uint8),
        shape_inv=inv,
        shape_min=np.zeros(n),
        shape_max=np.zeros(n),
        shape_closed=np.zeros(n, np.uint8),
        shape_triangle=shape_triangle,
        shape_material=np.arange(n, dtype=np.uint32),
        shape_eq_class=np.arange(n, dtype=np.uint32),
        tri_vertex_1=v1,
        tri_edge_1=e1,
        tri_edge_2=e2,
        tri_normal=normal,
        mat_color=colour,
        mat_params=params,
        mat_casts_shadow=np.ones(n, np.uint8),
        mat_pattern=mat_pattern,
        pat_type=pat_type,
        pat_color_a=rng.uniform(0.0, 1.0, (nq, 3)),
        pat_color_b=rng.uniform(0.0, 1.0, (nq, 3)),
        pat_inv=pat_inv,
        pat_child_a=np.full(nq, -1, np.int32),
        pat_child_b=np.full(nq, -1, np.int32),
        light_position=lights_pos,
        light_intensity=lights_int,
    )


def synthetic_camera(width: int, height: int, distance: float = 120.0, fov: float = 0.9) -> Camera:
    cam = Camera(width, height, fov)
    cam.set_transformation(P.view_transform((0.0, 0.0, -distance), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0)))
    return cam
