"""Pins the CPU oracle (oracle/rt_oracle.c) — and, with `-m gpu`, the DEVICE code through its known-answer probes
(rtgpu_debug_probe, csrc/rt_probe.cuh) — against the reference's OWN known-answer unit tests.

Every test names the reference test it replays (file:line under /root/reference/ray-tracer/src).
``==`` here is bit-equality of f64, as ``assert_eq!`` is in the reference; ``coarse`` is the
reference's ``coarse_eq`` (|a-b| < 8e-8, utils.rs:16-24).  The bit-exact cone / cylinder roots
discriminate fused-multiply-add placement.
"""
import math

import pytest

from oracle.oracle import Oracle
from ray_tracer_challenge_rs_b200 import (
    Camera,
    CheckerPattern,
    Cone,
    Cube,
    Cylinder,
    GradientPattern,
    Light,
    Material,
    Plane,
    RingPattern,
    Sphere,
    StripePattern,
    TestPattern,
    Triangle,
    World,
)
from ray_tracer_challenge_rs_b200 import primitives as P

EPS = 0.00000008
PI = math.pi
S2 = math.sqrt(2.0)
WHITE = (1.0, 1.0, 1.0)
BLACK = (0.0, 0.0, 0.0)
UP, DOWN, LEFT, RIGHT, FORWARD, BACKWARD = (0, 1, 0), (0, -1, 0), (-1, 0, 0), (1, 0, 0), (0, 0, 1), (0, 0, -1)


def coarse(a, b):
    return all(x == y or abs(x - y) < EPS for x, y in zip(a, b))


_BACKEND = Oracle  # what evaluates the replayed tests: the CPU oracle, or the device probes


@pytest.fixture(params=["oracle", pytest.param("device-persistent", marks=pytest.mark.gpu), pytest.param("device-wavefront", marks=pytest.mark.gpu)])
def backend(request, oracle_lib):
    """Every replayed test runs against the oracle (CPU suite) and against the device (`-m gpu`): the same inputs,
    the same expected values, bit-exact where the reference uses assert_eq!."""
    global _BACKEND
    if request.param == "oracle":
        _BACKEND = Oracle
    else:
        from ray_tracer_challenge_rs_b200.probe import DeviceProbe

        family = request.param.split("-")[1]
        _BACKEND = lambda flat: DeviceProbe(flat, family=family)  # noqa: E731
    yield request.param
    _BACKEND = Oracle


def make(flat):
    return _BACKEND(flat)


def shade(o, *args, **kwargs):
    """o.shade_entry(...).  reflected_color / refracted_color in isolation exist only on the CPU side (on the device
    they are phases of a node, covered through color_at and whole frames): those replays are skipped for the device."""
    from ray_tracer_challenge_rs_b200.probe import ProbeUnsupported

    try:
        return o.shade_entry(*args, **kwargs)
    except ProbeUnsupported as e:
        pytest.skip(str(e))


def oracle_for(*shapes, lights=None):
    world = World(list(lights) if lights is not None else [Light()], list(shapes))
    return make(world.flatten())


def fvec(v):
    return tuple(float(x) for x in v)


# ---------------------------------------------------------------------------------------------
# camera.rs


def test_pixel_size_horizontal_and_vertical():  # camera.rs:173-182
    assert Camera(200, 125, PI / 2.0).pixel_size == 0.009999999999999998
    assert Camera(125, 200, PI / 2.0).pixel_size == 0.009999999999999998


def test_ray_through_canvas_center_and_corner(backend):  # camera.rs:185-201
    o = oracle_for(Sphere())
    cam = Camera(201, 101, PI / 2.0)
    origin, direction = o.ray_for_pixel(cam, 100, 50)
    assert origin == (0.0, 0.0, 0.0) and coarse(direction, (0.0, 0.0, -1.0))
    origin, direction = o.ray_for_pixel(cam, 0, 0)
    assert origin == (0.0, 0.0, 0.0)
    assert direction == (0.6651864261194508, 0.3325932130597254, -0.6685123582500481)  # bit-exact


def test_ray_with_transformed_camera(backend):  # camera.rs:204-213
    o = oracle_for(Sphere())
    cam = Camera(201, 101, PI / 2.0)
    cam.set_transformation(P.mat_mul(P.rotation_y(PI / 4.0), P.translation(0, -2, 5)))
    origin, direction = o.ray_for_pixel(cam, 100, 50)
    assert origin == (0.0, 2.0, -5.0)
    assert coarse(direction, (S2 / 2.0, 0.0, -S2 / 2.0))


@pytest.mark.parametrize("threads", [1, 0])  # Camera::render and render_parallel
def test_rendering_default_world_11x11(backend, threads):  # camera.rs:216-249
    cam = Camera(11, 11, PI / 2.0)
    cam.set_transformation(P.view_transform((0, 0, -5), (0, 0, 0), UP))
    rgb, _, _ = make(World.default().flatten()).render(cam, threads=threads)
    assert coarse(rgb[5 + 5 * 11], (0.38066119308103435, 0.47582649135129296, 0.28549589481077575))


# ---------------------------------------------------------------------------------------------
# ray.rs / sphere.rs


@pytest.mark.parametrize(
    "origin, expected",
    [  # ray.rs:95-155
        ((0, 0, -5), [4.0, 6.0]),
        ((0, 1, -5), [5.0, 5.0]),
        ((0, 2, -5), []),
        ((0, 0, 0), [-1.0, 1.0]),
        ((0, 0, 5), [-6.0, -4.0]),
    ],
)
def test_ray_sphere(backend, origin, expected):
    assert oracle_for(Sphere()).intersect_shape(0, origin, FORWARD) == expected


def test_scaled_and_translated_sphere(backend):  # ray.rs:177-199
    s = Sphere()
    s.set_transformation(P.scaling(2, 2, 2))
    assert oracle_for(s).intersect_shape(0, (0, 0, -5), FORWARD) == [3.0, 7.0]
    s = Sphere()
    s.set_transformation(P.translation(5, 0, 0))
    assert oracle_for(s).intersect_shape(0, (0, 0, -5), FORWARD) == []


def test_sphere_normals(backend):  # sphere.rs:115-178
    o = oracle_for(Sphere())
    assert o.normal_at(0, (1, 0, 0)) == fvec(RIGHT)
    assert o.normal_at(0, (0, 1, 0)) == fvec(UP)
    assert o.normal_at(0, (0, 0, 1)) == fvec(FORWARD)
    t = math.sqrt(3.0) / 3.0
    n = o.normal_at(0, (t, t, t))
    assert n == (t, t, t)
    assert n == P.normalized(n)
    s = Sphere()
    s.set_transformation(P.translation(0, 1, 0))
    f = 0.7071067811865476  # FRAC_1_SQRT_2
    assert coarse(oracle_for(s).normal_at(0, (0.0, 1.0 + f, -f)), (0.0, f, -f))
    s = Sphere()
    s.set_transformation(P.mat_mul(P.scaling(1, 0.5, 1), P.rotation_z(PI / 5.0)))
    assert oracle_for(s).normal_at(0, (0, S2 / 2.0, -S2 / 2.0)) == (0.0, 0.9701425001453319, -0.24253562503633294)


# ---------------------------------------------------------------------------------------------
# plane.rs / cube.rs


def test_plane(backend):  # plane.rs:93-136
    o = oracle_for(Plane())
    for p in [(0, 0, 0), (10, 0, -10), (-5, 0, 150)]:
        assert o.normal_at(0, p) == fvec(UP)
    assert o.intersect_shape(0, (0, 10, 0), FORWARD) == []
    assert o.intersect_shape(0, (0, 1, 0), DOWN) == [1.0]
    assert o.intersect_shape(0, (0, -1, 0), UP) == [1.0]


@pytest.mark.parametrize(
    "origin, direction, t1, t2",
    [  # cube.rs:140-160
        ((5, 0.5, 0), LEFT, 4, 6),
        ((-5, 0.5, 0), RIGHT, 4, 6),
        ((0.5, 5, 0), DOWN, 4, 6),
        ((0.5, -5, 0), UP, 4, 6),
        ((0.5, 0, 5), BACKWARD, 4, 6),
        ((0.5, 0, -5), FORWARD, 4, 6),
        ((0, 0.5, 0), FORWARD, -1, 1),
    ],
)
def test_ray_intersects_cube(backend, origin, direction, t1, t2):
    assert oracle_for(Cube()).intersect_shape(0, origin, direction) == [float(t1), float(t2)]


@pytest.mark.parametrize(
    "origin, direction",
    [  # cube.rs:162-177 (the last case is the non-book "cube behind the ray" rule, cube.rs:79)
        ((-2, 0, 0), (0.2673, 0.5345, 0.8018)),
        ((0, -2, 0), (0.8018, 0.2673, 0.5345)),
        ((0, 0, -2), (0.5345, 0.8018, 0.2673)),
        ((2, 0, 2), BACKWARD),
        ((0, 2, 2), DOWN),
        ((2, 2, 0), LEFT),
        ((0, 0, 2), (0, 0, 1)),
    ],
)
def test_ray_misses_cube(backend, origin, direction):
    assert oracle_for(Cube()).intersect_shape(0, origin, direction) == []


@pytest.mark.parametrize(
    "point, normal",
    [  # cube.rs:179-193: un-normalised (p.x, 0, 0) etc.
        ((1, 0.5, -0.8), RIGHT),
        ((-1, -0.2, 0.9), LEFT),
        ((-0.4, 1, -0.1), UP),
        ((0.3, -1, -0.7), DOWN),
        ((-0.6, 0.3, 1), FORWARD),
        ((0.4, 0.4, -1), BACKWARD),
        ((1, 1, 1), RIGHT),
        ((-1, -1, -1), LEFT),
    ],
)
def test_cube_local_normal(backend, point, normal):
    assert oracle_for(Cube()).local_normal_at(0, point) == fvec(normal)


# ---------------------------------------------------------------------------------------------
# cylinder.rs


@pytest.mark.parametrize("origin, direction", [((1, 0, 0), UP), ((0, 1, 0), UP), ((0, 0, -5), (1, 1, 1))])
def test_ray_misses_cylinder(backend, origin, direction):  # cylinder.rs:184-195
    assert oracle_for(Cylinder()).intersect_shape(0, origin, P.normalized(direction)) == []


@pytest.mark.parametrize(
    "origin, direction, t1, t2",
    [  # cylinder.rs:197-215 — third row is bit-exact and discriminates the fma sites
        ((1, 0, -5), FORWARD, 5.0, 5.0),
        ((0, 0, -5), FORWARD, 4.0, 6.0),
        ((0.5, 0, -5), (0.1, 1, 1), 6.807981917027314, 7.088723439378867),
    ],
)
def test_ray_intersects_cylinder(backend, origin, direction, t1, t2):
    assert oracle_for(Cylinder()).intersect_shape(0, origin, P.normalized(direction)) == [t1, t2]


@pytest.mark.parametrize(
    "point, normal", [((1, 0, 0), RIGHT), ((0, 5, -1), BACKWARD), ((0, -2, 1), FORWARD), ((-1, 1, 0), LEFT)]
)
def test_cylinder_normal(backend, point, normal):  # cylinder.rs:217-226
    assert oracle_for(Cylinder()).local_normal_at(0, point) == fvec(normal)


@pytest.mark.parametrize(
    "origin, direction, count",
    [  # cylinder.rs:228-250
        ((0, 1.5, 0), (0.1, 1, 0), 0),
        ((0, 3, -5), FORWARD, 0),
        ((0, 0, -5), FORWARD, 0),
        ((0, 2, -5), FORWARD, 0),
        ((0, 1, -5), FORWARD, 0),
        ((0, 1.5, -2), FORWARD, 2),
    ],
)
def test_constrained_cylinder(backend, origin, direction, count):
    c = Cylinder(min=1.0, max=2.0)
    assert len(oracle_for(c).intersect_shape(0, origin, P.normalized(direction))) == count


@pytest.mark.parametrize(
    "origin, direction, count",
    [  # cylinder.rs:252-272
        ((0, 3, 0), DOWN, 2),
        ((0, 3, -2), (0, -1, 2), 2),
        ((0, 4, -2), (0, -1, 1), 2),
        ((0, 0, -2), (0, 1, 2), 2),
        ((0, -1, -2), (0, 1, 1), 2),
    ],
)
def test_closed_cylinder_caps(backend, origin, direction, count):
    c = Cylinder(min=1.0, max=2.0, closed=True)
    assert len(oracle_for(c).intersect_shape(0, origin, P.normalized(direction))) == count


@pytest.mark.parametrize(
    "point, normal",
    [((0, 1, 0), DOWN), ((0.5, 1, 0), DOWN), ((0, 1, 0.5), DOWN), ((0, 2, 0), UP), ((0.5, 2, 0), UP), ((0, 2, 0.5), UP)],
)
def test_cylinder_cap_normals(backend, point, normal):  # cylinder.rs:274-287
    c = Cylinder(min=1.0, max=2.0, closed=True)
    assert oracle_for(c).local_normal_at(0, point) == fvec(normal)


# ---------------------------------------------------------------------------------------------
# cone.rs


@pytest.mark.parametrize(
    "origin, direction, t1, t2",
    [  # cone.rs:191-209 — bit-exact; rows 2 and 3 discriminate the fma sites
        ((0, 0, -5), FORWARD, 5.0, 5.0),
        ((0, 0, -5), (1, 1, 1), 8.660254015492644, 8.660254060196127),
        ((1, 1, -5), (-0.5, -1, 1), 4.550055679356354, 49.44994432064365),
    ],
)
def test_ray_intersects_cone(backend, origin, direction, t1, t2):
    assert oracle_for(Cone()).intersect_shape(0, origin, P.normalized(direction)) == [t1, t2]


def test_cone_parallel_to_half(backend):  # cone.rs:211-220
    assert oracle_for(Cone()).intersect_shape(0, (0, 0, -1), P.normalized((0, 1, 1))) == [0.3535533905932738]


@pytest.mark.parametrize(
    "origin, direction, count", [((0, 0, -5), UP, 0), ((0, 0, -0.25), (0, 1, 1), 2), ((0, 0, -0.25), UP, 4)]
)
def test_cone_caps(backend, origin, direction, count):  # cone.rs:222-242
    c = Cone(min=-0.5, max=0.5, closed=True)
    assert len(oracle_for(c).intersect_shape(0, origin, P.normalized(direction))) == count


@pytest.mark.parametrize(
    "point, normal", [((0, 0, 0), (0, 0, 0)), ((1, 1, 1), (1, -S2, 1)), ((-1, -1, 0), (-1, 1, 0))]
)
def test_cone_normal(backend, point, normal):  # cone.rs:244-252
    got = oracle_for(Cone()).local_normal_at(0, point)
    assert got == fvec(normal)


# ---------------------------------------------------------------------------------------------
# triangle.rs


def _tri():
    return Triangle((0, 1, 0), (-1, 0, 0), (1, 0, 0))


def test_triangle_construction():  # triangle.rs:125-137
    t = _tri()
    assert t.edge_1 == (-1.0, -1.0, 0.0) and t.edge_2 == (1.0, -1.0, 0.0) and t.normal == (0.0, 0.0, -1.0)


def test_triangle_normal_and_intersections(backend):  # triangle.rs:140-223
    o = oracle_for(_tri())
    for p in [(0, 0.5, 0), (-0.5, 0.75, 0), (0.5, 0.25, 0)]:
        assert o.local_normal_at(0, p) == (0.0, 0.0, -1.0)
    assert o.intersect_shape(0, (0, -1, -2), UP) == []  # parallel
    assert o.intersect_shape(0, (1, 1, -2), FORWARD) == []  # p1-p3 edge
    assert o.intersect_shape(0, (-1, 1, -2), FORWARD) == []  # p1-p2 edge
    assert o.intersect_shape(0, (0, -1, -2), FORWARD) == []  # p2-p3 edge
    assert o.intersect_shape(0, (0, 0.5, -2), FORWARD) == [2.0]


# ---------------------------------------------------------------------------------------------
# intersections.rs / intersection.rs / computed_hit.rs


@pytest.mark.parametrize(
    "ts, expected",
    [([1, 2], 1.0), ([-1, 1], 1.0), ([-2, -1], None), ([5, 7, -3, 2], 2.0)],
)
def test_hit_rules(backend, ts, expected):  # intersections.rs:92-141
    o = oracle_for(Sphere())
    h = o.prepare_computations((0, 0, -5), FORWARD, k=-1, xs=[(t, 0) for t in ts])
    assert (h.distance if h else None) == expected


def test_prepare_computations_basics(backend):  # intersection.rs:122-158
    o = oracle_for(Sphere())
    h = o.prepare_computations((0, 0, -5), FORWARD, k=0, xs=[(4, 0)])
    assert h.distance == 4.0 and tuple(h.point) == (0, 0, -1)
    assert tuple(h.camera_direction) == (0, 0, -1) and tuple(h.normal) == (0, 0, -1) and not h.is_inside
    h = o.prepare_computations((0, 0, 0), FORWARD, k=0, xs=[(1, 0)])
    assert h.is_inside and tuple(h.point) == (0, 0, 1) and tuple(h.camera_direction) == (0, 0, -1)


def test_over_and_under_point(backend):  # intersection.rs:161-171, 242-254
    s = Sphere(Material.glass())
    s.set_transformation(P.translation(0, 0, 1))
    o = oracle_for(s)
    h = o.prepare_computations((0, 0, -5), FORWARD, k=0, xs=[(5, 0)])
    assert h.over_point[2] < -EPS / 2.0 and h.point[2] > h.over_point[2]
    assert h.under_point[2] > EPS / 2.0 and h.point[2] < h.under_point[2]


def test_reflection_vector(backend):  # intersection.rs:174-188 (bit-exact)
    o = oracle_for(Plane())
    h = o.prepare_computations((0, 1, -1), (0, -S2 / 2.0, S2 / 2.0), k=0, xs=[(S2, 0)])
    assert tuple(h.reflect_direction) == (0.0, S2 / 2.0, S2 / 2.0)


def _three_glass_spheres():
    a = Sphere(Material(transparency=1.0, refractive_index=1.5))
    a.set_transformation(P.scaling(2, 2, 2))
    b = Sphere(Material(transparency=1.0, refractive_index=2.0))
    b.set_transformation(P.translation(0, 0, -0.25))
    c = Sphere(Material(transparency=1.0, refractive_index=2.5))
    c.set_transformation(P.translation(0, 0, 0.25))
    return a, b, c


def test_refractive_indexes_hand_built_list(backend):  # intersection.rs:191-239
    o = oracle_for(*_three_glass_spheres())
    xs = [(2, 0), (2.75, 1), (3.25, 2), (4.75, 1), (5.25, 2), (6, 0)]
    n1 = [1.0, 1.5, 2.0, 2.5, 2.5, 1.5]
    n2 = [1.5, 2.0, 2.5, 2.5, 1.5, 1.0]
    for k in range(6):
        h = o.prepare_computations((0, 0, -4), FORWARD, k=k, xs=xs)
        assert (h.refractive_index_1, h.refractive_index_2) == (n1[k], n2[k])


def test_refractive_indexes_world_list(backend):
    """Same scene, but the list the world itself collects and sorts (world.rs:25-35)."""
    o = oracle_for(*_three_glass_spheres())
    xs = o.collect_intersections((0, 0, -4), FORWARD)
    assert [x[0] for x in xs] == [2.0, 2.75, 3.25, 4.75, 5.25, 6.0]
    assert [x[1] for x in xs] == [0, 1, 2, 1, 2, 0]
    n1 = [1.0, 1.5, 2.0, 2.5, 2.5, 1.5]
    n2 = [1.5, 2.0, 2.5, 2.5, 1.5, 1.0]
    for k in range(6):
        h = o.prepare_computations((0, 0, -4), FORWARD, k=k)
        assert (h.refractive_index_1, h.refractive_index_2) == (n1[k], n2[k])


def test_schlick(backend):  # computed_hit.rs:79-124
    o = oracle_for(Sphere(Material.glass()))
    h = o.prepare_computations((0, 0, S2 / 2.0), UP, k=1, xs=[(-S2 / 2.0, 0), (S2 / 2.0, 0)])
    assert h.schlick == 1.0
    h = o.prepare_computations((0, 0, 0), UP, k=1, xs=[(-1, 0), (1, 0)])
    assert h.schlick == 0.04000000000000001  # bit-exact: pins powi(5) order + the final fma
    h = o.prepare_computations((0, 0.99, -2), FORWARD, k=0, xs=[(1.8589, 0)])
    assert abs(h.schlick - 0.4887308101221217) < EPS


# ---------------------------------------------------------------------------------------------
# material.rs


@pytest.mark.parametrize(
    "eye, light, in_shadow, expected",
    [  # material.rs:201-282, all bit-exact
        (BACKWARD, (0, 0, -10), False, 1.9),
        ((0, S2 / 2.0, -S2 / 2.0), (0, 0, -10), False, 1.0),
        (BACKWARD, (0, 10, -10), False, 0.7363961030678927),
        ((0, -S2 / 2.0, -S2 / 2.0), (0, 10, -10), False, 1.6363961030678928),
        (BACKWARD, (0, 0, 10), False, 0.1),
        (BACKWARD, (0, 0, -10), True, 0.1),
    ],
)
def test_lighting(backend, eye, light, in_shadow, expected):
    o = oracle_for(Sphere())
    got = o.lighting(0, 0, light, WHITE, (0, 0, 0), eye, BACKWARD, in_shadow)
    assert got == (expected, expected, expected)


# ---------------------------------------------------------------------------------------------
# patterns


def _pattern_oracle(pattern, shape=None):
    shape = shape or Sphere()
    shape.material = Material(pattern=pattern)
    return oracle_for(shape)


def test_stripe_pattern(backend):  # stripe_pattern.rs:71-96
    o = _pattern_oracle(StripePattern(WHITE, BLACK))
    for p, c in [((0, 0, 0), WHITE), ((0.9, 0, 0), WHITE), ((1, 0, 0), BLACK), ((-0.1, 0, 0), BLACK),
                 ((-1, 0, 0), BLACK), ((-1.1, 0, 0), WHITE), ((0, 1, 0), WHITE), ((0, 2, 0), WHITE),
                 ((0, 0, 1), WHITE), ((0, 0, 2), WHITE)]:
        assert o.pattern_at_shape(0, 0, p) == c, p


def test_lighting_with_stripe_pattern(backend):  # stripe_pattern.rs:98-129
    s = Sphere(Material(pattern=StripePattern(WHITE, BLACK), ambient=1.0, diffuse=0.0, specular=0.0))
    o = oracle_for(s)
    assert o.lighting(0, 0, (0, 10, -10), WHITE, (0.9, 0, 0), BACKWARD, BACKWARD, False) == WHITE
    assert o.lighting(0, 0, (0, 10, -10), WHITE, (1.1, 0, 0), BACKWARD, BACKWARD, False) == BLACK


def test_stripe_with_transformations(backend):  # stripe_pattern.rs:131-157
    s = Sphere()
    s.set_transformation(P.scaling(2, 2, 2))
    assert _pattern_oracle(StripePattern(WHITE, BLACK), s).pattern_at_shape(0, 0, (1.5, 0, 0)) == WHITE
    pat = StripePattern(WHITE, BLACK)
    pat.set_transformation(P.scaling(2, 2, 2))
    assert _pattern_oracle(pat).pattern_at_shape(0, 0, (1.5, 0, 0)) == WHITE
    s = Sphere()
    s.set_transformation(P.scaling(2, 2, 2))
    pat = StripePattern(WHITE, BLACK)
    pat.set_transformation(P.translation(0.5, 0, 0))
    assert _pattern_oracle(pat, s).pattern_at_shape(0, 0, (2.5, 0, 0)) == WHITE


def test_test_pattern_transformations(backend):  # pattern.rs:99-125
    s = Sphere()
    s.set_transformation(P.scaling(2, 2, 2))
    assert _pattern_oracle(TestPattern(), s).pattern_at_shape(0, 0, (2, 3, 4)) == (1.0, 1.5, 2.0)
    pat = TestPattern()
    pat.set_transformation(P.scaling(2, 2, 2))
    assert _pattern_oracle(pat).pattern_at_shape(0, 0, (2, 3, 4)) == (1.0, 1.5, 2.0)
    s = Sphere()
    s.set_transformation(P.scaling(2, 2, 2))
    pat = TestPattern()
    pat.set_transformation(P.translation(0.5, 1, 1.5))
    assert _pattern_oracle(pat, s).pattern_at_shape(0, 0, (2.5, 3, 3.5)) == (0.75, 0.5, 0.25)


def test_gradient_pattern(backend):  # gradient_pattern.rs:67-84
    o = _pattern_oracle(GradientPattern(WHITE, BLACK))
    for x, v in [(0, 1.0), (0.25, 0.75), (0.5, 0.5), (0.75, 0.25), (1, 0.0)]:
        assert o.pattern_at_shape(0, 0, (x, 0, 0)) == (v, v, v)


def test_ring_pattern(backend):  # ring_pattern.rs:68-75
    o = _pattern_oracle(RingPattern(WHITE, BLACK))
    for p, c in [((0, 0, 0), WHITE), ((1, 0, 0), BLACK), ((0, 0, 1), BLACK), ((0.708, 0, 0.708), BLACK)]:
        assert o.pattern_at_shape(0, 0, p) == c


def test_checker_pattern(backend):  # checker_pattern.rs:67-89
    o = _pattern_oracle(CheckerPattern(WHITE, BLACK))
    for axis in range(3):
        for v, c in [(0, WHITE), (0.99, WHITE), (1.01, BLACK)]:
            p = [0.0, 0.0, 0.0]
            p[axis] = v
            assert o.pattern_at_shape(0, 0, p) == c


# ---------------------------------------------------------------------------------------------
# world.rs


def _default_world():
    return World.default()


def test_intersect_default_world(backend):  # world.rs:249-260
    xs = make(_default_world().flatten()).collect_intersections((0, 0, -5), FORWARD)
    assert [x[0] for x in xs] == [4.0, 4.5, 5.5, 6.0]


def test_shading_intersection(backend):  # world.rs:263-277
    o = make(_default_world().flatten())
    c = shade(o, (0, 0, -5), FORWARD, 0, 1, "shade_hit", xs=[(4.0, 0)])
    assert coarse(c, (0.38066119308103435, 0.47582649135129296, 0.28549589481077575))


def test_shading_intersection_from_inside(backend):  # world.rs:280-292
    w = _default_world()
    w.lights = [Light((0, 0.25, 0), WHITE)]
    c = shade(make(w.flatten()), (0, 0, 0), FORWARD, 0, 1, "shade_hit", xs=[(0.5, 1)])
    assert coarse(c, (0.9049844720832575,) * 3)


def test_color_at_miss_and_hit(backend):  # world.rs:295-315
    o = make(_default_world().flatten())
    assert o.color_at((0, 0, -5), UP) == BLACK
    assert coarse(o.color_at((0, 0, -5), FORWARD), (0.38066119308103435, 0.47582649135129296, 0.28549589481077575))


def test_color_with_intersection_behind_ray(backend):  # world.rs:318-334
    w = _default_world()
    w.shapes[0].material.ambient = 1.0
    w.shapes[1].material.ambient = 1.0
    assert make(w.flatten()).color_at((0, 0, 0.75), BACKWARD) == w.shapes[1].material.color


@pytest.mark.parametrize(
    "point, shadowed", [((0, 10, 0), False), ((-20, 20, -20), False), ((-2, 2, -2), False), ((10, -10, 10), True)]
)
def test_is_in_shadow(backend, point, shadowed):  # world.rs:337-366
    assert make(_default_world().flatten()).is_in_shadow(0, point) is shadowed


def test_shade_hit_in_shadow(backend):  # world.rs:369-383 (bit-exact 0.1)
    w = _default_world()
    w.lights = [Light((0, 0, -10), WHITE)]
    w.shapes.append(Sphere())
    s = Sphere()
    s.set_transformation(P.translation(0, 0, 10))
    w.shapes.append(s)
    c = shade(make(w.flatten()), (0, 0, 5), FORWARD, 0, 1, "shade_hit", xs=[(4, 3)])
    assert c == (0.1, 0.1, 0.1)


def _world_with_reflective_plane():
    w = _default_world()
    plane = Plane(Material(reflectiveness=0.5), P.translation(0, -1, 0))
    w.shapes.append(plane)
    return w


def test_reflected_color(backend):  # world.rs:386-427
    w = _default_world()
    w.shapes[0] = Sphere(Material(ambient=1.0), P.scaling(0.5, 0.5, 0.5))
    c = shade(make(w.flatten()), (0, 0, 0), FORWARD, 0, 1, "reflected", xs=[(1.0, 1)])
    assert c == BLACK
    o = make(_world_with_reflective_plane().flatten())
    ray = ((0, 0, -3), (0, -S2 / 2.0, S2 / 2.0))
    c = shade(o, *ray, 0, 1, "reflected", xs=[(S2, 2)])
    assert coarse(c, (0.19033061377890123, 0.23791326722362655, 0.14274796033417592))


def test_shade_hit_with_reflective_material(backend):  # world.rs:430-447
    o = make(_world_with_reflective_plane().flatten())
    c = shade(o, (0, 0, -3), (0, -S2 / 2.0, S2 / 2.0), 0, 1, "shade_hit", xs=[(S2, 2)])
    assert coarse(c, (0.8767560027604027, 0.9243386562051279, 0.8291733493156773))


def test_no_infinite_recursion(backend):  # world.rs:450-466
    lower = Plane(Material(reflectiveness=1.0), P.translation(0, -1, 0))
    upper = Plane(Material(reflectiveness=1.0), P.translation(0, 1, 0))
    o = oracle_for(lower, upper, lights=[Light((0, 0, 0), WHITE)])
    c = o.color_at((0, 0, 0), UP)
    assert all(math.isfinite(v) for v in c)


def test_reflected_color_at_max_depth(backend):  # world.rs:469-487
    o = make(_world_with_reflective_plane().flatten())
    assert shade(o, (0, 0, -3), (0, -S2 / 2.0, S2 / 2.0), 0, 0, "reflected", xs=[(S2, 2)]) == BLACK


def test_refracted_color_cutoffs(backend):  # world.rs:490-544
    o = make(_default_world().flatten())
    assert shade(o, (0, 0, -5), FORWARD, 0, 5, "refracted", xs=[(4, 0), (6, 0)]) == BLACK  # opaque
    w = _default_world()
    w.shapes[0].material.transparency = 1.0
    w.shapes[0].material.refractive_index = 1.5
    o = make(w.flatten())
    assert shade(o, (0, 0, -5), FORWARD, 0, 0, "refracted", xs=[(4, 0), (6, 0)]) == BLACK  # depth 0
    xs = [(-S2 / 2.0, 0), (S2 / 2.0, 0)]
    assert shade(o, (0, 0, S2 / 2.0), UP, 1, 5, "refracted", xs=xs) == BLACK  # total internal reflection


def test_refracted_color_with_refracted_ray(backend):  # world.rs:547-571 (needs TestPattern)
    w = _default_world()
    w.shapes[0].material.ambient = 1.0
    w.shapes[0].material.pattern = TestPattern()
    w.shapes[1].material.transparency = 1.0
    w.shapes[1].material.refractive_index = 1.5
    xs = [(-0.9899, 0), (-0.4899, 1), (0.4899, 1), (0.9899, 0)]
    c = shade(make(w.flatten()), (0, 0, 0.1), UP, 2, 5, "refracted", xs=xs)
    assert coarse(c, (0.0, 0.9988846813665367, 0.04721645191320928))


def _world_with_floor_and_ball(floor_material):
    w = _default_world()
    w.shapes.append(Plane(floor_material, P.translation(0, -1, 0)))
    ball = Sphere(Material(color=(1.0, 0.0, 0.0), ambient=0.5))
    ball.set_transformation(P.translation(0, -3.5, -0.5))
    w.shapes.append(ball)
    return w


def test_shade_hit_with_transparent_material(backend):  # world.rs:574-599
    w = _world_with_floor_and_ball(Material(transparency=0.5, refractive_index=1.5))
    c = shade(make(w.flatten()), (0, 0, -3), (0, -S2 / 2.0, S2 / 2.0), 0, 5, "shade_hit", xs=[(S2, 2)])
    assert coarse(c, (0.9364253889815014, 0.6864253889815014, 0.6864253889815014))


def test_shade_hit_with_reflective_and_transparent_material(backend):  # world.rs:602-629 (Schlick)
    w = _world_with_floor_and_ball(Material(reflectiveness=0.5, transparency=0.5, refractive_index=1.5))
    c = shade(make(w.flatten()), (0, 0, -3), (0, -S2 / 2.0, S2 / 2.0), 0, 5, "shade_hit", xs=[(S2, 2)])
    assert coarse(c, (0.9339151412754023, 0.696434227200244, 0.692430691912747))


# ---------------------------------------------------------------------------------------------
# canvas.rs quantisation


def test_quantise(backend):  # canvas.rs:117-123
    o = oracle_for(Sphere())
    assert o.quantise(0.0) == 0 and o.quantise(1.0) == 255 and o.quantise(1.5) == 255 and o.quantise(-0.5) == 0
    assert o.quantise(0.5) == 128  # 127.5 rounds half away from zero
    assert o.quantise(float("nan")) == 0
    assert o.quantise(0.3) == 77 and o.quantise(0.7) == 179  # 76.5 -> 77 (half away), 178.5 -> 179
