"""Hand-made worlds that exercise what no shipped scene does (cones,
triangles, gradient / ring / complex patterns, value-equal duplicate shapes, several lights,
casts_shadow = false, nested transparent shapes)."""
import math

from ray_tracer_challenge_rs_b200 import (
    Camera,
    CheckerPattern,
    ComplexPattern,
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
    Triangle,
    World,
)
from ray_tracer_challenge_rs_b200 import primitives as P


def _cam(w, h, fov, frm, to, up=(0, 1, 0)):
    cam = Camera(w, h, fov)
    cam.set_transformation(P.view_transform(frm, to, up))
    return cam


def all_shapes_world(w=160, h=96):
    """One of every shape type, every pattern type, two lights, reflective + transparent parts."""
    floor = Plane(Material(pattern=CheckerPattern((0.9, 0.9, 0.9), (0.2, 0.2, 0.25)), reflectiveness=0.3, specular=0.1))
    back = Plane(Material(pattern=RingPattern((0.8, 0.3, 0.3), (0.3, 0.3, 0.8)), specular=0.0),
                 P.mat_mul(P.translation(0, 0, 7), P.rotation_x(math.pi / 2)))
    ball = Sphere(Material(color=(0.9, 0.2, 0.2), reflectiveness=0.4, shininess=50.0), P.translation(-2.2, 1, 0.5))
    glass = Sphere(Material(color=(0.1, 0.1, 0.1), diffuse=0.1, reflectiveness=0.9, transparency=0.9, refractive_index=1.5,
                            shininess=300.0), P.mat_mul(P.translation(0.2, 1.0, -1.2), P.scaling(0.8, 0.8, 0.8)))
    air = Sphere(Material(color=(0.1, 0.1, 0.1), diffuse=0.1, transparency=1.0, refractive_index=1.0000034, reflectiveness=0.5),
                 P.mat_mul(P.translation(0.2, 1.0, -1.2), P.scaling(0.4, 0.4, 0.4)))
    grad = GradientPattern((1, 0, 0), (0, 0, 1))
    grad.set_transformation(P.mat_mul(P.translation(-1, 0, 0), P.scaling(2, 1, 1)))
    cube = Cube(Material(pattern=grad, reflectiveness=0.1), P.mat_mul(P.translation(2.4, 0.7, 1.0), P.mat_mul(P.rotation_y(0.6), P.scaling(0.7, 0.7, 0.7))))
    stripes = StripePattern((1, 1, 0.2), (0.1, 0.5, 0.1))
    stripes.set_transformation(P.scaling(0.2, 0.2, 0.2))
    cyl = Cylinder(Material(pattern=stripes, specular=0.6), P.mat_mul(P.translation(-0.8, 0, 2.5), P.scaling(0.5, 1, 0.5)), min=0.0, max=2.0, closed=True)
    open_cyl = Cylinder(Material(color=(0.3, 0.8, 0.8), reflectiveness=0.2), P.mat_mul(P.translation(3.5, 0, 3.5), P.scaling(0.4, 1, 0.4)), min=0.0, max=1.5, closed=False)
    complex_ = ComplexPattern(CheckerPattern((1, 1, 1), (0, 0, 0)), StripePattern((1, 0, 0), (0, 1, 0)))
    complex_.set_transformation(P.scaling(0.3, 0.3, 0.3))
    cone = Cone(Material(pattern=complex_, shininess=20.0), P.mat_mul(P.translation(1.2, 1.2, 2.8), P.scaling(0.6, 1.2, 0.6)), min=-1.0, max=0.0, closed=True)
    dcone = Cone(Material(color=(0.8, 0.6, 0.1), casts_shadow=False), P.mat_mul(P.translation(-3.2, 1.0, 3.0), P.scaling(0.5, 1.0, 0.5)), min=-1.0, max=1.0, closed=False)
    tri = Triangle((-1.5, 0.01, -2.5), (-0.2, 1.4, -2.2), (0.9, 0.02, -2.8))
    tri.material = Material(color=(0.2, 0.9, 0.3), reflectiveness=0.2)
    tri2 = Triangle((2.5, 0.0, -1.0), (3.5, 2.0, 0.0), (4.0, 0.0, -1.5))
    tri2.material = Material(color=(0.6, 0.2, 0.9), transparency=0.5, refractive_index=1.2)
    tri2.set_transformation(P.rotation_y(-0.2))
    lights = [Light((-6, 8, -8), (0.9, 0.9, 0.9)), Light((7, 5, -3), (0.35, 0.3, 0.3))]
    world = World(lights, [floor, back, ball, glass, air, cube, cyl, open_cyl, cone, dcone, tri, tri2])
    return world, _cam(w, h, 1.0, (0.5, 2.6, -7.5), (0.3, 0.9, 0.5))


def duplicate_glass_world(w=96, h=96):
    """Value-equal duplicate shapes (shape_eq_class groups them): two identical glass spheres at the
    same place, three identical glass cubes, inside a bigger glass sphere — the refraction container
    walk (intersection.rs:33-62) must treat equal shapes as ONE entry that toggles."""
    glass = dict(color=(0.05, 0.05, 0.05), diffuse=0.2, transparency=0.9, reflectiveness=0.6)
    outer = Sphere(Material(refractive_index=1.5, **glass), P.scaling(2, 2, 2))
    twin_a = Sphere(Material(refractive_index=2.0, **glass), P.translation(0.3, 0, 0))
    twin_b = Sphere(Material(refractive_index=2.0, **glass), P.translation(0.3, 0, 0))
    cubes = [Cube(Material(refractive_index=1.2, **glass), P.mat_mul(P.translation(-0.8, 0.2, 0.2), P.scaling(0.4, 0.4, 0.4))) for _ in range(3)]
    floor = Plane(Material(pattern=CheckerPattern((1, 1, 1), (0.1, 0.1, 0.1))), P.translation(0, -2.5, 0))
    world = World([Light((-5, 8, -6), (1, 1, 1))], [outer, twin_a, cubes[0], floor, twin_b, cubes[1], cubes[2]])
    return world, _cam(w, h, 0.9, (0, 1.0, -7), (0, 0, 0))


def mirror_box_world(w=96, h=64):
    """Two facing mirrors + a ball: every radiance path runs to the recursion limit (world.rs:450-466)."""
    lower = Plane(Material(reflectiveness=1.0, diffuse=0.1, specular=0.0), P.translation(0, -1, 0))
    upper = Plane(Material(reflectiveness=1.0, diffuse=0.1, specular=0.0), P.translation(0, 1, 0))
    ball = Sphere(Material(color=(0.9, 0.7, 0.1), reflectiveness=0.5), P.scaling(0.5, 0.5, 0.5))
    world = World([Light((0, 0, -3), (1, 1, 1))], [lower, upper, ball])
    return world, _cam(w, h, 1.2, (0, 0.2, -4), (0, 0, 0))


def no_light_world(w=32, h=16):
    world = World([], [Sphere(Material(reflectiveness=0.5)), Plane(Material(), P.translation(0, -1, 0))])
    return world, _cam(w, h, 1.0, (0, 0, -5), (0, 0, 0))


def empty_world(w=16, h=8):
    return World([Light()], []), _cam(w, h, 1.0, (0, 0, -5), (0, 0, 0))


def default_world(w=11, h=11):
    return World.default(), _cam(w, h, math.pi / 2, (0, 0, -5), (0, 0, 0))


def light_on_surface_world(w=128, h=96):
    """Lights that lie exactly ON shadow-casting surfaces (a plane, a cube face): for many pixels the shadow ray
    meets that surface at exactly the light's distance.  The reference's test is strict, `0 <= t < distance`
    (intersection.rs:77-79 via world.rs:98-112): such points are lit."""
    wall = Plane(Material(color=(0.8, 0.8, 0.9), specular=0.0), P.mat_mul(P.translation(0, 0, 4), P.rotation_x(math.pi / 2)))
    floor = Plane(Material(pattern=CheckerPattern((0.9, 0.9, 0.9), (0.3, 0.3, 0.3)), specular=0.0))
    box = Cube(Material(color=(0.9, 0.4, 0.2), reflectiveness=0.2), P.mat_mul(P.translation(1.5, 1.0, 1.0), P.scaling(1, 1, 1)))
    ball = Sphere(Material(color=(0.2, 0.6, 0.9)), P.translation(-1.5, 1.0, 0.5))
    lights = [Light((0.0, 2.0, 4.0), (0.7, 0.7, 0.7)),    # on the wall z = 4
              Light((0.5, 1.0, 2.0), (0.4, 0.4, 0.4)),    # on the cube face x = 0.5 ... and on its edge region y in [0, 2]
              Light((-3.0, 0.0, -1.0), (0.3, 0.3, 0.3))]  # on the floor y = 0
    world = World(lights, [wall, floor, box, ball])
    return world, _cam(w, h, 1.1, (0.3, 2.5, -6.0), (0.0, 1.0, 1.0))


SPECIAL_WORLDS = {
    "light_on_surface": light_on_surface_world,
    "all_shapes": all_shapes_world,
    "duplicate_glass": duplicate_glass_world,
    "mirror_box": mirror_box_world,
    "no_light": no_light_world,
    "empty": empty_world,
    "default": default_world,
}
