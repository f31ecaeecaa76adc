"""Host logic: the YAML loader + flattener (ray-tracer-cli/src/scene_loader.rs restated)."""
import os

import numpy as np
import pytest

from ray_tracer_challenge_rs_b200 import abi, load_scene_from_str
from ray_tracer_challenge_rs_b200 import primitives as P
from ray_tracer_challenge_rs_b200.fixtures import SHIPPED_SCENES, load_scene_fixture
from ray_tracer_challenge_rs_b200.flatten import camera_to_c
from ray_tracer_challenge_rs_b200.scene_loader import load_scene_description, parse_f64

from conftest import REFERENCE_DIR


@pytest.mark.parametrize("name", SHIPPED_SCENES)
def test_loader_reproduces_committed_fixtures(name, have_reference):
    """scenes/<name>.npz must be exactly what the loader yields from the reference YAML."""
    if not have_reference:
        pytest.skip("reference tree not present (GPU box)")
    world, camera = load_scene_description(os.path.join(REFERENCE_DIR, "scenes", f"{name}.yaml"))
    flat = world.flatten()
    fixture, fixture_camera = load_scene_fixture(name)
    for key, value in flat.to_arrays().items():
        assert np.array_equal(value, fixture.to_arrays()[key]), key
    a, b = camera_to_c(camera), camera_to_c(fixture_camera)
    assert bytes(a) == bytes(b)


def test_parse_f64():  # scene_loader.rs:378-398
    assert parse_f64(1) == 1.0
    assert parse_f64(-3) == -3.0
    assert parse_f64(0.5) == 0.5
    assert parse_f64("1e3") == 1000.0
    with pytest.raises(ValueError):
        parse_f64("abc")
    with pytest.raises(ValueError):
        parse_f64(None)


SCENE = """
- add: camera
  width: 40
  height: 20
  field-of-view: 1.2
  from: [0, 1.5, -5]
  to: [0, 1, 0]
  up: [0, 1, 0]
- add: light
  at: [-10, 10, -10]
  intensity: [1, 1, 1]
- define: base-material
  value:
    color: [0.2, 0.3, 0.4]
    reflective: 0.25
- define: shiny-material
  extend: base-material
  value:
    shininess: 50
    casts-shadow: false
- define: lift-transform
  value:
    - [translate, 0, 1, 0]
- define: big-object
  value:
    - lift-transform
    - [scale, 2, 2, 2]
- add: sphere
  material: shiny-material
  transform:
    - big-object
    - [rotate-y, 0.5]
    - [unknown-op, 1, 2, 3]
- add: cone
  min: -1
  max: 0
  closed: true
  material:
    pattern:
      type: rings
      colors: [[1, 0, 0], [0, 1, 0]]
      transform:
        - [scale, 0.25, 0.25, 0.25]
- add: cylinder
  transform:
    - [translate, 2, 0, 0]
- add: torus
- add: plane
"""


def test_loader_semantics():
    world, camera = load_scene_from_str(SCENE)
    assert (camera.horizontal_size, camera.vertical_size) == (40, 20)
    assert len(world.lights) == 1 and len(world.shapes) == 4  # `torus` silently ignored (:330)
    sphere, cone, cylinder, plane = world.shapes
    m = sphere.material
    assert m.color == (0.2, 0.3, 0.4) and m.reflectiveness == 0.25 and m.shininess == 50.0 and m.casts_shadow is False
    # named transform right-multiplied, inline ops left-multiplied (:204-233)
    lift = P.translation(0, 1, 0)
    big = P.mat_mul(P.scaling(2, 2, 2), P.mat_mul(P.identity(), lift))
    expected = P.mat_mul(P.rotation_y(0.5), P.mat_mul(P.identity(), big))
    assert sphere.transformation_inverse == P.inverse(expected)
    assert (cone.min, cone.max, cone.closed) == (-1.0, 0.0, True)
    assert cone.material.pattern.TYPE == abi.PATTERN_RING
    assert (cylinder.min, cylinder.max, cylinder.closed) == (P.F64_MIN, P.F64_MAX, False)
    assert plane.transformation_inverse == P.identity()

    flat = world.flatten()
    assert flat.n_shapes == 4 and flat.n_patterns == 1 and flat.n_materials == 3
    assert list(flat.shape_type) == [abi.SPHERE, abi.CONE, abi.CYLINDER, abi.PLANE]
    assert list(flat.shape_eq_class) == [0, 1, 2, 3]


def test_eq_class_groups_value_equal_shapes():
    from ray_tracer_challenge_rs_b200 import Material, Sphere, World

    a, b, c = Sphere(), Sphere(), Sphere(Material(ambient=0.2))
    flat = World([], [a, c, b]).flatten()
    assert list(flat.shape_eq_class) == [0, 1, 0]


def test_matrix_inverse_kat():  # primitives/matrix.rs:557-630 (bit-exact)
    m1 = P.matrix([[-5, 2, 6, -8], [1, -5, 1, 8], [7, 7, -6, -7], [1, -3, 7, 4]])
    assert P.determinant(m1) == 532.0
    assert P.cofactor(m1, 2, 3) == -160.0 and P.cofactor(m1, 3, 2) == 105.0
    inv = P.inverse(m1)
    assert inv[3][2] == -160.0 / 532.0 and inv[2][3] == 105.0 / 532.0
    assert inv == [
        [0.21804511278195488, 0.45112781954887216, 0.24060150375939848, -0.045112781954887216],
        [-0.8082706766917294, -1.4567669172932332, -0.44360902255639095, 0.5206766917293233],
        [-0.07894736842105263, -0.2236842105263158, -0.05263157894736842, 0.19736842105263158],
        [-0.5225563909774437, -0.8139097744360902, -0.3007518796992481, 0.30639097744360905],
    ]
    m3 = P.matrix([[9, 3, 0, 9], [-5, -2, 6, -3], [-4, 9, 6, 4], [-7, 6, 6, 2]])
    assert P.inverse(m3) == [
        [-0.004901960784313725, 0.10294117647058823, 0.27941176470588236, -0.38235294117647056],
        [-0.09313725490196079, -0.04411764705882353, 0.3088235294117647, -0.2647058823529412],
        [0.03839869281045752, 0.19362745098039216, 0.14460784313725492, -0.1715686274509804],
        [0.14705882352941177, -0.08823529411764706, -0.38235294117647056, 0.47058823529411764],
    ]


def test_view_transform_kat():  # primitives/transformations.rs:232-291
    assert P.view_transform((0, 0, 0), (0, 0, -1), (0, 1, 0)) == P.identity()
    assert P.view_transform((0, 0, 0), (0, 0, 1), (0, 1, 0)) == P.scaling(-1, 1, -1)
    assert P.view_transform((0, 0, 8), (0, 0, 0), (0, 1, 0)) == P.translation(0, 0, -8)
    got = P.view_transform((1, 3, 2), (4, -2, 8), (1, 1, 0))
    want = [
        [-0.5070925528371099, 0.5070925528371099, 0.6761234037828132, -2.366431913239846],
        [0.7677159338596801, 0.6060915267313263, 0.12121830534626524, -2.8284271247461894],
        [-0.35856858280031806, 0.5976143046671968, -0.7171371656006361, 0.0],
        [0.0, 0.0, 0.0, 1.0],
    ]
    assert all(P.coarse_eq(got[r][c], want[r][c]) for r in range(4) for c in range(4))
