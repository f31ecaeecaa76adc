"""The C++ host (`host/ray-tracer-cli`): its YAML-subset reader + loader + flattener must hand
rtgpu_render exactly the bytes the Python host does — for the reference's shipped scenes (when the
reference tree is present) and for a hand-written scene that is always present."""
import os
import struct
import subprocess

import numpy as np
import pytest

from ray_tracer_challenge_rs_b200 import build
from ray_tracer_challenge_rs_b200.fixtures import SHIPPED_SCENES
from ray_tracer_challenge_rs_b200.flatten import camera_to_c
from ray_tracer_challenge_rs_b200.scene_loader import load_scene_description

from conftest import REFERENCE_DIR

HERE = os.path.dirname(os.path.abspath(__file__))
SHOWCASE = os.path.join(HERE, "scenes", "showcase.yaml")


@pytest.fixture(scope="module")
def cli():
    return build.build_host()


def read_dump(path):
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        (name_len,) = struct.unpack_from("<I", data, pos)
        pos += 4
        name = data[pos:pos + name_len].decode()
        pos += name_len
        width, n = struct.unpack_from("<QQ", data, pos)
        pos += 16
        out[name] = data[pos:pos + width * n]
        pos += width * n
    return out


def assert_same_flat_scene(cli, yaml_path, tmp_path):
    dump = str(tmp_path / "scene.flat")
    subprocess.run([cli, yaml_path, "--dump-flat", dump], check=True)
    got = read_dump(dump)
    world, camera = load_scene_description(yaml_path)
    flat = world.flatten()
    for name, array in flat.to_arrays().items():
        assert got[name] == np.ascontiguousarray(array).tobytes(), name
    assert got["camera"] == bytes(camera_to_c(camera))
    return flat


def test_showcase_scene_flattens_identically(cli, tmp_path):
    flat = assert_same_flat_scene(cli, SHOWCASE, tmp_path)
    assert flat.shape_counts() == {"sphere": 2, "plane": 2, "cube": 2, "cylinder": 2, "cone": 1}
    assert flat.n_patterns == 4 and flat.n_lights == 2


@pytest.mark.parametrize("name", SHIPPED_SCENES)
def test_reference_scenes_flatten_identically(cli, tmp_path, name, have_reference):
    if not have_reference:
        pytest.skip("reference tree not present (GPU box)")
    assert_same_flat_scene(cli, os.path.join(REFERENCE_DIR, "scenes", f"{name}.yaml"), tmp_path)


def test_cpu_rendering_modes_are_refused(cli, tmp_path):
    out = str(tmp_path / "x.ppm")
    for mode in ("serial", "parallel", "bogus"):
        p = subprocess.run([cli, SHOWCASE, out, "-r", mode], capture_output=True, text=True)
        assert p.returncode == 2 and not os.path.exists(out)


def test_yaml_errors_are_reported(cli, tmp_path):
    bad = tmp_path / "bad.yaml"
    bad.write_text("- add: sphere\n  material: missing-material\n")
    p = subprocess.run([cli, str(bad), "--dump-flat", str(tmp_path / "d")], capture_output=True, text=True)
    assert p.returncode == 1 and "missing-material" in p.stderr


@pytest.mark.gpu
def test_cli_render_equals_python_host(cli, tmp_path):
    """End to end through the C++ host: PPM bytes == the Python host's PPM of the same scene."""
    out = str(tmp_path / "showcase.ppm")
    p = subprocess.run([cli, SHOWCASE, out, "--rendering-mode", "gpu", "--width", "160", "--height", "100"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "Image rendered in" in p.stdout
    world, camera = load_scene_description(SHOWCASE)
    canvas = camera.resized(160, 100).render_gpu(world)
    assert open(out).read() == canvas.to_ppm()
    png = str(tmp_path / "showcase.png")
    subprocess.run([cli, SHOWCASE, png, "-q", "--width", "160", "--height", "100"], check=True)
    from PIL import Image

    assert np.array_equal(np.asarray(Image.open(png).convert("RGB")), canvas.to_rgb8())
