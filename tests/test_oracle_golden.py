"""Pins the CPU oracle against the reference's own full-frame renders (rendered_images/*.png).

The fixtures (tests/golden/, scenes/) were generated from the reference tree by
tests/golden/make_golden.py: the flattened scenes, the sha256 of each golden RGB8 frame and every
64th golden row.  A render whose sha256 equals the golden's is byte-identical on every pixel.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle.oracle import Oracle
from ray_tracer_challenge_rs_b200.fixtures import SHIPPED_SCENES, load_scene_fixture

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "golden_index.json")) as f:
    INDEX = json.load(f)


@pytest.mark.slow
@pytest.mark.parametrize("name", SHIPPED_SCENES)
def test_oracle_matches_golden_png_full_frame(oracle_lib, name):
    flat, camera = load_scene_fixture(name)
    meta = INDEX[name]
    assert (camera.horizontal_size, camera.vertical_size) == (meta["width"], meta["height"])
    _, rgb8, stats = Oracle(flat).render(camera, max_depth=6, threads=0, want_rgb=False)
    frame = rgb8.reshape(meta["height"], meta["width"], 3)
    # sampled rows first: a failure here can be localised pixel by pixel
    with np.load(os.path.join(GOLDEN, f"{name}.rows.npz")) as z:
        rows, golden_rows = z["rows"], z["rgb8"]
    diff = np.abs(frame[rows].astype(np.int16) - golden_rows.astype(np.int16)).max(axis=2)
    bad = np.argwhere(diff != 0)
    assert bad.shape[0] == 0, f"{name}: {bad.shape[0]} mismatching sampled pixels, first (row, x): {[(int(rows[r]), int(x)) for r, x in bad[:10]]}"
    # then the whole frame through its checksum
    assert hashlib.sha256(frame.tobytes()).hexdigest() == meta["sha256_rgb8"]
    assert stats["rays_primary"] == meta["width"] * meta["height"]


def test_depth_5_is_not_the_parity_setting(oracle_lib):
    """MAX_REFLECTION_ITERATIONS = 6 (world.rs:15); at depth 5 refraction.yaml no
    longer reproduces its golden."""
    flat, camera = load_scene_fixture("refraction")
    with np.load(os.path.join(GOLDEN, "refraction.rows.npz")) as z:
        rows, golden_rows = z["rows"], z["rgb8"]
    row = int(rows[len(rows) // 2])
    pixels = np.arange(row * camera.horizontal_size, (row + 1) * camera.horizontal_size, dtype=np.uint64)
    o = Oracle(flat)
    _, rgb8_6, _ = o.render_pixels(camera, pixels, max_depth=6)
    _, rgb8_5, _ = o.render_pixels(camera, pixels, max_depth=5)
    golden = golden_rows[len(rows) // 2]
    assert (rgb8_6 == golden).all()
    assert (rgb8_5 != golden).any()
