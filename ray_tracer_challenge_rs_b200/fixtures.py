"""Committed scene fixtures (``scenes/<name>.npz``): the reference's shipped scenes, already
loaded and flattened (see ``tests/golden/make_golden.py``).  The GPU box has no
``/root/reference``, so tests and ``bench.py`` read scenes from here; users with YAML files call
:func:`..scene_loader.load_scene_description` instead."""
from __future__ import annotations

import os
from typing import List, Tuple

import numpy as np

from .flatten import FlatScene, camera_from_dict
from .scene import Camera

SCENES_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scenes")

SHIPPED_SCENES: List[str] = [
    "three_sphere_scene",
    "shadow_puppets",
    "cylinders",
    "metal",
    "table",
    "reflect_refract",
    "refraction",
    "cover",
]


def load_scene_fixture(name: str) -> Tuple[FlatScene, Camera]:
    path = os.path.join(SCENES_DIR, f"{name}.npz")
    with np.load(path) as z:
        flat = FlatScene.from_arrays(z)
        camera = camera_from_dict({k[len("camera_"):]: z[k] for k in z.files if k.startswith("camera_")})
    return flat, camera
