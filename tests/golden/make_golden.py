"""Generates the committed fixtures from the reference tree (run in the build container only).

    python tests/golden/make_golden.py [/root/reference]

Outputs
  scenes/<scene>.npz              the reference's scenes/<scene>.yaml, loaded by
                                  ray_tracer_challenge_rs_b200.scene_loader and flattened
                                  (bit-exact f64 arrays + the camera); the GPU box has no
                                  /root/reference, so tests and bench.py read these.
  tests/golden/<scene>.rows.npz   every 64th row (y % 64 == 32) of rendered_images/<scene>.png, RGB8
  tests/golden/golden_index.json  per scene: size, sha256 of the full golden RGB8 frame, row list

The golden PNGs are the reference's own renders at each scene's native camera size
(SURVEY.md Appendix B); the sha256 lets a test prove a full-frame render is byte-identical to the
golden without shipping 80 M pixels.
"""
import hashlib
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from ray_tracer_challenge_rs_b200.flatten import camera_to_dict  # noqa: E402
from ray_tracer_challenge_rs_b200.scene_loader import load_scene_description  # noqa: E402

SCENES = ["three_sphere_scene", "shadow_puppets", "cylinders", "metal", "table", "reflect_refract", "refraction", "cover"]
ROW_STRIDE, ROW_PHASE = 64, 32


def main(reference: str) -> None:
    Image.MAX_IMAGE_PIXELS = None
    index = {}
    os.makedirs(os.path.join(ROOT, "scenes"), exist_ok=True)
    for name in SCENES:
        world, camera = load_scene_description(os.path.join(reference, "scenes", f"{name}.yaml"))
        flat = world.flatten()
        arrays = flat.to_arrays()
        for k, v in camera_to_dict(camera).items():
            arrays[f"camera_{k}"] = np.asarray(v)
        np.savez_compressed(os.path.join(ROOT, "scenes", f"{name}.npz"), **arrays)

        img = np.asarray(Image.open(os.path.join(reference, "rendered_images", f"{name}.png")).convert("RGB"))
        h, w, _ = img.shape
        assert (w, h) == (camera.horizontal_size, camera.vertical_size), (name, w, h)
        rows = np.arange(ROW_PHASE, h, ROW_STRIDE, dtype=np.uint32)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"{name}.rows.npz"), rows=rows, rgb8=img[rows])
        index[name] = {
            "width": w,
            "height": h,
            "sha256_rgb8": hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest(),
            "row_stride": ROW_STRIDE,
            "row_phase": ROW_PHASE,
            "shapes": flat.shape_counts(),
            "lights": flat.n_lights,
        }
        print(name, w, h, index[name]["sha256_rgb8"][:16], flat.shape_counts())
    with open(os.path.join(ROOT, "tests", "golden", "golden_index.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
