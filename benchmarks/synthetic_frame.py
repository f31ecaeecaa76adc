#!/usr/bin/env python
"""One synthetic scaling scene (BASELINE.json configs[4]) rendered a few times on one GPU: kernel ms and Mrays/s.

    python benchmarks/synthetic_frame.py <n_shapes> <persistent|wavefront|auto> [width height] [frames]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ray_tracer_challenge_rs_b200.render import Renderer  # noqa: E402
from ray_tracer_challenge_rs_b200.synthetic import synthetic_camera, synthetic_scene  # noqa: E402

n = int(sys.argv[1])
family = None if len(sys.argv) < 3 or sys.argv[2] == "auto" else sys.argv[2]
w, h = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (7680, 4320)
frames = int(sys.argv[5]) if len(sys.argv) > 5 else 3
flat = synthetic_scene(n)
cam = synthetic_camera(w, h)
with Renderer(flat) as r:
    best = None
    for _ in range(frames):
        _, _, st = r.render(cam, want_rgb8=False, family=family)
        best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
print(n, best["family"], w, h, "kernel_ms", round(best["kernel_ms"], 2), "rays", best["rays"], "Mrays/s", round(best["rays"] / best["kernel_ms"] / 1e3))
