#!/usr/bin/env python
"""Probe: one wavefront-family frame of a synthetic scene.  wf_probe.py <shapes> <width> <height> <max_depth> [family]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ray_tracer_challenge_rs_b200.render import Renderer  # noqa: E402
from ray_tracer_challenge_rs_b200.synthetic import synthetic_camera, synthetic_scene  # noqa: E402

n, w, h, depth = (int(v) for v in sys.argv[1:5])
family = sys.argv[5] if len(sys.argv) > 5 else "wavefront"
flat = synthetic_scene(n)
cam = synthetic_camera(w, h)
with Renderer(flat) as r:
    for _ in range(2):
        t0 = time.perf_counter()
        _, _, st = r.render(cam, want_rgb8=False, max_depth=depth, family=family)
        print(n, w, h, depth, family, "kernel_ms", round(st["kernel_ms"], 3), "wall_s", round(time.perf_counter() - t0, 3), "rays", st["rays"], flush=True)
