#!/usr/bin/env python
"""A shipped scene at a very large frame size through both kernel families: same bits, memory holds.

    python benchmarks/large_frame.py [scene] [width] [height]      (default: cover 7680 4320)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from ray_tracer_challenge_rs_b200.fixtures import load_scene_fixture  # noqa: E402
from ray_tracer_challenge_rs_b200.render import Renderer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cover"
w = int(sys.argv[2]) if len(sys.argv) > 2 else 7680
h = int(sys.argv[3]) if len(sys.argv) > 3 else 4320
flat, camera = load_scene_fixture(name)
cam = camera.resized(w, h)
frames = {}
with Renderer(flat) as r:
    for family in ("persistent", "wavefront"):
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            rgb, _, st = r.render(cam, want_rgb8=False, family=family)
            best = st["kernel_ms"] if best is None else min(best, st["kernel_ms"])
        frames[family] = rgb
        print(f"{name} {w}x{h} {family}: kernel {best:.2f} ms, {st['rays'] / best / 1e3:.0f} Mrays/s, wall {time.perf_counter() - t0:.2f} s", flush=True)
same = np.array_equal(frames["persistent"].view(np.uint64), frames["wavefront"].view(np.uint64))
print("frames bit-identical:", same)
sys.exit(0 if same else 1)
