#!/usr/bin/env python
"""What one of N GPUs does in an N-way row-band split, timed on ONE GPU: renders shard 0 of N of the
1080p cover frame (kernel time).  Shows the strong-scaling floor without needing N devices."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ray_tracer_challenge_rs_b200.fixtures import load_scene_fixture  # noqa: E402
from ray_tracer_challenge_rs_b200.render import Renderer  # noqa: E402

flat, camera = load_scene_fixture(sys.argv[1] if len(sys.argv) > 1 else "cover")
family = sys.argv[2] if len(sys.argv) > 2 else None  # persistent | wavefront | (default) the library's measured choice
shards = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 4, 8, 16]
cam = camera.resized(1920, 1080)
out = []
with Renderer(flat) as r:
    for n in shards:
        best = 1e9
        for _ in range(6):
            _, _, st = r.render(cam, rows=(16, 0, n) if n > 1 else None, want_rgb8=False, family=family)
            best = min(best, st["kernel_ms"])
        out.append(f"1/{n}: {best:.3f} ms")
print("  ".join(out))
