#!/usr/bin/env python
"""Small frames through every kernel path — both families, both precisions, binned queues forced on, the
queue-overflow path, a BVH scene, a FULL scene (cylinders), row-band shards, the host render into pinned memory —
checked against each other.  Meant for a library built with -DRT_BOUNDS_CHECK=1 (profiles/tools/bounds_check.sh):
every table, queue, node and permutation index is then range-checked on the device and a bad one traps.
(compute-sanitizer is not available on this pool's GPU boxes.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from ray_tracer_challenge_rs_b200.fixtures import load_scene_fixture  # noqa: E402
from ray_tracer_challenge_rs_b200.render import PinnedArray, Renderer  # noqa: E402

small = len(sys.argv) > 1 and sys.argv[1] == "small"
w, h = (64, 36) if small else (160, 90)


def frames(scene, **env):
    for k, v in env.items():
        os.environ[k] = v
    flat, camera = load_scene_fixture(scene)
    cam = camera.resized(w, h)
    with Renderer(flat) as r:
        ref, ref8, _ = r.render(cam, family="persistent")
        for precision in ("f64", "f32"):
            for family in ("wavefront", "persistent"):
                a, a8, st = r.render(cam, family=family, precision=precision)
                if precision == "f64":
                    assert np.array_equal(a.view(np.uint64), ref.view(np.uint64)) and np.array_equal(a8, ref8), (scene, family)
        r.render(cam, family="wavefront", rows=(4, 1, 3))
        pin = PinnedArray((w * h, 3), np.float64)
        r.render(cam, family="wavefront", out_rgb=pin.array)
        assert np.array_equal(pin.array.view(np.uint64), ref.view(np.uint64))
        pin.close()
    for k in env:
        del os.environ[k]
    print("ok", scene, env, flush=True)


frames("cover", RTGPU_WF_BINS="1")
frames("cover", RTGPU_WF_BINS="1", RTGPU_WF_INITIAL_SCALE="0.02")
frames("cylinders", RTGPU_WF_BINS="1")
frames("refraction")
if not small:
    from ray_tracer_challenge_rs_b200.synthetic import synthetic_camera, synthetic_scene  # noqa: E402

    flat, cam = synthetic_scene(3000), synthetic_camera(w, h)
    with Renderer(flat) as r:
        a, _, _ = r.render(cam, family="wavefront")
        b, _, _ = r.render(cam, family="persistent")
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    print("ok synthetic 3000 shapes (BVH)", flush=True)
