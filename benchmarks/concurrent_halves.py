#!/usr/bin/env python
"""Experiment: the frame as two interleaved halves rendered CONCURRENTLY by two contexts on two streams of one device
(tails of one pipeline's level launches overlap the other's work) vs one pipeline.  Wavefront family, device-resident.

    python benchmarks/concurrent_halves.py [scene] [shard_count_of_parent]
"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ray_tracer_challenge_rs_b200.fixtures import load_scene_fixture  # noqa: E402
from ray_tracer_challenge_rs_b200.render import Renderer  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "cover"
parent = int(sys.argv[2]) if len(sys.argv) > 2 else 1  # emulate one of `parent` GPUs
flat, camera = load_scene_fixture(scene)
cam = camera.resized(1920, 1080)
torch.cuda.set_device(0)


def bench(n_pipes, frames=12):
    rs = [Renderer(flat) for _ in range(n_pipes)]
    streams = [torch.cuda.Stream() for _ in range(n_pipes)]
    rows = [(16, parent * j, parent * n_pipes) if parent * n_pipes > 1 else None for j in range(n_pipes)]
    outs = [torch.empty((max(1, r.rows_count(cam, rw)) * 1920, 3), dtype=torch.float64, device="cuda") for r, rw in zip(rs, rows)]

    def work(j):
        rs[j].render_device(cam, outs[j].data_ptr(), 0, 0, streams[j].cuda_stream, rows=rows[j], family="wavefront")

    times = []
    for it in range(frames):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ths = [threading.Thread(target=work, args=(j,)) for j in range(n_pipes)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
    for r in rs:
        r.close()
    return min(times[2:])


for n in (1, 2, 3, 4):
    print(f"{scene} 1/{parent} of the frame as {n} concurrent pipeline(s): {bench(n):.3f} ms wall (incl. launch + final sync)")
