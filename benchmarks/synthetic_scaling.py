#!/usr/bin/env python
"""BASELINE.json configs[4]: synthetic scaling scene (N spheres + triangles) rendered on the GPU, with a
parity spot-check against the CPU oracle on a random pixel sample (the oracle is brute force, O(P*S):
a full 8K frame of 10^6 shapes is out of its reach, as it is of the reference's).

    python benchmarks/synthetic_scaling.py --shapes 10000 100000 1000000 --width 7680 --height 4320
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from ray_tracer_challenge_rs_b200.render import Renderer  # noqa: E402
from ray_tracer_challenge_rs_b200.synthetic import synthetic_camera, synthetic_scene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", type=int, nargs="+", default=[10000, 100000, 1000000])
    ap.add_argument("--width", type=int, default=7680)
    ap.add_argument("--height", type=int, default=4320)
    ap.add_argument("--frames", type=int, default=5)
    ap.add_argument("--sample", type=int, default=512, help="pixels checked against the CPU oracle")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from oracle.oracle import Oracle, max_threads

    results = []
    for n in args.shapes:
        t0 = time.perf_counter()
        flat = synthetic_scene(n)
        cam = synthetic_camera(args.width, args.height)
        t_gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        with Renderer(flat) as r:
            t_upload = time.perf_counter() - t0
            best, stats, rgb = None, None, None
            for _ in range(args.frames):
                rgb, _, st = r.render(cam, want_rgb8=False)
                if best is None or st["kernel_ms"] < best:
                    best, stats = st["kernel_ms"], st
        rng = np.random.default_rng(n)
        px = rng.integers(0, args.width * args.height, args.sample).astype(np.uint64)
        t0 = time.perf_counter()
        ref, _, ost = Oracle(flat).render_pixels(cam, px)
        t_cpu = time.perf_counter() - t0
        got = rgb[px.astype(np.int64)]
        worst = float((np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)).max())
        rec = {
            "shapes": n, "width": args.width, "height": args.height, "kernel_ms": best, "rays": stats["rays"],
            "mrays_per_s": stats["rays"] / (best * 1e-3) / 1e6, "rays_per_pixel": stats["rays"] / stats["pixels"],
            "scene_build_s": t_gen, "pack_bvh_upload_s": t_upload,
            "oracle_sample_pixels": int(px.size), "oracle_sample_max_rel_diff": worst,
            "oracle_sample_rays_per_s": ost["rays"] / t_cpu, "oracle_threads": max_threads(),
            "oracle_full_frame_estimate_s": stats["rays"] / (ost["rays"] / t_cpu),
        }
        print(json.dumps(rec), flush=True)
        assert worst <= 1e-12, worst
        results.append(rec)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
