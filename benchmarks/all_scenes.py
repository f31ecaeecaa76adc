#!/usr/bin/env python
"""ms/frame and Mrays/s of every shipped scene on one GPU (kernel time, CUDA events inside the library),
beside the CPU oracle on all host threads.  BASELINE.json configs[2..3]: cover / cylinders / table /
shadow_puppets at 1920x1080, reflect_refract / refraction at 3840x2160.

    python benchmarks/all_scenes.py [--frames 5] [--no-cpu] [--out file.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from ray_tracer_challenge_rs_b200.fixtures import SHIPPED_SCENES, load_scene_fixture  # noqa: E402
from ray_tracer_challenge_rs_b200.render import Renderer  # noqa: E402

SIZES = {"reflect_refract": (3840, 2160), "refraction": (3840, 2160)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--out", default=None)
    ap.add_argument("--family", default=None, choices=[None, "persistent", "wavefront"], help="force a kernel family (default: the library's measured choice)")
    args = ap.parse_args()
    results = []
    for name in SHIPPED_SCENES:
        flat, camera = load_scene_fixture(name)
        w, h = SIZES.get(name, (1920, 1080))
        cam = camera.resized(w, h)
        with Renderer(flat) as r:
            times = []
            for _ in range(args.frames + 1):
                rgb, rgb8, st = r.render(cam, precision=args.precision, family=args.family)
                times.append(st["kernel_ms"])
        rec = {"scene": name, "family": st["family"], "width": w, "height": h, "kernel_ms": min(times[1:]), "rays": st["rays"],
               "mrays_per_s": st["rays"] / (min(times[1:]) * 1e-3) / 1e6, "rays_per_pixel": st["rays"] / st["pixels"]}
        if not args.no_cpu:
            from oracle.oracle import Oracle, max_threads

            orgb, orgb8, ost = Oracle(flat).render(cam)
            rec.update({"cpu_ms": ost["total_ms"], "cpu_threads": max_threads(), "speedup": ost["total_ms"] / rec["kernel_ms"],
                        "rgb8_mismatches": int((orgb8 != rgb8).any(axis=1).sum()) if args.precision == "f64" else None,
                        "f64_pixels_differing": int((orgb != rgb).any(axis=1).sum()) if args.precision == "f64" else None,
                        "counters_equal": all(ost[k] == st[k] for k in ("rays_primary", "rays_shadow", "rays_reflect", "rays_refract", "hit_nodes"))})
        print(json.dumps(rec), flush=True)
        results.append(rec)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
