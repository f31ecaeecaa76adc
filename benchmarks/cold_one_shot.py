#!/usr/bin/env python
"""Cold one-shot frame: what `ray-tracer-cli <scene> <out> --rendering-mode gpu` pays, in the region the reference's
CLI times (ray-tracer-cli/src/main.rs:17-22: the render call alone, scene already loaded).  Every measurement is a
FRESH process whose first CUDA call happens inside the timed rtgpu_render: driver + context initialisation, kernel
module load, scene pack + upload, buffers, kernel(s), D2H of the f64 Canvas.

    python benchmarks/cold_one_shot.py [--scene cover --width 1920 --height 1080 --runs 3] [--out file.json]

Prints one JSON object: per kernel family (auto = what a plain CLI call gets) the first call, the second call in the
same process, and — from a separate fresh process — the bare context creation (rtgpu_context_create: CUDA
initialisation + scene upload, no frame).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(args):
    import numpy as np  # noqa: F401

    from ray_tracer_challenge_rs_b200.fixtures import load_scene_fixture
    from ray_tracer_challenge_rs_b200.render import Renderer, last_family, render_gpu

    flat, camera = load_scene_fixture(args.scene)
    cam = camera.resized(args.width, args.height)
    family = None if args.family == "auto" else args.family
    if args.context_only:
        t0 = time.perf_counter()
        r = Renderer(flat)
        t1 = time.perf_counter()
        r.close()
        print(json.dumps({"context_ms": (t1 - t0) * 1e3}))
        return
    t0 = time.perf_counter()
    _, st1 = render_gpu(cam, flat, want_rgb8=False, return_stats=True, family=family)
    t1 = time.perf_counter()
    fam1 = last_family()
    _, st2 = render_gpu(cam, flat, want_rgb8=False, return_stats=True, family=family)
    t2 = time.perf_counter()
    print(json.dumps({"first_call_ms": (t1 - t0) * 1e3, "first_family": fam1, "first_kernel_ms": st1["kernel_ms"],
                      "second_call_ms": (t2 - t1) * 1e3, "second_family": last_family(), "rays": st1["rays"]}))


def run_child(args, family, context_only=False):
    cmd = [sys.executable, os.path.abspath(__file__), "--child", "--scene", args.scene, "--width", str(args.width), "--height", str(args.height),
           "--family", family] + (["--context-only"] if context_only else [])
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    if out.returncode != 0:
        raise SystemExit(out.stderr[-2000:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def measure(args):
    result = {"scene": args.scene, "width": args.width, "height": args.height, "runs": args.runs,
              "timed_region": "the rtgpu_render call of a fresh process (scene already loaded), as ray-tracer-cli/src/main.rs:17-22 times camera.render*"}
    ctx = [run_child(args, "auto", context_only=True)["context_ms"] for _ in range(args.runs)]
    result["context_create_ms"] = {"best": min(ctx), "all": ctx}
    for family in ("auto", "persistent", "wavefront"):
        runs = [run_child(args, family) for _ in range(args.runs)]
        best = min(runs, key=lambda r: r["first_call_ms"])
        result[family] = {"first_call_ms": best["first_call_ms"], "first_family": best["first_family"], "first_kernel_ms": best["first_kernel_ms"],
                          "second_call_ms": min(r["second_call_ms"] for r in runs), "second_family": best["second_family"],
                          "all_first_calls_ms": [r["first_call_ms"] for r in runs]}
    return result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="cover")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--runs", type=int, default=3)
    ap.add_argument("--family", default="auto")
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--context-only", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    if args.child:
        return child(args)
    result = measure(args)
    print(json.dumps(result))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(result, f, indent=1)


if __name__ == "__main__":
    main()
