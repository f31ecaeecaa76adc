"""Build recipe for ``librtgpu.so`` (the CUDA kernels + the C ABI), sm_100a only.

``-fmad=false`` is part of the numerics contract: the reference never contracts ``a*b+c`` and the
kernels call ``fma()`` exactly where the reference calls ``mul_add`` (csrc/rt_kernel.cuh).
nvcc cross-compiles without a GPU, so this runs in the CPU-only build container too.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from typing import List, Optional

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PACKAGE_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PACKAGE_DIR), "include")
OUTPUT = os.path.join(PACKAGE_DIR, "librtgpu.so")

SOURCES = ["rtgpu.cu"]
HEADERS = ["rt_kernel.cuh", "rt_wavefront.cuh", "rt_arith.cuh", "rt_bvh.h", "rt_scene.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def find_nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not os.path.exists(OUTPUT):
        return True
    out_m = os.path.getmtime(OUTPUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(INCLUDE, "rtgpu.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > out_m for d in deps)


def build(force: bool = False, extra_flags: Optional[List[str]] = None, verbose: bool = False, output: Optional[str] = None) -> str:
    """Compile librtgpu.so.  `extra_flags` + `output` build an A/B variant next to it (select it at
    run time with RTGPU_LIBRARY=<path>)."""
    if not force and not _stale() and not extra_flags and output is None:
        return OUTPUT
    output = output or OUTPUT
    os.makedirs(os.path.dirname(output), exist_ok=True)
    cmd = [find_nvcc()] + NVCC_FLAGS + (extra_flags or []) + ["-I", INCLUDE, "-o", output] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    proc = subprocess.run(cmd, capture_output=True, text=True, env=env)
    log = proc.stdout + proc.stderr
    with open(os.path.join(PACKAGE_DIR, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if verbose or proc.returncode != 0:
        print(log)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}); see {os.path.join(PACKAGE_DIR, 'build.log')}")
    return output


HOST_DIR = os.path.join(PACKAGE_DIR, "host")
HOST_CLI = os.path.join(HOST_DIR, "ray-tracer-cli")
HOST_SOURCES = ["main.cpp", "scene_loader.cpp", "canvas.cpp"]
HOST_HEADERS = ["rt_host.hpp", "scene_loader.hpp"]


def build_host(force: bool = False) -> str:
    """Compile the C++ host (`host/ray-tracer-cli`: YAML loader, World / Camera / Canvas mirror, flattener,
    `--rendering-mode gpu`) against librtgpu.so.  -ffp-contract=off: the host math is bit-faithful."""
    build()
    deps = [os.path.join(HOST_DIR, f) for f in HOST_SOURCES + HOST_HEADERS] + [os.path.join(INCLUDE, "rtgpu.h"), OUTPUT]
    if not force and os.path.exists(HOST_CLI) and all(os.path.getmtime(HOST_CLI) >= os.path.getmtime(d) for d in deps):
        return HOST_CLI
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not found")
    cmd = [gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-Wall", "-Wextra", "-Wno-comment", "-o", HOST_CLI] + \
          [os.path.join(HOST_DIR, s) for s in HOST_SOURCES] + ["-L", PACKAGE_DIR, "-lrtgpu", "-lz", "-Wl,-rpath,$ORIGIN/.."]
    env = dict(os.environ)
    env.pop("CXX", None)
    proc = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if proc.returncode != 0:
        raise RuntimeError("host build failed:\n" + proc.stdout + proc.stderr)
    return HOST_CLI


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose=True))
    print(build_host(force="--force" in sys.argv))
