#!/usr/bin/env python
"""Where a warm multi-device one-shot frame spends its wall time: RTGPU_TRACE=1 makes rtgpu_render print, per device,
when its host thread started, finished enqueueing and saw the stream drain (ms since the call began).

    python benchmarks/oneshot_trace.py [n_gpus] [band_rows] [frames] [rgb8]     (cover @ 1920x1080, f64, pinned Canvas)
"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RTGPU_TRACE"] = "1"

import torch  # noqa: E402

from ray_tracer_challenge_rs_b200 import abi  # noqa: E402
from ray_tracer_challenge_rs_b200.fixtures import load_scene_fixture  # noqa: E402
from ray_tracer_challenge_rs_b200.flatten import camera_to_c  # noqa: E402

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 1
band_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 4
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 12
rgb8 = len(sys.argv) > 4 and sys.argv[4] == "rgb8"  # the 8-bit Canvas alone (what the PPM / PNG writers consume): 1/8 of the bytes
flat, camera = load_scene_fixture("cover")
camera = camera.resized(1920, 1080)
lib = abi.load_library()
host = torch.empty((1920 * 1080, 3), dtype=torch.uint8 if rgb8 else torch.float64).pin_memory().numpy()
cscene, ccam = flat.as_c(), camera_to_c(camera)
opts = abi.RtgpuOpts(abi.PRECISION_F64, 6, n_gpus, band_rows, 0)
st = abi.RtgpuStats()
for k in range(frames):
    t0 = time.perf_counter()
    abi.check(lib, lib.rtgpu_render(C.byref(cscene), C.byref(ccam), C.byref(opts), None if rgb8 else host.ctypes.data,
                                     host.ctypes.data if rgb8 else None, C.byref(st)))
    sys.stderr.write("frame %d: %.3f ms wall\n" % (k, (time.perf_counter() - t0) * 1e3))
