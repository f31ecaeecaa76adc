#!/usr/bin/env python
"""What would binning each level's queue be worth?  RTGPU_SORT_EXPERIMENT=<key> makes the wavefront launcher time every
level launch and reorder the next level's queue ON THE HOST between launches (key 0: leave the order alone; 1: by hit
shape; 2: hit shape, then reflect / refract; 3: reflect / refract only; 4: shape, kind, direction octant).  The frame
is the same for every key (entries carry their parent link); only the level times matter.

    python benchmarks/sort_experiment.py [scene] [width] [height]
"""
import ctypes as C
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RTGPU_E2E_CHUNKS"] = "1"

import torch  # noqa: E402

from ray_tracer_challenge_rs_b200 import abi  # noqa: E402
from ray_tracer_challenge_rs_b200.fixtures import load_scene_fixture  # noqa: E402
from ray_tracer_challenge_rs_b200.flatten import camera_to_c  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "cover"
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
flat, camera = load_scene_fixture(scene)
camera = camera.resized(w, h)
lib = abi.load_library()
host = torch.empty((w * h, 3), dtype=torch.float64).pin_memory().numpy()
cscene, ccam = flat.as_c(), camera_to_c(camera)
opts = abi.RtgpuOpts(abi.PRECISION_F64, 6, 1, 16, abi.FLAG_WAVEFRONT)
abi.check(lib, lib.rtgpu_render(C.byref(cscene), C.byref(ccam), C.byref(opts), host.ctypes.data, None, None))
reference = hashlib.sha256(host.tobytes()).hexdigest()
for key in range(5):
    os.environ["RTGPU_SORT_EXPERIMENT"] = str(key)
    for frame in range(3):
        sys.stderr.write("--- key %d frame %d\n" % (key, frame))
        sys.stderr.flush()
        abi.check(lib, lib.rtgpu_render(C.byref(cscene), C.byref(ccam), C.byref(opts), host.ctypes.data, None, None))
    sys.stderr.write("key %d frame identical: %s\n" % (key, hashlib.sha256(host.tobytes()).hexdigest() == reference))
