#!/bin/bash
# BVH path: synthetic 1e5 shapes at 4K (a quarter of the 8K frame keeps the ncu replay short), both families
mkdir -p gpurun_out
cat > /tmp/syn.py <<'PY'
import sys, time
sys.path.insert(0, '.')
from ray_tracer_challenge_rs_b200.render import Renderer
from ray_tracer_challenge_rs_b200.synthetic import synthetic_camera, synthetic_scene
n = int(sys.argv[1]); fam = sys.argv[2]; w, h = int(sys.argv[3]), int(sys.argv[4])
flat = synthetic_scene(n); cam = synthetic_camera(w, h)
with Renderer(flat) as r:
    for _ in range(3):
        _, _, st = r.render(cam, want_rgb8=False, family=fam)
    print(n, fam, w, h, "kernel_ms", round(st["kernel_ms"], 2), "rays", st["rays"], "Mrays/s", round(st["rays"] / st["kernel_ms"] / 1e3))
PY
for fam in persistent wavefront; do python /tmp/syn.py 100000 $fam 7680 4320; python /tmp/syn.py 1000000 $fam 7680 4320; python /tmp/syn.py 10000 $fam 7680 4320; done
ncu --set full --import-source on --clock-control none -k regex:"render_kernel" -s 2 -c 1 -f -o gpurun_out/prof_r2_bvh_persistent python /tmp/syn.py 100000 persistent 3840 2160 > gpurun_out/ncu_bvh_p.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"wf_level" -s 14 -c 2 -f -o gpurun_out/prof_r2_bvh_wavefront python /tmp/syn.py 100000 wavefront 3840 2160 > gpurun_out/ncu_bvh_w.log 2>&1
ls -la gpurun_out/prof_r2_bvh_*
