#!/bin/bash
# BVH path: the synthetic scene at 8K through both families, then ncu --set full of the traversal kernels on 1e5 shapes
# at 4K (a quarter of the frame keeps the replay short).  One GPU; the plain commands run first.
mkdir -p gpurun_out
for fam in persistent wavefront; do for n in 10000 100000 1000000; do python benchmarks/synthetic_frame.py $n $fam 7680 4320; done; done
ncu --set full --import-source on --clock-control none -k regex:"render_kernel" -s 2 -c 1 -f -o gpurun_out/prof_${1:-r2}_bvh_persistent python benchmarks/synthetic_frame.py 100000 persistent 3840 2160 > gpurun_out/ncu_bvh_p.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"wf_level" -s 14 -c 2 -f -o gpurun_out/prof_${1:-r2}_bvh_wavefront python benchmarks/synthetic_frame.py 100000 wavefront 3840 2160 > gpurun_out/ncu_bvh_w.log 2>&1
ls -la gpurun_out/prof_${1:-r2}_bvh_*
