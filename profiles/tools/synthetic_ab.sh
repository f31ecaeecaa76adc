#!/bin/bash
# the synthetic scene at 8K for the in-tree library and every variant in build_variants/
for lib in ray_tracer_challenge_rs_b200/librtgpu.so build_variants/librtgpu_*.so; do
  [ -e $lib ] || continue
  for fam in persistent wavefront; do for n in 10000 100000 1000000; do echo "$(basename $lib) $(RTGPU_LIBRARY=$PWD/$lib python benchmarks/synthetic_frame.py $n $fam 7680 4320)"; done; done
done
