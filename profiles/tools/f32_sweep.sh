#!/bin/bash
for lib in ray_tracer_challenge_rs_b200/librtgpu.so build_variants/librtgpu_*.so; do
  echo "== $lib"
  for fam in wavefront; do
    RTGPU_FAMILY=$fam RTGPU_LIBRARY=$PWD/$lib python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k f32_fast 2>&1 | grep -o "f32 [a-z_]*: [0-9.]*% within 2 LSB, max [0-9]*" | awk -v f=$fam '{printf "%s %s%s/%s | ", f, $2, $3, $NF} END {print ""}'
  done
done
