#!/bin/bash
# kernel ms of every shipped scene, per family, for the in-tree library and every variant in build_variants/
run() { python benchmarks/all_scenes.py --no-cpu --frames 4 --family $2 | python -c "
import sys, json
print('$1 $2', ' '.join(f\"{json.loads(l)['scene']}={json.loads(l)['kernel_ms']:.3f}\" for l in sys.stdin if l.startswith('{')))"; }
for fam in wavefront persistent; do
  run tree $fam
  for lib in build_variants/librtgpu_*.so; do [ -e $lib ] || continue; name=$(basename $lib .so); RTGPU_LIBRARY=$PWD/$lib run ${name#librtgpu_} $fam; done
done
