#!/bin/bash
# The parity suites against build variants of the library (python /tmp/build_variants.py ... or build.build(extra_flags, output)):
# every variant must produce the same bits as the default build.
for lib in build_variants/librtgpu_*.so; do
  name=$(basename $lib .so)
  echo "== $name"
  RTGPU_LIBRARY=$PWD/$lib python -m pytest tests/test_gpu_parity.py tests/test_oracle_kat.py -m gpu -q -x 2>&1 | tail -2
done
