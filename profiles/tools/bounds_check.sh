#!/bin/bash
# The GPU parity tests, the device known-answer replays and benchmarks/all_paths_check.py against a library built with
# -DRT_BOUNDS_CHECK=1 (build_variants/librtgpu_bounds.so: python /tmp/build_variants.py bounds=-DRT_BOUNDS_CHECK=1, or
# build.build(extra_flags=["-DRT_BOUNDS_CHECK=1"], output=...)): every scene-table, queue, node and permutation index
# is range-checked on the device; a violation prints the failed condition and traps, which fails the test.
lib=$PWD/build_variants/librtgpu_bounds.so
[ -e $lib ] || { echo "build $lib first"; exit 1; }
RTGPU_LIBRARY=$lib python benchmarks/all_paths_check.py 2>&1 | tail -6
RTGPU_LIBRARY=$lib python -m pytest tests/test_gpu_parity.py tests/test_oracle_kat.py -m gpu -x -q 2>&1 | tail -3
