#!/bin/bash
# A/B over the library variants in build_variants/ (built with build.build(extra_flags=..., output=...)):
#   profiles/tools/ab.sh [bench args...]   -> one quick_bench line per variant, then the in-tree library
for lib in build_variants/librtgpu_*.so; do [ -e $lib ] || continue
  name=$(basename $lib .so); name=${name#librtgpu_}
  RTGPU_LIBRARY=$PWD/$lib profiles/tools/quick_bench.sh ab_$name "$@"
done
profiles/tools/quick_bench.sh ab_tree "$@"
