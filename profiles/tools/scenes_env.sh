#!/bin/bash
# kernel ms of every shipped scene (wavefront family) under each of the given environment settings:
#   profiles/tools/scenes_env.sh RTGPU_WF_BINS=0 RTGPU_WF_BINS=1
for setting in "$@"; do
  env $setting python benchmarks/all_scenes.py --no-cpu --frames 4 --family wavefront > gpurun_out/scenes_env.jsonl
  python - "$setting" <<'PY'
import json, sys
print(sys.argv[1], " ".join(f"{d['scene']}={d['kernel_ms']:.3f}" for d in (json.loads(l) for l in open("gpurun_out/scenes_env.jsonl") if l.startswith("{"))))
PY
done
