#!/bin/bash
# quick bench: prints ms_per_step, e2e ms, frame identity
label=$1; shift
python bench.py --family wavefront --steps 10 --warmup 3 "$@" > gpurun_out/bench_$label.json 2> gpurun_out/bench_$label.err || tail -c 1500 gpurun_out/bench_$label.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$label.json"))
f=d.get("frame") or {}
print("$label", "ms", round(d["ms_per_step"],4), "Mrays/s", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_frame"],3), "frac", round(d["roofline"]["frac"],4), "match_n1", f.get("frame_matches_n1"), "rgb8 diff", f.get("rgb8_pixels_differing_from_oracle"))
PY
