#!/bin/bash
mkdir -p gpurun_out
profiles/tools/wf_launches.sh --no-cold > gpurun_out/wf_launches_table.txt 2>&1; cp gpurun_out/wf_launches.csv gpurun_out/r2d_wavefront_launches.csv; mkdir -p profiles; cp gpurun_out/wf_launches.csv profiles/r2d_wavefront_launches.csv
head -c 200 gpurun_out/wf_launches_table.txt; echo
python bench.py --steps 2 --warmup 3 > gpurun_out/plain_default.json 2> gpurun_out/plain_default.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_default.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2d_n1.json 2> gpurun_out/bench_r2d_n1.err
python bench.py --steps 20 --warmup 3 --precision f32 --no-cold > gpurun_out/bench_r2d_f32.json 2>/dev/null
python bench.py --steps 20 --warmup 3 --family persistent --no-cold > gpurun_out/bench_r2d_persistent.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2d.json 2>/dev/null
python benchmarks/all_scenes.py --frames 5 > gpurun_out/all_scenes_r2d.json 2>/dev/null
python - <<'PY'
import json
for n in ("bench_r2d_n1", "bench_r2d_f32", "bench_r2d_persistent", "bench_ref_r2d"):
    d = json.load(open(f"gpurun_out/{n}.json"))
    print(n, d["ms_per_step"], (d.get("e2e") or {}).get("ms_per_frame"), d.get("gpu_launches"), (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("traffic"), d.get("family"))
PY
tail -9 gpurun_out/all_scenes_r2d.json | cut -c1-230
