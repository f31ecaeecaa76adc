#!/bin/bash
# Duration of every wavefront-family launch of one frame (ncu, time only): profiles/tools/wf_times.sh [bench args...]
mkdir -p gpurun_out
python bench.py --family wavefront --steps 2 --warmup 3 --no-cold "$@" > gpurun_out/plain_wf.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"wf_" -s 84 -c 21 --csv --log-file gpurun_out/wf_times.csv \
  python bench.py --family wavefront --steps 2 --warmup 3 --no-cold "$@" > gpurun_out/ncu_wf.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/wf_times.csv")) if len(r) > 10 and r[0].isdigit()]
tot = 0.0
for r in rows:
    name = r[4].split("(")[0].replace("void rt::", "")[:40]
    us = float(r[-1].replace(",", "")) / 1000
    tot += us
    print(f"{name:42s} {us:8.1f} us")
print("sum", round(tot, 1), "us")
PY
