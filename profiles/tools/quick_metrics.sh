#!/bin/bash
# Quick ncu pass for the render kernel: time, instruction-cache behaviour, FP64 pipe, issue rate, SIMT efficiency.
#   profiles/tools/quick_metrics.sh <label> [extra bench.py args]
# Follows B200_PROFILING.md: the same command must exit 0 without ncu first.
label=$1; shift
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 "$@" > gpurun_out/plain_$label.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_elapsed.avg,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:render_kernel -s 3 -c 1 --csv --log-file gpurun_out/quick_$label.csv python bench.py --steps 2 --warmup 3 "$@" > gpurun_out/ncu_$label.log 2>&1
python - "$label" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(f"gpurun_out/quick_{sys.argv[1]}.csv")) if len(r) > 10]
for r in rows[1:]:
    print(f"  {r[-3]:75s} {r[-1]:>16s} {r[-2]}")
PY
