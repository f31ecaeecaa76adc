#!/bin/bash
# ncu --set full (+ source counters) of the 7 level launches of one cover@1080p frame, wavefront family.
#   profiles/tools/wf_full.sh <label> [extra bench.py args]
# Follows B200_PROFILING.md: the same command exits 0 without ncu first; one GPU.
label=$1; shift
mkdir -p gpurun_out
python bench.py --family wavefront --steps 1 --warmup 3 "$@" > gpurun_out/plain_$label.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"wf_level" -s 21 -c 7 -f -o gpurun_out/prof_$label \
  python bench.py --family wavefront --steps 1 --warmup 3 "$@" > gpurun_out/ncu_$label.log 2>&1
ls -la gpurun_out/prof_$label.ncu-rep
