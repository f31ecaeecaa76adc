#!/bin/bash
# Per-launch metrics of the wavefront family for one frame (after the same command exited 0 without ncu).
# A binned frame is 7 level + 6 bin + 7 combine + 1 commit = 21 launches (WF_SKIP / WF_COUNT: 60 / 15 for an unbinned one).
mkdir -p gpurun_out
python bench.py --family wavefront --steps 2 --warmup 3 "$@" > gpurun_out/plain_wf.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_active.avg,sm__cycles_elapsed.avg,smsp__warps_active.avg.per_cycle_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio \
  --clock-control none -k regex:"wf_" -s ${WF_SKIP:-62} -c ${WF_COUNT:-21} --csv --log-file gpurun_out/wf_launches.csv python bench.py --family wavefront --steps 2 --warmup 3 "$@" > gpurun_out/ncu_wf.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/wf_launches.csv")) if len(r)>10]
per=collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[0], {"name": r[4][9:28]})[r[-3]] = r[-1]
f=lambda v,k: float(v.get(k,"0").replace(",",""))
for k,v in per.items():
    print(k, v["name"], "us", round(f(v,"gpu__time_duration.sum")/1000,1), "thr/inst", v.get("smsp__thread_inst_executed_per_inst_executed.ratio"), "Minst", round(f(v,"smsp__inst_executed.sum")/1e6,1), "issue%", v.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), "fp64%", v.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), "icc", v.get("sm__icc_request_hit_rate.pct"), "gcc%", v.get("gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed"), "dramW MB", round(f(v,"dram__bytes_write.sum")/1e6,1), "dramR MB", round(f(v,"dram__bytes_read.sum")/1e6,1), "warps", v.get("smsp__warps_active.avg.per_cycle_active"),
          "stalls", " ".join(f"{n.split('stalled_')[1].split('_per_')[0]}={val}" for n, val in v.items() if "stalled" in n))
PY
