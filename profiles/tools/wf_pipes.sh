#!/bin/bash
# Pipe utilisation and stall reasons per wavefront launch, for the in-tree library or RTGPU_LIBRARY:
#   profiles/tools/wf_pipes.sh <label> [bench args...]     -> gpurun_out/wf_pipes_<label>.csv + a table on stdout
label=$1; shift
mkdir -p gpurun_out
python bench.py --family wavefront --steps 2 --warmup 3 --no-cold "$@" > gpurun_out/plain_wf.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active,smsp__warps_active.avg.per_cycle_active,smsp__warps_eligible.avg.per_cycle_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio,smsp__average_warps_issue_stalled_misc_per_issue_active.ratio,smsp__average_warps_issue_stalled_drain_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio \
  --clock-control none -k regex:"wf_level" -s 28 -c 7 --csv --log-file gpurun_out/wf_pipes_$label.csv python bench.py --family wavefront --steps 2 --warmup 3 --no-cold "$@" > gpurun_out/ncu_wf.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/wf_pipes_$label.csv")) if len(r)>10]
per=collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[0], {})[r[-3]] = r[-1]
f=lambda v,k: float(v.get(k,"0").replace(",",""))
short=lambda n: n.replace("sm__inst_executed_pipe_","").replace(".avg.pct_of_peak_sustained_active","")
for k,v in per.items():
    print("$label", k, "us", round(f(v,"gpu__time_duration.sum")/1000,1), "thr/inst", v.get("smsp__thread_inst_executed_per_inst_executed.ratio"), "Minst", round(f(v,"smsp__inst_executed.sum")/1e6,1),
          "issue%", v.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), "warps", v.get("smsp__warps_active.avg.per_cycle_active"), "eligible", v.get("smsp__warps_eligible.avg.per_cycle_active"),
          "| pipes", " ".join(f"{short(n)}={val}" for n, val in v.items() if "pipe_" in n),
          "| stalls", " ".join(f"{n.split('stalled_')[1].split('_per_')[0]}={val}" for n, val in v.items() if "stalled" in n and float(val.replace(',','')) >= 0.05))
PY
