#!/usr/bin/env python
"""Markdown summary of profiles/r1_wavefront_launches.csv (the output of wf_launches.sh).

    python profiles/tools/wf_table.py <bench ms/frame> > profiles/r1_wavefront_launches.md
"""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
bench_ms = sys.argv[1] if len(sys.argv) > 1 else "?"
rows = [r for r in csv.reader(open(os.path.join(ROOT, "profiles", "r1_wavefront_launches.csv"))) if len(r) > 10]
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[0], {"name": r[4]})[r[-3]] = r[-1]


def f(v, k):
    return float(v.get(k, "0").replace(",", "") or 0)


def short(n):
    for k in ("wf_level_kernel", "wf_combine_kernel", "wf_commit_counters_kernel"):
        if k in n:
            return k
    return n[:30]


tot = sum(f(v, "gpu__time_duration.sum") for v in per.values())
print("# Wavefront family: the %d launches of one cover@1920x1080 f64 frame (ncu, --clock-control none)\n" % len(per))
print("Command: `profiles/tools/wf_launches.sh` (bench.py --family wavefront --steps 2 --warmup 3 exited 0 first; the capture is")
print("the last timed frame).  Times under ncu are serialised and cold-cache: use the shares, not the absolutes")
print("(bench.py measures %s ms for the frame with CUDA events; the launches below sum to %.2f ms).\n" % (bench_ms, tot / 1e6))
print("| # | kernel | µs | share | thr/inst | M warp-inst | issue % | FP64 pipe % | icc hit % | gcc inst % | DRAM W MB | DRAM R MB | stall wait | long_sb | no_inst |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for i, (k, v) in enumerate(per.items()):
    t = f(v, "gpu__time_duration.sum")
    g = lambda key: v.get(key, "")  # noqa: E731
    print(f"| {i} | {short(v['name'])} | {t / 1000:.1f} | {100 * t / tot:.1f} % | {g('smsp__thread_inst_executed_per_inst_executed.ratio')} | "
          f"{f(v, 'smsp__inst_executed.sum') / 1e6:.1f} | {g('smsp__issue_active.avg.pct_of_peak_sustained_active')} | "
          f"{g('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active')} | {g('sm__icc_request_hit_rate.pct')} | "
          f"{g('gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed')} | {f(v, 'dram__bytes_write.sum') / 1e6:.1f} | "
          f"{f(v, 'dram__bytes_read.sum') / 1e6:.1f} | {g('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio')} | "
          f"{g('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio')} | "
          f"{g('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio')} |")
share = lambda sub: 100 * sum(f(v, "gpu__time_duration.sum") for v in per.values() if sub in v["name"]) / tot  # noqa: E731
print("\nLevel kernels: %.1f %% of the frame; combine kernels %.1f %%; counter commit %.1f %%." % (share("wf_level"), share("combine"), share("commit")))
print("DRAM per frame: %.0f MB written, %.0f MB read (bench.py reports the sum as `roofline.traffic`)." % (
    sum(f(v, "dram__bytes_write.sum") for v in per.values()) / 1e6, sum(f(v, "dram__bytes_read.sum") for v in per.values()) / 1e6))
print("""
Reading: the level kernels are issue-bound on fixed-latency FP64 dependencies (`wait` ~2.5 cycles per issue with 3.9
warps per scheduler), FP64 pipe 33-38 % busy, instruction caches healthy on this scene (icc 97 %, GPC cache 30-56 %).
Instruction mix of the level-0 launch (ncu source view): DMUL + DFMA + DADD + DSETP 33 % of the warp instructions,
integer / move / address arithmetic 30 %, branches and convergence barriers 12 %, LDS 6 %.
On pattern-heavy scenes (table, metal) the same kernels run at icc 87-93 % / GPC cache 83-92 % and are instruction-supply bound.
The combine kernels move one 120-byte node record per interior node at ~3.8 TB/s; they are latency-bound at 5 % issue
(4x more CTAs changed nothing: 2.441 vs 2.445 ms).

`r1_final_launches.csv`: the `--metrics gpu__time_duration.sum` launch list of the default `python bench.py --steps 2 --warmup 3`.
Caveat: under ncu every launch is serialised and pays the profiler's per-launch overhead, so the family calibration —
which times whole frames — can see the multi-launch wavefront frame as slower and settle on the persistent kernel
there.  Outside the profiler the same command settles on the wavefront family (bench line: `config.family`), which is
what the table above profiles.
`r1d_wf_level_kernel_metrics.csv`: `ncu --set full` of the level-0 and level-1 launches (raw page; captured before the
last three micro-optimisations, 2.53 ms per frame at the time).""")
