#!/usr/bin/env python
"""Markdown summary of a per-launch capture (the output of wf_launches.sh).

    python profiles/tools/wf_table.py <bench ms/frame> [csv under profiles/] [title] > profiles/rN_wavefront_launches.md
"""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
bench_ms = sys.argv[1] if len(sys.argv) > 1 else "?"
csv_name = sys.argv[2] if len(sys.argv) > 2 else "r1_wavefront_launches.csv"
title = sys.argv[3] if len(sys.argv) > 3 else "cover@1920x1080 f64"
rows = [r for r in csv.reader(open(os.path.join(ROOT, "profiles", csv_name))) if len(r) > 10]
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[0], {"name": r[4]})[r[-3]] = r[-1]


def f(v, k):
    return float(v.get(k, "0").replace(",", "") or 0)


def short(n):
    for k in ("wf_level_kernel", "wf_combine_kernel", "wf_commit_counters_kernel", "wf_bin_kernel"):
        if k in n:
            return k
    return n[:30]


tot = sum(f(v, "gpu__time_duration.sum") for v in per.values())
print("# Wavefront family: the %d launches of one %s frame (ncu, --clock-control none; `profiles/%s`)\n" % (len(per), title, csv_name))
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
print("\nLevel kernels: %.1f %% of the frame; combine kernels %.1f %%; bin kernels %.1f %%; counter commit %.1f %%." % (share("wf_level"), share("combine"), share("wf_bin"), share("commit")))
print("DRAM per frame: %.0f MB written, %.0f MB read (bench.py reports the sum as `roofline.traffic`)." % (
    sum(f(v, "dram__bytes_write.sum") for v in per.values()) / 1e6, sum(f(v, "dram__bytes_read.sum") for v in per.values()) / 1e6))
reading = {
    "r1_wavefront_launches.csv": """
Reading: the level kernels are issue-bound on fixed-latency FP64 dependencies (`wait` ~2.5 cycles per issue with 3.9
warps per scheduler), FP64 pipe 33-38 % busy, instruction caches healthy on this scene (icc 97 %, GPC cache 30-56 %).
The combine kernels move one node record per interior node at ~3.8 TB/s; they are latency-bound at 5 % issue.""",
    "r2b_wavefront_launches.csv": """
Reading (round 2; the capture window starts at level 1 of one frame and ends with level 0 of the next: launch 14 is a
level-0 launch): with the node state parked in shared memory the level kernels run at 80 registers, 6 CTAs per SM
(5.9 warps per scheduler instead of 3.9): **issue slots 67-70 % busy (round 1: 54-57 %), FP64 pipe 47-52 % (33-40 %)**;
instruction counts and threads per instruction are those of round 1 — the divergence of the exact tests at depth
(17.6-25 threads per instruction at levels 1-6) is unchanged and is what is left.  The combine kernels (9.6 % of the
frame) are memory-latency bound at ~3.8 TB/s.""",
    "r2d_wavefront_launches.csv": """
Reading (end of round 2: single-precision pre-test, binned queues): a level's queue is consumed grouped by (hit shape,
reflected / refracted), which lifts the deeper levels from 17.9-25.3 to **21.9-29.3 threads per instruction**; the
pre-test no longer touches the FP64 pipe (26-31 % busy instead of 47-52 %: what is left there are the exact tests), issue
slots 63-73 % busy.  The six bin kernels (one pass over the (bin, rank) keys of a queue each) cost 7 µs apiece.""",
    "r2_synthetic_1e5_8k_launches.csv": """
Reading: the BVH path.  Level 0 is 70 % of the frame at 18.7 threads per instruction; the deeper levels run at 6-9.
After moving the box tests to single precision the FP64 pipe is 2-4 % busy (only the exact leaf tests use it) and the
kernels are bound by issue slots spent on partially filled warps, not by memory: DRAM carries 1.5 GB in the 49 ms of
level 0 (31 GB/s, 0.5 % of the HBM peak), L2 serves the 20 MB of nodes and shapes.""",
}
print(reading.get(csv_name, ""))
