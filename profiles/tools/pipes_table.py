#!/usr/bin/env python
"""Markdown table of a wf_pipes.sh capture: per level launch, pipe utilisation and the stall reasons above 0.05.

    python profiles/tools/pipes_table.py <csv> [title]
"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
title = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[0], {})[r[-3]] = r[-1]


def f(v, k):
    try:
        return float(v.get(k, "0").replace(",", ""))
    except ValueError:
        return 0.0


pipes = ("fp64", "fma", "fmaheavy", "alu", "lsu", "xu", "cbu", "uniform")
print(f"### {title}\n")
print("| level | µs | thr/inst | M warp-inst | issue % | warps / eligible per scheduler | " + " | ".join(p + " %" for p in pipes) + " | stalls (warps per issue) |")
print("|---|---|---|---|---|---|" + "---|" * len(pipes) + "---|")
for i, (k, v) in enumerate(per.items()):
    stalls = ", ".join(f"{n.split('stalled_')[1].split('_per_')[0]} {f(v, n):.2f}" for n in v if "stalled" in n and f(v, n) >= 0.05 and "selected" not in n)
    print(f"| {i} | {f(v, 'gpu__time_duration.sum') / 1000:.1f} | {v.get('smsp__thread_inst_executed_per_inst_executed.ratio')} | "
          f"{f(v, 'smsp__inst_executed.sum') / 1e6:.1f} | {v.get('smsp__issue_active.avg.pct_of_peak_sustained_active')} | "
          f"{v.get('smsp__warps_active.avg.per_cycle_active')} / {v.get('smsp__warps_eligible.avg.per_cycle_active')} | " +
          " | ".join(v.get(f"sm__inst_executed_pipe_{p}.avg.pct_of_peak_sustained_active", "") for p in pipes) + f" | {stalls} |")
print()
