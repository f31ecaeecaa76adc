#!/usr/bin/env python
"""Join an `ncu --page source --csv` export (SASS rows: executed counts + stall samples) with
`nvdisasm --print-line-info` of the same cubin, and aggregate per CUDA source line / per function.

    python profiles/tools/sass_hotspots.py <ncu_source.csv> <cubin> <kernel-mangled-substring> [top]
"""
import collections
import csv
import re
import subprocess
import sys


def sass_lines(cubin, kernel_sub):
    out = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout
    lines = out.splitlines()
    in_k = False
    cur = ("?", 0)
    res = []
    for ln in lines:
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            in_k = kernel_sub in m.group(1)
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            res.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return res


def main():
    ncu_csv, cubin, ksub = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(ncu_csv)))
    hdr = rows[1]
    ia, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    ncu = [(r[isrc].strip(), int(r[ia]), int(r[ismp])) for r in rows[2:] if len(r) > ia]
    sass = sass_lines(cubin, ksub)
    print(f"ncu rows {len(ncu)}, nvdisasm instrs {len(sass)}")
    n = min(len(ncu), len(sass))
    per_line = collections.defaultdict(lambda: [0, 0, 0])
    tot_i = sum(x[1] for x in ncu)
    tot_s = sum(x[2] for x in ncu)
    for k in range(n):
        key = sass[k][2]
        per_line[key][0] += ncu[k][1]
        per_line[key][1] += ncu[k][2]
        per_line[key][2] += 1
    print(f"total warp instrs {tot_i}, samples {tot_s}")
    print("by executed instructions:")
    for key, v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {key[0]}:{key[1]:<5d} instrs {v[0] / tot_i * 100:6.2f}%  samples {v[1] / tot_s * 100:6.2f}%  ({v[2]} SASS)")
    per_file = collections.defaultdict(lambda: [0, 0])
    for key, v in per_line.items():
        per_file[key[0]][0] += v[0]
        per_file[key[0]][1] += v[1]
    print("by file:")
    for f, v in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:28s} instrs {v[0] / tot_i * 100:6.2f}%  samples {v[1] / tot_s * 100:6.2f}%")


if __name__ == "__main__":
    main()
