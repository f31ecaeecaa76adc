#!/usr/bin/env python
"""Per-category split of one launch of wf_level_kernel<double,0,0,1> in an ncu report: share of warp instructions, of
stall samples, threads per instruction, for the pre-test loop / exact tests / node code, from source-line ranges that
are looked up in the CURRENT sources (function markers), so the table follows the code.

    python profiles/tools/categories.py <report.ncu-rep> <launch indices, e.g. 0,4,10> [library.so]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "profiles", "tools"))
import src_hotspots as H  # noqa: E402

KERNEL = "wf_level_kernelIdLb0ELb0ELb1"


def line_of(path, pattern):
    for i, ln in enumerate(open(path), 1):
        if re.search(pattern, ln):
            return i
    raise SystemExit(f"marker {pattern!r} not found in {path}")


def main():
    rep, idxs = sys.argv[1], [int(v) for v in sys.argv[2].split(",")]
    lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "ray_tracer_challenge_rs_b200", "librtgpu.so")
    k = os.path.join(ROOT, "ray_tracer_challenge_rs_b200", "csrc", "rt_kernel.cuh")
    exact0, trace0 = line_of(k, r"RT_DEV void exact_test\("), line_of(k, r"RT_DEV void trace_unified\(")
    pairs0 = line_of(k, r"pair-list trace: World::collect_intersections")
    consume0, solve0 = line_of(k, r"RT_DEV void consume\("), line_of(k, r"RT_DEV bool solve_quadratic\(")
    load_cull0 = line_of(k, r"RT_DEV void load_cull\(const double")
    cursor0 = line_of(k, r"struct CullCursor;")
    normal0 = line_of(k, r"RT_COLD V3<T> local_normal_at\(")
    node_fns0 = line_of(k, r"The pieces of a node both kernel families")

    def cat(key):
        f, l = key
        if f == "rt_arith.cuh":
            return "B exact tests: division / sqrt (rt_arith.cuh)"
        if f == "rt_wavefront.cuh":
            return "C node code (phases, lighting, queues: rt_wavefront.cuh)"
        if f == "rt_kernel.cuh":
            if trace0 <= l < pairs0 or cursor0 <= l < exact0 and l < trace0 or load_cull0 <= l < load_cull0 + 12:
                return "A pre-test loop (trace_unified, CullCursor)"
            if exact0 <= l < trace0 or consume0 <= l < load_cull0 or 160 <= l <= 200:
                return "B exact tests: transform, local_intersect, consume"
            if l >= normal0:
                return "C node code (normals, patterns, lighting functions)"
            return "C node code (vector helpers, normalisation)"
        return "other (" + f + ")"

    txt = subprocess.run(["ncu", "-i", os.path.abspath(rep), "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True, cwd="/tmp").stdout
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(txt)):
        if row and row[0] == "Kernel Name":
            cur = {"rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(row)
    sass = H.sass_lines(H.cubin_for(lib, KERNEL), KERNEL)
    print("| launch | category | warp instr | stall samples | threads / instr |")
    print("|---|---|---|---|---|")
    for idx in idxs:
        blk = blocks[idx]
        hdr = blk["rows"][0]
        ia, it, ismp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        rows = [r for r in blk["rows"][1:] if len(r) > it]
        tot, tots, tott = sum(int(r[ia]) for r in rows), sum(int(r[ismp]) for r in rows), sum(int(r[it]) for r in rows)
        per = collections.defaultdict(lambda: [0, 0, 0])
        for n, r in enumerate(rows[: len(sass)]):
            v = per[cat(sass[n][2])]
            v[0] += int(r[ia]); v[1] += int(r[it]); v[2] += int(r[ismp])
        print(f"| {idx} | all: {tot / 1e6:.1f} M warp instr, {tott / tot:.1f} threads / instr | | | |")
        for c, v in sorted(per.items()):
            print(f"| {idx} | {c} | {100 * v[0] / tot:.1f} % | {100 * v[2] / max(tots, 1):.1f} % | {v[1] / max(v[0], 1):.1f} |")


if __name__ == "__main__":
    main()
