#!/usr/bin/env python
"""Per-source-line view of one kernel launch in an ncu report (captured with --import-source on / -lineinfo).

    python profiles/tools/src_hotspots.py <report.ncu-rep> <launch-index-in-report> <library.so> <kernel-substring> [top]

Joins `ncu --page source --print-source sass --csv` (per SASS instruction: executed warp instructions, thread
instructions, stall samples) with `nvdisasm --print-line-info` of the cubin extracted from the library, and prints
  * per CUDA source line (innermost inlined location): share of warp instructions, of stall samples, threads / instruction;
  * per opcode class: share of warp instructions (FP64 pipe, integer / move, branch / barrier, shared / local / global memory).
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile


def cubin_for(lib, ksub):
    tmp = tempfile.mkdtemp(prefix="cubin_")
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    for c in glob.glob(os.path.join(tmp, "*.cubin")):
        syms = subprocess.run(["cuobjdump", "-elf", c], capture_output=True, text=True).stdout
        if ksub in syms:
            return c
    raise SystemExit("kernel not found in any cubin")


def sass_lines(cubin, ksub):
    out = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    if not out:
        out = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout
    in_k, cur, res, fresh = False, ("?", 0), [], True
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            in_k = ksub in m.group(1)
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            if fresh:  # -gi lists the inlining chain innermost first: keep the innermost location
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
                fresh = False
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            res.append((int(m.group(1), 16), m.group(2).strip(), cur))
            fresh = True
    return res


def classify(op):
    op = op.split()[0] if not op.startswith("@") else op.split()[1]
    base = op.split(".")[0]
    if base in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX", "MUFU"):
        return "fp64+mufu"
    if base in ("BRA", "BSSY", "BSYNC", "BREAK", "CALL", "RET", "EXIT", "WARPSYNC", "BAR", "NOP", "YIELD"):
        return "branch/barrier"
    if base in ("LDS", "STS", "ATOMS"):
        return "shared"
    if base in ("LDL", "STL"):
        return "local"
    if base in ("LDG", "STG", "LD", "ST", "ATOMG", "RED", "ATOM", "LDC", "LDCU"):
        return "global/const"
    if base in ("SHFL", "VOTE", "MATCH", "REDUX"):
        return "warp"
    if base.startswith("F") or base in ("I2F", "F2I", "F2F", "I2FP"):
        return "fp32/convert"
    return "integer/move"


def main():
    rep, idx, lib, ksub = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    txt = subprocess.run(["ncu", "-i", os.path.abspath(rep), "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True, cwd="/tmp").stdout
    # the export concatenates the launches: split on the "Kernel Name" rows
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(txt)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(row)
    blk = blocks[idx]
    hdr = blk["rows"][0]
    ia, it, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    ncu = [(r[isrc].strip(), int(r[ia]), int(r[it]), int(r[ismp])) for r in blk["rows"][1:] if len(r) > it]
    sass = sass_lines(cubin_for(lib, ksub), ksub)
    n = min(len(ncu), len(sass))
    print(f"launch {idx}: {blk['name'][:60]}  ncu rows {len(ncu)}, nvdisasm instrs {len(sass)}")
    tot_i, tot_t, tot_s = sum(x[1] for x in ncu), sum(x[2] for x in ncu), sum(x[3] for x in ncu)
    print(f"warp instrs {tot_i / 1e6:.1f} M, threads/instr {tot_t / max(tot_i, 1):.2f}, samples {tot_s}")
    per_line = collections.defaultdict(lambda: [0, 0, 0, 0])
    per_class = collections.defaultdict(lambda: [0, 0])
    for k in range(n):
        key = sass[k][2]
        v = per_line[key]
        v[0] += ncu[k][1]; v[1] += ncu[k][2]; v[2] += ncu[k][3]; v[3] += 1
        c = per_class[classify(ncu[k][0])]
        c[0] += ncu[k][1]; c[1] += ncu[k][3]
    print("opcode classes (share of warp instructions | of stall samples):")
    for c, v in sorted(per_class.items(), key=lambda kv: -kv[1][0]):
        print(f"  {c:16s} {v[0] / tot_i * 100:6.2f}%  {v[1] / max(tot_s, 1) * 100:6.2f}%")
    per_op = collections.Counter()
    for k in range(n):
        op = ncu[k][0]
        op = op.split()[0] if not op.startswith("@") else op.split()[1]
        per_op[op.split(".")[0]] += ncu[k][1]
    print("opcodes by executed warp instructions:", ", ".join(f"{o} {c / tot_i * 100:.1f}%" for o, c in per_op.most_common(28)))
    only = os.environ.get("HOT_CLASS")  # e.g. HOT_CLASS=integer/move: the source lines where that class executes most
    if only:
        cls_line = collections.Counter()
        for k in range(n):
            if classify(ncu[k][0]) == only:
                cls_line[sass[k][2]] += ncu[k][1]
        print(f"class {only} by source line (share of ALL warp instructions):")
        for key, c in cls_line.most_common(top):
            print(f"  {key[0]}:{key[1]:<5d} {c / tot_i * 100:6.2f}%")
    print("by source line (warp instrs | stall samples | threads per instr | #SASS):")
    for key, v in sorted(per_line.items(), key=lambda kv: -kv[1][2])[:top]:
        print(f"  {key[0]}:{key[1]:<5d} {v[0] / tot_i * 100:6.2f}%  {v[2] / max(tot_s, 1) * 100:6.2f}%  {v[1] / max(v[0], 1):5.1f}  ({v[3]})")
    per_fn = collections.defaultdict(lambda: [0, 0, 0])
    for key, v in per_line.items():
        f = per_fn[key[0]]
        f[0] += v[0]; f[1] += v[1]; f[2] += v[2]
    print("by file:")
    for f, v in sorted(per_fn.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:24s} {v[0] / tot_i * 100:6.2f}%  {v[2] / max(tot_s, 1) * 100:6.2f}%  {v[1] / max(v[0], 1):5.1f}")


if __name__ == "__main__":
    main()
