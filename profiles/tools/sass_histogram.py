#!/usr/bin/env python
"""SASS opcode histogram of librtgpu.so (whole library and the two hot kernels), so that claims about the
instruction stream are checkable from the repository: TMA staging (UBLKCP + SYNCS), FP64 mix, local-memory traffic.

    python profiles/tools/sass_histogram.py [library.so] > profiles/rN_sass_histogram.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HOT = {
    "rt::wf_level_kernel<double,FULL=0,BVH=0,SMEM=1>": "_ZN2rt15wf_level_kernelIdLb0ELb0ELb1E",
    "rt::render_kernel<double,8,FULL=0,BVH=0,SMEM=1>": "_ZN2rt13render_kernelIdLi8ELb0ELb0ELb1E",
    "rt::wf_level_kernel<double,FULL=1,BVH=1,SMEM=0>": "_ZN2rt15wf_level_kernelIdLb1ELb1ELb0E",
}


def histogram(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    per = collections.defaultdict(collections.Counter)
    cur = None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)", ln)
        if m and cur:
            per[cur][m.group(2)] += 1
    return per


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ray_tracer_challenge_rs_b200", "librtgpu.so")
    per = histogram(lib)
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print(f"# SASS opcode histogram of `{os.path.relpath(lib, ROOT)}` (cuobjdump -sass, sm_100a)\n")
    print(f"{len(per)} functions, {sum(total.values())} instructions.\n")
    print("Blackwell data movement: **UBLKCP {} (cp.async.bulk), SYNCS {} (mbarrier)**; tensor-core opcodes (HMMA / UTCMMA / "
          "tcgen05): {} — by design, the path is scalar FP64.\n".format(total["UBLKCP"], total["SYNCS"],
                                                                         sum(v for k, v in total.items() if "MMA" in k)))
    print("Programmatic dependent launch: **PREEXIT {} (griddepcontrol.launch_dependents), ACQBULK {} (griddepcontrol.wait)**.  "
          "Single precision beside the FP64 exact tests (bounding-sphere pre-test, BVH box tests): FFMA {}, FSETP {}; packed FFMA2 {}.\n".format(
              total["PREEXIT"], total["ACQBULK"], total["FFMA"], total["FSETP"], total["FFMA2"]))
    print("| scope | instr | DFMA | DMUL | DADD | DSETP | MUFU | FFMA | FSETP | LDS | STS | LDL | STL | LDG | STG | BRA | BSSY+BSYNC | UBLKCP | SYNCS | PREEXIT | ACQBULK |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")

    def row(name, c):
        cols = ["DFMA", "DMUL", "DADD", "DSETP", "MUFU", "FFMA", "FSETP", "LDS", "STS", "LDL", "STL", "LDG", "STG", "BRA"]
        print(f"| {name} | {sum(c.values())} | " + " | ".join(str(c[k]) for k in cols) +
              f" | {c['BSSY'] + c['BSYNC']} | {c['UBLKCP']} | {c['SYNCS']} | {c['PREEXIT']} | {c['ACQBULK']} |")

    row("whole library", total)
    for label, key in HOT.items():
        for fn, c in per.items():
            if key in fn:
                row(label, c)
    print("\nTop 25 opcodes, whole library: " + ", ".join(f"{k} {v}" for k, v in total.most_common(25)))


if __name__ == "__main__":
    main()
