"""Claims DESIGN.md makes about the instruction stream of the shipped library, checked with cuobjdump (no GPU needed):
the scene tables are staged with the TMA unit (cp.async.bulk -> UBLKCP, mbarrier -> SYNCS), and the path uses no
tensor-core instruction.  Round 1 lost the TMA staging to a preprocessor-order bug without any test noticing."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "profiles", "tools"))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")
def test_tma_staging_is_in_the_shipped_sass():
    from ray_tracer_challenge_rs_b200 import build
    import sass_histogram

    per = sass_histogram.histogram(build.build())
    smem_kernels = {fn: c for fn, c in per.items() if ("render_kernelI" in fn or "wf_level_kernelI" in fn) and fn.split("EEEv")[0].endswith("Lb1")}
    assert smem_kernels, "no SMEM kernel variants found"
    for fn, c in smem_kernels.items():
        assert c["UBLKCP"] >= 2 and c["SYNCS"] >= 2, (fn, c["UBLKCP"], c["SYNCS"])  # two bulk copies, mbarrier init + wait
    total = sum((c for c in per.values()), start=type(next(iter(per.values())))())
    assert not any("MMA" in op for op in total), "tensor-core instructions on a scalar FP64 path?"


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")
def test_launch_chaining_and_the_single_precision_pretest_are_in_the_shipped_sass():
    """every kernel of the wavefront chain releases its successor (griddepcontrol.launch_dependents ->
    PREEXIT) and waits for its predecessor (griddepcontrol.wait -> ACQBULK); the f64 level kernel's shape loop tests the
    cull spheres in single precision (FSETP + FFMA next to the DFMA of the exact tests)."""
    from ray_tracer_challenge_rs_b200 import build
    import sass_histogram

    per = sass_histogram.histogram(build.build())
    chain = {fn: c for fn, c in per.items() if any(k in fn for k in ("wf_level_kernelI", "wf_bin_kernel", "wf_combine_kernelI"))}
    assert len(chain) >= 10
    for fn, c in chain.items():
        assert c["PREEXIT"] >= 1 and c["ACQBULK"] >= 1, (fn, c["PREEXIT"], c["ACQBULK"])
    f64_level = [c for fn, c in per.items() if "wf_level_kernelIdLb0ELb0ELb1E" in fn]
    assert len(f64_level) == 1
    assert f64_level[0]["FSETP"] >= 3 and f64_level[0]["FFMA"] >= 8 and f64_level[0]["DFMA"] >= 100
