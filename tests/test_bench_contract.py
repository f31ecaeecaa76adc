"""bench.py's output contract: exactly one JSON line on stdout, with the keys the driver reads.

The CPU arm (`--impl reference`, the oracle on the host cores) runs everywhere; the B200 arm needs a device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def run_bench(*args):
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, f"stdout must hold exactly one line, got {len(lines)}: {proc.stdout[:500]}"
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line():
    line = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--width", "96", "--height", "54")
    assert line["impl"] == "reference"
    assert BASE_KEYS <= set(line)
    assert line["unit"] == "Mrays/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("cover.yaml")
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm must not shrink to one core because of it
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_reference_arm_ignores_omp_num_threads():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--width", "96",
                           "--height", "54"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert proc.returncode == 0, proc.stderr[-2000:]
    line = json.loads(proc.stdout.strip())
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert proc.returncode == 0 and proc.stdout.strip() == ""


@pytest.mark.gpu
def test_b200_arm_line_has_the_contract_keys():
    line = run_bench("--steps", "3", "--warmup", "3", "--width", "320", "--height", "180")
    assert BASE_KEYS | {"gpu_launches", "roofline", "clocks"} <= set(line)
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert line["family"] in ("persistent", "wavefront")
    assert set(line["config"]) == {"workload", "scene", "width", "height", "precision", "max_depth", "shapes", "lights", "cache"}  # = the reference arm's
    # counted by the library (rtgpu_context_launch_count): one kernel per frame of the persistent family; a level and a
    # combine kernel per recursion level, the status block and the counter commit for the wavefront family (no bin
    # kernels at this size)
    assert line["gpu_launches"] == 3 * {"persistent": 1, "wavefront": 2 * 7 + 2}[line["family"]]
    # pixel identity the driver can read: both legs' frames equal a 1-GPU render, and their RGB8 equals the oracle's
    assert line["frame"]["frame_matches_n1"] is True and len(line["frame"]["frame_sha256"]) == 64
    assert line["frame"]["rgb8_pixels_differing_from_oracle"] == 0
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-12
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] == 320 * 180 * 3 * 8 + 48
    assert line["e2e"]["value"] > 0 and line["e2e"]["value"] != line["value"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"])
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
