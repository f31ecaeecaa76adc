"""csrc/rt_bvh.h: the multi-threaded BVH build gives the tree the single-threaded build gives (every index of the
depth-first layout is known before the subtree below it exists), and the tree is valid.  Host code only: compiled with
g++ from tests/native/bvh_threads.cpp."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not installed")
    exe = str(tmp_path_factory.mktemp("bvh") / "bvh_threads")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "ray_tracer_challenge_rs_b200", "csrc"), "-o", exe,
                    os.path.join(ROOT, "tests", "native", "bvh_threads.cpp")], check=True, capture_output=True)
    return exe


@pytest.mark.parametrize("n,kind", [(1, 0), (2, 0), (3, 0), (7, 0), (1000, 0), (40000, 0), (200000, 0), (200000, 1)])
def test_tree_is_valid_and_independent_of_the_thread_count(harness, n, kind):
    out = subprocess.run([harness, str(n), str(kind)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
