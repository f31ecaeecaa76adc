"""The -march=native build of the CPU restatement (oracle/Makefile `native`; bench.py reports it as
cpu_baseline.native) keeps contraction off: it must render the same bits as the parity build."""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = """
import hashlib, sys
sys.path.insert(0, %r)
from oracle import oracle as O
from ray_tracer_challenge_rs_b200.fixtures import load_scene_fixture
flat, camera = load_scene_fixture("cover")
rgb, rgb8, stats = O.Oracle(flat).render(camera.resized(96, 54))
print(hashlib.sha256(rgb.tobytes()).hexdigest(), hashlib.sha256(rgb8.tobytes()).hexdigest(), stats["rays"])
"""


def frame_hashes(library=None):
    env = dict(os.environ)
    env.pop("RTORACLE_LIBRARY", None)
    if library:
        env["RTORACLE_LIBRARY"] = library
    out = subprocess.run([sys.executable, "-c", SCRIPT % ROOT], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-1500:]
    return out.stdout.strip().split()


def test_native_build_renders_the_same_bits():
    sys.path.insert(0, ROOT)
    from oracle import oracle as O

    native = O.build_native()
    assert os.path.exists(native)
    assert frame_hashes(native) == frame_hashes()
