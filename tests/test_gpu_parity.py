"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle.

Bars (BASELINE.json north_star):
  * f64 parity mode: 8-bit output within +-1 LSB on >= 99.9 % of pixels, every mismatch listed.
    What we actually assert is much tighter: RGB8 byte-identical, the f64 buffer equal except for
    the few pixels where CUDA's pow() and glibc's differ in the last bits (material.rs:110 is the
    only transcendental on the path), ray counters integer-equal.
  * shard / multi-GPU: N row-band shards == the unsharded frame, byte for byte.
  * at the reference's native sizes: sha256 of the GPU RGB8 frame == sha256 of the reference's PNG.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle.oracle import Oracle
from ray_tracer_challenge_rs_b200 import abi
from ray_tracer_challenge_rs_b200.fixtures import SHIPPED_SCENES, load_scene_fixture
from ray_tracer_challenge_rs_b200.render import Renderer, device_count, render_gpu

from worlds import SPECIAL_WORLDS

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "golden_index.json")) as f:
    INDEX = json.load(f)

COUNTERS = ("rays_primary", "rays_shadow", "rays_reflect", "rays_refract", "hit_nodes", "pixels")
F64_RTOL = 1e-12  # pow() ulp noise only; anything structural is orders of magnitude larger


def compare_with_oracle(flat, camera, max_depth=6, label="", family=None):
    canvas, gstats = render_gpu(camera, flat, max_depth=max_depth, return_stats=True, family=family)
    rgb, rgb8, ostats = Oracle(flat).render(camera, max_depth=max_depth, threads=0)
    g8 = canvas.to_rgb8().reshape(-1, 3)
    w = camera.horizontal_size
    # 1. bytes
    bad8 = np.argwhere((g8 != rgb8).any(axis=1)).ravel()
    listing = [(int(i % w), int(i // w), g8[i].tolist(), rgb8[i].tolist()) for i in bad8[:20]]
    assert bad8.size == 0, f"{label}: {bad8.size} RGB8 mismatches (x, y, gpu, oracle): {listing}"
    # 2. f64 colours
    diff = np.abs(canvas.pixels - rgb)
    scale = np.maximum(np.abs(rgb), 1.0)
    worst = float((diff / scale).max()) if diff.size else 0.0
    assert worst <= F64_RTOL, f"{label}: max relative f64 error {worst:.3e}"
    # 3. work counters
    for k in COUNTERS:
        assert gstats[k] == ostats[k], (label, k, gstats[k], ostats[k])
    n_differ = int((canvas.pixels != rgb).any(axis=1).sum())
    return n_differ, worst


@pytest.mark.parametrize("name", SHIPPED_SCENES)
def test_shipped_scene_small(name):
    flat, camera = load_scene_fixture(name)
    cam = camera.resized(384, 384 * camera.vertical_size // camera.horizontal_size)
    n_differ, worst = compare_with_oracle(flat, cam, label=name)
    print(f"{name}: {n_differ} pixels differ in f64 (max rel {worst:.2e})")


@pytest.mark.parametrize("name", sorted(SPECIAL_WORLDS))
def test_special_world(name):
    world, camera = SPECIAL_WORLDS[name]()
    compare_with_oracle(world.flatten(), camera, label=name)


def test_empty_shard_renders_nothing():
    """More shards than row bands: the shards without rows succeed, report zero work and leave the frame untouched
    (a fresh context has no staging buffers yet)."""
    flat, camera = load_scene_fixture("three_sphere_scene")
    cam = camera.resized(64, 40)  # 16-row bands: 3 bands, shards 3..7 of 8 are empty
    for family in ("persistent", "wavefront"):
        with Renderer(flat) as r:
            assert r.rows_count(cam, (16, 5, 8)) == 0
            rgb = np.full((64 * 40, 3), 7.0)
            rgb8 = np.full((64 * 40, 3), 9, np.uint8)
            _, _, st = r.render(cam, rows=(16, 5, 8), out_rgb=rgb, out_rgb8=rgb8, family=family)
            assert st["rays"] == 0 and st["pixels"] == 0
            assert (rgb == 7.0).all() and (rgb8 == 9).all()
            whole, whole8, _ = r.render(cam, family=family)
            parts, parts8 = np.zeros_like(whole), np.zeros_like(whole8)
            for index in range(8):
                r.render(cam, rows=(16, index, 8), out_rgb=parts, out_rgb8=parts8, family=family)
            assert np.array_equal(parts, whole) and np.array_equal(parts8, whole8)


def test_output_buffers_are_validated():
    flat, camera = load_scene_fixture("three_sphere_scene")
    cam = camera.resized(32, 16)
    with Renderer(flat) as r:
        with pytest.raises(ValueError):
            r.render(cam, out_rgb=np.zeros((32 * 16, 3), np.float32))  # f64 mode writes 8-byte elements
        with pytest.raises(ValueError):
            r.render(cam, out_rgb=np.zeros((32 * 8, 3)))  # half a frame
        with pytest.raises(ValueError):
            r.render(cam, out_rgb=np.zeros((32 * 16, 6))[:, ::2])  # not contiguous
        with pytest.raises(ValueError):
            r.render(cam, precision="f32", out_rgb=np.zeros((32 * 16, 3)))  # f32 mode writes 4-byte elements
        with pytest.raises(ValueError):
            r.render(cam, out_rgb8=np.zeros((32 * 16, 3), np.int8))


def test_one_shot_call_reuses_the_resident_scene():
    """rtgpu_render keeps one context per device: a second frame of a byte-identical scene description skips packing
    and upload, a changed description is uploaded again — the pixels say which scene was rendered."""
    flat_a, camera = load_scene_fixture("three_sphere_scene")
    flat_b, _ = load_scene_fixture("metal")
    cam = camera.resized(96, 64)
    a1 = render_gpu(cam, flat_a).to_rgb8()
    a2 = render_gpu(cam, flat_a).to_rgb8()
    b1 = render_gpu(cam, flat_b).to_rgb8()
    a3 = render_gpu(cam, flat_a).to_rgb8()
    assert np.array_equal(a1, a2) and np.array_equal(a1, a3) and not np.array_equal(a1, b1)
    assert np.array_equal(b1, Oracle(flat_b).render(cam)[1].reshape(b1.shape))
    assert np.array_equal(a3, Oracle(flat_a).render(cam)[1].reshape(a3.shape))


@pytest.mark.parametrize("max_depth", [0, 1, 3, 5, 7, 9])
def test_recursion_depths(max_depth):
    flat, camera = load_scene_fixture("refraction")
    compare_with_oracle(flat, camera.resized(96, 96), max_depth=max_depth, label=f"refraction@{max_depth}")
    world, camera = SPECIAL_WORLDS["mirror_box"]()
    compare_with_oracle(world.flatten(), camera, max_depth=max_depth, label=f"mirror_box@{max_depth}")


@pytest.mark.parametrize("size", [(1, 1), (7, 5), (33, 17), (130, 3), (3, 130)])
def test_ragged_sizes(size):
    flat, camera = load_scene_fixture("cover")
    compare_with_oracle(flat, camera.resized(*size), label=f"cover@{size}")


@pytest.mark.parametrize("name", SHIPPED_SCENES)
def test_native_size_matches_reference_png(name):
    """GPU render at the reference's native camera size == the reference's own PNG, byte for byte."""
    flat, camera = load_scene_fixture(name)
    meta = INDEX[name]
    canvas = render_gpu(camera, flat, want_rgb=False)
    frame = canvas.to_rgb8()
    with np.load(os.path.join(GOLDEN, f"{name}.rows.npz")) as z:
        rows, golden_rows = z["rows"], z["rgb8"]
    bad = np.argwhere((frame[rows] != golden_rows).any(axis=2))
    assert bad.shape[0] == 0, f"{name}: {bad.shape[0]} mismatching sampled pixels, first (row, x): {[(int(rows[r]), int(x)) for r, x in bad[:10]]}"
    assert hashlib.sha256(np.ascontiguousarray(frame).tobytes()).hexdigest() == meta["sha256_rgb8"]


def test_cover_1080p_against_oracle():
    """BASELINE.json config 2: cover.yaml at 1920x1080, f64 parity mode vs the CPU image."""
    flat, camera = load_scene_fixture("cover")
    n_differ, worst = compare_with_oracle(flat, camera.resized(1920, 1080), label="cover@1080p")
    print(f"cover@1080p: {n_differ} pixels differ in f64 (max rel {worst:.2e})")


# ---- BASELINE.json configs[2..4] at the sizes BASELINE.json states -------------------------------------------------
@pytest.mark.parametrize("name,family", [("reflect_refract", "wavefront"), ("refraction", "wavefront"), ("refraction", "persistent")])
def test_baseline_config2_4k_depth5(name, family):
    """configs[2]: reflect_refract.yaml + refraction.yaml at 3840x2160, max recursion depth 5, whole frame vs the oracle."""
    flat, camera = load_scene_fixture(name)
    n_differ, worst = compare_with_oracle(flat, camera.resized(3840, 2160), max_depth=5, label=f"{name}@4k/5", family=family)
    print(f"{name}@3840x2160 depth 5 ({family}): {n_differ} pixels differ in f64 (max rel {worst:.2e})")


@pytest.mark.parametrize("name", ["cylinders", "table", "shadow_puppets"])
def test_baseline_config3_1080p(name):
    """configs[3]: cube / cylinder / cone intersections and multi-light shadow rays at 1920x1080, whole frame vs the oracle."""
    flat, camera = load_scene_fixture(name)
    for family in ("persistent", "wavefront"):
        n_differ, worst = compare_with_oracle(flat, camera.resized(1920, 1080), label=f"{name}@1080p", family=family)
        print(f"{name}@1920x1080 ({family}): {n_differ} pixels differ in f64 (max rel {worst:.2e})")


def test_baseline_config4_synthetic_100k_at_8k():
    """configs[4]: 10^5 spheres + triangles with random materials and patterns at 7680x4320 (the BVH path), 4096
    sampled pixels against the brute-force oracle (a whole 8K frame of 10^5 shapes is a day of CPU time), plus the
    frame's sharding property: 8 interleaved row-band shards reproduce the unsharded frame byte for byte."""
    from ray_tracer_challenge_rs_b200.synthetic import synthetic_camera, synthetic_scene

    flat = synthetic_scene(100000)
    cam = synthetic_camera(7680, 4320)
    with Renderer(flat) as r:
        rgb, rgb8, st = r.render(cam)
        px = np.random.default_rng(4).integers(0, 7680 * 4320, 4096).astype(np.uint64)
        ref, ref8, ost = Oracle(flat).render_pixels(cam, px)
        idx = px.astype(np.int64)
        assert np.array_equal(rgb8[idx], ref8), "RGB8 of sampled pixels differs from the oracle"
        worst = float((np.abs(rgb[idx] - ref) / np.maximum(np.abs(ref), 1.0)).max())
        assert worst <= F64_RTOL, worst
        assert st["pixels"] == 7680 * 4320
        digest = hashlib.sha256(rgb8.tobytes()).hexdigest()
        parts8 = np.zeros_like(rgb8)
        for index in range(8):
            r.render(cam, rows=(16, index, 8), out_rgb8=parts8, want_rgb=False)
        assert hashlib.sha256(parts8.tobytes()).hexdigest() == digest
    print(f"synthetic 1e5 @ 8K: {st['rays']} rays, kernel {st['kernel_ms']:.1f} ms, sample max rel {worst:.2e}")


@pytest.mark.parametrize("band,count", [(16, 2), (16, 4), (4, 8), (5, 3)])
def test_row_band_shards_equal_whole_frame(band, count):
    """What N GPUs would each render (rtgpu_rows) assembles to exactly the unsharded frame."""
    flat, camera = load_scene_fixture("reflect_refract")
    cam = camera.resized(320, 214)
    with Renderer(flat) as r:
        whole, whole8, wstats = r.render(cam)
        rgb = np.full_like(whole, np.nan)
        rgb8 = np.zeros_like(whole8)
        totals = dict.fromkeys(COUNTERS, 0)
        for index in range(count):
            _, _, st = r.render(cam, rows=(band, index, count), out_rgb=rgb, out_rgb8=rgb8)
            for k in COUNTERS:
                totals[k] += st[k]
    assert np.array_equal(rgb.view(np.uint64), whole.view(np.uint64))
    assert np.array_equal(rgb8, whole8)
    assert totals == {k: wstats[k] for k in COUNTERS}


def test_pinned_output_is_written_zero_copy_and_equals_staged(monkeypatch):
    """Host-buffer renders into pinned memory (kernel writes the caller's Canvas directly) == staged copies,
    for the whole frame and for row-band shards."""
    from ray_tracer_challenge_rs_b200.render import PinnedArray

    flat, camera = load_scene_fixture("cover")
    cam = camera.resized(320, 180)
    n = 320 * 180
    with Renderer(flat) as r:
        staged, staged8, s0 = r.render(cam)
        pin, pin8 = PinnedArray((n, 3), np.float64), PinnedArray((n, 3), np.uint8)
        pin.array[:] = np.nan
        _, _, s1 = r.render(cam, out_rgb=pin.array, out_rgb8=pin8.array)
        assert np.array_equal(pin.array.view(np.uint64), staged.view(np.uint64)) and np.array_equal(pin8.array, staged8)
        pin.array[:] = np.nan
        pin8.array[:] = 0
        for index in range(3):
            r.render(cam, rows=(8, index, 3), out_rgb=pin.array, out_rgb8=pin8.array)
        assert np.array_equal(pin.array.view(np.uint64), staged.view(np.uint64)) and np.array_equal(pin8.array, staged8)
        monkeypatch.setenv("RTGPU_ZEROCOPY", "0")
        pin.array[:] = np.nan
        r.render(cam, out_rgb=pin.array, out_rgb8=pin8.array)
        assert np.array_equal(pin.array.view(np.uint64), staged.view(np.uint64))
        assert {k: s0[k] for k in COUNTERS} == {k: s1[k] for k in COUNTERS}
        pin.close()
        pin8.close()


def test_multi_gpu_one_shot_is_byte_identical():
    n = device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    flat, camera = load_scene_fixture("cover")
    cam = camera.resized(640, 360)
    one = render_gpu(cam, flat, n_gpus=1)
    for g in sorted({2, min(n, 4), n}):
        many = render_gpu(cam, flat, n_gpus=g, band_rows=8)
        assert np.array_equal(many.pixels.view(np.uint64), one.pixels.view(np.uint64)), g
        assert np.array_equal(many.to_rgb8(), one.to_rgb8()), g


# The f32 fast mode's STATED TOLERANCE (BASELINE.json north_star): the fraction of pixels whose
# 8-bit output is within 2 LSB of the f64 oracle's, per shipped scene at 384 pixels wide.  Measured values are
# 0.2 - 0.8 percentage points above these bars (profiles/r2_notes.md); the remaining outliers sit on silhouette and
# shadow edges, and in `refraction` on paths through five nested glass spheres.
F32_BARS = {
    "three_sphere_scene": 0.9999,
    "shadow_puppets": 0.999,
    "cylinders": 0.998,
    "metal": 0.9995,
    "table": 0.999,
    "reflect_refract": 0.999,
    "refraction": 0.985,
    "cover": 0.999,
}


@pytest.mark.parametrize("name", SHIPPED_SCENES)
def test_f32_fast_mode_tolerance(name):
    """f32 fast mode (offset scaled to the hit point, closest-approach discriminants for spheres and cylinders;
    SURVEY.md 0.6): the stated tolerance of every shipped scene is asserted, for both kernel families."""
    flat, camera = load_scene_fixture(name)
    cam = camera.resized(384, 384 * camera.vertical_size // camera.horizontal_size)
    _, rgb8, _ = Oracle(flat).render(cam, want_rgb=False)
    for family in ("persistent", "wavefront"):
        canvas = render_gpu(cam, flat, precision="f32", family=family)
        d = np.abs(canvas.to_rgb8().reshape(-1, 3).astype(int) - rgb8.astype(int)).max(axis=1)
        frac = float((d <= 2).mean())
        print(f"f32 {name}: {frac * 100:.3f}% within 2 LSB, max {int(d.max())}")
        assert frac >= F32_BARS[name], (name, family, frac)


# ---------------------------------------------------------------------------------------------
# BVH traversal (scenes with many bounded shapes).  RTGPU_BVH_MIN forces it on small scenes too.


@pytest.fixture
def force_bvh(monkeypatch):
    monkeypatch.setenv("RTGPU_BVH_MIN", "1")


@pytest.mark.parametrize("name", SHIPPED_SCENES)
def test_bvh_shipped_scene_small(name, force_bvh):
    flat, camera = load_scene_fixture(name)
    cam = camera.resized(384, 384 * camera.vertical_size // camera.horizontal_size)
    compare_with_oracle(flat, cam, label=f"bvh:{name}")


@pytest.mark.parametrize("name", sorted(SPECIAL_WORLDS))
def test_bvh_special_world(name, force_bvh):
    world, camera = SPECIAL_WORLDS[name]()
    compare_with_oracle(world.flatten(), camera, label=f"bvh:{name}")


@pytest.mark.parametrize("name", ["cover", "table", "cylinders", "refraction"])
def test_bvh_native_size_matches_reference_png(name, force_bvh):
    flat, camera = load_scene_fixture(name)
    canvas = render_gpu(camera, flat, want_rgb=False)
    assert hashlib.sha256(np.ascontiguousarray(canvas.to_rgb8()).tobytes()).hexdigest() == INDEX[name]["sha256_rgb8"]


def test_bvh_and_flat_traversal_agree_bit_for_bit(monkeypatch):
    flat, camera = load_scene_fixture("reflect_refract")
    cam = camera.resized(480, 320)
    monkeypatch.setenv("RTGPU_BVH_MIN", "0")
    a, sa = render_gpu(cam, flat, return_stats=True)
    monkeypatch.setenv("RTGPU_BVH_MIN", "1")
    b, sb = render_gpu(cam, flat, return_stats=True)
    assert np.array_equal(a.pixels.view(np.uint64), b.pixels.view(np.uint64))
    assert {k: sa[k] for k in COUNTERS} == {k: sb[k] for k in COUNTERS}


def test_uniform_lists_with_many_bounded_shapes(monkeypatch):
    monkeypatch.setenv("RTGPU_BVH_MIN", "0")  # no hierarchy, > 32 bounded shapes -> uniform lists
    from ray_tracer_challenge_rs_b200.synthetic import synthetic_camera, synthetic_scene

    compare_with_oracle(synthetic_scene(300, extent=4.0), synthetic_camera(96, 54, distance=11.0), label="uniform300")


@pytest.mark.parametrize("n_shapes,extent,size", [(40, 3.0, (160, 90)), (3000, 9.0, (192, 108)), (20000, 20.0, (128, 72))])
def test_synthetic_scene_against_oracle(n_shapes, extent, size):
    """BASELINE.json configs[4] at test scale: spheres + triangles, random materials / patterns, 2 lights."""
    from ray_tracer_challenge_rs_b200.synthetic import synthetic_camera, synthetic_scene

    flat = synthetic_scene(n_shapes, extent=extent)
    cam = synthetic_camera(*size, distance=2.6 * extent)
    n_differ, worst = compare_with_oracle(flat, cam, label=f"synthetic{n_shapes}")
    print(f"synthetic {n_shapes}: {n_differ} pixels differ in f64 (max rel {worst:.2e})")


@pytest.mark.parametrize("bvh_min", ["32", "1"])
def test_showcase_yaml_scene(bvh_min, monkeypatch):
    """tests/scenes/showcase.yaml through the loader: cones and cylinders from YAML, all four pattern kinds,
    two lights, define / extend — flat traversal and BVH."""
    from ray_tracer_challenge_rs_b200 import load_scene_description

    monkeypatch.setenv("RTGPU_BVH_MIN", bvh_min)
    world, camera = load_scene_description(os.path.join(os.path.dirname(os.path.abspath(__file__)), "scenes", "showcase.yaml"))
    compare_with_oracle(world.flatten(), camera, label=f"showcase:bvh_min={bvh_min}")


@pytest.mark.parametrize("seed", range(24))
def test_random_world_fuzz(seed, monkeypatch):
    """Seeded random worlds (tests/random_worlds.py): all shape and pattern kinds under rotations, shears and
    anisotropic scales, glass / mirror / matte materials, duplicates — flat traversal and, every other seed, BVH."""
    from random_worlds import random_world

    if seed % 2:
        monkeypatch.setenv("RTGPU_BVH_MIN", "1")
    world, cam = random_world(1000 + seed)
    compare_with_oracle(world.flatten(), cam, label=f"fuzz{seed}")


# ---------------------------------------------------------------------------------------------
# The wavefront kernel family (csrc/rt_wavefront.cuh): same pixels, one launch per recursion level.


@pytest.mark.parametrize("name", SHIPPED_SCENES)
def test_wavefront_shipped_scene_small(name):
    flat, camera = load_scene_fixture(name)
    cam = camera.resized(384, 384 * camera.vertical_size // camera.horizontal_size)
    compare_with_oracle(flat, cam, label=f"wavefront:{name}", family="wavefront")


@pytest.mark.parametrize("name", sorted(SPECIAL_WORLDS))
def test_wavefront_special_world(name):
    world, camera = SPECIAL_WORLDS[name]()
    compare_with_oracle(world.flatten(), camera, label=f"wavefront:{name}", family="wavefront")


@pytest.mark.parametrize("max_depth", [0, 1, 3, 5, 7, 9])
def test_wavefront_recursion_depths(max_depth):
    flat, camera = load_scene_fixture("refraction")
    compare_with_oracle(flat, camera.resized(96, 96), max_depth=max_depth, label=f"wavefront:refraction@{max_depth}", family="wavefront")


@pytest.mark.parametrize("seed", range(12))
def test_wavefront_random_world_fuzz(seed, monkeypatch):
    from random_worlds import random_world

    if seed % 2:
        monkeypatch.setenv("RTGPU_BVH_MIN", "1")
    world, cam = random_world(2000 + seed)
    compare_with_oracle(world.flatten(), cam, label=f"wavefront:fuzz{seed}", family="wavefront")


def test_wavefront_equals_persistent_bit_for_bit():
    """Both families perform the same operations per node: identical f64 frames, identical counters — also on a
    frame heavy enough (23 rays per pixel) to overflow the first-guess ray / node buffers and be re-rendered."""
    for name, size in (("refraction", (512, 512)), ("cover", (640, 360)), ("cylinders", (480, 240))):
        flat, camera = load_scene_fixture(name)
        cam = camera.resized(*size)
        a, sa = render_gpu(cam, flat, return_stats=True, family="persistent")
        b, sb = render_gpu(cam, flat, return_stats=True, family="wavefront")
        assert np.array_equal(a.pixels.view(np.uint64), b.pixels.view(np.uint64)), name
        assert np.array_equal(a.to_rgb8(), b.to_rgb8()), name
        assert {k: sa[k] for k in COUNTERS} == {k: sb[k] for k in COUNTERS}, name


@pytest.mark.parametrize("name", ["cover", "refraction", "table"])
def test_wavefront_native_size_matches_reference_png(name):
    flat, camera = load_scene_fixture(name)
    canvas = render_gpu(camera, flat, want_rgb=False, family="wavefront")
    assert hashlib.sha256(np.ascontiguousarray(canvas.to_rgb8()).tobytes()).hexdigest() == INDEX[name]["sha256_rgb8"]


def test_wavefront_shards_and_synthetic():
    flat, camera = load_scene_fixture("reflect_refract")
    cam = camera.resized(320, 214)
    with Renderer(flat) as r:
        whole, whole8, wstats = r.render(cam, family="wavefront")
        rgb = np.full_like(whole, np.nan)
        rgb8 = np.zeros_like(whole8)
        for index in range(3):
            r.render(cam, rows=(8, index, 3), out_rgb=rgb, out_rgb8=rgb8, family="wavefront")
    assert np.array_equal(rgb.view(np.uint64), whole.view(np.uint64)) and np.array_equal(rgb8, whole8)
    from ray_tracer_challenge_rs_b200.synthetic import synthetic_camera, synthetic_scene

    compare_with_oracle(synthetic_scene(3000, extent=9.0), synthetic_camera(192, 108, distance=23.4), label="wavefront:synthetic3000", family="wavefront")


def test_auto_family_measures_then_settles(monkeypatch):
    """family=None: a context times two frames of each family (W, P, W, P), then keeps one; every frame is the
    same bits whichever family rendered it.  Scenes without reflective / transparent materials never leave the
    persistent kernel."""
    monkeypatch.delenv("RTGPU_FAMILY", raising=False)
    monkeypatch.delenv("RTGPU_WAVEFRONT", raising=False)
    flat, camera = load_scene_fixture("cover")
    cam = camera.resized(320, 180)
    with Renderer(flat) as r:
        frames = [r.render(cam) for _ in range(8)]
        other = [r.render(cam.resized(160, 90))[2]["family"] for _ in range(2)]  # another frame shape: measured afresh
    families = [st["family"] for _, _, st in frames]
    assert families[:4] == ["wavefront", "persistent", "wavefront", "persistent"], families
    assert len(set(families[4:])) == 1, families
    assert other == ["wavefront", "persistent"], other
    for rgb, rgb8, st in frames[1:]:
        assert np.array_equal(rgb.view(np.uint64), frames[0][0].view(np.uint64))
        assert np.array_equal(rgb8, frames[0][1])
        assert {k: st[k] for k in COUNTERS} == {k: frames[0][2][k] for k in COUNTERS}
    flat, camera = load_scene_fixture("three_sphere_scene")
    with Renderer(flat) as r:
        assert [r.render(camera.resized(160, 80))[2]["family"] for _ in range(5)] == ["persistent"] * 5


@pytest.mark.timeout(180, method="thread")
def test_wavefront_bvh_large_frame_terminates_and_matches_persistent():
    """Regression: with 10^4 shapes at >= 720p the level-0 launch of one build never finished (full-mask warp
    votes in the BVH leaf batching; see trace_bvh).  Bit-equality with the persistent family on the same frame."""
    from ray_tracer_challenge_rs_b200.synthetic import synthetic_camera, synthetic_scene

    flat, cam = synthetic_scene(10000), synthetic_camera(1280, 720)
    with Renderer(flat) as r:
        a, _, sa = r.render(cam, want_rgb8=False, family="persistent")
        b, _, sb = r.render(cam, want_rgb8=False, family="wavefront")
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    assert {k: sa[k] for k in COUNTERS} == {k: sb[k] for k in COUNTERS}


def test_auto_family_survives_without_room_for_the_queues(monkeypatch):
    """RTGPU_WF_MAX_BYTES caps the wavefront family's buffers: automatic mode then stays with the persistent kernel
    (also when only the enlarged buffers of an overflowing frame do not fit); asking for the family explicitly fails."""
    monkeypatch.delenv("RTGPU_FAMILY", raising=False)
    monkeypatch.delenv("RTGPU_WAVEFRONT", raising=False)
    flat, camera = load_scene_fixture("refraction")
    cam = camera.resized(256, 256)
    with Renderer(flat) as r:
        want, want8, _ = r.render(cam, family="persistent")
    monkeypatch.setenv("RTGPU_WF_MAX_BYTES", "1000000")
    with Renderer(flat) as r:
        frames = [r.render(cam) for _ in range(4)]
        assert [st["family"] for _, _, st in frames] == ["persistent"] * 4
        assert all(np.array_equal(rgb.view(np.uint64), want.view(np.uint64)) and np.array_equal(rgb8, want8) for rgb, rgb8, _ in frames)
        with pytest.raises(abi.RtgpuError) as err:
            r.render(cam, family="wavefront")
        assert err.value.status == abi.ERR_OUT_OF_MEMORY
    # room for the first guess (1 ray + 2 nodes per pixel, two queues = 105 MB at 512x512) but not for what this frame needs
    # (23 rays per pixel: test_wavefront_equals_persistent_bit_for_bit relies on the same overflow)
    monkeypatch.delenv("RTGPU_WF_MAX_BYTES")
    monkeypatch.setenv("RTGPU_E2E_CHUNKS", "1")  # one launch sequence for the whole frame, as the sizes below assume
    cam = camera.resized(512, 512)
    with Renderer(flat) as r:
        want, _, _ = r.render(cam, family="persistent", want_rgb8=False)
    monkeypatch.setenv("RTGPU_WF_MAX_BYTES", str(300 * 1000 * 1000))
    with Renderer(flat) as r:
        frames = [r.render(cam, want_rgb8=False) for _ in range(4)]
    assert all(np.array_equal(rgb.view(np.uint64), want.view(np.uint64)) for rgb, _, _ in frames)
    assert [st["family"] for _, _, st in frames[1:]] == ["persistent"] * 3


@pytest.mark.parametrize("split", [None, "1/2", "2/3", "3/4", "5/8", "7/8"])
@pytest.mark.parametrize("height", [300, 257, 64 * 5 + 1])
def test_wavefront_chunked_host_render_unequal_parts(split, height, monkeypatch):
    """The chunked host render takes `take` of every `period` 16-row bands first and the rest second (unequal parts:
    the second part's kernels hide the first part's copy): every split, ragged heights included, is the frame."""
    from ray_tracer_challenge_rs_b200.render import PinnedArray

    if split:
        monkeypatch.setenv("RTGPU_E2E_SPLIT", split)
    flat, camera = load_scene_fixture("cover")
    cam = camera.resized(1024, height)
    with Renderer(flat) as r:
        want, want8, wstats = r.render(cam, family="persistent")
        pin, pin8 = PinnedArray((1024 * height, 3)), PinnedArray((1024 * height, 3), np.uint8)
        pin.array[:] = -1.0
        _, _, st = r.render(cam, family="wavefront", out_rgb=pin.array, out_rgb8=pin8.array)
        assert np.array_equal(pin.array.view(np.uint64), want.view(np.uint64)) and np.array_equal(pin8.array, want8)
        assert {k: st[k] for k in COUNTERS} == {k: wstats[k] for k in COUNTERS}
        pin.close()
        pin8.close()


def test_wavefront_chunked_host_render_overflows_and_recovers(monkeypatch):
    """Host-buffer renders of the wavefront family into PINNED memory run as two interleaved parts whose copies
    overlap; with a tiny first guess for the queues both halves overflow, the buffers grow, the frame is rendered
    again — same bits as the persistent kernel.  Pageable buffers take the single-sequence path."""
    from ray_tracer_challenge_rs_b200.render import PinnedArray

    flat, camera = load_scene_fixture("refraction")
    cam = camera.resized(640, 512)  # >= 2^18 pixels: chunked; 512 rows = 32 bands; a ragged case follows
    n = 640 * 512
    pin, pin8 = PinnedArray((n, 3), np.float64), PinnedArray((n, 3), np.uint8)
    with Renderer(flat) as r:
        want, want8, wstats = r.render(cam, family="persistent")
    monkeypatch.setenv("RTGPU_WF_INITIAL_SCALE", "0.02")
    with Renderer(flat) as r:
        for _ in range(2):  # the first call enlarges the buffers and renders again, the second finds them large enough
            pin.array[:] = np.nan
            pin8.array[:] = 0
            _, _, st = r.render(cam, family="wavefront", out_rgb=pin.array, out_rgb8=pin8.array)
            assert np.array_equal(pin.array.view(np.uint64), want.view(np.uint64)) and np.array_equal(pin8.array, want8)
            assert {k: st[k] for k in COUNTERS} == {k: wstats[k] for k in COUNTERS}
        pageable, pageable8, _ = r.render(cam, family="wavefront")
        assert np.array_equal(pageable.view(np.uint64), want.view(np.uint64)) and np.array_equal(pageable8, want8)
    pin.close()
    pin8.close()
    monkeypatch.delenv("RTGPU_WF_INITIAL_SCALE")
    flat, camera = load_scene_fixture("cover")
    cam = camera.resized(700, 411)  # 25 full bands + 11 rows: the last band is partial and belongs to the second half
    n = 700 * 411
    pin, pin8 = PinnedArray((n, 3), np.float64), PinnedArray((n, 3), np.uint8)
    with Renderer(flat) as r:
        a, a8, _ = r.render(cam, family="persistent")
        pin.array[:] = np.nan
        r.render(cam, family="wavefront", out_rgb=pin.array, out_rgb8=pin8.array)
    assert np.array_equal(a.view(np.uint64), pin.array.view(np.uint64)) and np.array_equal(a8, pin8.array)
    pin.close()
    pin8.close()


@pytest.mark.gpu
@pytest.mark.parametrize("scene,width,height", [("cover", 640, 360), ("reflect_refract", 333, 201), ("refraction", 200, 150), ("cylinders", 320, 200),
                                                ("metal", 97, 61)])
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_binned_queues_render_the_same_bits(monkeypatch, scene, width, height, precision):
    """rt_wavefront.cuh wf_bin_kernel: consuming a level's queue grouped by (hit shape, reflected / refracted) changes
    which entries share a warp, never a bit of the frame or a counter; forced on (RTGPU_WF_BINS=1) for frames and scenes
    the default gate would leave in arrival order, and together with the overflow -> enlarge -> render-again path."""
    flat, camera = load_scene_fixture(scene)
    cam = camera.resized(width, height)
    bits = np.uint64 if precision == "f64" else np.uint32
    monkeypatch.setenv("RTGPU_WF_BINS", "0")
    with Renderer(flat) as r:
        want, want8, wstats = r.render(cam, family="wavefront", precision=precision)
    monkeypatch.setenv("RTGPU_WF_BINS", "1")
    with Renderer(flat) as r:
        got, got8, stats = r.render(cam, family="wavefront", precision=precision)
        again, again8, _ = r.render(cam, family="wavefront", precision=precision)
    assert np.array_equal(got.view(bits), want.view(bits)) and np.array_equal(got8, want8)
    assert np.array_equal(again.view(bits), want.view(bits)) and np.array_equal(again8, want8)
    assert {k: stats[k] for k in COUNTERS} == {k: wstats[k] for k in COUNTERS}
    monkeypatch.setenv("RTGPU_WF_INITIAL_SCALE", "0.02")  # queues far too small at first: binned frames overflow, grow, render again
    with Renderer(flat) as r:
        got, got8, stats = r.render(cam, family="wavefront", precision=precision)
    assert np.array_equal(got.view(bits), want.view(bits)) and np.array_equal(got8, want8)
    assert {k: stats[k] for k in COUNTERS} == {k: wstats[k] for k in COUNTERS}


@pytest.mark.gpu
def test_launch_count_tells_binned_from_plain_frames(monkeypatch):
    """rtgpu_context_launch_count counts at the launch sites: a host-buffer frame of the wavefront family is a level and
    a combine kernel per recursion level plus the counter commit and the status block (16 at depth 6), a binned one
    six more; the persistent family one kernel plus the status block.  Below 2^18 pixels the default gate leaves the
    queues in arrival order.  (The first frame of a context may be rendered twice: its queues start small.)"""
    flat, camera = load_scene_fixture("cover")
    cam = camera.resized(320, 180)
    with Renderer(flat) as r:
        r.render(cam, family="wavefront", want_rgb8=False)
        n0 = r.launch_count()
        r.render(cam, family="wavefront", want_rgb8=False)
        n1 = r.launch_count()
        r.render(cam, family="persistent", want_rgb8=False)
        n2 = r.launch_count()
        monkeypatch.setenv("RTGPU_WF_BINS", "1")
        r.render(cam, family="wavefront", want_rgb8=False)
        n3 = r.launch_count()
    assert (n1 - n0, n2 - n1, n3 - n2) == (2 * 7 + 2, 1 + 1, 2 * 7 + 6 + 2)
