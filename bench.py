#!/usr/bin/env python
"""bench.py — the camera render pass on N B200s of one node, beside the CPU reference arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--scene cover] [--width 1920] [--height 1080] [--precision f64] [--max-depth 6]

A "step" = one full frame of the workload (default: BASELINE.json configs[1], the reference's
cover scene at 1920x1080, f64 parity mode, recursion depth 6).  Prints ONE JSON line:

  value        Mrays/s with everything resident on the device: the render kernel(s) only, CUDA events
               on the launching stream, summed over exactly K steps, max over ranks.  Rays =
               World::collect_intersections calls (primary + shadow + reflect + refract), counted on
               the device and integer-equal to the CPU oracle's count.
  e2e          the same metric through the reference-facing call rtgpu_render() (scene pack + H2D,
               kernel on N devices in row bands, D2H of the f64 Canvas into pinned host memory).
  roofline     FP64 (or FP32) FMA-pipe roofline of the render kernel: algorithmic flops per frame
               (SURVEY.md 8d table x device ray counters) / kernel time, against an FMA-chain peak
               measured live in this run (MEASURED_PEAKS.json has no FP64/FP32 entry).
  cpu_baseline the CPU oracle (restated reference, C + OpenMP, all host threads) on the same frame.

`--impl reference` times that CPU arm alone (the reference is Rust and cannot be built in this
image: the oracle port, validated byte-for-byte against the reference's golden renders, stands in).
Under torchrun (N > 1) every rank renders its own interleaved row bands; no collective touches the
data path — torch.distributed only carries the barrier and the max-over-ranks of the timings.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# algorithmic flops per ray-shape test (object-space transform + local_intersect,
# hit-path upper bound) and per hit node.
FLOP_PER_TEST = {"sphere": 60, "plane": 35, "cube": 45, "cylinder_open": 58, "cylinder_closed": 76,
                 "cone_open": 63, "cone_closed": 83, "triangle": 78}
FLOP_PER_HIT_NODE = 250
FLOP_PER_HIT_NODE_PER_LIGHT = 100
METRIC = "Mrays/s (primary+shadow+secondary)"


def flops_per_ray(flat) -> int:
    from ray_tracer_challenge_rs_b200 import abi

    total = 0
    for t, closed in zip(flat.shape_type.tolist(), flat.shape_closed.tolist()):
        if t == abi.SPHERE:
            total += FLOP_PER_TEST["sphere"]
        elif t == abi.PLANE:
            total += FLOP_PER_TEST["plane"]
        elif t == abi.CUBE:
            total += FLOP_PER_TEST["cube"]
        elif t == abi.CYLINDER:
            total += FLOP_PER_TEST["cylinder_closed" if closed else "cylinder_open"]
        elif t == abi.CONE:
            total += FLOP_PER_TEST["cone_closed" if closed else "cone_open"]
        else:
            total += FLOP_PER_TEST["triangle"]
    return total


def frame_flops(flat, stats) -> float:
    rays = stats["rays_primary"] + stats["rays_shadow"] + stats["rays_reflect"] + stats["rays_refract"]
    return float(rays) * flops_per_ray(flat) + float(stats["hit_nodes"]) * (FLOP_PER_HIT_NODE + FLOP_PER_HIT_NODE_PER_LIGHT * flat.n_lights)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.samples = []
        self._stop = threading.Event()
        self._thread = None
        # NVML directly when the binding is there (a query takes ~2 ms, so even a 15 ms timed region gets several
        # samples; initialised here, before the timed region); otherwise the nvidia-smi query of the profiling
        # recipe (a process spawn per sample)
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            reasons(handle)
            self._nvml = (pynvml, handle, reasons)
        except Exception:
            self._nvml = None

    def _run(self):
        if self._nvml is not None:
            pynvml, handle, reasons = self._nvml
            bits = (0x8, 0x40, 0x20, 0x4)  # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
            while not self._stop.is_set():
                try:
                    mask = int(reasons(handle))
                    self.samples.append([str(self.device_index), str(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)),
                                         str(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM)), "0"] +
                                        ["Active" if mask & bit else "Not Active" for bit in bits])
                except Exception:
                    break
                self._stop.wait(0.002)
            if self.samples:
                return
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i", str(self.device_index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                mx.append(float(s[2]))
                for name, v in zip(names, s[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_shapes(args) -> int:
    """--scene synthetic:<N> = BASELINE.json configs[4]: N spheres + triangles with random materials and patterns."""
    return int(args.scene.split(":", 1)[1]) if args.scene.startswith("synthetic:") else 0


def load_workload(args):
    if synthetic_shapes(args):
        from ray_tracer_challenge_rs_b200.synthetic import synthetic_camera, synthetic_scene

        return synthetic_scene(synthetic_shapes(args)), synthetic_camera(args.width, args.height)
    from ray_tracer_challenge_rs_b200.fixtures import load_scene_fixture

    flat, camera = load_scene_fixture(args.scene)
    return flat, camera.resized(args.width, args.height)


def oracle_sample(args, flat, camera):
    """Pixels the CPU arm renders per step: the whole frame for the shipped scenes; for the synthetic scene (brute
    force over N shapes per ray: a whole 8K frame of 10^5 shapes is a day of CPU time) a fixed pseudo-random sample
    sized for ~10 s on 16+ cores.  None = whole frame."""
    n = synthetic_shapes(args)
    if not n:
        return None
    total = camera.horizontal_size * camera.vertical_size
    count = int(min(total, max(512, 4.0e8 / n)))
    return np.random.default_rng(0xB200).choice(total, size=count, replace=False).astype(np.uint64)


def data_label(args) -> str:
    if synthetic_shapes(args):
        return (f"synthetic: {synthetic_shapes(args)} spheres + triangles with random materials and patterns, two lights (numpy PCG64, seed 0xB200; "
                "ray_tracer_challenge_rs_b200/synthetic.py) — BASELINE.json configs[4]")
    return (f"the reference's shipped scenes/{args.scene}.yaml (committed flattened fixture of the YAML), camera resized to "
            f"{args.width}x{args.height}; no dataset, no random geometry")


def quantise_rgb8(rgb: np.ndarray) -> np.ndarray:
    """Canvas::to_png_file's 8-bit quantisation (canvas.rs:117-123): clamp to [0, 1], * 255, round half away from zero."""
    v = np.nan_to_num(np.clip(rgb, 0.0, 1.0), nan=0.0) * 255.0
    r = np.floor(v)
    return (r + ((v - r) >= 0.5)).astype(np.uint8)


def workload_config(args, flat, extra=None) -> dict:
    what = (f"synthetic scaling scene, {synthetic_shapes(args)} shapes" if synthetic_shapes(args) else f"{args.scene}.yaml")
    which = "configs[4]" if synthetic_shapes(args) else "configs[1]" if (args.scene, args.width, args.height) == ("cover", 1920, 1080) else "configs[2..3]"
    cfg = {
        "workload": f"{what} @ {args.width}x{args.height}, {args.precision} {'parity' if args.precision == 'f64' else 'fast'} mode, "
                    f"max recursion depth {args.max_depth} (BASELINE.json {which})",
        "scene": args.scene, "width": args.width, "height": args.height, "precision": args.precision,
        "max_depth": args.max_depth, "shapes": flat.shape_counts(), "lights": flat.n_lights,
        "cache": "scene tables are ~3 KB staged in shared memory and the 50 MB frame is write-only, so L2 contents cannot help a step; "
                 "a 256 MiB buffer is overwritten between timed steps anyway (L2 flush, outside the event pairs)",
    }
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle on all host threads


def host_threads() -> int:
    """Host threads of the CPU arm: every core this process may run on.  Passed to the oracle explicitly —
    torchrun exports OMP_NUM_THREADS=1 to its workers, which must not shrink the CPU reference to one core."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def time_oracle(flat, camera, max_depth: int, steps: int, warmup: int, threads: int = 0, pixels=None):
    """Times the CPU oracle.  pixels = None: whole frames (returns the f64 frame); else only those pixel indices
    (returns their colours)."""
    from oracle import oracle as O

    if threads <= 0:
        threads = host_threads()
    if pixels is not None:
        orc = O.Oracle(flat)
        times, rgb, st = [], None, None
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            rgb, _, st = orc.render_pixels(camera, pixels, max_depth=max_depth, threads=threads)
            t1 = time.perf_counter()
            if i >= warmup:
                times.append(t1 - t0)
        return times, st, threads, rgb

    orc = O.Oracle(flat)
    n = camera.horizontal_size * camera.vertical_size
    from ray_tracer_challenge_rs_b200 import abi
    from ray_tracer_challenge_rs_b200.flatten import camera_to_c
    import ctypes as C

    cam = camera_to_c(camera)
    rgb = np.zeros((n, 3), np.float64)
    lib = O.lib()
    stats = abi.RtgpuStats()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        st = lib.rto_render(C.byref(orc._c), C.byref(cam), max_depth, None, threads, rgb.ctypes.data, None, C.byref(stats))
        t1 = time.perf_counter()
        assert st == 0
        if i >= warmup:
            times.append(t1 - t0)
    return times, stats.as_dict(), threads, rgb


def time_oracle_serial_sample(flat, camera, max_depth: int, stride: int = 16):
    """Serial (1 thread) rate of the oracle on every `stride`-th pixel of the frame — the reference's
    `--rendering-mode serial`, bounded to about a second of CPU work.  Returns (Mrays/s, pixels sampled)."""
    from oracle import oracle as O

    n = camera.horizontal_size * camera.vertical_size
    px = np.arange(0, n, stride, dtype=np.uint64)
    orc = O.Oracle(flat)
    t0 = time.perf_counter()
    _, _, st = orc.render_pixels(camera, px, max_depth=max_depth, threads=1)
    dt = time.perf_counter() - t0
    return st["rays"] / dt / 1e6, int(px.size)


def time_oracle_native(args):
    """the same CPU restatement built with -march=native ON THIS HOST (oracle/Makefile `native`),
    timed through this script's own reference arm in a child process; the favourable-to-CPU number beside the bit-parity
    build.  Returns a dict for cpu_baseline["native"]."""
    try:
        from oracle import oracle as O

        lib = O.build_native()
        env = dict(os.environ, RTORACLE_LIBRARY=lib)
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1", "--scene", args.scene, "--width", str(args.width),
               "--height", str(args.height), "--max-depth", str(args.max_depth)]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        line = json.loads(out.stdout.strip().splitlines()[-1])
        return {"value": line["value"], "unit": line["unit"], "cores": line["cpu_baseline"]["cores"], "ms_per_frame": line["ms_per_step"],
                "flags": "-O3 -march=native -ffp-contract=off (same bits as the parity build)"}
    except Exception as exc:  # no compiler on the box, ...: the parity build's number stands alone
        return {"unavailable": str(exc)[:200]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm runs once, on rank 0
    flat, camera = load_workload(args)
    sample = oracle_sample(args, flat, camera)
    times, stats, cores, _ = time_oracle(flat, camera, args.max_depth, args.steps, args.warmup, pixels=sample)
    total = sum(times)
    mrays = stats["rays"] * len(times) / total / 1e6
    sample_text = (f"the whole {args.width}x{args.height} frame" if sample is None else
                   f"{sample.size} pseudo-random pixels of the {args.width}x{args.height} frame (brute force over {flat.n_shapes} shapes per ray)")
    line = {
        "impl": "reference", "metric": METRIC, "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total / len(times) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": data_label(args),
        "config": workload_config(args, flat),
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_text}, {len(times)} times; restated reference (C + OpenMP, "
                                   "-ffp-contract=off), not rustc output — no Rust toolchain in this image",
                         "ms_per_frame": total / len(times) * 1e3, "best_ms_per_frame": min(times) * 1e3},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rays_per_frame": stats["rays"],
    }
    RESULT_LINE.append(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# B200 arm


RESULT_LINE = []  # filled by the arm that ran; printed by main() once stdout is back
CALIBRATION_FRAMES = 4  # 2 per kernel family (TUNE_RUNS in csrc/rtgpu.cu)


PROFILED_TRAFFIC = {  # (scene, width, height, precision, depth, family) -> committed ncu per-launch list of one frame
    ("cover", 1920, 1080, "f64", 6, "wavefront"): "profiles/r2d_wavefront_launches.csv",
    ("synthetic:100000", 7680, 4320, "f64", 6, "wavefront"): "profiles/r2_synthetic_1e5_8k_launches.csv",
}


def profiled_traffic(args, family: str):
    """DRAM bytes per frame (dram__bytes_read.sum + dram__bytes_write.sum over all launches of one frame) from the
    committed ncu launch list, for the configurations one was captured on; None for anything else.  (A profiler
    cannot run inside the timed bench: `traffic_model` beside it is computed in-run from the device's record counts.)"""
    rel = PROFILED_TRAFFIC.get((args.scene, args.width, args.height, args.precision, args.max_depth, family))
    path = os.path.join(ROOT, rel) if rel else None
    if not path or not os.path.exists(path):
        return None, None
    import csv

    total = 0.0
    with open(path) as f:
        for r in csv.reader(f):
            if len(r) > 10 and r[-3] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[-2], 1.0)
                total += float(r[-1].replace(",", "")) * scale
    return (total or None), f"{rel} (ncu, every launch of one frame)"


def hbm_peak_gbs():
    """MEASURED_PEAKS.json (driver-written) if present, else the profiling recipe's fallback for B200."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6550.0, "B200_PROFILING.md fallback (measured copy bandwidth of this pool's B200s)"




def run_b200(args):
    import torch
    import torch.distributed as dist

    from ray_tracer_challenge_rs_b200 import abi
    from ray_tracer_challenge_rs_b200.render import Renderer, last_family, measure_fma_peak

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    host_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # CPU-side barrier for the leg where rank 0 alone drives every device: an NCCL barrier would park
        # a spinning kernel on the other ranks' GPUs and steal SM time from the frame being measured
        host_group = dist.new_group(backend="gloo")

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=host_group)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(xs):
        t = torch.tensor(xs, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    flat, camera = load_workload(args)
    K, W = args.steps, max(args.warmup, 0)
    family = None if args.family == "auto" else args.family
    family_flag = {None: 0, "wavefront": abi.FLAG_WAVEFRONT, "persistent": abi.FLAG_PERSISTENT}[family]
    elem = 8 if args.precision == "f64" else 4
    tdtype = torch.float64 if args.precision == "f64" else torch.float32
    rows = (4, rank, world) if world > 1 else None  # one tile row per band: finest balance across ranks

    renderer = Renderer(flat, device=local_rank)
    my_rows = renderer.rows_count(camera, rows)
    d_out = torch.empty((max(my_rows, 1) * camera.horizontal_size, 3), dtype=tdtype, device="cuda")
    d_counters = torch.zeros(6, dtype=torch.int64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()

    def launch():
        renderer.render_device(camera, d_out.data_ptr(), 0, d_counters.data_ptr(), stream.cuda_stream, precision=args.precision,
                               max_depth=args.max_depth, rows=rows, family=family)

    # ---- device-resident timing ("value") ----
    # family "auto": the library times its first two frames of each kernel family (W, P, W, P) and keeps the
    # faster; those calibration frames come before the warm-up so that warm-up and timed steps run the settled one
    for _ in range(CALIBRATION_FRAMES if family is None else 0):
        launch()
    for _ in range(W):
        launch()
    barrier()
    d_counters.zero_()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    with ClockSampler(local_rank) as clocks:
        barrier()
        launches0 = renderer.launch_count()
        t_wall0 = time.perf_counter()
        for i in range(K):
            flush.fill_(i & 0xFF)  # L2 flush between timed steps (outside the event pair)
            starts[i].record(stream)
            launch()
            ends[i].record(stream)
        barrier()
        t_wall1 = time.perf_counter()
        launches_timed = renderer.launch_count() - launches0  # counted by the library at every kernel launch site
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    family_used = last_family()
    device_ms = max_over_ranks(sum(step_ms))
    gpu_launches = int(sum_over_ranks([float(launches_timed)])[0])
    counters = [int(v) // K for v in sum_over_ranks([float(v) for v in d_counters.tolist()])]
    stats = dict(zip(("rays_primary", "rays_shadow", "rays_reflect", "rays_refract", "hit_nodes", "pixels"), counters))
    rays = stats["rays_primary"] + stats["rays_shadow"] + stats["rays_reflect"] + stats["rays_refract"]
    value = rays * K / (device_ms * 1e-3) / 1e6
    clock_summary = clocks.summary()

    # ---- the frame the ranks just rendered, assembled on rank 0 from their shards (outside every timed region) ----
    import hashlib

    shard_host = d_out[: my_rows * camera.horizontal_size].cpu().numpy()
    frame_device = None
    if world == 1:
        frame_device = shard_host
    else:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((renderer.rows_list(camera, rows), shard_host), gathered, dst=0, group=host_group)
        if rank == 0:
            frame_device = np.zeros((camera.vertical_size, camera.horizontal_size, 3), shard_host.dtype)
            for row_ids, shard in gathered:
                frame_device[np.asarray(row_ids, dtype=np.int64)] = shard.reshape(len(row_ids), camera.horizontal_size, 3)
            frame_device = frame_device.reshape(-1, 3)

    # ---- what the wavefront family moved through HBM for one frame: record counts from the device (all ranks) ----
    records = {"queued_rays": 0, "node_records": 0, "ray_record_bytes": 0, "node_record_bytes": 0}
    if family_used == "wavefront":
        renderer.render(camera, precision=args.precision, max_depth=args.max_depth, rows=rows, want_rgb8=False, family="wavefront")
        records = renderer.frame_records()
    summed = sum_over_ranks([float(records["queued_rays"]), float(records["node_records"])])
    records["queued_rays"], records["node_records"] = int(summed[0]), int(summed[1])

    # ---- end to end through the reference-facing call, host buffers (rank 0 drives all N devices) ----
    n_px = camera.horizontal_size * camera.vertical_size
    e2e = None
    if rank == 0:
        host = torch.empty((n_px, 3), dtype=tdtype).pin_memory()
        host_np = host.numpy()
        lib = abi.load_library()
        import ctypes as C
        from ray_tracer_challenge_rs_b200.flatten import camera_to_c

        cscene, ccam = flat.as_c(), camera_to_c(camera)
        e2e_gpus = max(1, min(world, lib.rtgpu_device_count()))
        # one device: the band height only sets the granularity of the two overlapped chunks; several devices: the same
        # 4-row bands as the device-resident leg, so that no device ends up with a visibly dearer share of the rows
        e2e_band_rows = int(os.environ.get("RTGPU_BENCH_E2E_BAND_ROWS", "16" if e2e_gpus == 1 else "4"))
        opts = abi.RtgpuOpts(abi.PRECISION_F64 if args.precision == "f64" else abi.PRECISION_F32, args.max_depth, e2e_gpus, e2e_band_rows, family_flag)
        st = abi.RtgpuStats()

        def one_frame():
            abi.check(lib, lib.rtgpu_render(C.byref(cscene), C.byref(ccam), C.byref(opts), host_np.ctypes.data, None, C.byref(st)))

    if world > 1:
        renderer.close()  # rank 0's one-shot call owns every device for the e2e leg
    host_barrier()
    if rank == 0:
        for _ in range(W + (CALIBRATION_FRAMES if family is None else 0)):
            one_frame()
        t0 = time.perf_counter()
        for _ in range(K):
            one_frame()
        t1 = time.perf_counter()
        e2e_ms = (t1 - t0) * 1e3 / K
        scene_bytes = int(flat.n_shapes * 16 * 8 + flat.n_triangles * 12 * 8 + flat.n_materials * 12 * 8 + flat.n_patterns * 18 * 8 +
                          flat.n_lights * 6 * 8 + flat.n_shapes * 16 + flat.n_materials * 8 + flat.n_patterns * 16)
        e2e = {"value": st.as_dict()["rays"] / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_frame": e2e_ms,
               "h2d_bytes_per_step": scene_bytes * e2e_gpus + 256 * e2e_gpus, "d2h_bytes_per_step": n_px * 3 * elem + 48 * e2e_gpus,
               "n_gpus": e2e_gpus,
               "call": "rtgpu_render (scene pack + upload, kernel on N devices in %d-row bands, D2H of the full f64 Canvas into pinned host memory)" % e2e_band_rows,
               "kernel_ms_max_over_devices": st.kernel_ms, "family": last_family()}
        assert st.as_dict()["rays"] == rays, (st.as_dict(), stats)  # same kernel, same rays as the device-resident leg
        # ---- pixel identity across N (driver-readable): the N-GPU frames of both legs against a 1-GPU render ----
        frame_e2e = host_np.copy()
        opts1 = abi.RtgpuOpts(opts.precision, args.max_depth, 1, 16, family_flag)
        abi.check(lib, lib.rtgpu_render(C.byref(cscene), C.byref(ccam), C.byref(opts1), host_np.ctypes.data, None, None))
        sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
        frame_id = {"frame_sha256": sha(frame_device), "frame_sha256_e2e": sha(frame_e2e), "frame_sha256_n1": sha(host_np),
                    "frame_matches_n1": bool(sha(frame_device) == sha(host_np) and sha(frame_e2e) == sha(host_np)),
                    "frame_dtype": str(frame_e2e.dtype), "frame_shape": [camera.vertical_size, camera.horizontal_size, 3]}
        assert frame_id["frame_matches_n1"], frame_id
    host_barrier()

    if rank == 0:
        # ---- roofline of the render kernel: FP-pipe ----
        peak_tflops, _ = measure_fma_peak(args.precision, device=local_rank)
        flops = frame_flops(flat, stats)  # whole frame, all ranks
        traffic, traffic_source = profiled_traffic(args, family_used)
        # in-run model of the same traffic: every queued hit is one ray record written and read once, every node record
        # is written once and read once by the combine pass, every child colour is one 24-byte slot write, and the
        # frame is written once (counts from this run's device counters; record sizes from the library)
        # (a binned frame also moves 24 bytes of keys and permutation per queued hit; the model leaves them out)
        traffic_model = (2.0 * records["queued_rays"] * records["ray_record_bytes"] + 2.0 * records["node_records"] * records["node_record_bytes"] +
                         records["queued_rays"] * 3.0 * elem + n_px * 3.0 * elem)
        hbm_peak, hbm_peak_source = hbm_peak_gbs()
        achieved = flops / world / (device_ms / K * 1e-3) / 1e12  # per device: each renders 1/N of the frame in device_ms/K
        roofline = {
            "bound": "fp64_fma_pipe" if args.precision == "f64" else "fp32_fma_pipe", "achieved": achieved, "peak": peak_tflops,
            "unit": "TFLOP/s", "frac": achieved / peak_tflops, "traffic": traffic, "traffic_source": traffic_source,
            "kernel": "rt::wf_level_kernel x (max_depth + 1) = 87 % of the step (profiles/r2d_wavefront_launches.md); achieved = "
                      "frame flops / frame time" if family_used == "wavefront" else "rt::render_kernel (the whole step)",
            "peak_source": "measured live: 8 independent FMA chains per thread on every SM (rtgpu_measure_fma_peak); "
                           "MEASURED_PEAKS.json holds only HBM and bf16 peaks",
            "algorithmic_flops_per_frame": flops, "flops_per_ray": flops_per_ray(flat),
            "frac_note": "frac = the reference's brute-force flops (every shape tested exactly for every ray, SURVEY 8d) / time / peak: the "
                         "kernels skip most of those tests with a single-precision bounding-sphere pre-test, so frac is a speed relative to a "
                         "brute-force machine at peak, not pipe occupancy (profiled FP64 pipe: 26-31 % busy, issue slots 63-73 %, "
                         "profiles/r2d_wavefront_launches.md)",
            "traffic_model": traffic_model, "traffic_model_source": "in-run: device record counts (rtgpu_context_frame_records) x record sizes + the frame write",
            "hbm": {"achieved_gbs": (traffic or traffic_model) / (device_ms / K * 1e-3) / 1e9 / world, "peak_gbs": hbm_peak, "peak_source": hbm_peak_source,
                    "frac": (traffic or traffic_model) / (device_ms / K * 1e-3) / 1e9 / world / hbm_peak,
                    "note": "HBM is not the bound of this path: scene tables sit in shared memory / L2 (BVH scenes: L2 hit rate 84-91 %, "
                            "profiles/r2_notes.md); the traffic is queue and node records plus the frame write"},
        }
        if synthetic_shapes(args):
            roofline["note"] = ("algorithmic flops = the reference's brute force over every shape per ray (SURVEY 8d convention); the BVH skips almost all "
                                "of it, so `frac` says how much faster than a brute-force FP64 machine at peak the frame ran, not pipe utilisation")
        # ---- CPU baseline beside it (bounded: 3 frames) ----
        cpu, cold = None, None
        if world == 1:  # the CPU leg runs at N = 1 only (the other N reuse the reference arm the driver runs beside this one)
            sample = oracle_sample(args, flat, camera)
            if sample is None:
                times, ostats, cores, oracle_rgb = time_oracle(flat, camera, args.max_depth, steps=3, warmup=1)
                serial_mrays, serial_px = time_oracle_serial_sample(flat, camera, args.max_depth)
                cpu = {"value": ostats["rays"] / min(times) / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                       "sample": f"the same whole {args.width}x{args.height} frame, best of 3; restated reference (C + OpenMP), not rustc output",
                       "ms_per_frame": min(times) * 1e3,
                       "serial": {"value": serial_mrays, "unit": "Mrays/s", "cores": 1, "sample": f"every 16th pixel of the frame ({serial_px} pixels), one thread"},
                       "native": time_oracle_native(args)}
                device_rgb = frame_device
            else:
                times, ostats, cores, oracle_rgb = time_oracle(flat, camera, args.max_depth, steps=1, warmup=0, pixels=sample)
                cpu = {"value": ostats["rays"] / min(times) / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                       "sample": f"{sample.size} pseudo-random pixels of the frame, once (brute force over {flat.n_shapes} shapes per ray); restated reference "
                                 "(C + OpenMP), not rustc output",
                       "seconds": min(times), "whole_frame_estimate_s": rays / (ostats["rays"] / min(times))}
                device_rgb = frame_device[sample.astype(np.int64)]
                frame_id["oracle_sample_pixels"] = int(sample.size)
            # parity against the oracle's pixels of the same run: 8-bit output, and the work counters in parity mode
            dev8, ora8 = quantise_rgb8(device_rgb.astype(np.float64)), quantise_rgb8(oracle_rgb)
            diff = np.abs(dev8.astype(np.int16) - ora8.astype(np.int16)).max(axis=1)
            frame_id["rgb8_pixels_differing_from_oracle"] = int((diff > 0).sum())
            frame_id["rgb8_pixels_beyond_1lsb"] = int((diff > 1).sum())
            frame_id["oracle_rgb8_sha256"] = sha(ora8)
            frame_id["rgb8_sha256"] = sha(dev8)
            if args.precision == "f64":  # parity mode: the device ray count is integer-equal to the oracle's
                assert sample is not None or ostats["rays"] == rays, (ostats, stats)
                assert frame_id["rgb8_pixels_differing_from_oracle"] == 0, frame_id
            # ---- cold one-shot: a fresh process's first rtgpu_render (the region ray-tracer-cli/src/main.rs:17-22 times) ----
            if not args.no_cold and not synthetic_shapes(args):
                try:
                    out = subprocess.run([sys.executable, os.path.join(ROOT, "benchmarks", "cold_one_shot.py"), "--child", "--scene", args.scene, "--width",
                                          str(args.width), "--height", str(args.height), "--family", "auto"], capture_output=True, text=True, timeout=300)
                    cold = json.loads(out.stdout.strip().splitlines()[-1])
                    cold["note"] = ("fresh process, scene already loaded, pageable numpy Canvas: CUDA driver + context creation and module load dominate "
                                    "(0.4 - 4 s on these boxes, profiles/r2_notes.md); steady state is `e2e`")
                except Exception as exc:  # the cold probe must never cost the bench line
                    cold = {"error": repr(exc)}
        line = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": device_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.precision, "data": data_label(args),
            "config": workload_config(args, flat),  # identical to the reference arm's: the driver compares the two
            "parallelism": f"row bands of 4 rows, interleaved over {world} GPU(s); no collective in the data path",
            "family": family_used, "family_requested": args.family,
            "family_calibration_frames": CALIBRATION_FRAMES if family is None else 0,
            "frame": frame_id, "cold_one_shot": cold,
            "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clock_summary,
            "rays_per_frame": rays, "ms_per_frame": device_ms / K, "wall_ms_per_step_incl_flush": (t_wall1 - t_wall0) * 1e3 / K,
            "counters": stats,
        }
        RESULT_LINE.append(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default="cover")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--max-depth", type=int, default=6)
    ap.add_argument("--no-cold", action="store_true", help="skip the cold one-shot probe (a fresh subprocess, N = 1 only)")
    ap.add_argument("--family", default="auto", choices=["auto", "persistent", "wavefront"],
                    help="kernel family; auto = the library measures both on the first frames and keeps the faster")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # stdout carries exactly one JSON line: anything libraries print to fd 1 on the way (NCCL's version banner,
    # torchrun notices) goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_b200(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if RESULT_LINE:
        print(RESULT_LINE[0], flush=True)


if __name__ == "__main__":
    main()
