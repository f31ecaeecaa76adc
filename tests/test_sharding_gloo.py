"""N > 1 host logic on CPU: two processes over torch.distributed (gloo), each taking the row bands that
`rtgpu_rows_list` assigns to its rank — exactly what each rank of `bench.py` / each device of
`rtgpu_render` does.  There is no GPU here, so each rank renders its rows with the CPU oracle (the
checker) and the assembled frame must equal the unsharded one; the timing reduction (max over ranks)
and the counter reduction (sum over ranks) follow bench.py."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, band_rows, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle.oracle import Oracle
    from ray_tracer_challenge_rs_b200 import abi, build
    from ray_tracer_challenge_rs_b200.fixtures import load_scene_fixture

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    build.build()
    lib = abi.load_library()
    flat, camera = load_scene_fixture("reflect_refract")
    cam = camera.resized(96, 64)
    rows = abi.RtgpuRows(band_rows, rank, world)
    n = lib.rtgpu_rows_count(C.byref(rows), cam.vertical_size)
    mine = (C.c_uint32 * max(n, 1))()
    assert lib.rtgpu_rows_list(C.byref(rows), cam.vertical_size, mine, n) == n
    mine = np.asarray(mine[:n], dtype=np.int64)
    # the rows of this rank, rendered by the checker into a zero frame
    rgb, _, stats = Oracle(flat).render(cam, rows=(band_rows, rank, world), want_rgb8=False)
    frame = torch.from_numpy(rgb.reshape(cam.vertical_size, cam.horizontal_size, 3).copy())
    touched = torch.zeros(cam.vertical_size, dtype=torch.int64)
    touched[mine] = 1
    assert float(frame[touched == 0].abs().sum()) == 0.0  # nothing outside this rank's rows
    dist.all_reduce(frame, op=dist.ReduceOp.SUM)          # disjoint rows: the sum assembles the frame
    dist.all_reduce(touched, op=dist.ReduceOp.SUM)
    counters = torch.tensor([stats[k] for k in ("rays_primary", "rays_shadow", "rays_reflect", "rays_refract", "hit_nodes")], dtype=torch.int64)
    dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)  # bench.py: time = max over ranks
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        np.savez(out_path, frame=frame.numpy(), touched=touched.numpy(), counters=counters.numpy(), tmax=t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("band_rows", [4, 16])
def test_two_ranks_assemble_the_frame(tmp_path, band_rows, oracle_lib):
    import torch.multiprocessing as mp

    from oracle.oracle import Oracle
    from ray_tracer_challenge_rs_b200.fixtures import load_scene_fixture

    out = str(tmp_path / "assembled.npz")
    mp.spawn(_worker, args=(2, _free_port(), band_rows, out), nprocs=2, join=True)
    got = np.load(out)
    flat, camera = load_scene_fixture("reflect_refract")
    cam = camera.resized(96, 64)
    whole, _, stats = Oracle(flat).render(cam, want_rgb8=False)
    assert (got["touched"] == 1).all()
    assert np.array_equal(got["frame"].reshape(-1, 3), whole)
    assert got["counters"].tolist() == [stats[k] for k in ("rays_primary", "rays_shadow", "rays_reflect", "rays_refract", "hit_nodes")]
    assert float(got["tmax"][0]) == 2.0
