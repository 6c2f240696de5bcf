"""world_size-2 gloo test of the multi-GPU film path (CPU): strip ownership + one sum-reduce must
reproduce the single-process film bit for bit.  The render itself runs on the test-only host
emulation here; on the GPU box the same driver runs qz_render_device per rank (bench.py)."""
import ctypes
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

W, H, SPP, STRIP = 24, 20, 3, 4


def _render_shard(emu, scene, film, strip_rows, world, rank):
    from quetzalcoatlus_b200.harness import QzRegion, QzStats

    lib = emu.lib
    handle, cam = ctypes.c_void_p(scene.c_scene_handle()), scene.c_camera()
    planes = [np.zeros((H, W, 3), np.float32) for _ in range(3)]
    region = QzRegion(strip_rows, world, rank)
    st = QzStats()
    rc = lib.qz_render(handle, ctypes.byref(cam), SPP, 16, ctypes.byref(region), None, *(p.ctypes.data_as(ctypes.c_void_p) for p in planes),
                       ctypes.byref(st))
    assert rc == 0
    for k in range(3):
        film[k].copy_(torch.from_numpy(planes[k]))


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quetzalcoatlus_b200.distributed import owned_rows, render_sharded
    from quetzalcoatlus_b200.harness import Harness

    emu = Harness(ROOT / "tests" / "emu" / "_build" / "libqz_emu_harness.so", "qzh_")
    with emu.build_scene("cornell_box", W, H) as scene:
        film = render_sharded(lambda f, s, w, r: _render_shard(emu, scene, f, s, w, r), H, W, STRIP)
        mine = owned_rows(H, STRIP, world, rank)
        assert sorted(set(mine)) == mine and all((r // STRIP) % world == rank for r in mine)
    if rank == 0:
        np.save(out_path, film.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_film_equals_single_process(emu, tmp_path):
    from quetzalcoatlus_b200.distributed import owned_rows

    # ownership partitions the rows
    for world in (1, 2, 3, 8):
        rows = sorted(r for k in range(world) for r in owned_rows(H, STRIP, world, k))
        assert rows == list(range(H))
    out = tmp_path / "film.npy"
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    got = np.load(out)
    with emu.build_scene("cornell_box", W, H) as scene:
        want = scene.render(SPP, 16)
    for k, plane in enumerate((want.color, want.normal, want.albedo)):
        assert np.array_equal(got[k], plane)  # == treats +0 and -0 as equal; everything else bit-identical
