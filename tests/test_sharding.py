"""Multi-GPU host logic on CPU: partitions and the input broadcast with world_size 2 over gloo."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from mathmap_b200 import sharding


def test_frame_and_band_partitions_cover_everything():
    for world in (1, 2, 3, 8):
        frames = sorted(sum((sharding.frames_for_rank(240, r, world) for r in range(world)), []))
        assert frames == list(range(240))
        bands = [sharding.band_for_rank(0, 2160, r, world) for r in range(world)]
        assert bands[0][0] == 0 and bands[-1][1] == 2160
        assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
        for h in (1, 7, 8, 9, 2160, 16384):
            rows = sorted(sum((sharding.interleaved_rows_for_rank(h, r, world) for r in range(world)), []))
            assert rows == list(range(h))


def test_assemble_interleaved_round_trip():
    img = np.arange(37 * 5 * 4, dtype=np.uint8).reshape(37, 5, 4)
    for world in (1, 2, 4):
        parts = [img[sharding.interleaved_rows_for_rank(37, r, world)] for r in range(world)]
        assert np.array_equal(sharding.assemble_interleaved(parts, 37), img)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = torch.zeros((6, 5, 4), dtype=torch.uint8)
    if rank == 0:
        t = torch.arange(120, dtype=torch.uint8).reshape(6, 5, 4).clone()
    sharding.broadcast_drawable(t, src=0)
    mx = sharding.max_over_ranks(1.0 + rank)
    frames = sharding.frames_for_rank(10, rank, world)
    # every rank holds only its own band of a drawable (an odd number of rows: unequal bands), then all have all of it
    whole = torch.arange(7 * 3 * 4, dtype=torch.uint8).reshape(7, 3, 4)
    part = torch.zeros_like(whole)
    r0, r1 = sharding.band_for_rank(0, 7, rank, world)
    part[r0:r1] = whole[r0:r1]
    sharding.replicate_drawable_bands(part)
    assert torch.equal(part, whole)
    q.put((rank, int(t.sum()), mx, frames))
    dist.destroy_process_group()


def test_broadcast_and_timing_reduction_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = int(torch.arange(120, dtype=torch.uint8).sum())
    assert res[0][1] == expect and res[1][1] == expect
    assert res[0][2] == 2.0 and res[1][2] == 2.0
    assert res[0][3] == [0, 2, 4, 6, 8] and res[1][3] == [1, 3, 5, 7, 9]
