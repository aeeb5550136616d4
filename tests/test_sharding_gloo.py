"""world_size-2 gloo test of the N>1 path's only collective: the depth-map gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from doubletake_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, num_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.shard_frames(num_frames, rank, world)
        # frame i's "depth map" is filled with i + 0.5
        local = torch.stack([torch.full((1, 3, 4), i + 0.5) for i in mine]) if mine else torch.zeros(0, 1, 3, 4)
        full = sharding.gather_depth_maps(local, num_frames)
        expect = torch.stack([torch.full((1, 3, 4), i + 0.5) for i in range(num_frames)])
        q.put((rank, bool(torch.equal(full, expect))))
    finally:
        dist.destroy_process_group()


def _run(num_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in results), results


def test_gather_even_split():
    _run(6)


def test_gather_ragged_split():
    _run(5)
