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


def _worker_batched(rank, world, port, q):
    """bench.py --workload cfg4: every rank parks G keyframes of its round-robin shard of the tuple list and gathers them in one
    collective; gathered block g must hold global keyframes g*G*world .. (g+1)*G*world - 1 in tuple order."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tuples = sharding.synthetic_scannet_test_tuples()[:40]
        shard = sharding.shard_tuples(tuples, rank, world, by="frame")
        G = 4
        ok = True
        parked = torch.empty(G, 1, 3, 4)
        for i, t in enumerate(shard[: 3 * G]):
            parked[i % G] = float(t) + 0.25          # the "depth map" of keyframe t
            if (i + 1) % G == 0:
                full = sharding.gather_depth_maps(parked, G * world)
                block = (i + 1) // G - 1
                expect = torch.stack([torch.full((1, 3, 4), k + 0.25) for k in range(block * G * world, (block + 1) * G * world)])
                ok = ok and bool(torch.equal(full, expect))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_batched_gather_of_sharded_keyframes_is_in_tuple_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_batched, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in results), results
