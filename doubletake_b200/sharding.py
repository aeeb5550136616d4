"""Frame-level data parallelism for inference.

Reference keyframes are independent units (each frame's cost volume + decoder touches only its own K source views;
the reference itself runs inference on one GPU: README.md:125,307,326,344).  One process per GPU: global frame ``i``
goes to rank ``i % world_size``; every rank holds a full weight replica; the ONLY exchange is one gather of the
predicted depth maps per step (``all_gather`` over NCCL on NVLink; ``gloo`` on CPU for the host-logic tests).
Incremental mode is sequential inside a scan, so there the unit is the scan (``shard_scans``).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_frames(num_frames: int, rank: int, world_size: int):
    """Indices of the frames this rank owns: round-robin, as cfg 4 of BASELINE.json asks."""
    return list(range(rank, num_frames, world_size))


def shard_scans(scan_lengths, rank: int, world_size: int):
    """Greedy longest-first assignment of whole scans to ranks (incremental mode: a frame's hint depends on the
    fused depths of the earlier frames of its scan, reference test_incremental.py:186-269)."""
    order = sorted(range(len(scan_lengths)), key=lambda i: -scan_lengths[i])
    load = [0] * world_size
    mine = []
    for i in order:
        r = min(range(world_size), key=lambda j: (load[j], j))
        load[r] += scan_lengths[i]
        if r == rank:
            mine.append(i)
    return sorted(mine)


def read_frame_tuples(path, limit_to_scan_id=None, skip_to_frame=None, skip_frames=None):
    """Keyframe tuples of a reference tuple file (``scan_id frame_id_0 frame_id_1 ... frame_id_N-1`` per line, frame 0 the
    reference view; datasets/generic_mvs_dataset.py:29-37,149-180, incl. its scan filter / skip options).
    Returns a list of (scan_id, [frame ids]) in file order."""
    with open(path) as f:
        lines = f.read().splitlines()
    if limit_to_scan_id is not None:
        lines = [ln for ln in lines if ln.split(" ")[0] == limit_to_scan_id]
    if skip_to_frame is not None:
        lines = lines[skip_to_frame:]
    if skip_frames is not None:
        lines = lines[::skip_frames]
    out = []
    for ln in lines:
        scan_id, *frame_ids = ln.split(" ")
        out.append((scan_id, frame_ids))
    return out


def synthetic_scannet_test_tuples():
    """A stand-in for the reference's ScanNetv2 test tuple file with the SAME scans and per-scan keyframe counts (25 590 tuples
    over 100 scans, data/scannetv2_test_tuple_counts.json), for boxes where the reference tree does not exist: frame ids are
    synthetic (keyframe j of a scan reads frames 10 j, 10 j - 10, ...), the sharding sees exactly the real structure."""
    import json
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "scannetv2_test_tuple_counts.json")) as f:
        meta = json.load(f)
    out = []
    for scan_id, count in meta["scans"]:
        for j in range(count):
            out.append((scan_id, [f"{max(10 * (j + 1) - 10 * v, 0):06d}" for v in range(meta["views"])]))
    return out


def shard_tuples(tuples, rank: int, world_size: int, by: str = "frame"):
    """Indices (into ``tuples``) of the keyframes this rank processes, in processing order.

    ``by="frame"``: round-robin over keyframes (independent frames: test_no_hint / offline passes; BASELINE cfg 4).
    ``by="scan"``: whole scans, longest first (incremental mode: a frame's hint comes from the TSDF fused from the
    earlier frames of ITS scan, test_incremental.py:186-269); frames keep their order inside a scan."""
    if by == "frame":
        return shard_frames(len(tuples), rank, world_size)
    if by != "scan":
        raise ValueError("by must be 'frame' or 'scan'")
    scans, members = [], {}
    for i, (scan_id, _) in enumerate(tuples):
        if scan_id not in members:
            members[scan_id] = []
            scans.append(scan_id)
        members[scan_id].append(i)
    mine = shard_scans([len(members[s]) for s in scans], rank, world_size)
    return [i for k in mine for i in members[scans[k]]]


def gather_depth_maps(local_depth: torch.Tensor, num_frames: int, group=None):
    """Gather per-rank depth maps (n_local,1,H,W) into frame order (num_frames,1,H,W) on every rank.

    Ranks may own ceil or floor(num_frames / world) frames; shorter shards are zero-padded for the collective and
    trimmed afterwards.  This is the path's single collective."""
    if not dist.is_available() or not dist.is_initialized():
        return local_depth
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = (num_frames + world - 1) // world
    pad = per - local_depth.shape[0]
    send = local_depth
    if pad > 0:
        send = torch.cat([local_depth, local_depth.new_zeros((pad,) + tuple(local_depth.shape[1:]))], 0)
    send = send.contiguous()
    recv = torch.empty((world,) + tuple(send.shape), dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(recv.view(-1, *send.shape[1:]), send, group=group) if send.is_cuda else \
        dist.all_gather(list(recv.unbind(0)), send, group=group)
    # recv[r, j] is global frame j * world + r (round-robin).  Even shards: a transpose is the frame order (a free view
    # when every rank owns one frame); ragged shards: strided slice assignments -- no index tensors, so no host->device
    # copies or syncs ride on the collective.
    if per * world == num_frames:
        return recv.transpose(0, 1).reshape((num_frames,) + tuple(local_depth.shape[1:]))
    out = local_depth.new_empty((num_frames,) + tuple(local_depth.shape[1:]))
    for r in range(world):
        n = len(range(r, num_frames, world))
        if n:
            out[r::world] = recv[r, :n]
    return out
