"""Host logic of BASELINE cfg 4: reference tuple files -> per-rank keyframe shards (CPU only)."""
import os

import pytest

from doubletake_b200 import sharding

REF_TUPLES = "/root/reference/data_splits/ScanNetv2/standard_split/test_eight_view_deepvmvs.txt"


def write_tuples(path, scans):
    with open(path, "w") as f:
        for scan, n in scans:
            for i in range(n):
                f.write(" ".join([scan, f"{i * 10:06d}"] + [f"{max(i * 10 - j, 0):06d}" for j in range(1, 8)]) + "\n")


def test_round_robin_and_scan_sharding_cover_every_tuple_once(tmp_path):
    path = str(tmp_path / "t.txt")
    write_tuples(path, [("scene0001_00", 7), ("scene0002_00", 3), ("scene0003_01", 12), ("scene0004_00", 1)])
    tuples = sharding.read_frame_tuples(path)
    assert len(tuples) == 23 and tuples[0] == ("scene0001_00", ["000000"] + ["000000"] * 7)
    assert all(len(ids) == 8 for _, ids in tuples)
    for by in ("frame", "scan"):
        for world in (1, 2, 3, 8):
            shards = [sharding.shard_tuples(tuples, r, world, by=by) for r in range(world)]
            flat = sorted(i for s in shards for i in s)
            assert flat == list(range(len(tuples))), (by, world)
            if by == "frame":
                assert max(map(len, shards)) - min(map(len, shards)) <= 1
                assert shards[0] == list(range(0, len(tuples), world))
            else:
                for s in shards:  # whole scans, frames in order
                    scans_seen = [tuples[i][0] for i in s]
                    for scan in set(scans_seen):
                        idx = [i for i in s if tuples[i][0] == scan]
                        assert idx == [i for i, t in enumerate(tuples) if t[0] == scan]
    with pytest.raises(ValueError):
        sharding.shard_tuples(tuples, 0, 2, by="pixel")
    only = sharding.read_frame_tuples(path, limit_to_scan_id="scene0003_01", skip_to_frame=2, skip_frames=2)
    assert [t[1][0] for t in only] == [f"{i * 10:06d}" for i in range(2, 12, 2)]


@pytest.mark.skipif(not os.path.exists(REF_TUPLES), reason="reference tree not present (GPU box)")
def test_scannet_test_split_tuple_file():
    """The split BASELINE cfg 4 names: 25 590 eight-view tuples over 100 scans."""
    tuples = sharding.read_frame_tuples(REF_TUPLES)
    assert len(tuples) == 25590 and len({s for s, _ in tuples}) == 100 and all(len(ids) == 8 for _, ids in tuples)
    shards = [sharding.shard_tuples(tuples, r, 8, by="frame") for r in range(8)]
    assert sorted(map(len, shards)) == [3198] * 2 + [3199] * 6
    by_scan = [sharding.shard_tuples(tuples, r, 8, by="scan") for r in range(8)]
    assert sum(map(len, by_scan)) == 25590
    assert max(map(len, by_scan)) - min(map(len, by_scan)) < 0.05 * 25590 / 8  # longest-first keeps the load within 5 %
