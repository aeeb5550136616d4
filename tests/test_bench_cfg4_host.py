"""Host-side pieces of bench.py --workload cfg4 (SURVEY 8e / BASELINE configs[3]): the synthetic stand-in for the reference's ScanNetv2
test tuple file has the real file's structure, and sharding it covers every keyframe exactly once."""
import os

import pytest

from doubletake_b200 import sharding

REF_TUPLES = "/root/reference/data_splits/ScanNetv2/standard_split/test_eight_view_deepvmvs.txt"


def test_synthetic_tuples_have_the_split_structure():
    tuples = sharding.synthetic_scannet_test_tuples()
    assert len(tuples) == 25590 and len({s for s, _ in tuples}) == 100
    assert all(len(f) == 8 for _, f in tuples)


@pytest.mark.skipif(not os.path.exists(REF_TUPLES), reason="reference tree not present (GPU box)")
def test_synthetic_tuples_match_the_reference_file_scan_by_scan():
    ref = sharding.read_frame_tuples(REF_TUPLES)
    syn = sharding.synthetic_scannet_test_tuples()
    assert [s for s, _ in ref] == [s for s, _ in syn]


@pytest.mark.parametrize("by", ["frame", "scan"])
@pytest.mark.parametrize("world", [2, 8])
def test_shards_partition_the_split(by, world):
    tuples = sharding.synthetic_scannet_test_tuples()
    shards = [sharding.shard_tuples(tuples, r, world, by=by) for r in range(world)]
    assert sorted(i for sh in shards for i in sh) == list(range(len(tuples)))
    sizes = [len(sh) for sh in shards]
    assert max(sizes) - min(sizes) <= (1 if by == "frame" else 400)   # whole scans: longest-first greedy stays within one scan
