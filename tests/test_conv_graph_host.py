"""Host logic of the compiled-network runtime (csrc/conv_graph.cu): the dependency DAG derived from the descriptors'
buffers.  No GPU: dtb200_conv_graph_analyze makes no CUDA call, and the plans below are built over fabricated addresses
(the real CVEncoder / DepthDecoderPP / SkipDecoderRegression emit code, cfg-2 and cfg-3 shapes)."""
import ctypes as C
import itertools

import pytest
import torch

import doubletake_b200 as dt
from doubletake_b200 import _lib as L
from doubletake_b200 import networks as N


class FakeTensor:
    """Stands in for a device buffer: an address range and a shape."""

    _next = 1 << 30

    def __init__(self, shape):
        self.shape = tuple(shape)
        n = 1
        for s in shape:
            n *= s
        self.nbytes = n * 4
        self._ptr = FakeTensor._next
        FakeTensor._next += (self.nbytes + 255) // 256 * 256
        self.is_cuda = True

    def data_ptr(self):
        return self._ptr


class FakePlan(N.ConvPlan):
    def new(self, b, h, w, c):
        t = FakeTensor((b, h, w, c))
        self.keep.append(t)
        return N.Feature(t)

    def _pack(self, weight, src_channels):
        return FakeTensor((weight.numel() * 2,))


@pytest.fixture()
def fake_ptrs(monkeypatch):
    monkeypatch.setattr(L, "ptr", lambda t: None if t is None else t.data_ptr())
    monkeypatch.setattr(L, "f32", lambda t, device=None: t)


def build_cfg2_plan(math="tc3x", lanes=8):
    plan = FakePlan(torch.device("cpu"), math)
    plan.max_lanes = lanes
    enc = dt.CVEncoder(64, [48, 64, 160, 256], [64, 128, 256, 384])
    dec = dt.DepthDecoderPP([24, 64, 128, 256, 384])
    fcv = plan.input("cv", 1, 64, 120, 160)
    pri = [plan.input(f"p{i}", 1, c, 240 >> i, 320 >> i) for i, c in enumerate([24, 48, 64, 160, 256])]
    outs = enc.emit(plan, fcv, pri[1:])
    plan.outputs = dec.emit(plan, pri[:1] + outs)
    return plan.finalize()


def ranges(op):
    """Independent Python restatement of the byte ranges an op touches."""
    reads, writes = [], []
    for s in range(op.num_src):
        half = op.src_resample[s] != L.RESAMPLE_NONE
        h, w = (op.in_h // 2, op.in_w // 2) if half else (op.in_h, op.in_w)
        reads.append((op.src[s], op.src[s] + op.batch * h * w * op.src_c[s] * 4))
    out = op.batch * op.out_h * op.out_w * op.out_c * 4
    if op.residual and op.ksize:
        reads.append((op.residual, op.residual + out))
    writes.append((op.dst, op.dst + out))
    ws = int(L.lib().dtb200_conv_workspace_bytes(C.byref(op)))
    if ws and op.workspace:
        writes.append((op.workspace, op.workspace + ws))
    return reads, writes


def hit(a, b):
    return any(x[0] < y[1] and y[0] < x[1] for x in a for y in b)


def check_schedule(plan, info):
    n = len(plan.ops)
    acc = [ranges(op) for op in plan.ops]
    # reachability through the reported (reduced) dependency lists
    reach = [set() for _ in range(n)]
    for i, d in enumerate(info):
        assert all(0 <= j < i for j in d["deps"])
        for j in d["deps"]:
            reach[i] |= {j} | reach[j]
        assert d["level"] == (max(info[j]["level"] for j in d["deps"]) + 1 if d["deps"] else 0)
        assert 0 <= d["lane"] < plan.max_lanes
    for j, i in itertools.combinations(range(n), 2):
        (rj, wj), (ri, wi) = acc[j], acc[i]
        if hit(wj, ri) or hit(wj, wi) or hit(rj, wi):
            assert j in reach[i], f"op {i} conflicts with op {j} but does not depend on it"
    # ops on one lane execute in index order: that must never contradict a dependency (it cannot: edges point backwards)
    return reach


def test_small_diamond():
    """A -> (B, C) -> D, then D -> A: the WAR edges on A are implied by the chain through D's producer."""
    base = 1 << 20
    buf = {k: base + i * (1 << 16) for i, k in enumerate("ABCDW")}

    def op(src, dst, srcs2=None):
        p = L.ConvParams()
        p.math, p.batch, p.in_h, p.in_w, p.out_h, p.out_w, p.out_c = L.MATH_EXACT, 1, 8, 8, 8, 8, 16
        p.ksize, p.stride = 3, 1
        ss = [src] + ([srcs2] if srcs2 else [])
        p.num_src = len(ss)
        for i, s in enumerate(ss):
            p.src[i], p.src_c[i] = buf[s], 16
        p.weight, p.dst = buf["W"], buf[dst]
        return p

    ops = [op("A", "B"), op("A", "C"), op("B", "D", "C"), op("D", "A")]
    arr = (L.ConvParams * 4)(*ops)
    lane, level, off, deps = (C.c_int32 * 4)(), (C.c_int32 * 4)(), (C.c_int32 * 5)(), (C.c_int32 * 16)()
    L.check(L.lib().dtb200_conv_graph_analyze(arr, 4, 4, lane, level, off, deps, 16))
    dl = [list(deps[off[i]:off[i + 1]]) for i in range(4)]
    assert dl == [[], [], [0, 1], [2]]
    assert list(level) == [0, 0, 1, 2]
    assert lane[0] != lane[1]          # the two independent ops may overlap
    assert lane[2] in (lane[0], lane[1])  # the join continues one of its producers' lanes


def test_capacity_error_is_reported():
    p = L.ConvParams()
    arr = (L.ConvParams * 1)(p)
    assert L.lib().dtb200_conv_graph_analyze(arr, 1, 0, None, None, None, None, 0) == -1
    assert b"bad arguments" in L.lib().dtb200_last_error()


@pytest.mark.parametrize("math", ["tc3x", "exact"])
def test_cfg2_network_dag(fake_ptrs, math):
    plan = build_cfg2_plan(math)
    info = plan.analyze()
    check_schedule(plan, info)
    depth = max(d["level"] for d in info) + 1
    lanes = len({d["lane"] for d in info})
    # the UNet++ decoder is wide: the longest chain is far shorter than the op count, and several lanes are in use
    assert depth < 0.6 * len(plan.ops), (depth, len(plan.ops))
    assert 3 <= lanes <= plan.max_lanes
    # a BasicBlock's 1x1 skip projection and its conv1 read the same sources and must be independent of each other
    first_block = [i for i, op in enumerate(plan.ops) if op.num_src == 2 and op.src_c[1] == 48]
    assert len(first_block) == 2
    a, b = first_block
    assert a not in info[b]["deps"] and info[a]["level"] == info[b]["level"]


def test_single_lane_is_a_chain(fake_ptrs):
    plan = build_cfg2_plan("tc3x", lanes=1)
    info = plan.analyze()
    assert {d["lane"] for d in info} == {0}
    check_schedule(plan, info)


def test_split_k_slots_serialise_only_their_users(fake_ptrs):
    plan = build_cfg2_plan("tc3x")
    users = [i for i, op in enumerate(plan.ops) if int(L.lib().dtb200_conv_workspace_bytes(C.byref(op)))]
    assert users, "cfg 2 has split-K layers on the small maps"
    slots = {plan.ops[i].workspace for i in users}
    assert 1 < len(slots) <= plan.workspace_slots
    reach = check_schedule(plan, plan.analyze())
    by_slot = {}
    for i in users:
        by_slot.setdefault(plan.ops[i].workspace, []).append(i)
    for ops in by_slot.values():
        for j, i in zip(ops, ops[1:]):
            assert j in reach[i]


def test_skip_decoder_dag(fake_ptrs):
    plan = FakePlan(torch.device("cpu"), "tc3x")
    dec = dt.SkipDecoderRegression([64, 64, 128, 256, 384])
    feats = [plan.input(f"f{i}", 2, c, 192 >> i, 256 >> i) for i, c in enumerate([64, 64, 128, 256, 384])]
    plan.outputs = dec.emit(plan, feats)
    plan.finalize()
    check_schedule(plan, plan.analyze())
