"""Host side of the halo kernel's tail split (csrc/conv_tc.cu: halo_tail_plan, development switch bit 4; CPU only).
The plan the library derives (observable through dtb200_conv_workspace_bytes) must equal this restatement, and the item ->
(tile, unit range) map the kernel's roles walk must cover every (tile, unit) exactly once."""
import ctypes as C

import pytest

from doubletake_b200 import _lib as L

SMS, BM, BN = 148, 128, 64


def plan(out_h, out_w, batch, chunks):
    tiles = batch * -(-out_w // 8) * -(-out_h // 16)
    if tiles < SMS:
        return None
    rem = tiles % SMS
    units, max_parts = 3 * chunks, (SMS // rem if rem else 0)
    if rem == 0 or max_parts < 2 or units < 2:
        return dict(tiles=tiles, parts=0)
    want = min(units, max_parts)
    upp = -(-units // want)
    parts = -(-units // upp)
    if parts < 2:
        return dict(tiles=tiles, parts=0)
    return dict(tiles=tiles, rem=rem, units=units, upp=upp, parts=parts, full=tiles - rem, total=tiles - rem + rem * parts)


def conv_params(h, w, batch, src_c, out_c=64, k=3, stride=1):
    p = L.ConvParams()
    p.math, p.batch = L.MATH_TC3X, batch
    p.in_h = p.out_h = h
    p.in_w = p.out_w = w
    p.out_c, p.ksize, p.stride, p.num_src = out_c, k, stride, len(src_c)
    for i, c in enumerate(src_c):
        p.src_c[i] = c
    return p


@pytest.mark.parametrize("h,w,batch,src_c", [(240, 320, 1, (64,)), (240, 320, 1, (64, 64, 64)), (120, 160, 1, (64,)),
                                             (120, 160, 1, (64, 48)), (240, 320, 1, (24,)), (192, 256, 8, (64,)),
                                             (130, 165, 1, (40, 24)), (148 * 16, 8, 1, (64,))])
def test_workspace_bytes_follow_the_tail_plan(h, w, batch, src_c):
    lib = L.lib()
    chunks = sum(-(-c // 32) for c in src_c)
    pl = plan(h, w, batch, chunks)
    p = conv_params(h, w, batch, src_c)
    try:
        assert lib.dtb200_debug_set(0) == 0
        assert lib.dtb200_conv_workspace_bytes(C.byref(p)) == 0  # default: the halo kernel needs no scratch
        assert lib.dtb200_debug_set(16) == 0
        got = lib.dtb200_conv_workspace_bytes(C.byref(p))
        want = pl["rem"] * pl["parts"] * BM * BN * 4 if pl and pl["parts"] else 0
        assert got == want, (got, want, pl)
        assert lib.dtb200_debug_set(16 | 128) == 0  # halo kernel off: no tail plan either
        assert lib.dtb200_conv_workspace_bytes(C.byref(p)) == 0
    finally:
        lib.dtb200_debug_set(0)


@pytest.mark.parametrize("h,w,batch,chunks", [(240, 320, 1, 2), (240, 320, 1, 6), (120, 160, 1, 2), (120, 160, 1, 4),
                                              (130, 165, 1, 3), (240, 320, 1, 1)])
def test_items_cover_every_tile_unit_once(h, w, batch, chunks):
    pl = plan(h, w, batch, chunks)
    assert pl and pl["parts"] >= 2
    assert pl["rem"] * pl["parts"] <= SMS  # the tail is ONE round of short items
    seen = {}
    for item in range(pl["total"]):
        if item < pl["full"]:
            tile, u0, u1 = item, 0, pl["units"]
        else:
            j = item - pl["full"]
            tile, part = pl["full"] + j // pl["parts"], j % pl["parts"]
            u0, u1 = part * pl["upp"], min(pl["units"], part * pl["upp"] + pl["upp"])
        assert u0 < u1
        # the chunk walk of the A loader / splitters / MMA issuer: every chunk touched by [u0, u1), groups g_lo..g_hi
        ch_begin, ch_last = u0 // 3, (u1 - 1) // 3
        units = []
        for ch in range(ch_begin, ch_last + 1):
            g_lo = u0 - ch * 3 if ch == ch_begin else 0
            g_hi = (u1 - 1) - ch * 3 if ch == ch_last else 2
            units += [ch * 3 + g for g in range(g_lo, g_hi + 1)]
        assert units == list(range(u0, u1))  # == the B loader's walk
        for u in units:
            assert (tile, u) not in seen
            seen[(tile, u)] = item
    assert len(seen) == pl["tiles"] * pl["units"]
