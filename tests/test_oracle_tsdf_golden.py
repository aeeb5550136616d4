"""Pin the TSDF oracle (oracle/oracle_tsdf.py) and the host side of doubletake_b200.tsdf against fixtures produced by
executing the real reference tools/tsdf.py (oracle/make_golden_tsdf.py).  CPU only.  fp16 volumes: BIT-EXACT."""
import numpy as np
import pytest
import torch

import helpers as hp
from doubletake_b200 import tsdf as bt
from oracle import oracle_tsdf as ot

CASES = ["tsdf_room", "tsdf_room_mask_ext", "tsdf_near"]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint16)


def case_bounds(fx):
    b = fx["bounds"]
    return dict(xmin=b[0], xmax=b[1], ymin=b[2], ymax=b[3], zmin=b[4], zmax=b[5])


@pytest.mark.parametrize("name", CASES)
def test_oracle_integrate_is_bit_exact(name):
    fx = hp.load(name)
    seed, nf, ih, iw, batch, with_mask, ext = [int(v) for v in fx["meta"]]
    vol = ot.volume_from_bounds(case_bounds(fx), float(fx["voxel_size"]))
    assert np.array_equal(bits(vol["voxel_coords_3hwd"]), bits(fx["voxel_coords"]))
    assert np.array_equal(bits(vol["origin"]), bits(fx["origin"]))
    for s in range(0, nf, batch):
        ot.integrate_depth(vol, fx["depth"][s:s + batch], fx["cam_T_world"][s:s + batch], fx["K"][s:s + batch],
                           min_depth=float(fx["min_depth"]), max_depth=float(fx["max_depth"]),
                           depth_mask_b1hw=fx["mask"][s:s + batch] if with_mask else None, extended_neg_truncation=bool(ext))
        if s == 0:
            assert np.array_equal(bits(vol["tsdf_values"]), bits(fx["values_first"]))
            assert np.array_equal(bits(vol["tsdf_weights"]), bits(fx["weights_first"]))
    assert np.array_equal(bits(vol["tsdf_values"]), bits(fx["values"]))
    assert np.array_equal(bits(vol["tsdf_weights"]), bits(fx["weights"]))
    assert int((fx["weights"] > 0).sum()) > 10000  # the case really fuses something


@pytest.mark.parametrize("name", CASES)
def test_oracle_sampling_is_bit_exact(name):
    fx = hp.load(name)
    vol = dict(tsdf_values=fx["values"], tsdf_weights=fx["weights"], origin=fx["origin"], voxel_size=float(fx["voxel_size"]))
    for what in ("weights", "tsdf"):
        for mode in ("bilinear", "nearest"):
            got = ot.sample_volume(vol, fx["points"], what, mode)
            assert np.array_equal(got, fx[f"sample_{what}_{mode}"]), (what, mode)


def test_nonfinite_pixel_coordinates_follow_the_pinned_build():
    """tsdf_near holds voxels whose projected pixel overflows fp16: ATen's CPU build reads row / column 0 there (pinned),
    its CUDA builds read the zero padding -- the half-index CUDA semantics must differ from the CPU's on exactly such voxels
    and nowhere else; the fp32-index CUDA semantics (opmath_t coordinates) additionally moves every voxel whose fp16-rounded
    index falls on the other side of a .5 boundary."""
    fx = hp.load("tsdf_near")
    res = {}
    for sem in ("cpu", "cuda_half_index", "cuda"):
        vol = ot.volume_from_bounds(case_bounds(fx), float(fx["voxel_size"]))
        ot.integrate_depth(vol, fx["depth"], fx["cam_T_world"], fx["K"], min_depth=float(fx["min_depth"]),
                           max_depth=float(fx["max_depth"]), semantics=sem)
        res[sem] = vol["tsdf_weights"].copy()
    assert np.array_equal(bits(res["cpu"]), bits(fx["weights"]))
    diff = int((bits(res["cpu"]) != bits(res["cuda_half_index"])).sum())
    assert 0 < diff < 100
    assert int((bits(res["cuda"]) != bits(res["cuda_half_index"])).sum()) > 0


@pytest.mark.parametrize("name", CASES)
def test_host_mirror_grid_and_frame_constants(name):
    """doubletake_b200.tsdf on the host: same grid as the reference's from_bounds; the per-frame constants handed to the
    kernel (P, frustum box) equal the oracle's bit for bit."""
    fx = hp.load(name)
    vol = bt.TSDF.from_bounds(case_bounds(fx), float(fx["voxel_size"]))
    assert np.array_equal(bits(vol.voxel_coords_3hwd.numpy()), bits(fx["voxel_coords"]))
    assert np.array_equal(bits(vol.origin.numpy()), bits(fx["origin"]))
    assert tuple(vol.tsdf_values.shape) == fx["values"].shape and float(vol.tsdf_values.float().max()) == -1.0
    fuser = bt.TSDFFuser(vol, max_depth=float(fx["max_depth"]), semantics="aten_cpu")
    ih, iw = fx["depth"].shape[2:]
    for b in range(fx["depth"].shape[0]):
        K, T = torch.from_numpy(fx["K"][b]), torch.from_numpy(fx["cam_T_world"][b])
        P, lo, hi = fuser._frame_constants(T, K, ih, iw)
        trunc = 3.0 * float(fx["voxel_size"])
        invK = np.linalg.inv(fx["K"][b].astype(np.float32)).astype(np.float16)
        wTc = np.linalg.inv(fx["cam_T_world"][b].astype(np.float32)).astype(np.float16)
        olo, ohi = ot.frustum_bounds(invK, wTc, 0.01, float(fx["max_depth"]) + trunc + 0.1, ih, iw)
        oP = ot.matmul_h(fx["K"][b], fx["cam_T_world"][b])[:3]
        assert np.array_equal(np.float32(P), oP.astype(np.float32).ravel())
        assert np.array_equal(np.float32(lo), olo.astype(np.float32)) and np.array_equal(np.float32(hi), ohi.astype(np.float32))


def test_mirror_rejects_cpu_and_bad_arguments(tmp_path):
    vol = bt.TSDF.from_bounds(dict(xmin=0, xmax=0.5, ymin=0, ymax=0.5, zmin=0, zmax=0.5), 0.05)
    with pytest.raises(RuntimeError):
        bt.TSDFFuser(vol, use_gpu=False)
    with pytest.raises(KeyError):
        bt.TSDF.from_bounds(dict(xmin=0), 0.05)
    with pytest.raises(ValueError):
        vol.sample_tsdf(torch.zeros(4, 2))
    with pytest.raises(NotImplementedError):
        vol.to_mesh()
    path = str(tmp_path / "v.npz")
    vol.save_tsdf(path)
    back = bt.TSDF.from_file(path)
    assert torch.equal(back.voxel_coords_3hwd, vol.voxel_coords_3hwd) and back.voxel_size == vol.voxel_size
    assert back._origin_f32 is None  # a loaded grid is read from memory, not regenerated


@pytest.mark.parametrize("name", CASES)
def test_index_box_covers_every_voxel_of_the_frustum_box(name):
    """The launch scans only TSDFFuser._index_box: it must contain every voxel whose fp16 coordinate passes the reference's
    strict box test (and stay well below the whole volume for a camera inside a large one)."""
    fx = hp.load(name)
    vol = bt.TSDF.from_bounds(case_bounds(fx), float(fx["voxel_size"]))
    fuser = bt.TSDFFuser(vol, max_depth=float(fx["max_depth"]), semantics="aten_cpu")
    ih, iw = fx["depth"].shape[2:]
    c = fx["voxel_coords"].astype(np.float32)
    dims = c.shape[1:]
    for b in range(fx["depth"].shape[0]):
        _, lo, hi = fuser._frame_constants(torch.from_numpy(fx["cam_T_world"][b]), torch.from_numpy(fx["K"][b]), ih, iw)
        begin, end = fuser._index_box(lo, hi, dims)
        assert begin[2] % 8 == 0 and end[2] % 8 == 0 and all(0 <= bb <= ee <= d for bb, ee, d in zip(begin, end, dims))
        inside = np.ones(dims, bool)
        for a in range(3):
            inside &= (c[a] > np.float32(lo[a])) & (c[a] < np.float32(hi[a]))
        idx = np.argwhere(inside)
        assert len(idx) > 0
        assert all(idx[:, a].min() >= begin[a] and idx[:, a].max() < end[a] for a in range(3))
    big = bt.TSDF(None, torch.zeros(8, 8, 8), torch.zeros(8, 8, 8), 0.04, torch.tensor([-10.0, -10.0, -10.0]),
                  _origin_f32=torch.tensor([-10.0, -10.0, -10.0]))
    bf = bt.TSDFFuser(big, max_depth=3.0, semantics="aten_cpu")
    begin, end = bf._index_box(lo, hi, (504, 504, 504))
    assert np.prod([e - b for b, e in zip(begin, end)]) < 0.05 * 504 ** 3
    assert bf._index_box([float("nan")] * 3, hi, (504, 504, 504)) == ([0, 0, 0], [504, 504, 504])


@pytest.mark.parametrize("name", CASES)
def test_torch_restatement_is_bit_exact_on_cpu(name):
    """oracle/oracle_tsdf_torch.py executes the reference's torch ops in the reference's order; on the CPU it must land on
    the fixtures bit for bit.  The GPU suite then runs the SAME function on CUDA to pin the product's aten_cuda branch."""
    from oracle import oracle_tsdf_torch as ott

    fx = hp.load(name)
    seed, nf, ih, iw, batch, with_mask, ext = [int(v) for v in fx["meta"]]
    coords = torch.from_numpy(fx["voxel_coords"])
    values = -torch.ones(coords.shape[1:], dtype=torch.float16)
    weights = torch.zeros(coords.shape[1:], dtype=torch.float16)
    for s in range(0, nf, batch):
        ott.integrate_depth(coords, values, weights, float(fx["voxel_size"]), torch.from_numpy(fx["depth"][s:s + batch]),
                            torch.from_numpy(fx["cam_T_world"][s:s + batch]), torch.from_numpy(fx["K"][s:s + batch]),
                            min_depth=float(fx["min_depth"]), max_depth=float(fx["max_depth"]),
                            depth_mask_b1hw=torch.from_numpy(fx["mask"][s:s + batch]) if with_mask else None,
                            extended_neg_truncation=bool(ext))
    assert np.array_equal(bits(values.numpy()), bits(fx["values"]))
    assert np.array_equal(bits(weights.numpy()), bits(fx["weights"]))


def test_index_box_margin_follows_the_fp16_ulp_far_from_the_origin():
    """at |coord| >= 128 one fp16 ulp is 0.125 m = 3 voxels of 4 cm; the scanned index box must still contain
    every voxel whose ROUNDED coordinate passes the strict box test."""
    origin = np.array([250.0, -300.0, 120.0], np.float32)
    dims = (64, 64, 64)
    vs = 0.04
    vol = bt.TSDF(None, torch.zeros(dims), torch.zeros(dims), vs, torch.from_numpy(origin).half(),
                  _origin_f32=torch.from_numpy(origin))
    fuser = bt.TSDFFuser(vol, max_depth=3.0)
    coords = ot.generate_voxel_coords(origin, dims, vs).astype(np.float32)
    rng = np.random.default_rng(3)
    for _ in range(20):
        lo = (origin + rng.uniform(0.2, 1.0, 3)).astype(np.float16).astype(np.float32)
        hi = (lo + rng.uniform(0.3, 1.2, 3)).astype(np.float16).astype(np.float32)
        begin, end = fuser._index_box(lo.tolist(), hi.tolist(), dims)
        inside = np.ones(dims, bool)
        for a in range(3):
            inside &= (coords[a] > lo[a]) & (coords[a] < hi[a])
        idx = np.argwhere(inside)
        if len(idx):
            assert all(idx[:, a].min() >= begin[a] and idx[:, a].max() < end[a] for a in range(3)), (lo, hi, begin, end)


def test_raycast_oracle_recovers_a_fused_plane():
    """The hint renderer's CPU oracle (oracle_tsdf.raycast_hint = the product's marching rule, operation for operation):
    fuse a tilted plane four times, ray-cast it from the fusing camera and from a second, rotated and shifted one -- the
    rendered depth is the analytic ray/plane intersection to a fraction of a voxel, no pixel fakes a surface at a frustum
    boundary, and the default 0.025 confidence threshold of test_incremental.py:244 rejects these few-observation surfaces."""
    bounds = dict(xmin=-1.0, xmax=1.0, ymin=-0.8, ymax=0.8, zmin=0.2, zmax=2.8)
    vol = ot.volume_from_bounds(bounds, 0.04)
    ih, iw = 48, 64
    K = np.eye(4, dtype=np.float32)
    K[0, 0] = K[1, 1] = 60.0
    K[0, 2], K[1, 2] = iw / 2, ih / 2
    ys, xs = np.meshgrid(np.arange(ih), np.arange(iw), indexing="ij")
    rx, ry = (xs + 0.5 - K[0, 2]) / K[0, 0], (ys + 0.5 - K[1, 2]) / K[1, 1]
    depth = (2.0 / (1 - 0.2 * rx)).astype(np.float32)   # the plane z = 2 + 0.2 x seen from the origin
    T = np.eye(4, dtype=np.float32)
    for _ in range(4):
        ot.integrate_depth(vol, depth[None, None].astype(np.float16), T[None].astype(np.float16), K[None].astype(np.float16),
                           min_depth=0.5, max_depth=3.0)
    hint, mask, sw = ot.raycast_hint(vol, np.linalg.inv(K), T, ih, iw, weight_threshold=0.005)
    ok = mask > 0
    assert ok.mean() > 0.8 and np.isnan(hint[~ok]).all() and (sw[~ok] == 0).all()
    assert np.abs(hint[ok] - depth[ok]).max() < 0.25 * 0.04
    ang = np.deg2rad(8)
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], np.float32)
    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3], pose[:3, 3] = R, [0.15, 0.02, 0.1]
    hint2, mask2, _ = ot.raycast_hint(vol, np.linalg.inv(K), pose, ih, iw, weight_threshold=0.005)
    dirs = np.stack([rx, ry, np.ones_like(rx)], -1) @ R.T
    t = (2 + 0.2 * pose[0, 3] - pose[2, 3]) / (dirs[..., 2] - 0.2 * dirs[..., 0])
    ok2 = mask2 > 0
    assert ok2.mean() > 0.6 and np.abs(hint2[ok2] - t[ok2]).max() < 0.25 * 0.04
    _, mask3, _ = ot.raycast_hint(vol, np.linalg.inv(K), T, ih, iw)   # default threshold 0.025: 4 far observations are not enough
    assert mask3.sum() == 0
