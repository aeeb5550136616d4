"""CPU oracle for the hint-production step next to the hot path: TSDF integration of a predicted
depth map and sampling of the fused confidence.  TEST INFRASTRUCTURE ONLY -- never imported by ``doubletake_b200/``.

Restates reference ``tools/tsdf.py`` in explicit numpy arithmetic, one rounding per operation exactly where the reference's
fp16 torch ops round (the reference keeps voxel coordinates, TSDF values, weights, depth maps, intrinsics and extrinsics
in fp16, ``tools/fusers_helper.py:67-73``; every elementwise op computes in fp32 and rounds its result to fp16, the
(3x4)@(4xN) products accumulate in fp32 and round once):

* ``frustum_bounds``     tools/tsdf.py:15-50   (get_frustum_bounds)
* ``integrate_depth``    tools/tsdf.py:414-558 (TSDFFuser.integrate_depth), dense + masks instead of gather/scatter
* ``sample_volume``      tools/tsdf.py:277-337 (TSDF.sample_tsdf, CPU branch: fp32 grid_sample, align_corners=True)
* ``generate_voxel_coords`` / ``volume_from_bounds``  tools/tsdf.py:122-166

Pinned on fixtures produced by executing the reference itself (oracle/make_golden_tsdf.py -> tests/golden/tsdf_*.npz).
The open3d hash set of active voxels (tools/tsdf.py:523-531) only feeds the marching-cubes extension (row N3) and is not
restated.
"""
from __future__ import annotations

import numpy as np

F16 = np.float16
F32 = np.float32
VOX_MOD = 8
TRUNCATION_SIZE = 3.0
MAX_W = 100.0
UPDATE_RATE = 2.5


def h(x):
    """Round to fp16 (the result type of every reference op)."""
    return np.asarray(x, dtype=F32).astype(F16)


def f(x):
    return np.asarray(x).astype(F32)


def matmul_h(a, b):
    """fp16 @ fp16 -> fp16 with fp32 accumulation in k order (ATen cpublas gemm for reduced floating types)."""
    a32, b32 = f(a), f(b)
    acc = np.zeros(a32.shape[:-1] + b32.shape[-1:], dtype=F32)
    for k in range(a32.shape[-1]):
        acc = (acc + a32[..., :, k:k + 1] * b32[..., k:k + 1, :]).astype(F32)
    return acc.astype(F16)


def generate_voxel_coords(origin, dims, voxel_size):
    """tools/tsdf.py:155-166 then ``.half()`` (:143-145): fp32 origin + index * voxel_size, rounded once to fp16."""
    gx, gy, gz = np.meshgrid(np.arange(dims[0]), np.arange(dims[1]), np.arange(dims[2]), indexing="ij")
    grid = np.stack([gx, gy, gz], 0).astype(np.int64)
    # torch: int64 grid * python float -> float32 tensor (default dtype), + float32 origin
    coords = (f(origin).reshape(3, 1, 1, 1) + (grid.astype(F32) * F32(voxel_size)).astype(F32)).astype(F32)
    return coords.astype(F16)


def volume_from_bounds(bounds, voxel_size):
    """tools/tsdf.py:122-151: dims rounded up to multiples of 8, values -1, weights 0."""
    dims = [int(np.ceil((bounds[a + "max"] - bounds[a + "min"]) / voxel_size / VOX_MOD)) * VOX_MOD for a in "xyz"]
    origin = np.array([bounds["xmin"], bounds["ymin"], bounds["zmin"]], dtype=F32)
    coords = generate_voxel_coords(origin, dims, voxel_size)
    return dict(voxel_coords_3hwd=coords, tsdf_values=-np.ones(dims, F16), tsdf_weights=np.zeros(dims, F16),
                origin=origin.astype(F16), voxel_size=float(voxel_size))


def frustum_bounds(invK_44, world_T_cam_44, min_depth, max_depth, img_h, img_w):
    """tools/tsdf.py:15-50 on fp16 matrices."""
    corners = np.array([[0, 0, 1, 1], [img_w, 0, 1, 1], [0, img_h, 1, 1], [img_w, img_h, 1, 1]], dtype=F16).T  # (4,4)
    pts = matmul_h(invK_44, corners)
    near, far = pts.copy(), pts.copy()
    near[:3] = h(f(near[:3]) * F32(min_depth))
    far[:3] = h(f(far[:3]) * F32(max_depth))
    world = matmul_h(world_T_cam_44, np.concatenate([near, far], 1))  # (4,8)
    return world.min(1)[:3], world.max(1)[:3]


def nearest_sample_h(depth_hw, px, py, semantics="cpu"):
    """F.grid_sample(mode="nearest", padding_mode="zeros", align_corners=False) of an fp16 map at fp16 normalised
    coordinates (tools/tsdf.py:476-483).  ATen's fp16 behaviour differs between its two builds, and both are restated:

    * ``"cpu"`` (what the golden fixtures pin: the reference executed on CPU): ``((g + 1) * size - 1) / 2`` in c10::Half
      arithmetic, one fp16 rounding per operation; a NaN / +-inf index converts to integer 0 (x86 build), i.e. such
      voxels read row / column 0 instead of the zero padding.
    * ``"cuda"`` (ATen GridSampler.cu with opmath_t coordinates; pinned on the GPU box against torch's own CUDA ops through
      oracle_tsdf_torch.py): the same formula in fp32, never rounded back to fp16; the float -> int conversion saturates,
      so +-inf is out of bounds and NaN is 0.
    * ``"cuda_half_index"`` (older ATen CUDA builds with a scalar_t index; restated from source, unpinned): as "cuda" but
      the un-normalised index is rounded to fp16 once before nearbyint."""
    H, W = depth_hw.shape

    def unnormalize(g, size):
        if semantics == "cpu":
            i = h(f(h(f(h(f(g) + F32(1))) * F32(size))) - F32(1))
            return f(h(f(i) / F32(2)))
        i32 = ((((f(g) + F32(1)) * F32(size)).astype(F32) - F32(1)).astype(F32) / F32(2)).astype(F32)
        return f(h(i32)) if semantics == "cuda_half_index" else i32

    with np.errstate(invalid="ignore", over="ignore"):
        xn, yn = np.rint(unnormalize(px, W)), np.rint(unnormalize(py, H))  # nearbyint: ties to even
    if semantics == "cpu":
        xn = np.where(np.isfinite(xn), xn, F32(0))
        yn = np.where(np.isfinite(yn), yn, F32(0))
    else:
        xn = np.where(np.isnan(xn), F32(0), xn)
        yn = np.where(np.isnan(yn), F32(0), yn)
    ok = (xn >= 0) & (xn <= W - 1) & (yn >= 0) & (yn <= H - 1)
    xi = np.where(ok, xn, 0).astype(np.int64)
    yi = np.where(ok, yn, 0).astype(np.int64)
    return np.where(ok, depth_hw[yi, xi], F16(0)).astype(F16)


def integrate_depth(vol, depth_b1hw, cam_T_world_b44, K_b44, min_depth=0.5, max_depth=5.0, depth_mask_b1hw=None,
                    extended_neg_truncation=False, semantics="cpu"):
    """TSDFFuser.integrate_depth (tools/tsdf.py:414-558) for a batch of fp16 depth maps, in place on ``vol``."""
    coords = vol["voxel_coords_3hwd"]
    dims = coords.shape[1:]
    N = int(np.prod(dims))
    truncation = TRUNCATION_SIZE * vol["voxel_size"]
    depth_b1hw = np.asarray(depth_b1hw, F16)
    if depth_mask_b1hw is not None:
        depth_b1hw = np.where(np.asarray(depth_mask_b1hw, bool), depth_b1hw, F16(-1))
    img_h, img_w = depth_b1hw.shape[2:]
    hom = np.concatenate([coords.reshape(3, N), np.ones((1, N), F16)], 0)  # (4,N)
    values = vol["tsdf_values"].reshape(N)
    weights = vol["tsdf_weights"].reshape(N)
    for b in range(depth_b1hw.shape[0]):
        K, T = np.asarray(K_b44[b], F16), np.asarray(cam_T_world_b44[b], F16)
        depth_max = max_depth + truncation + 0.1
        invK = np.linalg.inv(f(K)).astype(F32).astype(F16)           # torch.inverse(K.float()).half()
        world_T_cam = np.linalg.inv(f(T)).astype(F32).astype(F16)
        lo, hi = frustum_bounds(invK, world_T_cam, 0.01, depth_max, img_h, img_w)
        c = coords.reshape(3, N)
        in_box = np.all((c > lo[:, None]) & (c < hi[:, None]), 0)
        P = matmul_h(K, T)[:3]                                        # (3,4) fp16
        cam = matmul_h(P, hom)                                        # (3,N) fp16
        vz = cam[2]
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            px = h(f(cam[0]) / f(vz))
            py = h(f(cam[1]) / f(vz))
            gx = h(f(h(f(h(F32(2) * f(px))) / F32(img_w))) - F32(1))  # 2 * pix / img_size - 1
            gy = h(f(h(f(h(F32(2) * f(py))) / F32(img_h))) - F32(1))
            sd = nearest_sample_h(depth_b1hw[b, 0], gx, gy, semantics)
            conf = h(f(h(f(sd) - F32(min_depth))) / F32(max_depth - min_depth))
            conf = h(F32(1) - f(conf))
            conf = np.clip(conf, F16(0.25), F16(1.0))
            conf = h(f(conf) * f(conf))
            dist = h(f(sd) - f(vz))
            tsdf = np.clip(h(f(dist) / F32(truncation)), F16(-1), F16(1))
            trunc_check = F16(-truncation * 1.5 if extended_neg_truncation else -truncation)  # compared in fp16
            valid = in_box & (vz > 0) & (dist > trunc_check) & (sd > 0) & (vz < F16(max_depth)) & (conf > 0)
            new_w = h(f(h(f(conf) * F32(UPDATE_RATE))) / F32(MAX_W))
            total = h(f(weights) + f(new_w))
            num = h(f(h(f(values) * f(weights))) + f(h(f(tsdf) * f(new_w))))
            new_v = h(f(num) / f(total))
        values[valid] = new_v[valid]
        weights[valid] = np.minimum(total, F16(1.0))[valid]
    vol["tsdf_values"] = values.reshape(dims)
    vol["tsdf_weights"] = weights.reshape(dims)
    return vol


def sample_volume(vol, world_points_N3, what="weights", mode="bilinear"):
    """TSDF.sample_tsdf (tools/tsdf.py:277-337), CPU branch: the fp16 volume is read as fp32 and sampled with fp32
    coordinates, ``align_corners=True``, zeros padding; ``bilinear`` on a 5-D input is trilinear."""
    volume = f(vol["tsdf_values"] if what == "tsdf" else vol["tsdf_weights"])
    dims = np.array(volume.shape, dtype=F32)
    p = f(world_points_N3)
    v = (p - f(vol["origin"]).reshape(1, 3)).astype(F32)
    v = (v / F32(vol["voxel_size"])).astype(F32)
    v = (v / (dims.reshape(1, 3) - F32(1))).astype(F32)
    g = (v * F32(2) - F32(1)).astype(F32)
    # grid_sample(x = last volume axis): coordinates are swapped to (z, y, x) by the reference, so axis a of the volume is
    # addressed by component a of the point; align_corners=True: index = (g + 1) / 2 * (size - 1)
    idx = [((g[:, a] + F32(1)) / F32(2) * (dims[a] - F32(1))).astype(F32) for a in range(3)]
    if mode == "nearest":
        r = [np.rint(i) for i in idx]
        ok = np.ones(len(p), bool)
        for a in range(3):
            ok &= (r[a] >= 0) & (r[a] <= dims[a] - 1)
        ii = [np.where(ok, r[a], 0).astype(np.int64) for a in range(3)]
        return np.where(ok, volume[ii[0], ii[1], ii[2]], F32(0)).astype(F32)
    lo = [np.floor(i) for i in idx]
    out = np.zeros(len(p), F32)
    # ATen grid_sampler_3d corner order: tnw, tne, tsw, tse, bnw, bne, bsw, bse with x = last axis
    x, y, z = idx[2], idx[1], idx[0]
    x0, y0, z0 = lo[2], lo[1], lo[0]
    x1, y1, z1 = x0 + 1, y0 + 1, z0 + 1
    corners = [
        (x0, y0, z0, (x1 - x) * (y1 - y) * (z1 - z)), (x1, y0, z0, (x - x0) * (y1 - y) * (z1 - z)),
        (x0, y1, z0, (x1 - x) * (y - y0) * (z1 - z)), (x1, y1, z0, (x - x0) * (y - y0) * (z1 - z)),
        (x0, y0, z1, (x1 - x) * (y1 - y) * (z - z0)), (x1, y0, z1, (x - x0) * (y1 - y) * (z - z0)),
        (x0, y1, z1, (x1 - x) * (y - y0) * (z - z0)), (x1, y1, z1, (x - x0) * (y - y0) * (z - z0)),
    ]
    for cx, cy, cz, w in corners:
        ok = (cx >= 0) & (cx <= dims[2] - 1) & (cy >= 0) & (cy <= dims[1] - 1) & (cz >= 0) & (cz <= dims[0] - 1)
        xi, yi, zi = [np.where(ok, c, 0).astype(np.int64) for c in (cx, cy, cz)]
        out = (out + np.where(ok, volume[zi, yi, xi] * w.astype(F32), F32(0)).astype(F32)).astype(F32)
    return out


def _index_of(vol, pts_N3):
    """The index arithmetic of sample_volume (one rounded fp32 op per step)."""
    dims = np.array(vol["tsdf_values"].shape, dtype=F32)
    v = (f(pts_N3) - f(vol["origin"]).reshape(1, 3)).astype(F32)
    v = (v / F32(vol["voxel_size"])).astype(F32)
    v = (v / (dims.reshape(1, 3) - F32(1))).astype(F32)
    g = (v * F32(2) - F32(1)).astype(F32)
    return np.stack([((g[:, a] + F32(1)) / F32(2) * (dims[a] - F32(1))).astype(F32) for a in range(3)], 1), dims


def raycast_hint(vol, invK_44, world_T_cam_44, height, width, z_near=0.05, z_far=10.0, weight_threshold=0.025, max_steps=4096):
    """CPU restatement of the product's TSDF ray caster (csrc/tsdf.cu::tsdf_raycast_kernel) -- the reference has no ray
    caster (it meshes the TSDF and rasterises the mesh, test_incremental.py:202-252), so this oracle pins OUR marching rule
    operation for operation; what it shares with the reference is pinned elsewhere: the confidence at the hit point is
    sample_volume (= TSDF.sample_tsdf, fixtures) and the threshold / NaN / mask rules are test_incremental.py:238-252.
    Vectorised over pixels: every ray carries its own state, finished rays are masked out.
    Returns (depth_hint with NaN, mask float, sampled_weights) of shape (height, width)."""
    invK, Rt = f(invK_44), f(world_T_cam_44)
    ys, xs = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    px, py = (xs.reshape(-1).astype(F32) + F32(0.5)), (ys.reshape(-1).astype(F32) + F32(0.5))
    fma = lambda a, b, c: (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(F32)  # noqa: E731  (one rounding)
    ray = np.stack([(fma(py, invK[i, 1], (invK[i, 0] * px).astype(F32)) + invK[i, 2]).astype(F32) for i in range(3)], 1)
    vs = F32(vol["voxel_size"])
    big, small = (F32(2) * vs).astype(F32), (F32(0.5) * vs).astype(F32)
    n = len(px)

    def point(z, sel):
        c = (z[:, None] * ray[sel]).astype(F32)
        out = []
        for i in range(3):
            acc = (Rt[i, 0] * c[:, 0]).astype(F32)
            acc = fma(c[:, 1], Rt[i, 1], acc)
            acc = fma(c[:, 2], Rt[i, 2], acc)
            out.append((acc + Rt[i, 3]).astype(F32))
        return np.stack(out, 1)

    def sample(what, idx):
        volume = f(vol["tsdf_values"] if what == "tsdf" else vol["tsdf_weights"])
        dims = volume.shape
        x, y, z = idx[:, 2], idx[:, 1], idx[:, 0]
        x0, y0, z0 = np.floor(x), np.floor(y), np.floor(z)
        x1, y1, z1 = x0 + 1, y0 + 1, z0 + 1
        out = np.zeros(len(idx), F32)
        lo = np.full(len(idx), F32(3e38), F32)
        for cx, cy, cz, wx, wy, wz in ((x0, y0, z0, x1 - x, y1 - y, z1 - z), (x1, y0, z0, x - x0, y1 - y, z1 - z),
                                       (x0, y1, z0, x1 - x, y - y0, z1 - z), (x1, y1, z0, x - x0, y - y0, z1 - z),
                                       (x0, y0, z1, x1 - x, y1 - y, z - z0), (x1, y0, z1, x - x0, y1 - y, z - z0),
                                       (x0, y1, z1, x1 - x, y - y0, z - z0), (x1, y1, z1, x - x0, y - y0, z - z0)):
            ok = (cx >= 0) & (cx <= dims[2] - 1) & (cy >= 0) & (cy <= dims[1] - 1) & (cz >= 0) & (cz <= dims[0] - 1)
            xi, yi, zi = [np.where(ok, c, 0).astype(np.int64) for c in (cx, cy, cz)]
            w = ((wx.astype(F32) * wy.astype(F32)).astype(F32) * wz.astype(F32)).astype(F32)
            out = (out + np.where(ok, (volume[zi, yi, xi] * w).astype(F32), F32(0))).astype(F32)
            lo = np.minimum(lo, np.where(ok, volume[zi, yi, xi], F32(0)))
        return out, lo

    z = np.full(n, F32(z_near), F32)
    z_prev, v_prev, w_prev = np.zeros(n, F32), np.full(n, F32(-1), F32), np.zeros(n, F32)
    hit = np.full(n, F32(-1), F32)
    active = np.ones(n, bool)
    for _ in range(max_steps):
        active &= z <= F32(z_far)
        if not active.any():
            break
        a = np.nonzero(active)[0]
        idx, dims = _index_of(vol, point(z[a], a))
        inside = np.all((idx >= 0) & (idx <= dims.reshape(1, 3) - 1), 1)
        # "observed" sample: all eight surrounding voxels carry a weight > 0 (corners in the zero padding count as 0)
        w = (inside & (sample("weights", idx)[1] > 0)).astype(F32)
        v = np.where(w > 0, sample("tsdf", idx)[0], F32(-1)).astype(F32)
        cross = (v_prev[a] > 0) & (v <= 0) & (w_prev[a] > 0) & (w > 0)
        with np.errstate(divide="ignore", invalid="ignore"):
            zs = (z_prev[a] + (((z[a] - z_prev[a]).astype(F32) * v_prev[a]).astype(F32) / (v_prev[a] - v).astype(F32)).astype(F32)).astype(F32)
        hit[a[cross]] = zs[cross]
        active[a[cross]] = False
        keep = ~cross
        z_prev[a[keep]], v_prev[a[keep]], w_prev[a[keep]] = z[a[keep]], v[keep], w[keep]
        z[a[keep]] = (z[a[keep]] + np.where(np.abs(v[keep]) >= F32(0.99), big, small)).astype(F32)
    found = hit > 0
    sw = np.zeros(n, F32)
    if found.any():
        idx, _ = _index_of(vol, point(hit[found], found))
        sw[found] = sample("weights", idx)[0]
    depth = (hit * ray[:, 2]).astype(F32)
    valid = found & ~(sw < F32(weight_threshold))
    hint = np.where(valid, depth, np.float32("nan")).astype(F32)
    return (hint.reshape(height, width), valid.astype(F32).reshape(height, width),
            np.where(valid, sw, F32(0)).astype(F32).reshape(height, width))
