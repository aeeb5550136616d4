"""Generate tests/golden/tsdf_*.npz by EXECUTING THE REAL REFERENCE ``tools/tsdf.py`` on CPU (build container only).

Run:  python oracle/make_golden_tsdf.py       (needs /root/reference; never runs on the GPU box)

``TSDF.from_bounds``, ``TSDFFuser.integrate_depth`` (fp16, ``use_gpu=False``) and ``TSDF.sample_tsdf`` are the reference's
own code; open3d / trimesh / pytorch3d / skimage and the compiled marching-cubes extension are import-only stubs
(oracle/ref_stubs): the integrate / sample arithmetic never calls them (the open3d hash set only records active voxels
for marching cubes).  Inputs are seeded and stored in the fixture.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_stubs"))
sys.path.insert(0, "/root/reference/src")
_m = types.ModuleType("doubletake.utils.pytorch3d_extras")  # the compiled marching-cubes wrapper (row N3)
_m.marching_cubes = None
sys.modules["doubletake.utils.pytorch3d_extras"] = _m

from doubletake.tools import tsdf as ref_tsdf  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_grad_enabled(False)


def scene(seed, n_frames, img_h, img_w, with_mask):
    """Seeded camera track in front of a wavy wall: fp16 depth maps (zeros = holes), extrinsics, intrinsics."""
    g = torch.Generator().manual_seed(seed)
    K = torch.eye(4)
    K[0, 0] = K[1, 1] = 0.9 * img_w
    K[0, 2], K[1, 2] = img_w / 2 - 0.5, img_h / 2 - 0.5
    ys, xs = torch.meshgrid(torch.arange(img_h, dtype=torch.float32), torch.arange(img_w, dtype=torch.float32), indexing="ij")
    depths, exts, masks = [], [], []
    for i in range(n_frames):
        ang = (torch.rand(3, generator=g) - 0.5) * 0.3
        t = (torch.rand(3, generator=g) - 0.5) * torch.tensor([0.5, 0.3, 0.3])
        cx, sx, cy, sy, cz, sz = ang[0].cos(), ang[0].sin(), ang[1].cos(), ang[1].sin(), ang[2].cos(), ang[2].sin()
        Rx = torch.tensor([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        Ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rz = torch.tensor([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        T = torch.eye(4)
        T[:3, :3] = Rz @ Ry @ Rx
        T[:3, 3] = t
        d = 1.6 + 0.35 * torch.sin(xs / img_w * 6.0 + i) * torch.cos(ys / img_h * 4.0) + 0.05 * torch.rand(img_h, img_w, generator=g)
        holes = torch.rand(img_h, img_w, generator=g) < 0.05
        d[holes] = 0.0
        depths.append(d[None, None])
        exts.append(T[None])
        masks.append((torch.rand(img_h, img_w, generator=g) > 0.1)[None, None])
    return (torch.cat(depths).half(), torch.cat(exts).half(), K[None].expand(n_frames, 4, 4).contiguous().half(),
            torch.cat(masks) if with_mask else None)


def case(name, bounds, voxel_size, seed, n_frames, img_h, img_w, batch, with_mask=False, extended=False, max_depth=3.0):
    vol = ref_tsdf.TSDF.from_bounds(bounds, voxel_size)
    fuser = ref_tsdf.TSDFFuser(vol, max_depth=max_depth, use_gpu=False)
    depth, ext, K, mask = scene(seed, n_frames, img_h, img_w, with_mask)
    snap = {}
    for s in range(0, n_frames, batch):
        fuser.integrate_depth(depth[s:s + batch], ext[s:s + batch], K[s:s + batch],
                              depth_mask_b1hw=None if mask is None else mask[s:s + batch], extended_neg_truncation=extended)
        if s == 0:
            snap = dict(values_first=vol.tsdf_values.numpy().copy(), weights_first=vol.tsdf_weights.numpy().copy())
    g = torch.Generator().manual_seed(seed + 1)
    lo = torch.tensor([bounds["xmin"], bounds["ymin"], bounds["zmin"]])
    hi = torch.tensor([bounds["xmax"], bounds["ymax"], bounds["zmax"]])
    pts = lo + (hi - lo) * (torch.rand(4096, 3, generator=g) * 1.2 - 0.1)  # 10 % margin: some points fall outside
    samples = {}
    for what in ("weights", "tsdf"):
        for mode in ("bilinear", "nearest"):
            samples[f"sample_{what}_{mode}"] = vol.sample_tsdf(pts, what_to_sample=what, sampling_method=mode).numpy()
    out = dict(
        meta=np.array([seed, n_frames, img_h, img_w, batch, int(with_mask), int(extended)], np.int64),
        bounds=np.array([bounds[k] for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")], np.float64),
        voxel_size=np.float64(voxel_size), max_depth=np.float64(max_depth), min_depth=np.float64(fuser.min_depth),
        depth=depth.numpy(), cam_T_world=ext.numpy(), K=K.numpy(),
        voxel_coords=vol.voxel_coords_3hwd.numpy(), origin=vol.origin.numpy(),
        values=vol.tsdf_values.numpy(), weights=vol.tsdf_weights.numpy(), points=pts.numpy(), **snap, **samples)
    if mask is not None:
        out["mask"] = mask.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    w = vol.tsdf_weights.float()
    print(f"{name}: dims {tuple(vol.tsdf_values.shape)}, touched voxels {int((w > 0).sum())}, max weight {float(w.max()):.4f}")


if __name__ == "__main__":
    room = dict(xmin=-1.2, xmax=1.2, ymin=-0.9, ymax=0.9, zmin=0.6, zmax=2.6)
    case("tsdf_room", room, 0.04, 5001, n_frames=6, img_h=96, img_w=128, batch=2)
    case("tsdf_room_mask_ext", room, 0.05, 5002, n_frames=4, img_h=72, img_w=96, batch=1, with_mask=True, extended=True)
    case("tsdf_near", dict(xmin=-0.4, xmax=0.4, ymin=-0.3, ymax=0.3, zmin=-0.1, zmax=1.4), 0.02, 5003, n_frames=3,
         img_h=96, img_w=128, batch=3, max_depth=1.0)
