"""Device-agnostic torch restatement of the reference's TSDF integration -- TEST INFRASTRUCTURE ONLY, never imported by
``doubletake_b200/``.

Why a second TSDF oracle: ATen evaluates fp16 ``grid_sample`` differently in its CPU and CUDA builds (index arithmetic in
fp16 vs fp32, non-finite indices -> 0 vs out of bounds).  ``oracle/oracle_tsdf.py`` restates both in numpy, but only the
CPU behaviour can be pinned on fixtures generated in the (GPU-less) build container.  The reference fuser always runs on
CUDA (``use_gpu=True``, tools/fusers_helper.py), so the CUDA behaviour is the one that matters.  This module performs the
SAME torch ops in the SAME order as ``TSDFFuser.integrate_depth`` (tools/tsdf.py:414-558, ``project_to_camera`` :401-412,
``get_frustum_bounds`` :15-50) on whatever device its inputs live on:

* on the CPU it must reproduce the golden fixtures bit for bit (tests/test_oracle_tsdf_golden.py) -- that proves it is a
  faithful restatement;
* on the GPU box it runs on CUDA, where ATen's own CUDA kernels decide every rounding, and the product's
  ``semantics="aten_cuda"`` kernel is compared with it (tests/test_gpu_tsdf.py).

The open3d hash set (:523-531) is not restated (it only feeds marching cubes).
"""
from __future__ import annotations

import torch
import torch.nn.functional as TF

TRUNCATION_SIZE = 3.0
MAX_W = 100.0
UPDATE_RATE = 2.5


def frustum_bounds(invK_44, world_T_cam_44, min_depth, max_depth, img_h, img_w):
    """tools/tsdf.py:15-50: the 8 corners of the view frustum in world space, their per-axis min / max."""
    dev, dt = invK_44.device, invK_44.dtype
    corners = torch.tensor([[0, 0, 1, 1], [img_w, 0, 1, 1], [0, img_h, 1, 1], [img_w, img_h, 1, 1]], dtype=dt, device=dev).T
    pts = invK_44 @ corners
    near, far = pts.clone(), pts.clone()
    near[:3] *= min_depth
    far[:3] *= max_depth
    world = world_T_cam_44 @ torch.cat([near, far], 1)
    return world.min(1).values[:3], world.max(1).values[:3]


def frame_constants(cam_T_world_44, K_44, img_h, img_w, depth_max):
    """The per-frame 4x4 work of integrate_depth (tools/tsdf.py:444-450 and the K @ T product of :405) on the device the
    matrices live on: (P_34, box_lo_3, box_hi_3), all fp16."""
    invK = torch.inverse(K_44[None].float()).half()[0]
    world_T_cam = torch.inverse(cam_T_world_44[None].float()).half()[0]
    lo, hi = frustum_bounds(invK, world_T_cam, 0.01, depth_max, img_h, img_w)
    return torch.matmul(K_44[None], cam_T_world_44[None])[0, :3], lo, hi


def integrate_depth(coords_3hwd, values, weights, voxel_size, depth_b1hw, cam_T_world_b44, K_b44, min_depth=0.5, max_depth=5.0,
                    depth_mask_b1hw=None, extended_neg_truncation=False, host_frame_constants=False):
    """In place on ``values`` / ``weights`` (fp16, same device as ``coords_3hwd``).  Every tensor op below is the
    reference's, line for line in meaning; intermediate names follow tools/tsdf.py.
    ``host_frame_constants``: evaluate the per-frame 4x4 constants (two fp32 inverses, the frustum corners, K @ T) on the
    CPU and move the fp16 results over -- a 4x4 fp16 product may round differently in cuBLAS than in ATen's CPU gemm, and one
    flipped fp16 ulp in P moves every projected voxel; with the constants shared, a CUDA run isolates the per-voxel
    arithmetic (projection, grid_sample, confidence, update) that the product's kernel restates."""
    dev = coords_3hwd.device
    truncation = TRUNCATION_SIZE * voxel_size
    dims = coords_3hwd.shape[1:]
    hom_14N = torch.cat([coords_3hwd, torch.ones_like(coords_3hwd[:1])], 0).flatten(1).unsqueeze(0)  # (1,4,N) fp16
    depth_b1hw = depth_b1hw.to(dev)
    img_h, img_w = depth_b1hw.shape[2:]
    img_size = torch.tensor([img_w, img_h], dtype=torch.float16, device=dev).view(1, 1, 1, 2)
    if depth_mask_b1hw is not None:
        depth_b1hw = depth_b1hw.clone()
        depth_b1hw[~depth_mask_b1hw.to(dev)] = -1
    flat_v, flat_w = values.view(-1), weights.view(-1)
    for b in range(len(depth_b1hw)):
        T_144 = cam_T_world_b44[b:b + 1].to(dev)
        K_144 = K_b44[b:b + 1].to(dev)
        depth_11hw = depth_b1hw[b:b + 1]
        depth_max = max_depth + truncation + 0.1
        if host_frame_constants:
            P_34, lo, hi = (t.to(dev) for t in frame_constants(cam_T_world_b44[b].cpu(), K_b44[b].cpu(), img_h, img_w, depth_max))
        else:
            P_34, lo, hi = frame_constants(T_144[0], K_144[0], img_h, img_w, depth_max)
        c = hom_14N[0, :3]
        in_box = torch.logical_and(c > lo.view(3, 1), c < hi.view(3, 1)).all(0)
        idx = in_box.nonzero().squeeze(1)
        vox_14N = hom_14N[..., in_box]
        P_134 = P_34[None]
        cam_13N = torch.matmul(P_134, vox_14N)
        cam_13N[:, :2] = cam_13N[:, :2] / cam_13N[:, 2, None]
        vz = cam_13N[:, 2:3]
        pix = cam_13N[:, :2].reshape(1, 2, 1, -1).permute(0, 2, 3, 1)
        pix = 2 * pix / img_size - 1
        sd = TF.grid_sample(input=depth_11hw, grid=pix, mode="nearest", padding_mode="zeros", align_corners=False).flatten(2)
        conf = torch.clamp(1.0 - (sd - min_depth) / (max_depth - min_depth), min=0.25, max=1.0) ** 2
        dist = sd - vz
        tsdf = torch.clamp(dist / truncation, min=-1.0, max=1.0)
        trunc_check = -truncation * 1.5 if extended_neg_truncation else -truncation
        valid = ((vz > 0) & (dist > trunc_check) & (sd > 0) & (vz > 0) & (vz < max_depth) & (conf > 0))[0, 0]
        sel = idx[valid]
        old_v, old_w = flat_v[sel], flat_w[sel]
        new_v = tsdf[0, 0][valid]
        new_w = conf[0, 0][valid] * UPDATE_RATE / MAX_W
        total = old_w + new_w
        flat_v[sel] = (old_v * old_w + new_v * new_w) / total
        flat_w[sel] = torch.clamp(total, max=1.0)
    return values.view(dims), weights.view(dims)
