"""CPU oracle for the plane-sweep depth hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional, fp32, CPU restatement (torch ops only; no import of /root/reference, so it travels to the GPU
box) of the reference algorithm.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product package
``doubletake_b200`` never does.

Pinned against golden vectors produced by executing the real reference modules in the build container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``; checked by ``tests/test_oracle_golden.py``).  The
reference ships no tests or golden vectors of its own, so those executed-reference fixtures
are the pin.

Every function cites the reference lines (relative to /root/reference/src/doubletake/) it restates.
Weights are passed as plain ``{name: tensor}`` dicts using the reference's state_dict key names.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------------------------------------


def depth_planes(min_depth, max_depth, num_planes):
    """modules/cost_volume.py:96-130 with the ramp buffer of :60-61.  Returns (D,) fp32 plane depths."""
    ramp = torch.linspace(0, 1, num_planes, dtype=torch.float32)
    mn = torch.as_tensor(min_depth, dtype=torch.float32).reshape(())
    mx = torch.as_tensor(max_depth, dtype=torch.float32).reshape(())
    return torch.exp(torch.log(mn) + torch.log(mx / mn) * ramp)


def pose_measures(src_poses):
    """utils/geometry_utils.py:187-199 (DVMVS pose distance) on (B,K,4,4) -> three (B,K) tensors."""
    R = src_poses[..., :3, :3]
    t = src_poses[..., :3, 3]
    trace = R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]
    r_m = torch.sqrt(2 * (1 - torch.minimum(torch.full_like(trace, 3.0), trace) / 3))
    t_m = torch.norm(t, dim=-1)
    return torch.sqrt(t_m**2 + r_m**2), r_m, t_m


def pixel_grid(H, W):
    """utils/geometry_utils.py:34-48: homogeneous pixel centres (x+0.5, y+0.5, 1), row-major, shape (3, H*W)."""
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    return torch.stack([xs.flatten() + 0.5, ys.flatten() + 0.5, torch.ones(H * W)], 0).float()


def backproject(depth, cur_invK, pix):
    """utils/geometry_utils.py:55-63: X = d * (invK[:3,:3] @ pix), homogeneous.  depth scalar, -> (B,4,N)."""
    rays = torch.matmul(cur_invK[:, :3, :3], pix[None])
    X = depth * rays
    return torch.cat([X, torch.ones_like(X[:, :1])], 1)


def project(X_b4N, src_Ks, src_extrinsics, eps=1e-8):
    """utils/geometry_utils.py:77-93 for all K views: returns pixel coords (B,K,2,N) and depth z' (B,K,N)."""
    P = torch.matmul(src_Ks, src_extrinsics)  # (B,K,4,4)
    q = torch.matmul(P[:, :, :3], X_b4N[:, None])  # (B,K,3,N)
    z = q[:, :, 2:3]
    zp = z + eps
    scale = torch.where(z.abs() > eps, 1.0 / zp, torch.ones_like(zp))
    return q[:, :, :2] * scale, zp[:, :, 0]


def warp_sources(src_feats, pix_bk2N, H, W):
    """modules/mesh_hint_volume.py:238-249: uv = 2*pix*[1/W,1/H] - 1, bilinear grid_sample, zeros padding,
    align_corners=False.  -> (B,K,C,H,W)."""
    B, K, C = src_feats.shape[:3]
    uv_scale = torch.tensor([1 / W, 1 / H], dtype=torch.float32, device=src_feats.device).view(1, 1, 1, 2)
    grid = pix_bk2N.reshape(B * K, 2, H, W).permute(0, 2, 3, 1)
    grid = 2 * grid * uv_scale - 1
    warped = F.grid_sample(
        src_feats.reshape(B * K, C, H, W), grid, mode="bilinear", padding_mode="zeros", align_corners=False
    )
    return warped.view(B, K, C, H, W)


def bounds_mask(pix_bk2N, H, W):
    """modules/cost_volume.py:73-94: 2-pixel inset bounds test per view."""
    x, y = pix_bk2N[:, :, 0], pix_bk2N[:, :, 1]
    return (x > 2) & (x < W - 2) & (y > 2) & (y < H - 2)


# --------------------------------------------------------------------------------------------------------
# cost volumes
# --------------------------------------------------------------------------------------------------------


def _lowest_cost(volume, planes):
    """modules/cost_volume.py:317-320,356-361: argmax over planes (first max wins) -> plane depth."""
    idx = torch.argmax(volume, 1)
    return planes[idx], idx


def cost_volume_dot(cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth,
                    num_planes):
    """``CostVolumeManager.forward`` (modules/cost_volume.py:219-363): per plane, warp K source maps, dot with
    the current features, multiply by the depth-validity mask, sum over views.
    Returns dict(volume (B,D,H,W), lowest_cost (B,H,W), index (B,H,W) int64, planes (D,))."""
    B, K, C, H, W = src_feats.shape
    # evaluated on the CPU (bit-identical planes / grid), then moved: the oracle also runs on CUDA tensors (bench.py times
    # it there as the "reference on the same B200" arm)
    planes = depth_planes(min_depth, max_depth, num_planes).to(cur_feats.device)
    pix = pixel_grid(H, W).to(cur_feats.device)
    vols = []
    for d in planes:
        X = backproject(d, cur_invK, pix)
        uv, zp = project(X, src_Ks, src_extrinsics)
        warped = warp_sources(src_feats, uv, H, W)
        mask = (zp > 0).float().view(B, K, H, W)
        dot = (warped * cur_feats[:, None]).sum(2) * mask
        vols.append(dot.sum(1, keepdim=True))
    volume = torch.cat(vols, 1)
    lowest, idx = _lowest_cost(volume, planes)
    return dict(volume=volume, lowest_cost=lowest, index=idx, planes=planes)


def _mlp(x, w, prefix):
    """modules/networks.py:120-135 with disable_final_activation=True: Linear, LeakyReLU(0.01), Linear,
    LeakyReLU(0.01), Linear."""
    x = F.leaky_relu(F.linear(x, w[f"{prefix}.net.0.weight"], w[f"{prefix}.net.0.bias"]), 0.01)
    x = F.leaky_relu(F.linear(x, w[f"{prefix}.net.2.weight"], w[f"{prefix}.net.2.bias"]), 0.01)
    return F.linear(x, w[f"{prefix}.net.4.weight"], w[f"{prefix}.net.4.bias"])


def feature_volume(cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth,
                   num_planes, weights, hint=None, mask_mode="fast"):
    """``FeatureVolumeManager`` (modules/feature_volume.py:81-356) when ``hint is None``, else
    ``FeatureMeshHintVolumeManager`` (modules/mesh_hint_volume.py:84-393).

    ``weights``: dict with ``mlp.net.{0,2,4}.{weight,bias}`` (+ ``hint_mlp.*`` for the hint variant).
    ``hint``: dict with depth_hint_b1hw / sampled_weights_b1hw / depth_hint_mask_b1hw at any resolution
    (nearest-resized to matching res as mesh_hint_volume.py:186-204 does).
    ``mask_mode``: "fast" -> (B,H,W) any-view mask of the last plane (feature_volume.py:247-259,
    mesh_hint_volume.py:818-822); "slow_hint" -> (B,K,H,W) per-view mask of the last plane
    (mesh_hint_volume.py:273-287).
    """
    B, K, C, H, W = src_feats.shape
    # evaluated on the CPU (bit-identical planes / grid), then moved: the oracle also runs on CUDA tensors (bench.py times
    # it there as the "reference on the same B200" arm)
    planes = depth_planes(min_depth, max_depth, num_planes).to(cur_feats.device)
    pix = pixel_grid(H, W).to(cur_feats.device)
    comb, r_m, t_m = pose_measures(src_poses)

    def bc(x):  # (B,K) -> (B,K,H,W)
        return x[:, :, None, None].expand(B, K, H, W)

    t_src = src_poses[:, :, :3, 3]  # (B,K,3)   utils/geometry_utils.py:178-180

    if hint is not None:
        hd = F.interpolate(hint["depth_hint_b1hw"], size=(H, W), mode="nearest")
        hw = F.interpolate(hint["sampled_weights_b1hw"], size=(H, W), mode="nearest").clone()
        hm = F.interpolate(hint["depth_hint_mask_b1hw"], size=(H, W), mode="nearest").bool()
        hw[~hm] = 0

    vols = []
    overall = None
    for d in planes:
        X = backproject(d, cur_invK, pix)  # (B,4,N)
        uv, zp = project(X, src_Ks, src_extrinsics)
        warped = warp_sources(src_feats, uv, H, W)
        mask_b = (zp > 0).view(B, K, H, W)
        mask = mask_b.float()
        if mask_mode == "slow_hint":
            overall = mask_b & bounds_mask(uv, H, W).view(B, K, H, W)
        else:
            overall = mask_b.any(1) & bounds_mask(uv, H, W).view(B, K, H, W).any(1)

        Xw = X[:, :3]  # (B,3,N)
        ray_cur = F.normalize(Xw, dim=1)  # feature_volume.py:262-273
        ray_src = F.normalize(Xw[:, None] - t_src[..., None], dim=2)  # (B,K,3,N) geometry_utils.py:178-182
        angle = F.cosine_similarity(ray_cur[:, None].expand_as(ray_src), ray_src, dim=2, eps=1e-5)
        dot = (warped * cur_feats[:, None]).sum(2) * mask

        feats = torch.cat(
            [
                warped.reshape(B, K * C, H, W),
                cur_feats,
                mask,
                zp.view(B, K, H, W),
                d.expand(B, 1, H, W),
                dot,
                angle.view(B, K, H, W),
                ray_cur.view(B, 3, H, W),
                ray_src.reshape(B, 3 * K, H, W),
                bc(comb),
                bc(r_m),
                bc(t_m),
            ],
            1,
        )  # channel order: mesh_hint_volume.py:343-367
        score = _mlp(feats.permute(0, 2, 3, 1), weights, "mlp")  # (B,H,W,1)
        if hint is not None:
            h = torch.abs(hd - d)
            h[~hm] = -1
            hin = torch.cat([score, h.permute(0, 2, 3, 1), hw.permute(0, 2, 3, 1)], -1)
            score = _mlp(hin, weights, "hint_mlp")  # mesh_hint_volume.py:373-386
        vols.append(score.squeeze(-1).unsqueeze(1))
    volume = torch.cat(vols, 1)
    lowest, idx = _lowest_cost(volume, planes)
    return dict(volume=volume, lowest_cost=lowest, index=idx, planes=planes, mask=overall)


# --------------------------------------------------------------------------------------------------------
# conv stacks
# --------------------------------------------------------------------------------------------------------


def basic_block(x, w, p, stride=1):
    """modules/layers.py:77-94: conv3x3+b -> LReLU(0.2) -> conv3x3+b -> (+x | +downsample(x)) -> LReLU(0.2).
    The skip projection is 1x1 when stride==1 and a strided 3x3 otherwise (layers.py:67-74)."""
    out = F.conv2d(x, w[f"{p}.conv1.weight"], w[f"{p}.conv1.bias"], stride=stride, padding=1)
    out = F.leaky_relu(out, 0.2)
    out = F.conv2d(out, w[f"{p}.conv2.weight"], w[f"{p}.conv2.bias"], padding=1)
    if f"{p}.downsample.0.weight" in w:
        dw = w[f"{p}.downsample.0.weight"]
        pad = 1 if dw.shape[-1] == 3 else 0
        x = F.conv2d(x, dw, w[f"{p}.downsample.0.bias"], stride=stride, padding=pad)
    return F.leaky_relu(out + x, 0.2)


def cv_encoder(cost_volume, img_feats, w, prefix="convs"):
    """``CVEncoder.forward`` (modules/networks.py:110-117): 4 stages of ds_conv_i (stride 1 then 2) ->
    cat [x, prior_i] -> two BasicBlocks."""
    outs = []
    x = cost_volume
    for i in range(len(img_feats)):
        x = basic_block(x, w, f"{prefix}.ds_conv_{i}", stride=1 if i == 0 else 2)
        x = torch.cat([x, img_feats[i]], 1)
        x = basic_block(x, w, f"{prefix}.conv_{i}.0")
        x = basic_block(x, w, f"{prefix}.conv_{i}.1")
        outs.append(x)
    return outs


def _up2(x):
    """utils/generic_utils.py:95-104: bilinear x2, align_corners=False."""
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)


def depth_decoder_pp(feats, w, prefix="convs"):
    """``DepthDecoderPP.forward`` (modules/networks.py:65-85): UNet++ grid; node(i,j) =
    in_conv_ij(cat[right_conv_{i,j-1}(node(i,j-1)), up(diag_conv_{i+1,j-1}(node(i+1,j-1))),
    up(up_conv_{i+1,j}(node(i+1,j))) if i+j != 4]).  Only the last head written per scale survives
    (networks.py:81), so the four surviving heads are evaluated at nodes (3,1),(2,2),(1,3),(0,4)."""
    prev = list(feats)
    out = {}
    for j in range(1, 5):
        col = []
        for i in range(4 - j, -1, -1):
            parts = [basic_block(prev[i], w, f"{prefix}.right_conv_{i}{j - 1}")]
            parts.append(_up2(basic_block(prev[i + 1], w, f"{prefix}.diag_conv_{i + 1}{j - 1}")))
            if i + j != 4:
                parts.append(_up2(basic_block(col[-1], w, f"{prefix}.up_conv_{i + 1}{j}")))
            x = torch.cat(parts, 1)
            x = basic_block(x, w, f"{prefix}.in_conv_{i}{j}.0")
            x = basic_block(x, w, f"{prefix}.in_conv_{i}{j}.conv_0")
            col.append(x)
            if i + j == 4:
                h = x
                if i != 0:
                    h = basic_block(h, w, f"{prefix}.output_{i}.0")
                h = F.conv2d(h, w[f"{prefix}.output_{i}.1.weight"], w[f"{prefix}.output_{i}.1.bias"])
                out[f"log_depth_pred_s{i}_b1hw"] = h
        prev = col[::-1]
    return out


def skip_decoder_regression(feats, w, prefix=""):
    """``SkipDecoderRegression.forward`` (modules/networks_fast.py:79-95,134-141): 4x [ConvBlock(ELU) ->
    nearest x2 -> cat skip -> ConvBlock(ELU)], heads 1x1 C->128-ELU-128-ELU-1."""
    def key(s):
        return f"{prefix}{s}"

    def conv_block(x, p):
        x = F.elu(F.conv2d(x, w[key(f"{p}.conv1.weight")], w[key(f"{p}.conv1.bias")], padding=1))
        return F.elu(F.conv2d(x, w[key(f"{p}.conv2.weight")], w[key(f"{p}.conv2.bias")], padding=1))

    out = {}
    x = feats[-1]
    for n in range(1, 5):
        x = conv_block(x, f"block{n}.pre_concat_conv")
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        x = torch.cat([x, feats[-1 - n]], 1)
        x = conv_block(x, f"block{n}.post_concat_conv")
        s = 4 - n
        out[f"feature_s{s}_b1hw"] = x
        h = F.elu(F.conv2d(x, w[key(f"out{n}.0.weight")], w[key(f"out{n}.0.bias")]))
        h = F.elu(F.conv2d(h, w[key(f"out{n}.2.weight")], w[key(f"out{n}.2.bias")]))
        out[f"log_depth_pred_s{s}_b1hw"] = F.conv2d(h, w[key(f"out{n}.4.weight")], w[key(f"out{n}.4.bias")])
    return out


# --------------------------------------------------------------------------------------------------------
# model forward (hot-path body)
# --------------------------------------------------------------------------------------------------------


def relative_poses(cur_data, src_data):
    """experiment_modules/doubletake_model.py:341-349."""
    src_cam_T_cur_cam = src_data["cam_T_world_b44"] @ cur_data["world_T_cam_b44"].unsqueeze(1)
    cur_cam_T_src_cam = cur_data["cam_T_world_b44"].unsqueeze(1) @ src_data["world_T_cam_b44"]
    return src_cam_T_cur_cam, cur_cam_T_src_cam


def depth_model_forward(matching_cur_feats, matching_src_feats, prior_feats, cur_data, src_data, weights,
                        num_planes, decoder="unet_pp", hint=True, min_depth=0.25, max_depth=5.0,
                        matching_scale=1, volume="mlp", mask_mode="slow_hint"):
    """Body of ``DepthModelCVHint.forward`` / ``DepthModel.forward`` between the encoders and the return
    (experiment_modules/doubletake_model.py:341-349,374-423; sr_depth_model.py:351-433) with the two image
    encoders replaced by their outputs (matching features + 5 prior maps), which sit upstream of the boundary.
    ``weights`` uses the full-model key prefixes ``cost_volume.``, ``cost_volume_net.``, ``depth_decoder.``."""
    ext, poses = relative_poses(cur_data, src_data)
    src_K = src_data[f"K_s{matching_scale}_b44"]
    cur_invK = cur_data[f"invK_s{matching_scale}_b44"]
    cvw = {k[len("cost_volume."):]: v for k, v in weights.items() if k.startswith("cost_volume.")}
    if volume == "dot":
        cv = cost_volume_dot(matching_cur_feats, matching_src_feats, ext, poses, src_K, cur_invK, min_depth,
                             max_depth, num_planes)
        cv["mask"] = None
    else:
        cv = feature_volume(matching_cur_feats, matching_src_feats, ext, poses, src_K, cur_invK, min_depth,
                            max_depth, num_planes, cvw, hint=cur_data if hint else None, mask_mode=mask_mode)
    encw = {k[len("cost_volume_net."):]: v for k, v in weights.items() if k.startswith("cost_volume_net.")}
    decw = {k[len("depth_decoder."):]: v for k, v in weights.items() if k.startswith("depth_decoder.")}
    cv_feats = cv_encoder(cv["volume"], prior_feats[matching_scale:], encw)
    feats = list(prior_feats[:matching_scale]) + cv_feats
    if decoder == "unet_pp":
        out = depth_decoder_pp(feats, decw)
    else:
        out = skip_decoder_regression(feats, decw)
    for k in list(out.keys()):
        out[k] = out[k].float()
        out[k.replace("log_", "")] = torch.exp(out[k])  # doubletake_model.py:410-418
    out["lowest_cost_bhw"] = cv["lowest_cost"]
    out["overall_mask_bhw"] = cv["mask"]
    out["cost_volume"] = cv["volume"]
    out["lowest_cost_index"] = cv["index"]
    return out
