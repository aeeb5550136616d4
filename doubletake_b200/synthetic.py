"""Deterministic synthetic inputs for the plane-sweep hot path.

Everything is generated on the CPU with a seeded ``torch.Generator`` so the reference (in the build
container), the oracle and the CUDA path all see identical tensors.  Shapes follow the contract of
``DepthModelCVHint.forward`` (reference ``experiment_modules/doubletake_model.py:330-349``) and
``FeatureMeshHintVolumeManager.build_cost_volume`` (``modules/mesh_hint_volume.py:84-96``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F


@dataclass
class WorkloadConfig:
    """One row of SURVEY.md §8's config table."""

    name: str
    batch: int
    num_src: int
    image_h: int
    image_w: int
    planes: int
    feat_ch: int = 16
    hint: bool = True
    # image-prior channels at strides 2,4,8,16,32 (timm tf_efficientnetv2_s / resnet18d)
    prior_ch: tuple = (24, 48, 64, 160, 256)
    decoder: str = "unet_pp"  # or "skip"
    seed: int = 1000

    @property
    def match_h(self):
        return self.image_h // 4

    @property
    def match_w(self):
        return self.image_w // 4

    @property
    def mlp_in(self):
        return 26 * self.num_src + 20


CONFIGS = {
    # cfg 1: PR1 correctness reference
    "cfg1": WorkloadConfig("cfg1", 1, 2, 512, 640, 32, hint=False, seed=1001),
    # cfg 2: DoubleTake, 640x480 image, 64 planes, 7 src views, hint on  (the headline metric)
    "cfg2": WorkloadConfig("cfg2", 1, 7, 480, 640, 64, hint=True, seed=1002),
    # cfg 3: DoubleTake-small throughput
    "cfg3": WorkloadConfig(
        "cfg3", 8, 5, 384, 512, 48, hint=True, prior_ch=(64, 64, 128, 256, 512), decoder="skip", seed=1003
    ),
    # cfg 4: ScanNetv2 test-split shapes (the reference's default 512x384 input, options.py:69-70), DoubleTake, one keyframe per step
    "cfg4": WorkloadConfig("cfg4", 1, 7, 384, 512, 64, hint=True, seed=1004),
    # cfg 5: synthetic stress
    "cfg5": WorkloadConfig("cfg5", 4, 9, 768, 1024, 96, hint=True, seed=1005),
    # small variants for tests / golden fixtures
    "tiny": WorkloadConfig("tiny", 1, 2, 192, 256, 16, hint=True, seed=1100),
    "tiny_small": WorkloadConfig(
        "tiny_small", 2, 3, 128, 192, 48, hint=True, prior_ch=(64, 64, 128, 256, 512), decoder="skip", seed=1101
    ),
}


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def smooth_features(shape, g, white=False):
    """N(0,1) noise -> 5x5 box blur -> per-(sample, channel) instance norm (the matching encoder ends in
    InstanceNorm2d, reference ``modules/networks.py:184-185``).  ``white=True`` skips the blur: the
    adversarial arg-max case."""
    x = torch.randn(shape, generator=g, dtype=torch.float32)
    lead = x.shape[:-3]
    x = x.reshape(-1, 1, x.shape[-2], x.shape[-1])
    if not white:
        x = F.avg_pool2d(F.pad(x, (2, 2, 2, 2), mode="replicate"), 5, stride=1)
    mean = x.mean(dim=(-1, -2), keepdim=True)
    var = x.var(dim=(-1, -2), keepdim=True, unbiased=False)
    x = (x - mean) / torch.sqrt(var + 1e-5)
    return x.reshape(*lead, *shape[-3:]).contiguous()


def scannet_intrinsics(image_h, image_w, scale):
    """4x4 K at stride 2**(scale+1) of the image: ScanNet colour intrinsics scaled the way
    ``datasets/scannet_dataset.py:469-479`` does (K_s0 = depth res = image/2, each further scale halves)."""
    K = torch.eye(4, dtype=torch.float32)
    K[0, 0] = 570.92
    K[1, 1] = 570.92
    K[0, 2] = 319.5
    K[1, 2] = 239.5
    K[0] *= (image_w / 2) / 640.0
    K[1] *= (image_h / 2) / 480.0
    for _ in range(scale):
        K[:2] /= 2.0
    K[2, 2] = 1.0
    K[3, 3] = 1.0
    return K


def _axis_angle_to_R(axis, angle):
    axis = axis / axis.norm()
    x, y, z = axis.tolist()
    c, s = math.cos(angle), math.sin(angle)
    C = 1 - c
    return torch.tensor(
        [
            [c + x * x * C, x * y * C - z * s, x * z * C + y * s],
            [y * x * C + z * s, c + y * y * C, y * z * C - x * s],
            [z * x * C - y * s, z * y * C + x * s, c + z * z * C],
        ],
        dtype=torch.float64,
    )


def relative_poses(batch, num_src, g):
    """src_cam_T_cur_cam (B,K,4,4) in the DVMVS keyframe regime (``tools/keyframe_buffer.py:12-23``):
    rotation angle U(0,12deg), translation length U(0.05,0.30) m; sorted by pose distance ascending
    (``datasets/generic_mvs_dataset.py:730-738``).  Returns (src_extrinsics, src_poses) fp32."""
    ext = torch.zeros(batch, num_src, 4, 4, dtype=torch.float64)
    for b in range(batch):
        mats = []
        for _ in range(num_src):
            axis = torch.randn(3, generator=g, dtype=torch.float64)
            angle = float(torch.rand(1, generator=g)) * math.radians(12.0)
            tdir = torch.randn(3, generator=g, dtype=torch.float64)
            tdir = tdir / tdir.norm()
            tlen = 0.05 + 0.25 * float(torch.rand(1, generator=g))
            T = torch.eye(4, dtype=torch.float64)
            T[:3, :3] = _axis_angle_to_R(axis, angle)
            T[:3, 3] = tdir * tlen
            mats.append(T)

        def dist(T):
            tr = min(3.0, float(T[:3, :3].trace()))
            r = math.sqrt(max(0.0, 2 * (1 - tr / 3)))
            t = float(T[:3, 3].norm())
            return math.sqrt(t * t + r * r)

        mats.sort(key=dist)
        ext[b] = torch.stack(mats)
    poses = torch.linalg.inv(ext)
    return ext.float().contiguous(), poses.float().contiguous()


def depth_hint(batch, h, w, g, empty=False):
    """Rendered-depth hint at depth resolution (image/2): smooth depth field in [0.5,4.5] m, blobby validity
    mask (~70 % valid), NaN where invalid (``test_incremental.py:215-218``), TSDF confidence U(0,1) zeroed
    where invalid.  ``empty=True`` is the first-keyframe case (``test_incremental.py:260-269``)."""
    if empty:
        hint = torch.full((batch, 1, h, w), float("nan"), dtype=torch.float32)
        mask = torch.zeros(batch, 1, h, w, dtype=torch.float32)
        weights = torch.zeros(batch, 1, h, w, dtype=torch.float32)
    else:
        low = torch.rand(batch, 1, max(2, h // 16), max(2, w // 16), generator=g)
        hint = 0.5 + 4.0 * F.interpolate(low, size=(h, w), mode="bilinear", align_corners=True)
        noise = torch.rand(batch, 1, max(2, h // 8), max(2, w // 8), generator=g)
        mask = (F.interpolate(noise, size=(h, w), mode="bilinear", align_corners=True) > 0.38).float()
        weights = torch.rand(batch, 1, h, w, generator=g) * mask
        hint = torch.where(mask.bool(), hint, torch.full_like(hint, float("nan")))
    return {
        "depth_hint_b1hw": hint.contiguous(),
        "depth_hint_mask_b1hw": mask.contiguous(),
        "depth_hint_mask_b_b1hw": mask.bool().contiguous(),
        "sampled_weights_b1hw": weights.contiguous(),
    }


def cost_volume_inputs(cfg: WorkloadConfig, white=False, empty_hint=False, seed=None, match_hw=None):
    """Keyword arguments for ``*VolumeManager.forward`` / ``build_cost_volume`` at matching resolution."""
    g = _gen(cfg.seed if seed is None else seed)
    H, W = match_hw if match_hw is not None else (cfg.match_h, cfg.match_w)
    B, K, C = cfg.batch, cfg.num_src, cfg.feat_ch
    cur_feats = smooth_features((B, C, H, W), g, white)
    src_feats = smooth_features((B, K, C, H, W), g, white)
    ext, poses = relative_poses(B, K, g)
    K1 = scannet_intrinsics(H * 4, W * 4, 1)
    src_Ks = K1[None, None].expand(B, K, 4, 4).contiguous()
    cur_invK = torch.linalg.inv(K1.double()).float()[None].expand(B, 4, 4).contiguous()
    out = dict(
        cur_feats=cur_feats,
        src_feats=src_feats,
        src_extrinsics=ext,
        src_poses=poses,
        src_Ks=src_Ks,
        cur_invK=cur_invK,
        min_depth=torch.tensor(0.25).view(1, 1, 1, 1),
        max_depth=torch.tensor(5.0).view(1, 1, 1, 1),
    )
    if cfg.hint:
        out["cv_depth_hint_dict"] = depth_hint(B, 2 * H, 2 * W, g, empty=empty_hint)
    return out


def prior_features(cfg: WorkloadConfig, g=None):
    """Image-prior encoder outputs (the encoder itself is upstream of the boundary): N(0,1)*0.5 maps at strides
    2..32 with the channel layout of the timm backbone the config names."""
    g = g or _gen(cfg.seed + 7)
    feats = []
    for i, ch in enumerate(cfg.prior_ch):
        s = 2 ** (i + 1)
        feats.append(0.5 * torch.randn(cfg.batch, ch, cfg.image_h // s, cfg.image_w // s, generator=g))
    return feats


def seeded_state_dict(shapes: dict, seed: int, scale: float = 1.0):
    """Deterministic weights for a module given its ``{name: shape}`` map: each tensor is
    U(-b, b) with b = scale/sqrt(fan_in) (PyTorch's default Linear/Conv bound), drawn in sorted key order.
    Used so the reference (fixture generation), the oracle and the CUDA path load identical parameters without
    committing 116 MB of weights."""
    g = _gen(seed)
    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        if name.endswith("weight"):
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
        else:
            fan_in = max(1, shape[0])
        bound = scale / math.sqrt(fan_in)
        out[name] = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
    return out
