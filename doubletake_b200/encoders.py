"""B200 mirror of the reference's matching-feature encoder.

``ResnetMatchingEncoder`` (reference modules/networks.py:138-189): ResNet-18 stem + layer1 (antialiased_cnns or torchvision
flavour) -> 1x1 conv -> InstanceNorm -> LeakyReLU(0.2) -> 3x3 conv with replicate padding -> InstanceNorm.  Same constructor
arguments, ``forward(image) -> (N, C, H/4, W/4)`` and ``state_dict`` keys (``net.0``, ``net.1``, ``net.3.1.filt``,
``net.4.{0,1}.*``, ``net.5``, ``net.8``), so reference checkpoints load unchanged.  The module only HOLDS parameters; the
forward runs hand-written CUDA through the C ABI: stem conv (BatchNorm folded), blur / max pooling, the fused conv
descriptors for layer1 and the two convs, two InstanceNorm launches.  The replicate-padded 3x3 conv is a zero-padded conv on a
map that the first InstanceNorm writes with a replicated 1-pixel border; the second InstanceNorm crops it again.

``forward_views`` returns the features in the layouts the cost-volume kernels read -- current view NCHW, source views
channels-last (presented as a permuted ``(B, K, C, H, W)`` view, which the managers use without a transpose) -- so the 8.6 MB
per-frame source transpose of round 1 disappears when the encoder runs here.  Inference only (BatchNorm uses its running
statistics); no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from .networks import ConvPlan


class _BlurPool(nn.Module):
    """Holds antialiased_cnns.BlurPool's ``filt`` buffer ((C,1,4,4) binomial / 64) for checkpoint compatibility; the kernel
    regenerates the filter."""

    def __init__(self, channels, filt_size=4):
        super().__init__()
        a = torch.tensor([1.0, 3.0, 3.0, 1.0])
        filt = a[:, None] * a[None, :]
        self.register_buffer("filt", (filt / filt.sum())[None, None].repeat(channels, 1, 1, 1))


class _ResBlock(nn.Module):
    """torchvision / antialiased_cnns BasicBlock (stride 1): keys conv1, bn1, conv2, bn2."""

    def __init__(self, planes):
        super().__init__()
        self.conv1 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)


def _fold_bn(conv_w, bn):
    """Inference BatchNorm folded into the preceding bias-free conv: w' = w * g / sqrt(var + eps), b' = beta - mean * g / sqrt(...)."""
    scale = bn.weight.detach() / torch.sqrt(bn.running_var.detach() + bn.eps)
    return conv_w.detach() * scale.view(-1, 1, 1, 1), bn.bias.detach() - bn.running_mean.detach() * scale


def _conv_holder(weight, bias, k, pad):
    m = nn.Conv2d(weight.shape[1], weight.shape[0], k, padding=pad, bias=True)
    m.weight = nn.Parameter(weight.contiguous(), requires_grad=False)
    m.bias = nn.Parameter(bias.contiguous(), requires_grad=False)
    return m


class ResnetMatchingEncoder(nn.Module):
    """reference modules/networks.py:138-189 (18- or 34-layer variants share stem and layer1 width; only 18 is built here)."""

    def __init__(self, num_layers=18, num_ch_out=16, pretrained=False, antialiased=True, math="exact"):
        super().__init__()
        if num_layers != 18:
            raise ValueError("doubletake_b200.ResnetMatchingEncoder implements the reference's 18-layer matching encoder")
        if pretrained:
            raise ValueError("no pretrained weights in this build: load a reference checkpoint's state_dict instead")
        if num_ch_out % 4 != 0 or num_ch_out > 64:
            raise ValueError("num_ch_out must be a multiple of 4 and at most 64")
        self.num_ch_enc = np.array([64, 64])
        self.num_ch_out = num_ch_out
        self.antialiased = antialiased
        self.math = math
        pool = nn.Sequential(nn.MaxPool2d(2, 1), _BlurPool(64)) if antialiased else nn.MaxPool2d(3, 2, 1)
        self.net = nn.Sequential(
            nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True), pool,
            nn.Sequential(_ResBlock(64), _ResBlock(64)),
            nn.Conv2d(64, 128, (1, 1)), nn.InstanceNorm2d(128), nn.LeakyReLU(0.2, True),
            nn.Conv2d(128, num_ch_out, (3, 3), padding=1, padding_mode="replicate"), nn.InstanceNorm2d(num_ch_out))
        self._plans = {}
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._plans.clear())

    def _apply(self, fn, recurse=True):
        self._plans = {}
        return super()._apply(fn, recurse)

    def _versions(self):
        return tuple(t._version for t in list(self.parameters()) + list(self.buffers()))

    def _build(self, N, H, W, dev):
        if H % 4 or W % 4:
            raise ValueError("image height and width must be multiples of 4")
        net = self.net
        h2, w2, h4, w4 = H // 2, W // 2, H // 4, W // 4
        st = {"versions": self._versions()}
        w0, b0 = _fold_bn(net[0].weight, net[1])
        st["stem_w"], st["stem_b"] = L.f32(w0, dev), L.f32(b0, dev)
        st["stem_pack"] = torch.empty(147 * 64, dtype=torch.float32, device=dev)
        st["stem_out"] = torch.empty((N, h2, w2, 64), dtype=torch.float32, device=dev)
        # plan A: layer1 (two residual blocks, BatchNorm folded, ReLU = LeakyReLU with slope 0) and the 1x1 conv
        pa = ConvPlan(dev, self.math)
        x = pa.new(N, h4, w4, 64)
        st["pool_out"] = x
        holders = []
        for blk in net[4]:
            w1, b1 = _fold_bn(blk.conv1.weight, blk.bn1)
            w2_, b2 = _fold_bn(blk.conv2.weight, blk.bn2)
            c1, c2 = _conv_holder(w1.to(dev), b1.to(dev), 3, 1), _conv_holder(w2_.to(dev), b2.to(dev), 3, 1)
            holders += [c1, c2]
            t = pa.conv([(x, L.RESAMPLE_NONE)], c1, L.ACT_LEAKY, 0.0)
            x = pa.conv([(t, L.RESAMPLE_NONE)], c2, L.ACT_LEAKY, 0.0, residual=x)
        c3 = _conv_holder(net[5].weight.detach().to(dev), net[5].bias.detach().to(dev), 1, 0)
        holders.append(c3)
        f128 = pa.conv([(x, L.RESAMPLE_NONE)], c3)
        pa.finalize()
        # plan B: the replicate-padded 3x3 conv as a zero-padded conv on the map enlarged by its replicated border;
        # output channels padded to 64 (the tensor-core / SIMT tiles are 64 channels wide)
        pb = ConvPlan(dev, self.math)
        padded = pb.new(N, h4 + 2, w4 + 2, 128)
        w8 = torch.zeros((64, 128, 3, 3), dtype=torch.float32, device=dev)
        b8 = torch.zeros(64, dtype=torch.float32, device=dev)
        w8[: self.num_ch_out], b8[: self.num_ch_out] = net[8].weight.detach().to(dev), net[8].bias.detach().to(dev)
        c8 = _conv_holder(w8, b8, 3, 1)
        holders.append(c8)
        out64 = pb.conv([(padded, L.RESAMPLE_NONE)], c8)
        pb.finalize()
        st.update(plan_a=pa, plan_b=pb, f128=f128, padded=padded, out64=out64, holders=holders,
                  stats1=torch.empty(N * 128 * 2, dtype=torch.float32, device=dev),
                  stats2=torch.empty(N * self.num_ch_out * 2, dtype=torch.float32, device=dev))
        return st

    @staticmethod
    def _layout(f):
        return L.LAYOUT_SPLIT16 if f.fmt == "split16" else L.LAYOUT_F32

    def _run(self, images):
        """(N,3,H,W) fp32 CUDA -> ((N,C,h,w) NCHW, (N,h,w,C) channels-last), both fp32, freshly allocated."""
        if not images.is_cuda:
            raise RuntimeError("doubletake_b200 encoders run on CUDA only (no CPU fallback)")
        images = L.f32(images)
        N, _, H, W = images.shape
        dev = images.device
        key = (N, H, W, str(dev), self.math)
        st = self._plans.get(key)
        if st is None or st["versions"] != self._versions():
            st = self._plans[key] = self._build(N, H, W, dev)
        lib, s = L.lib(), L.stream()
        h4, w4 = H // 4, W // 4
        L.check(lib.dtb200_encoder_stem(L.ptr(images), L.ptr(st["stem_w"]), L.ptr(st["stem_b"]), L.ptr(st["stem_pack"]),
                                        L.ptr(st["stem_out"]), N, H, W, s))
        pool = st["pool_out"]
        L.check(lib.dtb200_encoder_pool(L.ptr(st["stem_out"]), L.ptr(pool.t), self._layout(pool), N, H // 2, W // 2, 64,
                                        0 if self.antialiased else 1, s))
        st["plan_a"].run()
        p = L.InstanceNormParams()
        f128, padded = st["f128"], st["padded"]
        p.src, p.src_layout, p.src_channels, p.src_border = L.ptr(f128.t), self._layout(f128), 128, 0
        p.batch, p.height, p.width, p.channels = N, h4, w4, 128
        p.eps, p.act, p.act_slope = 1e-5, L.ACT_LEAKY, 0.2
        p.dst, p.dst_layout, p.dst_border = L.ptr(padded.t), self._layout(padded), 1
        p.dst_nchw, p.stats = None, L.ptr(st["stats1"])
        L.check(lib.dtb200_instance_norm(C.byref(p), s))
        st["plan_b"].run()
        out64 = st["out64"]
        nchw = torch.empty((N, self.num_ch_out, h4, w4), dtype=torch.float32, device=dev)
        nhwc = torch.empty((N, h4, w4, self.num_ch_out), dtype=torch.float32, device=dev)
        q = L.InstanceNormParams()
        q.src, q.src_layout, q.src_channels, q.src_border = L.ptr(out64.t), self._layout(out64), 64, 1
        q.batch, q.height, q.width, q.channels = N, h4, w4, self.num_ch_out
        q.eps, q.act, q.act_slope = 1e-5, L.ACT_NONE, 0.0
        q.dst, q.dst_layout, q.dst_border = L.ptr(nhwc), L.LAYOUT_F32, 0
        q.dst_nchw, q.stats = L.ptr(nchw), L.ptr(st["stats2"])
        L.check(lib.dtb200_instance_norm(C.byref(q), s))
        return nchw, nhwc

    def forward(self, input_image):
        return self._run(input_image)[0]

    def forward_views(self, cur_image_b3hw, src_image_bk3hw):
        """All views of a batch in one pass -> (cur (B,C,h,w) NCHW, src (B,K,C,h,w) as a permuted view of channels-last
        storage): the layouts ``CostVolumeManager._run`` consumes without staging copies."""
        B, K = src_image_bk3hw.shape[:2]
        allv = torch.cat([cur_image_b3hw[:, None], src_image_bk3hw], 1).flatten(0, 1)
        nchw, nhwc = self._run(allv)
        nchw = nchw.view(B, K + 1, *nchw.shape[1:])
        nhwc = nhwc.view(B, K + 1, *nhwc.shape[1:])
        cur = nchw[:, 0].contiguous()
        src = nhwc[:, 1:].contiguous().permute(0, 1, 4, 2, 3)
        return cur, src
