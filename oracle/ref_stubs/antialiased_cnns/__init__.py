"""Restatement of the part of the third-party package ``antialiased_cnns`` (Zhang, "Making Convolutional Networks
Shift-Invariant Again", ICML 2019; PyPI antialiased-cnns, un-pinned in the reference's environment.yml:29, latest 0.3)
that the reference's ``ResnetMatchingEncoder`` (modules/networks.py:138-189) uses: ``resnet18(pretrained, filter_size=4,
pool_only=True)`` up to ``layer1`` -- a torchvision ResNet-18 whose ``maxpool`` is ``MaxPool2d(2, stride 1)`` followed by
``BlurPool(64, filt_size=4, stride=2)`` (reflection padding (1, 2, 1, 2), binomial [1, 3, 3, 1] x [1, 3, 3, 1] / 64 depthwise
filter).  TEST INFRASTRUCTURE ONLY: the package is absent from this image and cannot be installed (no network), so
oracle/make_golden_encoder.py executes the reference class on top of this restatement; the fixture it produces is
"reference code over a restated dependency" -- parity UNPINNED for the BlurPool step itself, pinned for everything else
(the torchvision variant, ``antialiased=False``, is executed against the real torchvision).
The strided stages (layer2..4), which the matching encoder never touches, are left as torchvision's.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class BlurPool(nn.Module):
    def __init__(self, channels, pad_type="reflect", filt_size=4, stride=2, pad_off=0):
        super().__init__()
        self.filt_size, self.stride, self.channels = filt_size, stride, channels
        lo, hi = int(1.0 * (filt_size - 1) / 2), int(np.ceil(1.0 * (filt_size - 1) / 2))
        self.pad_sizes = [lo + pad_off, hi + pad_off, lo + pad_off, hi + pad_off]
        a = {1: [1.0], 2: [1.0, 1.0], 3: [1.0, 2.0, 1.0], 4: [1.0, 3.0, 3.0, 1.0], 5: [1.0, 4.0, 6.0, 4.0, 1.0]}[filt_size]
        a = np.array(a)
        filt = torch.Tensor(a[:, None] * a[None, :])
        filt = filt / torch.sum(filt)
        self.register_buffer("filt", filt[None, None, :, :].repeat((channels, 1, 1, 1)))
        self.pad = {"reflect": nn.ReflectionPad2d, "replicate": nn.ReplicationPad2d, "zero": nn.ZeroPad2d}[pad_type](self.pad_sizes)

    def forward(self, inp):
        return F.conv2d(self.pad(inp), self.filt, stride=self.stride, groups=inp.shape[1])


def _resnet(fn, pretrained, filter_size, pool_only):
    if pretrained:
        raise RuntimeError("restated antialiased_cnns: no pretrained weights in this image (no network)")
    if not pool_only:
        raise NotImplementedError("only pool_only=True (the package default) is restated")
    net = fn(weights=None)
    net.maxpool = nn.Sequential(nn.MaxPool2d(kernel_size=2, stride=1), BlurPool(64, filt_size=filter_size, stride=2))
    return net


def resnet18(pretrained=False, filter_size=4, pool_only=True, **kwargs):
    from torchvision.models import resnet18 as tv

    return _resnet(tv, pretrained, filter_size, pool_only)


def resnet34(pretrained=False, filter_size=4, pool_only=True, **kwargs):
    from torchvision.models import resnet34 as tv

    return _resnet(tv, pretrained, filter_size, pool_only)


def _unsupported(*a, **k):
    raise NotImplementedError("restated antialiased_cnns: only resnet18 / resnet34 up to layer1")


resnet50 = resnet101 = resnet152 = _unsupported
