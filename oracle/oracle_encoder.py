"""CPU oracle of the matching-feature encoder -- TEST INFRASTRUCTURE ONLY, never imported by
``doubletake_b200/``.

Functional restatement of reference ``ResnetMatchingEncoder.forward`` (modules/networks.py:138-189) from a ``state_dict``
with the reference's key names (``net.0`` conv1, ``net.1`` bn1, ``net.4.{0,1}.{conv1,bn1,conv2,bn2}`` layer1, ``net.5`` 1x1
conv, ``net.8`` 3x3 replicate-padded conv; the two InstanceNorm2d carry no parameters):

    conv 7x7 / 2 (3 -> 64, no bias) -> BatchNorm (eval) -> ReLU
    -> pool: torchvision  MaxPool2d(3, stride 2, padding 1)                                (``antialiased=False``)
             antialiased  MaxPool2d(2, stride 1) -> BlurPool(4, stride 2, reflect (1,2,1,2))  (``antialiased=True``, default)
    -> 2 x BasicBlock(64): conv3x3 -> BN -> ReLU -> conv3x3 -> BN -> + x -> ReLU            (torchvision / antialiased_cnns)
    -> conv 1x1 (64 -> 128) -> InstanceNorm2d(128) -> LeakyReLU(0.2)
    -> conv 3x3 (128 -> C, replicate padding) -> InstanceNorm2d(C)

Pinned on fixtures produced by executing the reference class itself (oracle/make_golden_encoder.py): against the real
torchvision for the ``antialiased=False`` variant, against a restatement of the absent third-party ``antialiased_cnns``
(oracle/ref_stubs/antialiased_cnns) for the default variant.
"""
import torch
import torch.nn.functional as F


def _bn(x, w, p):
    return F.batch_norm(x, w[p + ".running_mean"], w[p + ".running_var"], w[p + ".weight"], w[p + ".bias"], False, 0.1, 1e-5)


def blur_pool(x, filt_size=4, stride=2):
    a = torch.tensor({3: [1.0, 2.0, 1.0], 4: [1.0, 3.0, 3.0, 1.0]}[filt_size])
    filt = a[:, None] * a[None, :]
    filt = (filt / filt.sum())[None, None].repeat(x.shape[1], 1, 1, 1).to(x)
    lo, hi = (filt_size - 1) // 2, -(-(filt_size - 1) // 2)
    return F.conv2d(F.pad(x, (lo, hi, lo, hi), mode="reflect"), filt, stride=stride, groups=x.shape[1])


def matching_encoder(image_b3hw, w, antialiased=True):
    x = F.conv2d(image_b3hw, w["net.0.weight"], None, stride=2, padding=3)
    x = F.relu(_bn(x, w, "net.1"))
    if antialiased:
        x = blur_pool(F.max_pool2d(x, 2, 1))
    else:
        x = F.max_pool2d(x, 3, 2, 1)
    for blk in ("net.4.0", "net.4.1"):
        y = F.relu(_bn(F.conv2d(x, w[blk + ".conv1.weight"], None, padding=1), w, blk + ".bn1"))
        y = _bn(F.conv2d(y, w[blk + ".conv2.weight"], None, padding=1), w, blk + ".bn2")
        x = F.relu(y + x)
    x = F.conv2d(x, w["net.5.weight"], w["net.5.bias"])
    x = F.leaky_relu(F.instance_norm(x, eps=1e-5), 0.2)
    x = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), w["net.8.weight"], w["net.8.bias"])
    return F.instance_norm(x, eps=1e-5)


def encoder_state(seed, num_ch_out=16):
    """Seeded parameters AND buffers with the reference's key names: conv / BN affine weights like
    ``synthetic.seeded_state_dict``, BN running statistics non-trivial (mean U(-0.5, 0.5), var U(0.5, 1.5))."""
    from doubletake_b200 import synthetic as syn

    shapes = {"net.0.weight": (64, 3, 7, 7), "net.1.weight": (64,), "net.1.bias": (64,),
              "net.5.weight": (128, 64, 1, 1), "net.5.bias": (128,), "net.8.weight": (num_ch_out, 128, 3, 3),
              "net.8.bias": (num_ch_out,)}
    for b in ("net.4.0", "net.4.1"):
        for c in ("conv1", "conv2"):
            shapes[f"{b}.{c}.weight"] = (64, 64, 3, 3)
        for n in ("bn1", "bn2"):
            shapes[f"{b}.{n}.weight"] = (64,)
            shapes[f"{b}.{n}.bias"] = (64,)
    sd = syn.seeded_state_dict(shapes, seed, 1.5)
    g = torch.Generator().manual_seed(seed + 1)
    for k in list(sd):
        if k.endswith(".weight") and sd[k].ndim == 1:   # BN gamma around 1
            sd[k] = 1.0 + 0.5 * sd[k] * 8.0
            base = k[: -len(".weight")]
            sd[base + ".running_mean"] = torch.rand(64, generator=g) - 0.5
            sd[base + ".running_var"] = torch.rand(64, generator=g) + 0.5
    return sd
