"""The plug-in idiom against the REAL reference modules (build container only: skipped where /root/reference is absent,
e.g. on the GPU box).  CPU-only: checks that `install` / `to_b200` rebuild the reference's modules as B200 modules with
identical parameters, ready to be called from the reference's own forward."""
import contextlib
import io
import os
import sys

import pytest
import torch

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def _ref_modules():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, "oracle", "ref_stubs"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    from doubletake.modules.mesh_hint_volume import FeatureMeshHintVolumeManager
    from doubletake.modules.networks import CVEncoder, DepthDecoderPP
    from doubletake.modules.networks_fast import SkipDecoderRegression
    return FeatureMeshHintVolumeManager, CVEncoder, DepthDecoderPP, SkipDecoderRegression


@pytest.mark.parametrize("decoder", ["unet_pp", "skip"])
def test_install_swaps_modules_and_keeps_weights(decoder):
    import doubletake_b200 as dt
    FMH, CVE, PP, Skip = _ref_modules()

    class Holder(torch.nn.Module):  # stands in for the LightningModule: only the three hot-path attributes matter
        pass

    m = Holder()
    with contextlib.redirect_stdout(io.StringIO()):
        m.cost_volume = FMH(24, 32, num_depth_bins=16, mlp_channels=[0, 128, 128, 1], matching_dim_size=16, num_source_views=3)
    prior = [24, 48, 64, 160, 256] if decoder == "unet_pp" else [64, 64, 128, 256, 512]
    m.cost_volume_net = CVE(num_ch_cv=16, num_ch_enc=prior[1:], num_ch_outs=[64, 128, 256, 384])
    dec_in = prior[:1] + m.cost_volume_net.num_ch_enc
    m.depth_decoder = PP(dec_in) if decoder == "unet_pp" else Skip(dec_in)
    want = {k: v.clone() for k, v in m.state_dict().items()}
    dt.install(m, math="exact")
    assert isinstance(m.cost_volume, dt.FeatureMeshHintVolumeManager)
    assert isinstance(m.cost_volume_net, dt.CVEncoder)
    assert isinstance(m.depth_decoder, dt.DepthDecoderPP if decoder == "unet_pp" else dt.SkipDecoderRegression)
    got = m.state_dict()
    assert set(got) == set(want)
    for k in want:
        assert torch.equal(got[k], want[k]), k
    # strict loading of a reference checkpoint into the swapped model works too
    m.load_state_dict(want, strict=True)
