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
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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


def test_tsdf_mirror_has_the_reference_signatures():
    """doubletake_b200.tsdf against the REAL reference classes (tools/tsdf.py, imported with the stubs of oracle/ref_stubs):
    every mirrored method takes the reference's parameters, in the reference's order, with the reference's defaults."""
    import inspect
    import types

    if not os.path.exists("/root/reference/src/doubletake/tools/tsdf.py"):
        pytest.skip("reference tree not present (GPU box)")
    for pth in (os.path.join(ROOT, "oracle", "ref_stubs"), "/root/reference/src"):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    if "doubletake.utils.pytorch3d_extras" not in sys.modules:
        m = types.ModuleType("doubletake.utils.pytorch3d_extras")
        m.marching_cubes = None
        sys.modules["doubletake.utils.pytorch3d_extras"] = m
    from doubletake.tools import tsdf as ref
    from doubletake_b200 import tsdf as ours

    def params(fn):
        return [(n, p.default) for n, p in inspect.signature(fn).parameters.items()]

    for cls, names in (("TSDF", ["__init__", "from_file", "from_mesh", "generate_voxel_coords", "cuda", "cpu", "save_tsdf",
                                 "sample_tsdf"]),
                       ("TSDFFuser", ["integrate_depth"])):
        for name in names:
            want = params(getattr(getattr(ref, cls), name))
            got = params(getattr(getattr(ours, cls), name))
            assert got[: len(want)] == want, (cls, name, got, want)  # extensions may only be appended, with defaults
            assert all(d is not inspect.Parameter.empty for _, d in got[len(want):]), (cls, name)
    # from_bounds / TSDFFuser.__init__: same leading parameters, one appended keyword each (lazy_grid, semantics)
    assert params(ours.TSDF.from_bounds)[:2] == params(ref.TSDF.from_bounds)
    assert params(ours.TSDFFuser.__init__)[:5] == params(ref.TSDFFuser.__init__)
    assert params(ours.get_frustum_bounds) == params(ref.get_frustum_bounds)
    assert ours.TSDF.VOX_MOD == ref.TSDF.VOX_MOD
    f = ours.TSDFFuser(ours.TSDF.from_bounds(dict(xmin=0, xmax=0.4, ymin=0, ymax=0.4, zmin=0, zmax=0.4), 0.05))
    rf = ref.TSDFFuser(ref.TSDF.from_bounds(dict(xmin=0, xmax=0.4, ymin=0, ymax=0.4, zmin=0, zmax=0.4), 0.05), use_gpu=False)
    assert (f.truncation, f.maxW, f.min_depth, f.max_depth, tuple(f.shape)) == \
           (rf.truncation, rf.maxW, rf.min_depth, rf.max_depth, tuple(rf.shape))


@pytest.mark.parametrize("antialiased", [True, False])
def test_matching_encoder_mirror_has_the_reference_state_dict(antialiased):
    """doubletake_b200.ResnetMatchingEncoder vs the reference class (modules/networks.py:138-189; antialiased_cnns is the
    restated stub): identical state_dict keys and shapes, strict round trip of the weights."""
    import doubletake_b200 as dt
    _ref_modules()
    from doubletake.modules.networks import ResnetMatchingEncoder as Ref

    with contextlib.redirect_stdout(io.StringIO()):
        ref = Ref(18, 16, pretrained=False, antialiased=antialiased)
    ours = dt.ResnetMatchingEncoder(18, 16, antialiased=antialiased)
    rs, os_ = ref.state_dict(), ours.state_dict()
    assert {k: tuple(v.shape) for k, v in rs.items()} == {k: tuple(v.shape) for k, v in os_.items()}
    ours.load_state_dict(rs, strict=True)
    assert all(torch.equal(rs[k], ours.state_dict()[k]) for k in rs)
    assert list(ours.num_ch_enc) == list(ref.num_ch_enc) and ours.num_ch_out == ref.num_ch_out
