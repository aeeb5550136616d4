"""Host-side logic that needs no GPU: checkpoint-key compatibility with the reference, plane generation, plan
structure, sharding arithmetic, loud failure on CPU tensors."""
import json

import pytest
import torch

import helpers as hp
import doubletake_b200 as dt
from doubletake_b200 import sharding, synthetic as syn
from oracle import oracle_torch as orc


def _param_shapes(m):
    return {k: list(v.shape) for k, v in m.named_parameters()}


def test_network_state_dict_keys_match_reference():
    fx = hp.load("net_pp_d64")
    enc = dt.CVEncoder(64, [48, 64, 160, 256], [64, 128, 256, 384])
    dec = dt.DepthDecoderPP([24] + enc.num_ch_enc)
    assert _param_shapes(enc) == json.loads(str(fx["enc_shapes"]))
    assert _param_shapes(dec) == json.loads(str(fx["dec_shapes"]))
    fx = hp.load("net_skip_d48")
    enc = dt.CVEncoder(48, [64, 128, 256, 512], [64, 128, 256, 384])
    dec = dt.SkipDecoderRegression([64] + enc.num_ch_enc)
    assert _param_shapes(enc) == json.loads(str(fx["enc_shapes"]))
    assert _param_shapes(dec) == json.loads(str(fx["dec_shapes"]))


def test_manager_state_dict_keys_match_reference():
    fx = hp.load("model_tiny_pp")
    mgr = dt.FeatureMeshHintVolumeManager(48, 64, 16, matching_dim_size=16, num_source_views=2)
    assert _param_shapes(mgr) == json.loads(str(fx["cv_shapes"]))
    sd = mgr.state_dict()
    # buffers the reference registers exist so strict loading works
    for k in ("linear_ramp_1d11", "backprojector.pix_coords_13N", "projector.eps"):
        assert k in sd
    assert tuple(sd["backprojector.pix_coords_13N"].shape) == (1, 3, 48 * 64)
    assert dt.FeatureVolumeManager(8, 8, 4, num_source_views=7).mlp.net[0].weight.shape == (128, 202)
    with pytest.raises(ValueError):
        dt.FeatureVolumeManager(8, 8, 4, mlp_channels=[0, 64, 64, 1])


def test_depth_planes_match_oracle_bitwise():
    mgr = dt.CostVolumeManager(4, 6, num_depth_bins=64)
    planes = mgr.generate_depth_planes(2, torch.tensor(0.25).view(1, 1, 1, 1), torch.tensor(5.0).view(1, 1, 1, 1))
    assert planes.shape == (2, 64, 4, 6)
    assert torch.equal(planes[0, :, 0, 0], orc.depth_planes(0.25, 5.0, 64))


def test_cpu_tensors_fail_loudly():
    cfg = syn.WorkloadConfig("t", 1, 2, 0, 0, 4, hint=False, seed=1)
    inp = syn.cost_volume_inputs(cfg, match_hw=(8, 8))
    mgr = dt.CostVolumeManager(8, 8, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        mgr(**inp)
    enc = dt.CVEncoder(16, [48, 64, 160, 256], [64, 128, 256, 384])
    with pytest.raises(RuntimeError):
        enc(torch.zeros(1, 16, 32, 32), [torch.zeros(1, 48, 32, 32), torch.zeros(1, 64, 16, 16),
                                         torch.zeros(1, 160, 8, 8), torch.zeros(1, 256, 4, 4)])


def test_model_rejects_training_phase_and_bad_options():
    opts = dt.HotPathOptions(image_height=64, image_width=96, matching_num_depth_bins=16, model_num_views=3)
    model = dt.DepthModelCVHint(opts)
    with pytest.raises(NotImplementedError):
        model("train", {}, {})
    with pytest.raises(ValueError):
        dt.DepthModelCVHint(dt.HotPathOptions(depth_decoder_name="nope"))
    with pytest.raises(ValueError):
        dt.DepthModel(dt.HotPathOptions(feature_volume_type="mlp_mesh_hint_feature_volume"))
    assert isinstance(dt.DepthModel(dt.HotPathOptions(feature_volume_type="simple_cost_volume")).cost_volume,
                      dt.CostVolumeManager)


def test_shard_frames_round_robin_partition():
    for n, w in ((25590, 8), (7, 4), (3, 8), (16, 2)):
        seen = []
        for r in range(w):
            idx = sharding.shard_frames(n, r, w)
            assert all(i % w == r for i in idx)
            seen += idx
        assert sorted(seen) == list(range(n))


def test_shard_scans_balances_and_covers():
    lengths = [300, 10, 250, 40, 40, 500, 5, 120]
    owned = [sharding.shard_scans(lengths, r, 3) for r in range(3)]
    assert sorted(sum(owned, [])) == list(range(len(lengths)))
    loads = [sum(lengths[i] for i in o) for o in owned]
    assert max(loads) - min(loads) <= max(lengths)
