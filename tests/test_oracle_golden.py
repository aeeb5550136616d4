"""Pin the CPU oracle (oracle/oracle_torch.py) against fixtures produced by executing the real reference
(oracle/make_golden.py).  CPU only.  Tolerances: cost volume 1e-5 relative, arg-max plane
identical (near-ties must be proven), depth 1e-4 relative (north_star)."""
import json
import sys
import os

import numpy as np
import pytest
import torch

import helpers as hp
from doubletake_b200 import synthetic as syn

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle_torch as orc  # noqa: E402

torch.set_grad_enabled(False)

VOLUME_CASES = [
    ("cv_dot_small", "dot"), ("cv_dot_portrait_white", "dot"), ("fv_mlp_small", "mlp"),
    ("fv_hint_small", "hint"), ("fv_hint_empty", "hint"),
    ("cfg1_dot", "dot"), ("cfg1_mlp", "mlp"), ("cfg1_hint", "hint"),
]


@pytest.mark.parametrize("name,kind", VOLUME_CASES)
def test_volume_matches_reference(name, kind):
    fx = hp.load(name)
    inp, weights, m = hp.volume_case_inputs(fx, kind)
    args = (inp["cur_feats"], inp["src_feats"], inp["src_extrinsics"], inp["src_poses"], inp["src_Ks"],
            inp["cur_invK"], inp["min_depth"], inp["max_depth"], m["D"])
    if kind == "dot":
        out = orc.cost_volume_dot(*args)
    else:
        out = orc.feature_volume(*args, weights, hint=inp.get("cv_depth_hint_dict"),
                                 mask_mode="slow_hint" if kind == "hint" else "fast")
    s = m["sub"]
    assert np.array_equal(out["planes"].numpy(), fx["out.planes"])
    assert hp.rel_err(out["volume"][:, :, ::s, ::s], fx["out.volume"]) < 1e-5
    n_bad, n_unexplained = hp.argmax_mismatch_report(out["volume"], fx["out.index"])
    assert n_unexplained == 0 and n_bad <= 2, (n_bad, n_unexplained)
    if n_bad == 0:
        assert np.array_equal(out["lowest_cost"].numpy(), fx["out.lowest_cost"])
    if kind != "dot":
        assert np.array_equal(out["mask"].numpy(), fx["out.mask"])
        fast = orc.feature_volume(*args, weights, hint=inp.get("cv_depth_hint_dict"), mask_mode="fast")
        assert np.array_equal(fast["mask"].numpy(), fx["out.fast_mask"])


def _weights(fx, which, seed, scale):
    shapes = json.loads(str(fx[which]))
    return syn.seeded_state_dict(shapes, seed, scale)


@pytest.mark.parametrize("name,decoder", [("net_pp_d64", "unet_pp"), ("net_pp_d16_b2", "unet_pp"),
                                           ("net_skip_d48", "skip")])
def test_conv_stacks_match_reference(name, decoder):
    fx = hp.load(name)
    cfg, cv, priors, seed = hp.network_case_inputs(fx, decoder)
    encw = _weights(fx, "enc_shapes", seed + 1, 1.5)
    decw = _weights(fx, "dec_shapes", seed + 2, 1.5)
    cvf = orc.cv_encoder(cv, priors[1:], encw)
    for i, f in enumerate(cvf):
        assert hp.rel_err(f, fx[f"out.cv_feat_{i}"]) < 1e-5
    feats = priors[:1] + cvf
    out = orc.depth_decoder_pp(feats, decw) if decoder == "unet_pp" else orc.skip_decoder_regression(feats, decw)
    for i in range(4):
        k = f"log_depth_pred_s{i}_b1hw"
        assert float((out[k] - torch.from_numpy(fx["out." + k])).abs().max()) < 1e-4, k


@pytest.mark.parametrize("name,cfgname,empty", [("model_tiny_pp", "tiny", False), ("model_tiny_skip", "tiny_small", False),
                                                ("model_tiny_pp_emptyhint", "tiny", True)])
def test_model_forward_matches_reference(name, cfgname, empty):
    fx = hp.load(name)
    cfg = syn.CONFIGS[cfgname]
    inp = syn.cost_volume_inputs(cfg, empty_hint=empty)
    priors = syn.prior_features(cfg)
    cur_data = {k[7:]: torch.from_numpy(v) for k, v in fx.items() if k.startswith("in.cur.")}
    src_data = {k[7:]: torch.from_numpy(v) for k, v in fx.items() if k.startswith("in.src.")}
    cur_data.update(inp["cv_depth_hint_dict"])
    cur_data["invK_s1_b44"] = inp["cur_invK"]
    src_data["K_s1_b44"] = inp["src_Ks"]
    w = {}
    for pre, which, off, sc in (("cost_volume.", "cv_shapes", 10, 1.0), ("cost_volume_net.", "enc_shapes", 11, 1.5),
                                ("depth_decoder.", "dec_shapes", 12, 1.5)):
        for k, v in _weights(fx, which, cfg.seed + off, sc).items():
            w[pre + k] = v
    out = orc.depth_model_forward(inp["cur_feats"], inp["src_feats"], priors, cur_data, src_data, w, cfg.planes,
                                  decoder=cfg.decoder, hint=True)
    assert hp.rel_err(out["cost_volume"], fx["out.cost_volume"]) < 1e-5
    n_bad, n_unexplained = hp.argmax_mismatch_report(out["cost_volume"], fx["out.index"])
    assert n_unexplained == 0 and n_bad <= 2
    assert np.array_equal(out["overall_mask_bhw"].numpy(), fx["out.overall_mask_bhw"])
    for i in range(4):
        k = f"depth_pred_s{i}_b1hw"
        ref = torch.from_numpy(fx["out." + k])
        assert float(((out[k] - ref).abs() / ref.abs()).max()) < 1e-4, k
