"""Generate tests/golden/*.npz by EXECUTING THE REAL REFERENCE modules on CPU (build container only).

Run:  python oracle/make_golden.py            (needs /root/reference; never runs on the GPU box)

The reference has no tests or golden vectors, so parity is pinned on the reference's own
modules executed here: ``CostVolumeManager``, ``FeatureVolumeManager``, ``FeatureMeshHintVolumeManager`` (+ their
``Fast*`` variants for the mask contract), ``CVEncoder``, ``DepthDecoderPP``, ``SkipDecoderRegression`` and a
composition that follows ``DepthModelCVHint.forward`` (experiment_modules/doubletake_model.py:341-349,374-423)
around them.  ``kornia`` / ``timm`` / ``antialiased_cnns`` are import-only stubs (oracle/ref_stubs): none of
them is called on this path.  Inputs come from ``doubletake_b200.synthetic`` (seeded, CPU) and are either
stored in the fixture or regenerated from the recorded seed; weights are regenerated from seeds
(``synthetic.seeded_state_dict``) so 116 MB of parameters never enter the repo.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_stubs"))
sys.path.insert(0, "/root/reference/src")

from doubletake.modules.cost_volume import CostVolumeManager  # noqa: E402
from doubletake.modules.feature_volume import FeatureVolumeManager  # noqa: E402
from doubletake.modules.mesh_hint_volume import FeatureMeshHintVolumeManager  # noqa: E402
from doubletake.modules.networks import CVEncoder, DepthDecoderPP  # noqa: E402
from doubletake.modules.networks_fast import SkipDecoderRegression  # noqa: E402

from doubletake_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_grad_enabled(False)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def np_(x):
    return x.detach().cpu().numpy()


def param_shapes(module):
    return {k: tuple(v.shape) for k, v in module.named_parameters()}


def load_seeded(module, seed, scale=1.0):
    sd = syn.seeded_state_dict(param_shapes(module), seed, scale)
    missing, unexpected = module.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    return sd


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


def flat_inputs(inp):
    out = {}
    for k, v in inp.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                out[f"in.hint.{kk}"] = np_(vv.float() if vv.dtype == torch.bool else vv)
        else:
            out[f"in.{k}"] = np_(v)
    return out


# ------------------------------------------------------------------------------------------------- volumes
def volume_case(name, kind, cfg, H, W, white=False, empty_hint=False, store_inputs=True, subsample=1,
                weight_seed=4242):
    inp = syn.cost_volume_inputs(cfg, white=white, empty_hint=empty_hint, match_hw=(H, W))
    K, D, C = cfg.num_src, cfg.planes, cfg.feat_ch
    extra = {}
    if kind == "dot":
        mgr = CostVolumeManager(H, W, num_depth_bins=D)
        inp.pop("cv_depth_hint_dict", None)
    elif kind == "mlp":
        mgr = quiet(FeatureVolumeManager, H, W, num_depth_bins=D, mlp_channels=[0, 128, 128, 1],
                    matching_dim_size=C, num_source_views=K)
        inp.pop("cv_depth_hint_dict", None)
        load_seeded(mgr, weight_seed)
    else:
        mgr = quiet(FeatureMeshHintVolumeManager, H, W, num_depth_bins=D, mlp_channels=[0, 128, 128, 1],
                    matching_dim_size=C, num_source_views=K)
        load_seeded(mgr, weight_seed)
    mgr.eval()
    vol, lowest, planes, mask = mgr(**inp, return_mask=True)
    if kind != "dot":
        fast = quiet(mgr.to_fast).eval()
        fvol, flow, _, fmask = fast(**inp, return_mask=True)
        extra["out.fast_mask"] = np_(fmask)
        extra["out.fast_vs_slow_maxabs"] = np.array(float((fvol - vol).abs().max()))
        assert torch.equal(torch.argmax(fvol, 1), torch.argmax(vol, 1)) or kind != "dot"
    idx = torch.argmax(vol, 1)
    arrays = dict(
        meta=np.array([cfg.batch, K, C, H, W, D, cfg.seed, weight_seed, int(white), int(empty_hint), subsample]),
        **{"out.index": np_(idx).astype(np.uint8), "out.lowest_cost": np_(lowest),
           "out.planes": np_(planes[0, :, 0, 0]), "out.volume": np_(vol[:, :, ::subsample, ::subsample])},
        **extra,
    )
    if mask is not None:
        arrays["out.mask"] = np_(mask)
    if store_inputs:
        arrays.update(flat_inputs(inp))
    save(name, **arrays)


# ------------------------------------------------------------------------------------------------- networks
def network_case(name, decoder, D, image_hw, prior_ch, batch, seed):
    ih, iw = image_hw
    cfg = syn.WorkloadConfig(name, batch, 2, ih, iw, D, prior_ch=tuple(prior_ch), decoder=decoder, seed=seed)
    priors = syn.prior_features(cfg)
    g = torch.Generator().manual_seed(seed + 11)
    cv = torch.randn(batch, D, ih // 4, iw // 4, generator=g)
    enc = CVEncoder(num_ch_cv=D, num_ch_enc=list(prior_ch[1:]), num_ch_outs=[64, 128, 256, 384]).eval()
    load_seeded(enc, seed + 1, scale=1.5)
    dec_in = list(prior_ch[:1]) + enc.num_ch_enc
    dec = (DepthDecoderPP(dec_in) if decoder == "unet_pp" else SkipDecoderRegression(dec_in)).eval()
    load_seeded(dec, seed + 2, scale=1.5)
    cvf = enc(cv, priors[1:])
    out = dec(priors[:1] + cvf)
    arrays = {"meta": np.array([batch, D, ih, iw, seed]), "prior_ch": np.array(prior_ch),
              "enc_shapes": np.array(json.dumps(param_shapes(enc))), "dec_shapes": np.array(json.dumps(param_shapes(dec)))}
    for i, f in enumerate(cvf):
        arrays[f"out.cv_feat_{i}"] = np_(f)
    for k, v in out.items():
        if k.startswith("feature_") and not k.startswith("feature_s3"):
            continue  # keep fixtures small; the log-depth heads cover these maps
        arrays[f"out.{k}"] = np_(v)
    save(name, **arrays)


# ------------------------------------------------------------------------------------------------- model
def model_case(name, cfg, empty_hint=False):
    """Follows DepthModelCVHint.forward (doubletake_model.py:341-349,374-423) around the reference modules."""
    H, W, K, D = cfg.match_h, cfg.match_w, cfg.num_src, cfg.planes
    g = torch.Generator().manual_seed(cfg.seed + 3)
    inp = syn.cost_volume_inputs(cfg, empty_hint=empty_hint)
    # absolute poses: random world pose for cur, src derived so that src_cam_T_cur_cam == ext
    ext = inp["src_extrinsics"]
    _, cur_pose = syn.relative_poses(cfg.batch, 1, g)
    cur_world_T_cam = cur_pose[:, 0]
    cur_cam_T_world = torch.linalg.inv(cur_world_T_cam.double()).float()
    src_cam_T_world = ext @ cur_cam_T_world[:, None]
    src_world_T_cam = torch.linalg.inv(src_cam_T_world.double()).float()
    cur_data = {"cam_T_world_b44": cur_cam_T_world, "world_T_cam_b44": cur_world_T_cam,
                "invK_s1_b44": inp["cur_invK"], **inp["cv_depth_hint_dict"]}
    src_data = {"cam_T_world_b44": src_cam_T_world, "world_T_cam_b44": src_world_T_cam, "K_s1_b44": inp["src_Ks"]}
    priors = syn.prior_features(cfg)

    cost_volume = quiet(FeatureMeshHintVolumeManager, H, W, num_depth_bins=D, mlp_channels=[0, 128, 128, 1],
                        matching_dim_size=cfg.feat_ch, num_source_views=K).eval()
    load_seeded(cost_volume, cfg.seed + 10)
    cost_volume_net = CVEncoder(num_ch_cv=D, num_ch_enc=list(cfg.prior_ch[1:]), num_ch_outs=[64, 128, 256, 384]).eval()
    load_seeded(cost_volume_net, cfg.seed + 11, scale=1.5)
    dec_in = list(cfg.prior_ch[:1]) + cost_volume_net.num_ch_enc
    depth_decoder = (DepthDecoderPP(dec_in) if cfg.decoder == "unet_pp" else SkipDecoderRegression(dec_in)).eval()
    load_seeded(depth_decoder, cfg.seed + 12, scale=1.5)

    # ---- doubletake_model.py:341-349
    src_cam_T_cur_cam = src_data["cam_T_world_b44"] @ cur_data["world_T_cam_b44"].unsqueeze(1)
    cur_cam_T_src_cam = cur_data["cam_T_world_b44"].unsqueeze(1) @ src_data["world_T_cam_b44"]
    # ---- :374-390
    min_depth = torch.tensor(0.25).type_as(src_data["K_s1_b44"]).view(1, 1, 1, 1)
    max_depth = torch.tensor(5.0).type_as(src_data["K_s1_b44"]).view(1, 1, 1, 1)
    cv, lowest_cost, _, overall_mask = cost_volume(
        cur_feats=inp["cur_feats"], src_feats=inp["src_feats"], src_extrinsics=src_cam_T_cur_cam,
        src_poses=cur_cam_T_src_cam, src_Ks=src_data["K_s1_b44"], cur_invK=cur_data["invK_s1_b44"],
        min_depth=min_depth, max_depth=max_depth, return_mask=True, cv_depth_hint_dict=cur_data)
    # ---- :398-406
    feats = cost_volume_net(cv, priors[1:])
    cur_feats = priors[:1] + feats
    depth_outputs = depth_decoder(cur_feats)
    # ---- :410-423
    for k in list(depth_outputs.keys()):
        log_depth = depth_outputs[k].float()
        depth_outputs[k] = log_depth
        depth_outputs[k.replace("log_", "")] = torch.exp(log_depth)
    arrays = {"meta": np.array([cfg.batch, K, cfg.feat_ch, H, W, D, cfg.seed, int(empty_hint)]),
              "cv_shapes": np.array(json.dumps(param_shapes(cost_volume))),
              "enc_shapes": np.array(json.dumps(param_shapes(cost_volume_net))),
              "dec_shapes": np.array(json.dumps(param_shapes(depth_decoder)))}
    for k, v in depth_outputs.items():
        if k.startswith("feature_"):
            continue
        arrays[f"out.{k}"] = np_(v)
    arrays["out.lowest_cost_bhw"] = np_(lowest_cost)
    arrays["out.overall_mask_bhw"] = np_(overall_mask)
    arrays["out.index"] = np_(torch.argmax(cv, 1)).astype(np.uint8)
    arrays["out.cost_volume"] = np_(cv)
    for k in ("cam_T_world_b44", "world_T_cam_b44"):
        arrays[f"in.cur.{k}"] = np_(cur_data[k])
        arrays[f"in.src.{k}"] = np_(src_data[k])
    save(name, **arrays)


def main():
    W = syn.WorkloadConfig
    # small, inputs stored
    volume_case("cv_dot_small", "dot", W("s", 1, 2, 0, 0, 32, hint=False, seed=2001), 48, 64)
    volume_case("cv_dot_portrait_white", "dot", W("s", 2, 3, 0, 0, 16, hint=False, seed=2002), 40, 24, white=True)
    volume_case("fv_mlp_small", "mlp", W("s", 1, 2, 0, 0, 16, hint=False, seed=2003), 48, 64)
    volume_case("fv_hint_small", "hint", W("s", 1, 3, 0, 0, 16, hint=True, seed=2004), 48, 64)
    volume_case("fv_hint_empty", "hint", W("s", 2, 2, 0, 0, 8, hint=True, seed=2005), 24, 32, empty_hint=True)
    # cfg 1 (BASELINE.json configs[0]): 128x160 feats, K=2, D=32 -- inputs regenerated from the seed,
    # volume stored at every 4th pixel, arg-max index stored in full.
    c1 = syn.CONFIGS["cfg1"]
    volume_case("cfg1_dot", "dot", c1, 128, 160, store_inputs=False, subsample=4)
    volume_case("cfg1_mlp", "mlp", c1, 128, 160, store_inputs=False, subsample=4)
    c1h = W("cfg1h", 1, 2, 512, 640, 32, hint=True, seed=1001)
    volume_case("cfg1_hint", "hint", c1h, 128, 160, store_inputs=False, subsample=4)
    # conv stacks
    network_case("net_pp_d64", "unet_pp", 64, (128, 192), (24, 48, 64, 160, 256), 1, 3001)
    network_case("net_pp_d16_b2", "unet_pp", 16, (64, 96), (24, 48, 64, 160, 256), 2, 3002)
    network_case("net_skip_d48", "skip", 48, (128, 192), (64, 64, 128, 256, 512), 2, 3003)
    # model forward composition
    model_case("model_tiny_pp", syn.CONFIGS["tiny"])
    model_case("model_tiny_skip", syn.CONFIGS["tiny_small"])
    model_case("model_tiny_pp_emptyhint", syn.CONFIGS["tiny"], empty_hint=True)


if __name__ == "__main__":
    main()
