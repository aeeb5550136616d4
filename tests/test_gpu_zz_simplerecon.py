"""GPU parity of the SimpleRecon variants (DepthModel, reference experiment_modules/sr_depth_model.py:32-435: dot-product or
metadata-MLP volume, no hint) through the whole forward, against the oracle composition.  (Added after the round's last GPU
run: each component is covered by the fixture tests; the file sorts last so that `pytest -x` reaches it after them.)"""
import pytest
import torch

import doubletake_b200 as dt
from doubletake_b200 import synthetic as syn
from oracle import oracle_torch as orc

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda"


@pytest.mark.parametrize("fvt,volume", [("simple_cost_volume", "dot"), ("mlp_feature_volume", "mlp")])
@pytest.mark.parametrize("math", ["exact", "tc3x"])
def test_simplerecon_forward_matches_oracle(fvt, volume, math):
    cfg = syn.WorkloadConfig("sr", 1, 3, 192, 256, 24, hint=False, seed=3100)
    opts = dt.HotPathOptions(feature_volume_type=fvt, matching_num_depth_bins=cfg.planes, model_num_views=cfg.num_src + 1,
                             image_height=cfg.image_h, image_width=cfg.image_w)
    vm = math if volume == "mlp" else "exact"  # the dot-product volume has one (exact) kernel
    model = dt.DepthModel(opts, math=math, volume_math=vm)
    shapes = {k: tuple(v.shape) for k, v in model.named_parameters()}
    sd = syn.seeded_state_dict(shapes, 3101, 1.3)
    model.load_state_dict(sd, strict=False)
    model = model.to(DEV)
    inp = syn.cost_volume_inputs(cfg)
    priors = syn.prior_features(cfg)
    eye = torch.eye(4).expand(cfg.batch, 4, 4).contiguous()
    cur = {"cam_T_world_b44": eye, "world_T_cam_b44": eye, "invK_s1_b44": inp["cur_invK"]}
    src = {"cam_T_world_b44": inp["src_extrinsics"], "world_T_cam_b44": inp["src_poses"], "K_s1_b44": inp["src_Ks"]}
    ref = orc.depth_model_forward(inp["cur_feats"], inp["src_feats"], priors, cur, src, sd, cfg.planes, hint=False,
                                  volume=volume, mask_mode="fast")
    cur_d = {k: v.to(DEV) for k, v in cur.items()}
    src_d = {k: v.to(DEV) for k, v in src.items()}
    cur_d["image_prior_feats"] = [p.to(DEV) for p in priors]
    cur_d["matching_feats_bchw"] = inp["cur_feats"].to(DEV)
    src_d["matching_feats_bkchw"] = inp["src_feats"].to(DEV)
    out = model("test", cur_d, src_d, return_mask=True)
    for i in range(4):
        got, want = out[f"depth_pred_s{i}_b1hw"].cpu(), ref[f"depth_pred_s{i}_b1hw"]
        assert got.shape == want.shape
        assert float(((got - want).abs() / want.abs()).max()) < 1e-4, i
    mism = out["lowest_cost_bhw"].cpu() != ref["lowest_cost_bhw"]
    assert float(mism.float().mean()) < 5e-3, int(mism.sum())  # near-ties only (volume-level proof: test_gpu_cost_volume.py)
    if volume == "dot":
        assert out["overall_mask_bhw"] is None  # CostVolumeManager returns no mask (cost_volume.py:315)
    else:
        assert torch.equal(out["overall_mask_bhw"].cpu(), ref["overall_mask_bhw"])
