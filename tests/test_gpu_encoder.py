"""GPU parity of the matching-feature encoder (doubletake_b200.ResnetMatchingEncoder through the C ABI)
against fixtures produced by executing the reference's ResnetMatchingEncoder, and end to end from images through
DepthModelCVHint.forward against the oracle."""
import pytest
import torch

import helpers as hp
import doubletake_b200 as dt
from doubletake_b200 import _lib as L
from doubletake_b200 import synthetic as syn
from oracle import oracle_encoder as oe
from oracle import oracle_torch as orc

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda"
# instance-normalised features are O(1) with |max| ~ 4.5: absolute tolerance
TOL = {"exact": 5e-5, "tc3x": 1e-4, "tch": 1e-4}


@pytest.mark.parametrize("math", ["exact", "tc3x", "tch"])
@pytest.mark.parametrize("name", ["enc_tv", "enc_aa", "enc_aa_odd"])
def test_encoder_matches_reference_fixture(name, math):
    fx = hp.load(name)
    n, h, w, seed, wseed, aa = [int(v) for v in fx["meta"]]
    images = torch.rand(n, 3, h, w, generator=torch.Generator().manual_seed(seed)) * 2 - 1
    enc = dt.ResnetMatchingEncoder(18, 16, antialiased=bool(aa), math=math)
    missing, unexpected = enc.load_state_dict(oe.encoder_state(wseed), strict=False)
    assert not unexpected and all(k.endswith("filt") or k.endswith("num_batches_tracked") for k in missing)
    enc = enc.to(DEV)
    ref = torch.from_numpy(fx["feats"])
    before = L.launch_count()
    got = enc(images.to(DEV))                                   # batched
    assert L.launch_count() > before and got.shape == ref.shape
    assert float((got.cpu() - ref).abs().max()) < TOL[math]
    one = torch.cat([enc(images[i:i + 1].to(DEV)) for i in range(n)], 0)   # unbatched, as the reference runs it at test time
    assert float((one.cpu() - ref).abs().max()) < TOL[math]
    # per-image InstanceNorm statistics: batching changes nothing but (in the tensor-core modes) the split-K plan
    assert float((one - got).abs().max()) < (1e-6 if math == "exact" else TOL[math])


def test_forward_views_layouts_and_model_from_images():
    """Images in, depth out: DepthModelCVHint with the B200 matching encoder attached (the image-prior encoder stays an
    injected callable) against the oracle encoder + oracle forward; the source features arrive channels-last, so the cost
    volume runs without its staging transpose."""
    cfg = syn.CONFIGS["tiny"]
    opts = dt.HotPathOptions(matching_num_depth_bins=cfg.planes, model_num_views=cfg.num_src + 1, image_height=cfg.image_h,
                             image_width=cfg.image_w)
    g = torch.Generator().manual_seed(5)
    cur_img = torch.rand(1, 3, cfg.image_h, cfg.image_w, generator=g) * 2 - 1
    src_img = torch.rand(1, cfg.num_src, 3, cfg.image_h, cfg.image_w, generator=g) * 2 - 1
    esd = oe.encoder_state(91)
    inp = syn.cost_volume_inputs(cfg)
    priors = syn.prior_features(cfg)
    eye = torch.eye(4).expand(cfg.batch, 4, 4).contiguous()
    cur = {"cam_T_world_b44": eye, "world_T_cam_b44": eye, "invK_s1_b44": inp["cur_invK"], **inp["cv_depth_hint_dict"]}
    src = {"cam_T_world_b44": inp["src_extrinsics"], "world_T_cam_b44": inp["src_poses"], "K_s1_b44": inp["src_Ks"]}
    for math in ("exact", "tch"):
        enc = dt.ResnetMatchingEncoder(18, 16, math=math)
        enc.load_state_dict(esd, strict=False)
        model = dt.DepthModelCVHint(opts, encoder=lambda image: [p.to(DEV) for p in priors], matching_model=enc, math=math,
                                    volume_math=math)
        shapes = {k: tuple(v.shape) for k, v in model.named_parameters() if not k.startswith("matching_model.")}
        sd = syn.seeded_state_dict(shapes, 2024, 1.3)
        model.load_state_dict(sd, strict=False)
        model = model.to(DEV)
        mc, ms = model.matching_model.forward_views(cur_img.to(DEV), src_img.to(DEV))
        assert mc.shape == (1, 16, cfg.match_h, cfg.match_w) and ms.shape == (1, cfg.num_src, 16, cfg.match_h, cfg.match_w)
        assert mc.is_contiguous() and ms.permute(0, 1, 3, 4, 2).is_contiguous()   # staged channels-last
        rc = oe.matching_encoder(cur_img, esd)
        rs = torch.stack([oe.matching_encoder(src_img[:, k], esd) for k in range(cfg.num_src)], 1)
        assert float((mc.cpu() - rc).abs().max()) < TOL[math] and float((ms.cpu() - rs).abs().max()) < TOL[math]
        cur_d = {k: v.to(DEV) for k, v in cur.items()}
        cur_d["image_b3hw"] = cur_img.to(DEV)
        src_d = {k: v.to(DEV) for k, v in src.items()}
        src_d["image_b3hw"] = src_img.to(DEV)
        out = model("test", cur_d, src_d, return_mask=True)
        ref = orc.depth_model_forward(rc, rs, priors, cur, src, sd, cfg.planes, hint=True)
        got, want = out["depth_pred_s0_b1hw"].cpu(), ref["depth_pred_s0_b1hw"]
        assert float(((got - want).abs() / want.abs()).max()) < 1e-4
        assert float((out["lowest_cost_bhw"].cpu() == ref["lowest_cost_bhw"]).float().mean()) > 0.995
