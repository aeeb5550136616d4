"""Where does the depth error come from?  Runs one BASELINE configuration through DepthModelCVHint with a given pair of
arithmetic modes (conv stack / cost volume) and prints, against the CPU oracle: the cost-volume error, and per output scale
the max / 99.99th percentile / mean relative depth error with the location of the maximum (development tool, GPU only).

    python tools/parity_diag.py --cfg cfg5 --math tch --volume-math tch [--batch 1]
"""
import argparse
import dataclasses
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import doubletake_b200 as dt  # noqa: E402
from doubletake_b200 import synthetic as syn  # noqa: E402
from oracle import oracle_torch as orc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="cfg5")
ap.add_argument("--math", default="tch")
ap.add_argument("--volume-math", default="tch")
ap.add_argument("--batch", type=int, default=0)
args = ap.parse_args()
torch.set_grad_enabled(False)
DEV = "cuda"
cfg = syn.CONFIGS[args.cfg]
if args.batch:
    cfg = dataclasses.replace(cfg, batch=args.batch)
fam = "efficientnet" if cfg.prior_ch[0] == 24 else "resnet18d"
opts = dt.HotPathOptions(image_encoder_name=fam, depth_decoder_name=cfg.decoder, matching_num_depth_bins=cfg.planes,
                         model_num_views=cfg.num_src + 1, image_height=cfg.image_h, image_width=cfg.image_w)
inp = syn.cost_volume_inputs(cfg)
priors = syn.prior_features(cfg)
eye = torch.eye(4).expand(cfg.batch, 4, 4).contiguous()
cur = {"cam_T_world_b44": eye, "world_T_cam_b44": eye, "invK_s1_b44": inp["cur_invK"], **inp["cv_depth_hint_dict"]}
src = {"cam_T_world_b44": inp["src_extrinsics"], "world_T_cam_b44": inp["src_poses"], "K_s1_b44": inp["src_Ks"]}
shapes = {k: tuple(v.shape) for k, v in dt.DepthModelCVHint(opts).named_parameters()}
sd = syn.seeded_state_dict(shapes, 2024, 1.3)
ref = orc.depth_model_forward(inp["cur_feats"], inp["src_feats"], priors, cur, src, sd, cfg.planes, hint=True, decoder=cfg.decoder)
cur_d = {k: v.to(DEV) for k, v in cur.items()}
src_d = {k: v.to(DEV) for k, v in src.items()}
cur_d["image_prior_feats"] = [p.to(DEV) for p in priors]
cur_d["matching_feats_bchw"] = inp["cur_feats"].to(DEV)
src_d["matching_feats_bkchw"] = inp["src_feats"].to(DEV)
for math, vmath in [tuple(x.split("+")) for x in (args.math + "+" + args.volume_math).split(",")] if "," not in args.math else []:
    pass
pairs = [(m, v) for m in args.math.split(",") for v in args.volume_math.split(",")]
for math, vmath in pairs:
    model = dt.DepthModelCVHint(opts, math=math, volume_math=vmath)
    model.load_state_dict(sd, strict=False)
    model = model.to(DEV)
    out = model("test", cur_d, src_d, return_mask=True)
    cv = model.cost_volume._run(cur_d["matching_feats_bchw"], src_d["matching_feats_bkchw"], *model._relative_poses(cur_d, src_d, torch.device(DEV)),
                                src_d["K_s1_b44"], cur_d["invK_s1_b44"], torch.tensor(0.25).view(1, 1, 1, 1), torch.tensor(5.0).view(1, 1, 1, 1),
                                cur_d, None, True)
    torch.cuda.synchronize()
    vol, rvol = cv["volume"].cpu(), ref["cost_volume"]
    print(f"== conv {math} / volume {vmath}: volume max|err| / max|vol| = {float((vol - rvol).abs().max() / rvol.abs().max()):.3e}, "
          f"mean|err| = {float((vol - rvol).abs().mean()):.3e}, max|vol| = {float(rvol.abs().max()):.3f}")
    for i in range(4):
        got, want = out[f"depth_pred_s{i}_b1hw"].cpu(), ref[f"depth_pred_s{i}_b1hw"]
        rel = ((got - want).abs() / want.abs()).flatten()
        k = int(rel.argmax())
        q = float(torch.quantile(rel[:: max(1, rel.numel() // 2000000)], 0.9999))
        lg = (out[f"log_depth_pred_s{i}_b1hw"].cpu() - ref[f"log_depth_pred_s{i}_b1hw"]).abs()
        print(f"   s{i}: depth rel err max {float(rel.max()):.3e} (flat index {k} of {rel.numel()}, ref depth {float(want.flatten()[k]):.4f}) "
              f"p99.99 {q:.3e} mean {float(rel.mean()):.3e}; log-depth abs err max {float(lg.max()):.3e}, max|log depth| {float(ref[f'log_depth_pred_s{i}_b1hw'].abs().max()):.2f}")
