"""Time the cost-volume kernel alone at cfg-2 size (development tool, GPU only)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import doubletake_b200 as dt  # noqa: E402
from doubletake_b200 import synthetic as syn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--math", default="tc3x")
ap.add_argument("--kind", default="hint")
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
cfg = syn.CONFIGS["cfg2"]
inp = syn.cost_volume_inputs(cfg)
dev = "cuda"
if args.kind == "dot":
    mgr = dt.CostVolumeManager(cfg.match_h, cfg.match_w, cfg.planes).to(dev)
    inp.pop("cv_depth_hint_dict")
else:
    mgr = dt.FeatureMeshHintVolumeManager(cfg.match_h, cfg.match_w, cfg.planes, num_source_views=cfg.num_src, math=args.math).to(dev)
ci = {k: ({kk: vv.to(dev) for kk, vv in v.items()} if isinstance(v, dict) else (v if "depth" in k and "hint" not in k else v.to(dev)))
      for k, v in inp.items()}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    mgr(**ci, return_mask=True)
torch.cuda.synchronize()
ts = []
for _ in range(args.reps):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    mgr(**ci, return_mask=True)
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ts.sort()
print(f"{args.kind} {args.math}: median {ts[len(ts)//2]*1e3:.1f} us  min {ts[0]*1e3:.1f} us")
