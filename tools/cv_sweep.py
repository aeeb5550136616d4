"""North-star sweep of the fused warp+match cost-volume kernels (SURVEY.md 8 ambiguity note).

Runs, with the FEATURE MAP itself at 120x160 (what a 640x480 image gives) and at 480x640 (16x the work, so that the
bandwidth measurement is not launch-latency dominated), over D in {8, 16, 64} planes and K in {1, 2, 7} source views:
  dot/ldg   CostVolumeManager, bilinear taps gathered with __ldg (L1-cached)            csrc/cost_volume.cu cv_dot_kernel
  dot/tma   the same with per-(tile, plane, view) TMA boxes staged in shared memory     csrc/cost_volume.cu cv_dot_tma_kernel
  mlp/tch   FeatureMeshHintVolumeManager, metadata MLP + hint MLP on tcgen05            csrc/cost_volume_tch.cu
and prints kernel time (CUDA-graph replay of one manager call, L2 flushed between replays), algorithmic bytes per launch
(4[(K+1)CHW + 3HW(hint) + DHW + HW] + HW), achieved GB/s and the fraction of the measured HBM copy peak.

    python tools/cv_sweep.py [--reps 5] [--sizes 120x160,480x640] [--once]
--once: a single plain launch per point (for `ncu --metrics dram__bytes_read.sum,...`, see tools/gpu_call26.sh).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import doubletake_b200 as dt  # noqa: E402
from doubletake_b200 import synthetic as syn  # noqa: E402
import helpers as hp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--sizes", default="120x160,480x640")
ap.add_argument("--planes", default="8,16,64")
ap.add_argument("--views", default="1,2,7")
ap.add_argument("--kinds", default="dot/ldg,dot/tma,mlp/tch")
ap.add_argument("--once", action="store_true")
args = ap.parse_args()
torch.set_grad_enabled(False)
dev = "cuda"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print(f"# HBM copy peak {peak:.1f} GB/s (MEASURED_PEAKS.json); time = median of {args.reps} CUDA-graph replays, L2 flushed between replays")
print(f"{'kernel':9s} {'feat map':>9s} {'D':>3s} {'K':>2s} {'time us':>9s} {'alg MB':>8s} {'GB/s':>8s} {'% HBM':>6s}")
for size in args.sizes.split(","):
    H, W = [int(v) for v in size.split("x")]
    for K in [int(v) for v in args.views.split(",")]:
        cfg = syn.WorkloadConfig("sweep", 1, K, 0, 0, 64, hint=True, seed=4000 + K)
        inp = syn.cost_volume_inputs(cfg, match_hw=(H, W))
        ci = {k: ({kk: vv.to(dev) for kk, vv in v.items()} if isinstance(v, dict) else (v if k in ("min_depth", "max_depth") else v.to(dev)))
              for k, v in inp.items()}
        staged = ci["src_feats"].permute(0, 1, 3, 4, 2).contiguous()
        ci["src_feats"] = staged.permute(0, 1, 4, 2, 3)   # channels-last staged: the managers use it without a transpose
        weights = hp.volume_weights("hint", K, 16, 7)
        for D in [int(v) for v in args.planes.split(",")]:
            for kind in args.kinds.split(","):
                os.environ["DTB200_CV_DOT_VARIANT"] = "tma" if kind == "dot/tma" else "ldg"
                if kind.startswith("dot"):
                    mgr = dt.CostVolumeManager(H, W, D).to(dev)
                    call = {k: v for k, v in ci.items() if k != "cv_depth_hint_dict"}
                    hint = False
                else:
                    mgr = dt.FeatureMeshHintVolumeManager(H, W, D, num_source_views=K, math="tch")
                    mgr.load_state_dict(weights, strict=False)
                    mgr = mgr.to(dev)
                    call, hint = ci, True

                def run():
                    return mgr(**call, return_mask=False)

                run()
                torch.cuda.synchronize()
                if args.once:
                    run()
                    torch.cuda.synchronize()
                    print(f"{kind:9s} {size:>9s} {D:3d} {K:2d} once")
                    continue
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    out = run()
                ts = []
                for _ in range(args.reps):
                    flush.zero_()
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    g.replay()
                    e.record()
                    torch.cuda.synchronize()
                    ts.append(s.elapsed_time(e) * 1e3)
                us = sorted(ts)[len(ts) // 2]
                alg = 4 * ((K + 1) * 16 * H * W + (3 * H * W if hint else 0) + D * H * W + H * W) + H * W
                gbs = alg / us / 1e3
                print(f"{kind:9s} {size:>9s} {D:3d} {K:2d} {us:9.1f} {alg / 1e6:8.2f} {gbs:8.1f} {100 * gbs / peak:6.2f}")
                del g
