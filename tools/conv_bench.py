"""Per-layer timing of representative DoubleTake conv layers through the C ABI (development tool, GPU only).

    python tools/conv_bench.py [--math tc3x] [--only NAME] [--reps 20]
"""
import argparse
import os
import sys

os.environ.setdefault("DTB200_DEVELOPMENT", "1")  # lets --debug reach the timing knock-outs (wrong results by design)

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import doubletake_b200 as dt  # noqa: E402
from doubletake_b200 import _lib as L  # noqa: E402

# name, H, W (conv-input size), [(channels, resample)], out_c, k, stride
LAYERS = [
    ("s0_64_64_3x3", 240, 320, [(64, 0)], 64, 3, 1),
    ("s0_cat192_64_3x3_up", 240, 320, [(64, 0), (64, 1), (64, 1)], 64, 3, 1),
    ("s0_cat192_64_1x1_up", 240, 320, [(64, 0), (64, 1), (64, 1)], 64, 1, 1),
    ("s0_24_64_3x3", 240, 320, [(24, 0)], 64, 3, 1),
    ("s1_64_64_3x3", 120, 160, [(64, 0)], 64, 3, 1),
    ("s1_cat112_64_3x3", 120, 160, [(64, 0), (48, 0)], 64, 3, 1),
    ("s1_64_128_3x3_s2", 120, 160, [(64, 0)], 128, 3, 2),
    ("s2_128_128_3x3", 60, 80, [(128, 0)], 128, 3, 1),
    ("s2_cat384_128_3x3_up", 60, 80, [(128, 0), (128, 1), (128, 1)], 128, 3, 1),
    ("s3_256_256_3x3", 30, 40, [(256, 0)], 256, 3, 1),
    ("s3_cat416_256_3x3", 30, 40, [(256, 0), (160, 0)], 256, 3, 1),
    ("s4_384_384_3x3", 15, 20, [(384, 0)], 384, 3, 1),
    ("s4_cat640_384_3x3", 15, 20, [(384, 0), (256, 0)], 384, 3, 1),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--math", default="tc3x")
    ap.add_argument("--only", default=None)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--burst", type=int, default=1, help="plan replays per event pair (CUDA events tick in ~2 us steps on this driver)")
    ap.add_argument("--no-flush", action="store_true", help="L2-warm timing (inside a plan the producer's output is still in L2)")
    ap.add_argument("--debug", default="0", help="comma-separated dtb200_debug_set values, one pass per value")
    args = ap.parse_args()
    dev = torch.device("cuda")
    torch.zeros(1, device=dev)
    for flags in [int(x) for x in args.debug.split(",")]:
        L.check(L.lib().dtb200_debug_set(flags))
        if flags:
            print("debug flags", flags)
        run(args, dev)


def run(args, dev):
    torch.manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    total_ms = 0.0
    for name, H, W, srcs, oc, k, stride in LAYERS:
        if args.only and args.only not in name:
            continue
        plan = dt.ConvPlan(dev, args.math)
        feats = []
        for i, (c, r) in enumerate(srcs):
            h, w = (H // 2, W // 2) if r else (H, W)
            f = plan.new(1, h, w, c)
            plan.set_feature(f, torch.randn(1, c, h, w, device=dev))
            feats.append((f, r))
        conv = nn.Conv2d(sum(c for c, _ in srcs), oc, k, stride=stride, padding=k // 2)
        plan.conv(feats, conv, L.ACT_LEAKY, 0.2)
        plan.finalize()
        for _ in range(3):
            plan.run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            if not args.no_flush:
                flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(args.burst):
                plan.run()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) / args.burst)
        ts.sort()
        ms = ts[len(ts) // 2]
        fl = plan.flops()
        total_ms += ms
        print(f"{name:28s} {ms * 1e3:9.1f} us  {fl / ms / 1e9:8.1f} TFLOP/s  ({fl / 1e9:.2f} GFLOP)")
    print(f"total {total_ms:.3f} ms")


if __name__ == "__main__":
    main()
