"""Where does the conv stack's time go?  Per-op timing of the REAL cfg-2 plan (CVEncoder + DepthDecoderPP, the one bench.py
runs) through the C ABI, one descriptor at a time, next to the host-side dependency analysis (development tool, GPU only).

    python tools/plan_profile.py [--reps 10] [--top 40] [--csv gpurun_out/plan_ops.csv]

Prints: every op (kernel family, shape, tiles / persistent rounds, isolated time, TFLOP/s), totals per family, the sum of
isolated times vs the critical path through the DAG (lower bound of the graph replay with unlimited overlap) vs the measured
graph replay.  Isolated times are L2-warm on purpose: inside a step the producer's output is still in the 126 MB L2.
"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import doubletake_b200 as dt  # noqa: E402
from doubletake_b200 import _lib as L  # noqa: E402
from doubletake_b200 import synthetic as syn  # noqa: E402


def describe(op):
    cin = [op.src_c[i] for i in range(op.num_src)]
    if op.ksize == 0:
        return "resample", f"x2 {op.in_h // 2}x{op.in_w // 2}->{op.out_h}x{op.out_w} c{op.out_c}"
    kind = "head" if op.out_c % 64 else ("halo" if (op.ksize == 3 and op.stride == 1 and op.out_c == 64) else "tap-major")
    res = "+res" if op.residual else ""
    return kind, f"{op.ksize}x{op.ksize}/s{op.stride} {op.in_h}x{op.in_w} {'+'.join(map(str, cin))}->{op.out_c}{res}"


def flops(op):
    if op.ksize == 0:
        return 0
    return 2 * op.batch * op.out_h * op.out_w * op.out_c * sum(op.src_c[i] for i in range(op.num_src)) * op.ksize ** 2


def tiles(op):
    if op.ksize == 0 or op.out_c % 64:
        return 0
    bn = 128 if op.out_c % 128 == 0 else 64
    if op.ksize == 3 and op.stride == 1 and bn == 64:
        t = -(-op.out_w // 8) * -(-op.out_h // 16)
        if t >= 148:
            return op.batch * t
    best = min(-(-op.out_w // w) * -(-op.out_h // (128 // w)) for w in (128, 64, 32, 16, 8))
    return op.batch * best * (op.out_c // bn)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--csv", default=None)
    args = ap.parse_args()
    torch.set_grad_enabled(False)
    dev = torch.device("cuda")
    cfg = syn.CONFIGS["cfg2"]
    opts = dt.HotPathOptions(matching_num_depth_bins=cfg.planes, model_num_views=cfg.num_src + 1,
                             image_height=cfg.image_h, image_width=cfg.image_w)
    math = os.environ.get("DTB200_PLAN_MATH", "tch")
    model = dt.DepthModelCVHint(opts, math=math, volume_math=math)
    shapes = {k: tuple(v.shape) for k, v in model.named_parameters()}
    model.load_state_dict(syn.seeded_state_dict(shapes, 2024, 1.3), strict=False)
    model = model.to(dev)
    priors = [p.to(dev) for p in syn.prior_features(cfg)]
    cv = torch.randn(cfg.batch, cfg.planes, cfg.match_h, cfg.match_w, device=dev)
    plan = model._network_plan(cv.shape, priors)
    plan.load_inputs({"cv": cv, **{f"prior{i}": f for i, f in enumerate(priors)}})
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()

    def timed(fn):
        ts = []
        for _ in range(args.reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        return sorted(ts)[len(ts) // 2] * 1e3  # us

    graph_us = timed(plan.run)
    stream = L.stream()
    rows = []
    for i, op in enumerate(plan.ops):
        us = timed(lambda: L.check(L.lib().dtb200_conv2d(C.byref(op), stream)))
        kind, text = describe(op)
        t = tiles(op)
        rows.append(dict(i=i, kind=kind, text=text, us=us, gflop=flops(op) / 1e9, tiles=t, rounds=t / 148.0 if t else 0.0))
    info = plan.analyze()
    # critical path: longest chain of isolated times through the reduced dependency DAG
    finish = [0.0] * len(rows)
    for i, r in enumerate(rows):
        finish[i] = r["us"] + max([finish[d] for d in info[i]["deps"]], default=0.0)
    total = sum(r["us"] for r in rows)
    print(f"{len(rows)} ops; graph replay {graph_us:.1f} us; sum of isolated op times {total:.1f} us; critical path {max(finish):.1f} us")
    fam = {}
    for r in rows:
        f = fam.setdefault(r["kind"], [0, 0.0, 0.0])
        f[0] += 1
        f[1] += r["us"]
        f[2] += r["gflop"]
    for k, (n, us, gf) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:10s} {n:4d} ops {us:9.1f} us ({us / total * 100:5.1f} %)  {gf:8.2f} GFLOP  {gf / max(us, 1e-9) * 1e-3:7.1f} TFLOP/s")
    print(f"top {args.top} ops by isolated time:")
    for r in sorted(rows, key=lambda r: -r["us"])[: args.top]:
        tf = r["gflop"] / max(r["us"], 1e-9) * 1e-3
        print(f"  #{r['i']:3d} {r['kind']:10s} {r['text']:44s} {r['us']:8.1f} us {tf:7.1f} TFLOP/s  tiles {r['tiles']:5d} ({r['rounds']:.2f} rounds)"
              f"  lane {info[r['i']]['lane']} level {info[r['i']]['level']}")
    if args.csv:
        with open(args.csv, "w") as f:
            f.write("op,kind,shape,us,gflop,tiles,rounds,lane,level,deps\n")
            for r in rows:
                a = info[r["i"]]
                f.write(f"{r['i']},{r['kind']},{r['text']},{r['us']:.2f},{r['gflop']:.4f},{r['tiles']},{r['rounds']:.3f},{a['lane']},"
                        f"{a['level']},{' '.join(map(str, a['deps']))}\n")


if __name__ == "__main__":
    main()
