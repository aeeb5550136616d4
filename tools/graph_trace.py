"""In-graph timeline of the cfg-2 conv plan (development tool, GPU only): every tch conv kernel stamps %globaltimer at its first
CTA's start and its last CTA's end (dtb200_debug_trace), the DAG CUDA graph is replayed, and the stamps are printed per op next to
the dependency analysis.  Answers what per-kernel timing cannot: which kernels really overlap, where the replay idles, what the
dependent-launch gaps cost.

    python tools/graph_trace.py [--reps 5] [--csv gpurun_out/graph_trace.csv]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import doubletake_b200 as dt  # noqa: E402
from doubletake_b200 import _lib as L  # noqa: E402
from doubletake_b200 import synthetic as syn  # noqa: E402
from plan_profile import describe, flops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--csv", default=None)
    ap.add_argument("--workload", default="cfg2")
    args = ap.parse_args()
    torch.set_grad_enabled(False)
    dev = torch.device("cuda")
    cfg = syn.CONFIGS[args.workload]
    opts = dt.HotPathOptions(matching_num_depth_bins=cfg.planes, model_num_views=cfg.num_src + 1,
                             image_height=cfg.image_h, image_width=cfg.image_w)
    model = dt.DepthModelCVHint(opts, math="tch", volume_math="tch")
    shapes = {k: tuple(v.shape) for k, v in model.named_parameters()}
    model.load_state_dict(syn.seeded_state_dict(shapes, 2024, 1.3), strict=False)
    model = model.to(dev)
    priors = [p.to(dev) for p in syn.prior_features(cfg)]
    cv = torch.randn(cfg.batch, cfg.planes, cfg.match_h, cfg.match_w, device=dev)
    cap = 1024
    buf = torch.empty((cap, 2), dtype=torch.int64, device=dev)
    L.check(L.lib().dtb200_debug_trace(L.ptr(buf), cap))       # before the plan is captured: slots are baked into the graph
    plan = model._network_plan(cv.shape, priors)
    plan.load_inputs({"cv": cv, **{f"prior{i}": f for i, f in enumerate(priors)}})
    plan.run()                                                 # the first run captures the graph (slots in op order)
    torch.cuda.synchronize()
    L.check(L.lib().dtb200_debug_trace(None, 0))               # later launches (other plans) are not traced
    n = len(plan.ops)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    runs = []
    for rep in range(args.reps + 2):
        buf[:, 0] = 2 ** 63 - 1
        buf[:, 1] = 0
        flush.zero_()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        plan.run()
        e.record()
        torch.cuda.synchronize()
        runs.append((s.elapsed_time(e) * 1e3, buf[:n].cpu().clone()))
    us, stamps = sorted(runs[2:], key=lambda r: r[0])[len(runs[2:]) // 2]
    used = int((buf[:, 1] > 0).sum())
    assert used == n, f"{used} stamped launches for {n} ops: the op <-> slot mapping assumes one kernel per op"
    t0 = int(stamps[:, 0].min())
    start = (stamps[:, 0] - t0).double() / 1e3
    end = (stamps[:, 1] - t0).double() / 1e3
    info = plan.analyze()
    span = float(end.max())
    print(f"{n} ops; replay {us:.1f} us by CUDA events, {span:.1f} us first start -> last end by %globaltimer")
    # busy time: union of the kernels' [start, end] intervals, and the same for the 148-CTA (whole-GPU) kernels alone
    def union(iv):
        iv = sorted(iv)
        tot, cur_s, cur_e = 0.0, None, None
        for a, b in iv:
            if cur_e is None or a > cur_e:
                if cur_e is not None:
                    tot += cur_e - cur_s
                cur_s, cur_e = a, b
            else:
                cur_e = max(cur_e, b)
        return tot + (cur_e - cur_s if cur_e is not None else 0.0)
    kinds = [describe(op)[0] for op in plan.ops]
    alliv = [(float(start[i]), float(end[i])) for i in range(n)]
    print(f"  some kernel running: {union(alliv):.1f} us; idle inside the replay: {span - union(alliv):.1f} us")
    for kind in sorted(set(kinds)):
        iv = [alliv[i] for i in range(n) if kinds[i] == kind]
        print(f"  {kind:10s} {len(iv):3d} ops: sum of durations {sum(b - a for a, b in iv):8.1f} us, union {union(iv):8.1f} us")
    # dependency gap: start of an op minus the latest end of its dependencies
    gaps = []
    for i in range(n):
        deps = info[i]["deps"]
        if deps:
            gaps.append(float(start[i]) - max(float(end[d]) for d in deps))
    gaps_t = torch.tensor(gaps)
    print(f"  start - latest dependency end: median {float(gaps_t.median()):.2f} us, mean {float(gaps_t.mean()):.2f} us, "
          f"negative (programmatic early start) {int((gaps_t < 0).sum())} of {len(gaps)}")
    print("   op kind       shape                                          start      end      dur   gap-after-deps  lane")
    for i in sorted(range(n), key=lambda i: float(start[i])):
        deps = info[i]["deps"]
        gap = float(start[i]) - max(float(end[d]) for d in deps) if deps else 0.0
        print(f"  {i:3d} {kinds[i]:10s} {describe(plan.ops[i])[1]:44s} {float(start[i]):8.1f} {float(end[i]):8.1f} {float(end[i] - start[i]):8.1f} {gap:10.2f}"
              f"      {info[i]['lane']}")
    if args.csv:
        with open(args.csv, "w") as f:
            f.write("op,kind,shape,start_us,end_us,lane,level,deps\n")
            for i in range(n):
                f.write(f"{i},{kinds[i]},{describe(plan.ops[i])[1]},{float(start[i]):.2f},{float(end[i]):.2f},{info[i]['lane']},{info[i]['level']},"
                        f"{' '.join(map(str, info[i]['deps']))}\n")


if __name__ == "__main__":
    main()
