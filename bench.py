#!/usr/bin/env python
"""Benchmark of the DoubleTake hot path on B200: depth frames/s at 640x480 image x 64 planes x 7 source views
(BASELINE.json metric; cfg 2 = configs[1]).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU arm: the oracle port of the reference on the host cores

One "step" = one pass of the hot path (relative poses -> fused hint cost volume -> CVEncoder -> DepthDecoderPP -> exp)
over one batch (B=1 frame per rank) of synthetic input.  `value` = frames/s with inputs resident in HBM, device-timed
with CUDA events (one event pair per step, L2 flushed between steps outside the timed intervals), max over ranks.
`e2e` = the same metric through the public API `DepthModelCVHint.forward` fed from pinned HOST buffers (H2D of the
step's inputs and D2H of the depth map inside the timed region).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from doubletake_b200 import synthetic as syn  # noqa: E402

METRIC = "depth frames/sec at 640x480x64planes x7views"
UNIT = "frames/s"
L2_FLUSH_BYTES = 256 << 20
# BASELINE.json configs: [1] = cfg2 is the headline metric's configuration (the default); cfg3 / cfg5 are the other two
# single-GPU throughput configurations, benchmarked with --workload and committed under profiles/
WORKLOADS = {
    "cfg2": (METRIC, "cfg2: DoubleTake 640x480 image, 120x160x16 matching feats, 64 planes, 7 src views, "
                     "rendered-depth hint on, batch 1 per GPU, CVEncoder+DepthDecoderPP (effnetv2-s priors)"),
    "cfg3": ("depth frames/sec at 512x384x48planes x5views, batch 8 (DoubleTake-small)",
             "cfg3: DoubleTake-small 512x384 image, 96x128x16 matching feats, 48 planes, 5 src views, hint on, batch 8 per GPU, "
             "CVEncoder+SkipDecoderRegression (resnet18d priors)"),
    "cfg4": ("depth frames/sec over the ScanNetv2 test split shapes (512x384, 64 planes, 7 views), keyframes sharded round-robin",
             "cfg4: DoubleTake 512x384 image, 96x128x16 matching feats, 64 planes, 7 src views, hint on, one keyframe per rank per step "
             "taken from this rank's shard of the 25 590-tuple / 100-scan ScanNetv2 test split (synthetic tensors per tuple), NCCL "
             "depth-map gather batched every --gather-every keyframes"),
    "cfg5": ("depth frames/sec at 1024x768x96planes x9views, batch 4 (stress)",
             "cfg5: synthetic stress 1024x768 image, 192x256x16 matching feats, 96 planes, 9 src views, hint on, batch 4 per GPU, "
             "CVEncoder+DepthDecoderPP (effnetv2-s priors)"),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def algorithmic_bytes_cost_volume(cfg, hint=True):
    """4*[(K+1)*C*H*W + 3*H*W + D*H*W + H*W] + H*W per frame."""
    H, W, K, C, D = cfg.match_h, cfg.match_w, cfg.num_src, cfg.feat_ch, cfg.planes
    return cfg.batch * (4 * ((K + 1) * C * H * W + (3 * H * W if hint else 0) + D * H * W + H * W) + H * W)


def mlp_flops(cfg):
    """2*H*W*D*(F*128 + 128*128 + 128 + 36 + 144 + 12) per frame."""
    H, W, D, F = cfg.match_h, cfg.match_w, cfg.planes, cfg.mlp_in
    return cfg.batch * 2 * H * W * D * (F * 128 + 128 * 128 + 128 + 36 + 144 + 12)


# --------------------------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------------------------
def host_inputs(cfg, seed):
    """(cur_data, src_data) on the HOST, as DepthModelCVHint.forward consumes them, with the upstream encoders'
    outputs (matching features, 5 image-prior maps) supplied precomputed."""
    inp = syn.cost_volume_inputs(cfg, seed=seed)
    priors = syn.prior_features(cfg, syn._gen(seed + 7))
    eye = torch.eye(4).expand(cfg.batch, 4, 4).contiguous()
    cur = {"cam_T_world_b44": eye, "world_T_cam_b44": eye.clone(), "invK_s1_b44": inp["cur_invK"],
           "matching_feats_bchw": inp["cur_feats"], "image_prior_feats": priors}
    cur.update({k: v for k, v in inp["cv_depth_hint_dict"].items() if v.dtype != torch.bool})
    src = {"cam_T_world_b44": inp["src_extrinsics"], "world_T_cam_b44": inp["src_poses"], "K_s1_b44": inp["src_Ks"],
           "matching_feats_bkchw": inp["src_feats"]}
    return cur, src


def tree_map(fn, d):
    out = {}
    for k, v in d.items():
        if isinstance(v, (list, tuple)):
            out[k] = [fn(x) for x in v]
        else:
            out[k] = fn(v)
    return out


def tree_bytes(d):
    n = 0
    for v in d.values():
        for x in (v if isinstance(v, (list, tuple)) else [v]):
            n += x.numel() * x.element_size()
    return n


def model_options(cfg):
    import doubletake_b200 as dt
    return dt.HotPathOptions(image_encoder_name="efficientnet" if cfg.prior_ch[0] == 24 else "resnet18d",
                             depth_decoder_name=cfg.decoder, matching_num_depth_bins=cfg.planes,
                             model_num_views=cfg.num_src + 1, image_height=cfg.image_h, image_width=cfg.image_w)


def model_weights(model, seed=2024):
    shapes = {k: tuple(v.shape) for k, v in model.named_parameters()}
    return syn.seeded_state_dict(shapes, seed, 1.3)


# --------------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled every 20 ms in a child process.  It is started BEFORE the warm-up (the tool needs a few hundred
    ms to come up, longer than a short timed region) and every line carries nvidia-smi's own timestamp, so ``stop`` keeps
    exactly the samples taken between the wall-clock marks of the timed regions."""

    FIELDS = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    @staticmethod
    def _epoch(stamp):
        # "2026/10/17 01:48:12.345" (local time of the box, like time.time() through mktime)
        try:
            main, _, ms = stamp.partition(".")
            return time.mktime(time.strptime(main, "%Y/%m/%d %H:%M:%S")) + (float("0." + ms) if ms else 0.0)
        except Exception:
            return None

    def stop(self, windows=()):
        """``windows``: (t0, t1) pairs of time.time() marks; samples outside every window are dropped (with a 50 ms
        margin).  If no sample falls inside a window (a region shorter than the polling period) all samples are kept
        and ``window`` says so."""
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 10:
                    continue
                try:
                    sm, mx = float(parts[2]), float(parts[3])
                except ValueError:
                    continue
                rows.append((self._epoch(parts[0]), sm, mx,
                             {n for n, v in zip(names, parts[6:10]) if v.lower().startswith("active")}))
            os.unlink(self.path)
        except Exception:
            pass
        self.rows = rows
        return self.window_stats(windows)

    def window_stats(self, windows):
        """Median SM clock / throttle reasons over the samples inside the given (t0, t1) wall-clock windows."""
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        rows = getattr(self, "rows", [])
        inside = [r for r in rows if r[0] is not None and any(t0 - 0.05 <= r[0] <= t1 + 0.05 for t0, t1 in windows)]
        out["window"] = "timed regions" if inside else "whole run (no sample fell inside the timed regions)"
        use = inside or rows
        if use:
            sm = sorted(r[1] for r in use)
            reasons = set().union(*[r[3] for r in use])
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(r[2] for r in use), reasons=sorted(reasons), samples=len(use))
        return out


# --------------------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist

    import doubletake_b200 as dt
    from doubletake_b200 import _lib as L
    from doubletake_b200 import sharding

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: doubletake_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its banner there)
        dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)

    cfg = syn.CONFIGS[args.workload]
    metric, workload_desc = WORKLOADS[args.workload]
    opts = model_options(cfg)
    volume_math = args.volume_math or ("exact" if args.math == "exact" else "tch")
    model = dt.DepthModelCVHint(opts, math=args.math, volume_math=volume_math)
    model.load_state_dict(model_weights(model), strict=False)
    model = model.to(dev)

    # a ring of distinct input sets (different seeds per rank and slot), resident in HBM and mirrored in pinned host memory
    n_sets = 4
    host_sets, dev_sets = [], []
    for s in range(n_sets):
        cur, src = host_inputs(cfg, cfg.seed + 100 * rank + s)
        cur, src = tree_map(lambda t: t.pin_memory(), cur), tree_map(lambda t: t.pin_memory(), src)
        host_sets.append((cur, src))
        dev_sets.append((tree_map(lambda t: t.to(dev), cur), tree_map(lambda t: t.to(dev), src)))
    h2d_bytes = tree_bytes(host_sets[0][0]) + tree_bytes(host_sets[0][1])
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    frames_per_step = cfg.batch * world
    depth_shape = (cfg.batch, 1, cfg.image_h // 2, cfg.image_w // 2)
    host_out = torch.empty(depth_shape, dtype=torch.float32).pin_memory()

    # cfg 4: this rank's keyframes of the tuple list (round-robin, or whole scans for the incremental variant); a step is the
    # next keyframe of the shard, its synthetic tensors picked by the tuple; depth maps are parked and gathered every G keyframes
    shard = None
    if args.workload == "cfg4":
        tuples = sharding.synthetic_scannet_test_tuples()
        shard = sharding.shard_tuples(tuples, rank, world, by=args.shard_by)
        G = max(1, args.gather_every)
        parked = torch.empty((G,) + depth_shape[1:], dtype=torch.float32, device=dev)

    def step_resident(i):
        if shard is not None:
            t = shard[i % len(shard)]
            cur, src = dev_sets[(t * 2654435761 >> 7) % n_sets]
            out = model("test", cur, src, return_mask=True)
            depth = out["depth_pred_s0_b1hw"]
            if world > 1:
                parked[i % G].copy_(depth[0])
                if (i + 1) % G == 0:
                    depth = sharding.gather_depth_maps(parked, G * world)
            return depth
        cur, src = dev_sets[i % n_sets]
        out = model("test", cur, src, return_mask=True)
        depth = out["depth_pred_s0_b1hw"]
        if world > 1:
            depth = sharding.gather_depth_maps(depth, frames_per_step)
        return depth

    def step_e2e(i):
        # pinned HOST dicts go straight into the public API: forward() stages them on its copy stream (matching features
        # and hint first, prior maps while the cost volume runs) -- every byte is copied inside the timed region
        if shard is not None:
            t = shard[i % len(shard)]
            cur, src = host_sets[(t * 2654435761 >> 7) % n_sets]
        else:
            cur, src = host_sets[i % n_sets]
        out = model("test", cur, src, return_mask=True)
        depth = out["depth_pred_s0_b1hw"]
        if world > 1:
            if shard is not None:
                parked[i % G].copy_(depth[0])
                if (i + 1) % G == 0:
                    sharding.gather_depth_maps(parked, G * world)
            else:
                depth = sharding.gather_depth_maps(depth, frames_per_step)
        host_out.copy_(depth[: cfg.batch], non_blocking=True)
        return depth

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # before the warm-up: nvidia-smi takes longer to come up than a short timed region lasts

    # ---- warm-up (compiles the conv plan, packs weights)
    for i in range(max(args.warmup, 3)):
        step_resident(i)
        step_e2e(i)
    barrier()

    # ---- device-timed region: one event pair per step, L2 flushed between steps outside the timed intervals
    launches0 = L.launch_count()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    w_timed0 = time.time()
    for i in range(args.steps):
        flush.zero_()
        starts[i].record()
        step_resident(i)
        ends[i].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    w_timed1 = time.time()
    launches = L.launch_count() - launches0
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    dev_ms = reduce_max(dev_ms)

    # ---- end-to-end region: host buffers -> public API -> host result, wall clock between device syncs
    barrier()
    t0 = time.perf_counter()
    w_e2e0 = time.time()
    for i in range(args.steps):
        step_e2e(i)
    barrier()
    e2e_s = reduce_max(time.perf_counter() - t0)
    w_e2e1 = time.time()

    # ---- sustained figure: the same resident step back to back for >= --sustain seconds (no L2 flush: 4 rotating input
    # sets), one event pair around the whole loop, with its own clock record (power-capped clocks differ from the burst)
    sustained = None
    if args.sustain > 0:
        n_sus = max(args.steps, int(args.sustain / max(dev_ms / args.steps / 1e3, 1e-6)) + 1)
        barrier()
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w_sus0 = time.time()
        s_ev.record()
        for i in range(n_sus):
            step_resident(i)
        e_ev.record()
        barrier()
        w_sus1 = time.time()
        sus_ms = reduce_max(s_ev.elapsed_time(e_ev))
        sustained = {"value": round(frames_per_step * n_sus / (sus_ms / 1e3), 3), "unit": UNIT, "steps": n_sus,
                     "seconds": round(sus_ms / 1e3, 3), "ms_per_step": round(sus_ms / n_sus, 4)}
    if rank == 0:
        clocks = sampler.stop([(w_timed0, w_timed1), (w_e2e0, w_e2e1)])
        if sustained is not None:
            sustained["clocks"] = sampler.window_stats([(w_sus0, w_sus1)])
    else:
        clocks = None

    # ---- per-kernel timing for the roofline (rank 0, N=1 semantics): cost-volume kernel and the conv plan, alone
    roof = None
    cpu_base = None
    if rank == 0:
        roof = kernel_rooflines(model, dev_sets, cfg, flush, L)
        ref_gpu = reference_on_gpu(cfg, model, dev) if (world == 1 and not args.no_reference_gpu) else None
        if world == 1 and not args.no_cpu_baseline:
            cpu_base = cpu_baseline(args.workload, budget_s=args.cpu_budget)

    if rank == 0:
        ms_per_step = dev_ms / args.steps
        value = frames_per_step / (ms_per_step / 1e3)
        line = {
            "metric": metric, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"exact": "f32", "tc3x": "f32 via 3xTF32 (tf32 big/small split, fp32 accumulate)",
                      "tch": "f32 via 2-term fp16 split (big + small/2048, three kind::f16 MMAs per product, fp32 accumulate)"}[args.math],
            "data": "synthetic",
            "config": {"workload": workload_desc, "math": args.math, "volume_math": volume_math, "frames_per_step": frames_per_step,
                       "l2": "256 MiB L2 flush between timed steps (outside the per-step event pairs); 4 rotating input sets",
                       "weights": "random-init (seeded), reference architecture",
                       "wall_ms_per_step_incl_flush": round(1e3 * t_wall / args.steps, 4),
                       **({"tuples": len(tuples), "shard_by": args.shard_by, "keyframes_this_rank": len(shard),
                           "gather_every": G, "full_split_seconds_at_this_rate": round(len(tuples) / value, 1)}
                          if shard is not None else {})},
            "e2e": {"value": round(frames_per_step * args.steps / e2e_s, 3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": host_out.numel() * 4},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "sustained": sustained,
            "reference_gpu": ref_gpu,
            "roofline": roof["dominant"] if roof else None,
            "roofline_kernels": roof["all"] if roof else None,
            "cpu_baseline": cpu_base,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic():
    """DRAM bytes per step per kernel family from the committed ncu launch list of this same command
    (profiles/r*_traffic.json, produced by tools/traffic_from_launches.py; the newest file wins); {} when absent."""
    import glob
    out = {}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json"))):
        with open(path) as f:
            out.update(json.load(f))
    return out


def kernel_rooflines(model, dev_sets, cfg, flush, L, reps=9):
    """Live CUDA-event timing of the two kernel families on the launching stream (torch's current stream)."""
    peaks = measured_peaks()
    traffic = ncu_traffic().get(model.math, {})
    cur, src = dev_sets[0]
    dev = flush.device
    ext, pose = model._relative_poses(cur, src, dev)
    mn = torch.tensor(model.run_opts.min_matching_depth).view(1, 1, 1, 1)
    mx = torch.tensor(model.run_opts.max_matching_depth).view(1, 1, 1, 1)

    def time_fn(fn):
        # median of `reps` cold-L2 launches after two untimed ones (the first calls pack weights / capture the graph; a mean
        # let one late lazy initialisation report a 0.85 ms kernel as 1.7 - 7 ms in round 2)
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        return sorted(ts)[len(ts) // 2]

    cv = {}

    def run_cv():
        cv["out"] = model.cost_volume._run(cur["matching_feats_bchw"], src["matching_feats_bkchw"], ext, pose,
                                           src["K_s1_b44"], cur["invK_s1_b44"], mn, mx, cur, None, True)

    cv_ms = time_fn(run_cv)
    # the dot-product volume (CostVolumeManager, the SimpleRecon-style matcher) on the same features: the kernel the north star's
    # HBM target names; not part of the DoubleTake step, timed for the record
    import doubletake_b200 as dt

    dot_mgr = dt.CostVolumeManager(cfg.match_h, cfg.match_w, cfg.planes).to(dev)

    def run_dot():
        dot_mgr._run(cur["matching_feats_bchw"], src["matching_feats_bkchw"], ext, pose, src["K_s1_b44"], cur["invK_s1_b44"], mn, mx,
                     None, None, False)

    dot_ms = time_fn(run_dot)
    plan = model._network_plan(cv["out"]["volume"].shape, cur["image_prior_feats"])
    conv_ms = time_fn(plan.run)
    n_conv = len(plan.ops)
    conv_flops = plan.flops()
    cv_bytes = algorithmic_bytes_cost_volume(cfg, hint=True)
    cv_flops = mlp_flops(cfg)
    # both families are timed ALONE here (one cold-L2 launch / graph replay per event pair), so by the profiling recipe the
    # denominator is the BURST bf16 figure; the sustained one is reported next to it
    tens_peak = peaks["tf_burst"]
    tens_sustained = peaks["tf_sustained"]
    entries = {
        "cost_volume_mlp_hint": {
            "bound": "tensor", "achieved": round(cv_flops / (cv_ms * 1e-3) / 1e12, 3), "peak": tens_peak, "unit": "TFLOP/s",
            "frac": round(cv_flops / (cv_ms * 1e-3) / 1e12 / tens_peak, 5), "traffic": traffic.get("cost_volume"),
            "ms_per_launch": round(cv_ms, 4), "launches_per_step": 2,
            "hbm_view": {"bound": "hbm", "achieved": round(cv_bytes / (cv_ms * 1e-3) / 1e9, 2), "peak": peaks["hbm"],
                         "unit": "GB/s", "frac": round(cv_bytes / (cv_ms * 1e-3) / 1e9 / peaks["hbm"], 5),
                         "algorithmic_bytes": cv_bytes},
            "algorithmic_flops": cv_flops, "peak_source": peaks["src"] + " bf16 burst (kernel timed alone)",
            "frac_of_sustained_peak": round(cv_flops / (cv_ms * 1e-3) / 1e12 / tens_sustained, 5)},
        "conv_stack": {
            "bound": "tensor", "achieved": round(conv_flops / (conv_ms * 1e-3) / 1e12, 3), "peak": tens_peak,
            "unit": "TFLOP/s", "frac": round(conv_flops / (conv_ms * 1e-3) / 1e12 / tens_peak, 5),
            "traffic": traffic.get("conv_stack"),
            "ms_per_launch": round(conv_ms / n_conv, 5), "launches_per_step": n_conv, "ms_all_launches": round(conv_ms, 4),
            "algorithmic_flops": conv_flops, "peak_source": peaks["src"] + " bf16 burst (graph replay timed alone)",
            "frac_of_sustained_peak": round(conv_flops / (conv_ms * 1e-3) / 1e12 / tens_sustained, 5)},
    }
    dot_bytes = algorithmic_bytes_cost_volume(cfg, hint=False)
    entries["cost_volume_dot"] = {
        "bound": "hbm", "achieved": round(dot_bytes / (dot_ms * 1e-3) / 1e9, 2), "peak": peaks["hbm"], "unit": "GB/s",
        "frac": round(dot_bytes / (dot_ms * 1e-3) / 1e9 / peaks["hbm"], 5), "traffic": None, "ms_per_launch": round(dot_ms, 4),
        "launches_per_step": 0, "algorithmic_bytes": dot_bytes, "peak_source": peaks["src"] + " HBM copy",
        "note": "CostVolumeManager (dot-product matcher) at this workload's shapes; instruction-issue bound, not HBM bound "
                "(profiles/r02c_cv_sweep.txt); not part of the timed step"}
    # context for `frac`: an fp32-accurate tensor-core path issues 3 MMAs per product -- TF32 ones at half the bf16 rate
    # (ceiling = peak / 6) or, with the fp16 split, f16 ones at the full rate (ceiling = peak / 3); reported next to the
    # contract's frac-of-bf16-peak, never instead of it
    maths = {"conv_stack": model.math, "cost_volume_mlp_hint": model.volume_math}
    for name, e in entries.items():
        div = {"tc3x": 6.0, "tch": 3.0}.get(maths.get(name))
        if div:
            e["split_ceiling"] = round(tens_peak / div, 1)
            e["frac_of_split_ceiling"] = round(e["achieved"] / (tens_peak / div), 4)
            e["split"] = {"tc3x": "3xTF32", "tch": "2-term fp16, 3 MMAs"}[maths.get(name)]
    dom = "conv_stack" if conv_ms >= cv_ms else "cost_volume_mlp_hint"
    d = dict(entries[dom])
    d["kernel"] = dom
    return {"dominant": d, "all": entries}


# --------------------------------------------------------------------------------------------------------------
# reference arms: the oracle port of the reference's path (a) in eager fp32 torch on the same B200, (b) on the host cores
# --------------------------------------------------------------------------------------------------------------
def oracle_step_fn(cfg_s, device="cpu"):
    """One pass of the reference algorithm (oracle port: the reference's torch ops in the reference's order) over the
    workload ``cfg_s`` with every tensor on ``device``."""
    from oracle import oracle_torch as orc
    import doubletake_b200 as dt

    inp = syn.cost_volume_inputs(cfg_s)
    priors = [p.to(device) for p in syn.prior_features(cfg_s)]
    shapes = {k: tuple(v.shape) for k, v in dt.DepthModelCVHint(model_options(cfg_s)).named_parameters()}
    w = {k: v.to(device) for k, v in syn.seeded_state_dict(shapes, 2024, 1.3).items()}
    eye = torch.eye(4).expand(cfg_s.batch, 4, 4).contiguous()
    cur = {"cam_T_world_b44": eye, "world_T_cam_b44": eye, "invK_s1_b44": inp["cur_invK"], **inp["cv_depth_hint_dict"]}
    src = {"cam_T_world_b44": inp["src_extrinsics"], "world_T_cam_b44": inp["src_poses"], "K_s1_b44": inp["src_Ks"]}
    cur = {k: v.to(device) for k, v in cur.items()}
    src = {k: v.to(device) for k, v in src.items()}
    mc, ms = inp["cur_feats"].to(device), inp["src_feats"].to(device)

    def step():
        return orc.depth_model_forward(mc, ms, priors, cur, src, w, cfg_s.planes, hint=True, decoder=cfg_s.decoder)

    return step


def reference_on_gpu(cfg, model, dev, reps=3):
    """BASELINE.md 3.3 / SURVEY 8d: the reference's own algorithm in eager fp32 PyTorch on the SAME B200 -- the bar a GPU
    user of the reference has today (the per-plane Python loop of the slow manager, cuDNN / cuBLAS kernels, TF32 off).
    Timed like test_no_hint.py:161-175 (synchronise, wall clock around forward), 1 warm-up + `reps` passes, median."""
    prev = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        step = oracle_step_fn(cfg, dev)
        step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            step()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev
    ts.sort()
    med = ts[len(ts) // 2]
    return {"value": round(cfg.batch / med, 3), "unit": UNIT, "ms_per_step": round(med * 1e3, 2), "reps": reps,
            "kind": "port: oracle/oracle_torch.py (the reference's torch ops in the reference's order) in eager fp32 on this GPU, "
                    "encoders' outputs supplied like in the B200 arm"}


def cpu_sample_config(workload, width_div):
    c = syn.CONFIGS[workload]
    return syn.WorkloadConfig(f"{workload}_w{width_div}", c.batch, c.num_src, c.image_h, c.image_w // width_div, c.planes,
                              hint=True, prior_ch=c.prior_ch, decoder=c.decoder, seed=c.seed)


def cpu_step_fn(cfg_s):
    """One pass of the reference algorithm (oracle port) over a 1/width_div-width crop of the frame: identical per-pixel
    work (all planes, all views, full conv stack), fewer pixels."""
    return oracle_step_fn(cfg_s, "cpu")


def pick_cpu_sample(workload, total_steps, budget_s):
    """Probe a 1/10-width crop, then choose the largest crop (width stays a multiple of 32) whose total run fits
    the budget."""
    ncpu = os.cpu_count() or 1
    probe = cpu_step_fn(cpu_sample_config(workload, 16 if workload == "cfg5" else 10))
    # "all the host threads it can use": torch's CPU ops on this path stop scaling (and regress) well below the
    # core count of the GPU hosts, so probe a few thread counts and keep the fastest
    best = None
    for n in sorted({ncpu, max(1, ncpu // 2), max(1, ncpu // 4), min(ncpu, 32), min(ncpu, 16), min(ncpu, 8)}, reverse=True):
        torch.set_num_threads(n)
        probe()
        t0 = time.perf_counter()
        probe()
        t = time.perf_counter() - t0
        if best is None or t < best[0]:
            best = (t, n)
    t10, cores = best
    probe_div = 16 if workload == "cfg5" else 10
    torch.set_num_threads(cores)
    for div in (1, 2, 4, 8, 16, 32):
        if syn.CONFIGS[workload].image_w % (32 * div) and div != 1:
            continue
        if t10 * (probe_div / div) * total_steps <= budget_s or div == 32:
            return div, cores
    return 32, cores


def cpu_baseline(workload, budget_s=25.0, reps=3):
    """1 warm-up + >= 3 timed passes, median (and min) reported."""
    torch.set_grad_enabled(False)
    c = syn.CONFIGS[workload]
    div, cores = pick_cpu_sample(workload, reps + 1, budget_s)
    step = cpu_step_fn(cpu_sample_config(workload, div))
    step()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return {"value": round((c.batch / div) / med, 5), "unit": UNIT, "cores": cores, "kind": "port",
            "min_s": round(ts[0], 3), "median_s": round(med, 3), "reps": reps,
            "sample": f"oracle (torch CPU restatement of the reference path), 1 warm-up + {reps} timed passes (median) over a "
                      f"1/{div}-width crop of the {workload} batch ({c.image_h}x{c.image_w // div} image, {c.planes} planes, "
                      f"{c.num_src} views, hint, full conv stack, batch {c.batch}) = {c.batch / div:.3f} frame(s) in {med:.2f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    torch.set_grad_enabled(False)
    warm = max(args.warmup, 1)
    c = syn.CONFIGS[args.workload]
    metric, workload_desc = WORKLOADS[args.workload]
    div, cores = pick_cpu_sample(args.workload, args.steps + warm, args.ref_budget)
    step = cpu_step_fn(cpu_sample_config(args.workload, div))
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt_s = time.perf_counter() - t0
    value = args.steps * (c.batch / div) / dt_s
    sample = (f"each step = oracle port of the reference path over a 1/{div}-width crop of the {args.workload} batch "
              f"({c.image_h}x{c.image_w // div} image, {c.planes} planes, {c.num_src} views, hint, full conv stack, batch {c.batch}) "
              f"= {c.batch / div:.3f} frame(s)")
    line = {
        "impl": "reference", "metric": metric, "value": round(value, 5), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": round(1e3 * dt_s / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc, "math": "f32 (reference algorithm, torch CPU ops, host threads)",
                   "sample": sample},
        "cpu_baseline": {"value": round(value, 5), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--math", default="tch", choices=["exact", "tc3x", "tch"],
                    help="tch: tcgen05 kind::f16 with the 2-term fp16 split of activations and weights (fp32-class, parity-green, "
                         "default); tc3x: 3xTF32 split; exact: fp32 CUDA cores")
    ap.add_argument("--volume-math", default=None, choices=["exact", "tc3x", "tch"],
                    help="cost-volume MLP arithmetic; default: tch (kind::f16, 2-term fp16 split) unless --math exact")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS),
                    help="cfg2 = the headline metric's configuration (default); cfg3 / cfg5 = BASELINE.json configs[2] / [4]")
    ap.add_argument("--gather-every", type=int, default=16, help="cfg4: keyframes per rank between two depth-map gathers")
    ap.add_argument("--shard-by", default="frame", choices=["frame", "scan"],
                    help="cfg4: round-robin keyframes, or whole scans per rank (the incremental mode's constraint)")
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of back-to-back steps for the `sustained` figure (0 = off)")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip the eager-PyTorch-on-this-GPU reference timing")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    ap.add_argument("--ref-budget", type=float, default=150.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
