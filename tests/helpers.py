"""Shared test helpers: golden-fixture loading, seeded input/weight regeneration, comparison metrics."""
import os

import numpy as np
import torch

from doubletake_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

MLP_KEYS = lambda F: {  # noqa: E731
    "mlp.net.0.weight": (128, F), "mlp.net.0.bias": (128,),
    "mlp.net.2.weight": (128, 128), "mlp.net.2.bias": (128,),
    "mlp.net.4.weight": (1, 128), "mlp.net.4.bias": (1,),
}
HINT_KEYS = {
    "hint_mlp.net.0.weight": (12, 3), "hint_mlp.net.0.bias": (12,),
    "hint_mlp.net.2.weight": (12, 12), "hint_mlp.net.2.bias": (12,),
    "hint_mlp.net.4.weight": (1, 12), "hint_mlp.net.4.bias": (1,),
}


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def volume_weights(kind, K, C, seed):
    """Same parameters make_golden.py loaded into the reference manager (sorted-key seeded draw)."""
    F = (C + 10) * K + C + 4
    shapes = dict(MLP_KEYS(F))
    if kind == "hint":
        shapes.update(HINT_KEYS)
    return syn.seeded_state_dict(shapes, seed)


def volume_case_inputs(fx, kind):
    """Rebuild the kwargs make_golden.volume_case used, from the fixture (stored inputs or recorded seed)."""
    B, K, C, H, W, D, seed, wseed, white, empty, sub = [int(v) for v in fx["meta"]]
    if "in.cur_feats" in fx:
        inp = {k[3:]: torch.from_numpy(v) for k, v in fx.items() if k.startswith("in.") and not k.startswith("in.hint.")}
        hint = {k[8:]: torch.from_numpy(v) for k, v in fx.items() if k.startswith("in.hint.")}
        if hint:
            inp["cv_depth_hint_dict"] = hint
    else:
        cfg = syn.WorkloadConfig("fx", B, K, 0, 0, D, feat_ch=C, hint=(kind == "hint"), seed=seed)
        inp = syn.cost_volume_inputs(cfg, white=bool(white), empty_hint=bool(empty), match_hw=(H, W))
    if kind != "hint":
        inp.pop("cv_depth_hint_dict", None)
    weights = volume_weights(kind, K, C, wseed) if kind != "dot" else None
    return inp, weights, dict(B=B, K=K, C=C, H=H, W=W, D=D, sub=sub)


def rel_err(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def argmax_mismatch_report(volume, ref_index, tol=1e-5):
    """Indices must match; a mismatch is tolerated only on a proven near-tie: the two candidate planes'
    values in OUR volume differ by <= tol * max|volume| (SURVEY.md §7 'bit-exact arg-max').
    Returns (n_mismatch, n_unexplained)."""
    volume = torch.as_tensor(volume)
    ref_index = torch.as_tensor(ref_index).long()
    ours = torch.argmax(volume, 1)
    bad = ours != ref_index
    n_bad = int(bad.sum())
    if n_bad == 0:
        return 0, 0
    v_ours = torch.gather(volume, 1, ours.unsqueeze(1)).squeeze(1)
    v_ref = torch.gather(volume, 1, ref_index.unsqueeze(1)).squeeze(1)
    gap = (v_ours - v_ref).abs()[bad]
    scale = float(volume.abs().max())
    argmax_mismatch_report.last = dict(n_bad=n_bad, max_rel_gap=float(gap.max()) / scale, scale=scale)
    return n_bad, int((gap > tol * scale).sum())


def network_case_inputs(fx, decoder):
    B, D, ih, iw, seed = [int(v) for v in fx["meta"]]
    prior_ch = tuple(int(v) for v in fx["prior_ch"])
    cfg = syn.WorkloadConfig("fx", B, 2, ih, iw, D, prior_ch=prior_ch, decoder=decoder, seed=seed)
    priors = syn.prior_features(cfg)
    g = torch.Generator().manual_seed(seed + 11)
    cv = torch.randn(B, D, ih // 4, iw // 4, generator=g)
    return cfg, cv, priors, seed


def log_argmax(tag, n_bad, n_total, extra=""):
    """One line per parity case with the arg-max mismatch COUNT (record it, do not hide it behind an allowance):
    printed (pytest -s / -rP shows it) and appended to gpurun_out/argmax_mismatch.log when that directory exists."""
    line = f"[argmax] {tag}: {n_bad} of {n_total} pixels differ from the reference arg-max {extra}".rstrip()
    print(line)
    out = os.path.join(os.path.dirname(GOLDEN), "..", "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "argmax_mismatch.log"), "a") as f:
            f.write(line + "\n")
