"""GPU parity of the fused conv plans (CVEncoder, DepthDecoderPP, SkipDecoderRegression) through the C ABI."""
import json

import pytest
import torch

import helpers as hp
import doubletake_b200 as dt
from doubletake_b200 import _lib as L
from doubletake_b200 import synthetic as syn
from oracle import oracle_torch as orc

pytestmark = pytest.mark.gpu
# feature-map tolerance (relative to the map's max): exact = fp32 FMA in a fixed order; tc3x = 3xTF32 on tcgen05, whose
# fp32 TMEM accumulation truncates instead of rounding (measured ~2e-5 over K up to 5760).  The contract that matters,
# depth within 1e-4 relative, is asserted on the network outputs below and in test_gpu_model.py for both modes.
FEAT_TOL = {"exact": 1e-5, "tc3x": 5e-5, "tch": 5e-5}
torch.set_grad_enabled(False)
DEV = "cuda"


def build(fx, decoder, D, prior_ch, seed, math="exact"):
    enc = dt.CVEncoder(D, list(prior_ch[1:]), [64, 128, 256, 384], math=math)
    dec_in = list(prior_ch[:1]) + enc.num_ch_enc
    dec = (dt.DepthDecoderPP(dec_in, math=math) if decoder == "unet_pp" else dt.SkipDecoderRegression(dec_in, math=math))
    encw = syn.seeded_state_dict(json.loads(str(fx["enc_shapes"])), seed + 1, 1.5)
    decw = syn.seeded_state_dict(json.loads(str(fx["dec_shapes"])), seed + 2, 1.5)
    enc.load_state_dict(encw)
    dec.load_state_dict(decw)
    return enc.to(DEV), dec.to(DEV), encw, decw


@pytest.mark.parametrize("math", ["exact", "tc3x", "tch"])
@pytest.mark.parametrize("name,decoder", [("net_pp_d64", "unet_pp"), ("net_pp_d16_b2", "unet_pp"),
                                           ("net_skip_d48", "skip")])
def test_conv_stacks_match_reference_fixture(name, decoder, math):
    fx = hp.load(name)
    cfg, cv, priors, seed = hp.network_case_inputs(fx, decoder)
    enc, dec, _, _ = build(fx, decoder, cfg.planes, cfg.prior_ch, seed, math=math)
    cvf = enc(cv.to(DEV), [p.to(DEV) for p in priors[1:]])
    for i, f in enumerate(cvf):
        assert f.shape == fx[f"out.cv_feat_{i}"].shape
        assert hp.rel_err(f.cpu(), fx[f"out.cv_feat_{i}"]) < FEAT_TOL[math]
    out = dec([priors[0].to(DEV)] + cvf)
    for i in range(4):
        k = f"log_depth_pred_s{i}_b1hw"
        ref = torch.from_numpy(fx["out." + k])
        assert out[k].shape == ref.shape
        assert float((out[k].cpu() - ref).abs().max()) < 1e-4, k
    if decoder == "skip":
        assert hp.rel_err(out["feature_s3_b1hw"].cpu(), fx["out.feature_s3_b1hw"]) < FEAT_TOL[math]


@pytest.mark.parametrize("math", ["exact", "tc3x", "tch"])
@pytest.mark.parametrize("ih,iw,B", [(96, 160, 1), (160, 96, 2)])
def test_odd_sizes_against_oracle(ih, iw, B, math):
    """Sizes whose /32 maps are odd (3x5, 5x3): partial 8x8 tiles, stride-2 convs on odd inputs."""
    fx = hp.load("net_pp_d64")
    prior_ch = (24, 48, 64, 160, 256)
    cfg = syn.WorkloadConfig("t", B, 2, ih, iw, 64, prior_ch=prior_ch, seed=51)
    priors = syn.prior_features(cfg)
    cv = torch.randn(B, 64, ih // 4, iw // 4, generator=torch.Generator().manual_seed(3))
    enc, dec, encw, decw = build(fx, "unet_pp", 64, prior_ch, 700, math=math)
    cvf = enc(cv.to(DEV), [p.to(DEV) for p in priors[1:]])
    out = dec([priors[0].to(DEV)] + cvf)
    rcvf = orc.cv_encoder(cv, priors[1:], encw)
    ref = orc.depth_decoder_pp(priors[:1] + rcvf, decw)
    for i in range(4):
        assert hp.rel_err(cvf[i].cpu(), rcvf[i]) < FEAT_TOL[math]
        k = f"log_depth_pred_s{i}_b1hw"
        assert float((out[k].cpu() - ref[k]).abs().max()) < 1e-4


@pytest.mark.parametrize("math", ["exact", "tc3x", "tch"])
def test_single_conv_features_against_torch(math):
    """Each fused feature of dtb200_conv2d in isolation: concat of 3 sources, bilinear / nearest x2 on load, stride 2,
    1x1, residual, LeakyReLU / ELU -- against F.conv2d on CPU."""
    import torch.nn as nn
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    B, H, W = 2, 10, 12
    a = torch.randn(B, 16, H, W, generator=g)
    b = torch.randn(B, 24, H // 2, W // 2, generator=g)
    c = torch.randn(B, 8, H // 2, W // 2, generator=g)
    res = torch.randn(B, 64, H, W, generator=g)
    conv = nn.Conv2d(48, 64, 3, padding=1)
    plan = dt.ConvPlan(torch.device(DEV), math)
    fa, fb, fc = plan.input("a", *a.shape), plan.input("b", *b.shape), plan.input("c", *c.shape)
    fr = plan.input("r", *res.shape)
    o1 = plan.conv([(fa, L.RESAMPLE_NONE), (fb, L.RESAMPLE_BILINEAR_UP2), (fc, L.RESAMPLE_NEAREST_UP2)], conv,
                   L.ACT_LEAKY, 0.2, residual=fr)
    conv2 = nn.Conv2d(64, 128, 3, stride=2, padding=1)
    o2 = plan.conv([(o1, L.RESAMPLE_NONE)], conv2, L.ACT_ELU)
    conv3 = nn.Conv2d(128, 64, 1)
    o3 = plan.conv([(o2, L.RESAMPLE_NONE)], conv3)
    head = nn.Conv2d(64, 1, 1)
    o4 = plan.conv([(o3, L.RESAMPLE_NONE)], head)
    plan.finalize()
    plan.load_inputs({"a": a.to(DEV), "b": b.to(DEV), "c": c.to(DEV), "r": res.to(DEV)})
    plan.run()
    x = torch.cat([a, F.interpolate(b, scale_factor=2, mode="bilinear", align_corners=False),
                   F.interpolate(c, scale_factor=2, mode="nearest")], 1)
    r1 = F.leaky_relu(conv(x) + res, 0.2)
    r2 = F.elu(conv2(r1))
    r3 = conv3(r2)
    r4 = head(r3)
    for o, r in ((o1, r1), (o2, r2), (o3, r3), (o4, r4)):
        got = plan.output_nchw(o).cpu()
        assert got.shape == r.shape
        assert hp.rel_err(got, r) < FEAT_TOL[math]


@pytest.mark.parametrize("math", ["exact", "tc3x", "tch"])
def test_graph_mode_is_bit_identical_to_stream_order(math, monkeypatch):
    """The compiled-network runtime (csrc/conv_graph.cu) only reorders INDEPENDENT launches: outputs must not change by a
    bit against the same descriptors launched in program order, run after run."""
    fx = hp.load("net_pp_d64")
    cfg, cv, priors, seed = hp.network_case_inputs(fx, "unet_pp")
    outs = {}
    for mode in ("sequence", "graph"):
        monkeypatch.setenv("DTB200_CONV_MODE", mode)
        enc, dec, _, _ = build(fx, "unet_pp", cfg.planes, cfg.prior_ch, seed, math=math)
        runs = []
        for _ in range(3):
            cvf = enc(cv.to(DEV), [p.to(DEV) for p in priors[1:]])
            out = dec([priors[0].to(DEV)] + cvf)
            runs.append([f.clone() for f in cvf] + [out[f"log_depth_pred_s{i}_b1hw"].clone() for i in range(4)])
        for r in runs[1:]:
            assert all(torch.equal(a, b) for a, b in zip(runs[0], r)), mode
        outs[mode] = runs[0]
        plan = next(iter(dec._plans.values()))
        if mode == "graph":
            info = plan.graph_info()
            assert info["ops"] == len(plan.ops) and info["kernels"] >= info["ops"]
            assert info["lanes"] > 1 and info["depth"] < info["ops"]
        else:
            assert plan.graph_info() is None
    assert all(torch.equal(a, b) for a, b in zip(outs["sequence"], outs["graph"]))


@pytest.mark.parametrize("math", ["tc3x", "tch"])
@pytest.mark.parametrize("B,H,W,chans,oc", [(1, 130, 165, (40, 24), 64), (2, 72, 88, (64,), 64), (1, 120, 160, (64, 48), 64)])
def test_large_map_3x3_conv_against_torch(B, H, W, chans, oc, math):
    """3x3 / stride-1 convs on maps with more tiles than SMs (the shapes the halo-tile tensor-core kernel serves): ragged
    right / bottom tiles, a concat with a channel count that is not a multiple of 32, bias + residual + LeakyReLU."""
    import torch.nn as nn
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(7)
    xs = [torch.randn(B, c, H, W, generator=g) for c in chans]
    res = torch.randn(B, oc, H, W, generator=g)
    conv = nn.Conv2d(sum(chans), oc, 3, padding=1)
    plan = dt.ConvPlan(torch.device(DEV), math)
    fs = [plan.input(f"x{i}", *x.shape) for i, x in enumerate(xs)]
    fr = plan.input("r", *res.shape)
    o = plan.conv([(f, L.RESAMPLE_NONE) for f in fs], conv, L.ACT_LEAKY, 0.2, residual=fr)
    plan.finalize()
    plan.load_inputs({**{f"x{i}": x.to(DEV) for i, x in enumerate(xs)}, "r": res.to(DEV)})
    plan.run()
    want = F.leaky_relu(conv(torch.cat(xs, 1)) + res, 0.2)
    got = plan.output_nchw(o).cpu()
    assert got.shape == want.shape
    assert hp.rel_err(got, want) < FEAT_TOL[math]
    # per-pixel check too: a wrong tap or a shifted row shows up as a large error on few pixels, not in the max norm only
    assert float((got - want).abs().max()) < 5e-5 * float(want.abs().max())


@pytest.mark.parametrize("chans,with_res", [((64,), True), ((24,), False), ((64,), False)])
def test_resident_weights_halo_variant_against_torch(chans, with_res):
    """The resident-weights halo kernel (development switch bit 4: weights of a <= 64-channel 3x3 layer stay in shared memory,
    two MMA-issuing warps) on a map with >= 3 tiles per SM and ragged right / bottom tiles; the default (streaming) kernel on
    the same inputs must agree with it to fp32 round-off."""
    import torch.nn as nn
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(11)
    H, W, oc = 200, 330, 64
    x = torch.randn(1, chans[0], H, W, generator=g)
    res = torch.randn(1, oc, H, W, generator=g)
    conv = nn.Conv2d(chans[0], oc, 3, padding=1)
    want = conv(x) + res if with_res else conv(x)
    want = F.leaky_relu(want, 0.2)
    outs = {}
    try:
        for flags in (16, 0):
            L.check(L.lib().dtb200_debug_set(flags))
            plan = dt.ConvPlan(torch.device(DEV), "tch")
            fx = plan.input("x", *x.shape)
            fr = plan.input("r", *res.shape) if with_res else None
            o = plan.conv([(fx, L.RESAMPLE_NONE)], conv, L.ACT_LEAKY, 0.2, residual=fr)
            plan.finalize()
            plan.load_inputs({"x": x.to(DEV), **({"r": res.to(DEV)} if with_res else {})})
            before = L.launch_count()
            plan.run()
            torch.cuda.synchronize()
            assert L.launch_count() > before
            outs[flags] = plan.output_nchw(o).cpu()
    finally:
        L.check(L.lib().dtb200_debug_set(0))
    for got in outs.values():
        assert hp.rel_err(got, want) < FEAT_TOL["tch"]
        assert float((got - want).abs().max()) < 5e-5 * float(want.abs().max())
    assert hp.rel_err(outs[16], outs[0]) < 2e-6


def test_standalone_mlp_and_basic_block_forward():
    """the reference's MLP and BasicBlock are callable modules; so are the mirrors (one-plan evaluation)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    mlp = dt.MLP([202, 128, 128, 1], disable_final_activation=True)
    x = torch.randn(5, 7, 202, generator=g)
    want = mlp.net[4](F.leaky_relu(mlp.net[2](F.leaky_relu(mlp.net[0](x), 0.01)), 0.01))
    got = mlp.to(DEV)(x.to(DEV))
    assert got.shape == (5, 7, 1) and hp.rel_err(got.cpu(), want) < 1e-5
    for cin, cout, stride in ((64, 64, 1), (48, 64, 1), (64, 128, 2)):
        blk = dt.BasicBlock(cin, cout, stride)
        xb = torch.randn(2, cin, 20, 28, generator=g)
        t = F.leaky_relu(blk.conv1(xb), 0.2)
        skip = xb if blk.downsample is None else blk.downsample[0](xb)
        want = F.leaky_relu(blk.conv2(t) + skip, 0.2)
        got = blk.to(DEV)(xb.to(DEV))
        assert got.shape == want.shape and hp.rel_err(got.cpu(), want) < 1e-5
