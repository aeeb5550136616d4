"""B200 drop-ins for the reference's cost-volume encoder and depth decoders.

``CVEncoder`` (reference modules/networks.py:88-117), ``DepthDecoderPP`` (modules/networks.py:20-85),
``SkipDecoderRegression`` (modules/networks_fast.py:98-141) and ``BasicBlock`` (modules/layers.py:33-94) keep the
reference's constructor arguments, ``forward`` signatures (NCHW tensors in and out) and ``state_dict`` keys, so
checkpoints load unchanged and the modules are swapped in as ``model.cost_volume_net`` / ``model.depth_decoder``.

The modules only HOLD parameters.  ``forward`` compiles, once per input shape, a ``ConvPlan`` -- an array of fused
convolution descriptors (include/doubletake_b200.h: ``dtb200_conv_params``) over channels-last buffers in which
``torch.cat``, the x2 up-samples, bias, residual add and activation are folded into the convolutions -- and replays
it through ONE native call (``dtb200_conv2d_sequence``).  No PyTorch op touches an activation; no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L


# ----------------------------------------------------------------------------------------------------------------
# plan builder
# ----------------------------------------------------------------------------------------------------------------
class Feature:
    """A channels-last activation buffer (B,H,W,C) inside a plan.  ``fmt``: "f32" = plain fp32 NHWC; "split16" = the
    tensor-core layout of math "tch", (B,H,W,2,C) fp16 big | small planes in the same bytes (include/doubletake_b200.h)."""

    __slots__ = ("t", "b", "h", "w", "c", "fmt")

    def __init__(self, t, fmt="f32"):
        self.t = t
        self.b, self.h, self.w, self.c = t.shape
        self.fmt = fmt


class ConvPlan:
    """Records fused-conv descriptors over pre-allocated NHWC buffers; ``run()`` launches them from native code."""

    def __init__(self, device, math="exact"):
        self.device = device
        self.math = L.MATH_NAMES[math]
        self.ops = []
        self.keep = []  # tensors referenced by raw pointers
        self.inputs = {}  # name -> Feature filled from an NCHW tensor before run()
        self._packed = {}
        self._tracked = []  # (parameter, version at pack time): stale() tells the owner to rebuild the plan
        self._upsampled = {}
        self._array = None
        self.workspace = None
        self._graph = None
        # execution mode: "graph" = dependency-DAG CUDA graph (independent layers overlap), "sequence" = stream order
        self.mode = os.environ.get("DTB200_CONV_MODE", "graph")
        self.max_lanes = int(os.environ.get("DTB200_CONV_LANES", "8"))
        self.workspace_slots = int(os.environ.get("DTB200_CONV_WS_SLOTS", "6"))

    def new(self, b, h, w, c, fmt=None):
        """A fresh buffer.  In a "tch" plan every map with a multiple of 8 channels is split16 (same bytes as fp32)."""
        if fmt is None:
            fmt = "split16" if (self.math == L.MATH_TCH and c % 8 == 0) else "f32"
        t = torch.empty((b, h, w, c), dtype=torch.float32, device=self.device)
        self.keep.append(t)
        return Feature(t, fmt)

    def input(self, name, b, c, h, w):
        if self.math == L.MATH_TCH and c % 8 != 0:
            raise ValueError(f"math='tch' needs input maps with a multiple of 8 channels, {name} has {c}")
        f = self.new(b, h, w, c)
        self.inputs[name] = f
        return f

    def set_feature(self, f, x_nchw):
        """Fill a plan buffer from an NCHW fp32 CUDA tensor in the buffer's own layout."""
        if tuple(x_nchw.shape) != (f.b, f.c, f.h, f.w):
            raise ValueError(f"expected {(f.b, f.c, f.h, f.w)}, got {tuple(x_nchw.shape)}")
        if not x_nchw.is_cuda:
            raise RuntimeError("doubletake_b200 networks run on CUDA only (no CPU fallback)")
        if f.fmt == "split16":
            L.nchw_to_split16(x_nchw, f.t)
        else:
            L.nchw_to_nhwc(x_nchw, out=f.t)

    def _pack(self, weight, src_channels):
        """Device copy of an OIHW weight in the layout the math mode consumes, for this concat split of its inputs."""
        key = (weight.data_ptr(), weight._version, tuple(src_channels))
        if key not in self._packed:
            self._tracked.append((weight, weight._version))
            w = L.f32(weight.detach(), self.device)
            oc, ic, k, _ = w.shape
            split = (C.c_int32 * len(src_channels))(*src_channels)
            n = int(L.lib().dtb200_packed_conv_weight_floats_srcs(self.math, oc, len(src_channels), split, k))
            packed = torch.empty(n, dtype=torch.float32, device=self.device)
            L.check(L.lib().dtb200_pack_conv_weight_srcs(self.math, L.ptr(w), L.ptr(packed), oc, len(src_channels), split, k,
                                                         L.stream()))
            self._packed[key] = packed
            self.keep.append(packed)
        return self._packed[key]

    def upsampled(self, f: Feature, mode):
        """Materialise the x2 up-sampled map once (resample descriptor, ksize = 0) and share it between consumers."""
        key = (f.t.data_ptr(), mode)  # buffers live as long as the plan (self.keep), so the pointer is a stable key
        if key not in self._upsampled:
            out = self.new(f.b, 2 * f.h, 2 * f.w, f.c)
            op = L.ConvParams()
            op.math, op.batch = self.math, f.b
            op.in_h = op.out_h = 2 * f.h
            op.in_w = op.out_w = 2 * f.w
            op.out_c, op.ksize, op.stride, op.num_src = f.c, 0, 1, 1
            op.src[0], op.src_c[0], op.src_resample[0] = L.ptr(f.t), f.c, mode
            op.dst = L.ptr(out.t)
            self.ops.append(op)
            self._upsampled[key] = out
        return self._upsampled[key]

    def conv(self, srcs, conv: nn.Conv2d, act=L.ACT_NONE, slope=0.0, residual=None):
        """srcs: list of (Feature, resample).  Returns the output Feature."""
        if self.math in (L.MATH_TC3X, L.MATH_TCH) and conv.out_channels % 64 == 0:
            # tensor-core path: interpolate once, not once per tap per consumer
            srcs = [(self.upsampled(f, r), L.RESAMPLE_NONE) if r != L.RESAMPLE_NONE else (f, r) for f, r in srcs]
        k = conv.kernel_size[0]
        stride = conv.stride[0]
        f0, r0 = srcs[0]
        in_h = f0.h * (2 if r0 != L.RESAMPLE_NONE else 1)
        in_w = f0.w * (2 if r0 != L.RESAMPLE_NONE else 1)
        pad = k // 2
        out_h = (in_h + 2 * pad - k) // stride + 1
        out_w = (in_w + 2 * pad - k) // stride + 1
        total_c = 0
        op = L.ConvParams()
        op.math, op.batch = self.math, f0.b
        op.in_h, op.in_w, op.out_h, op.out_w, op.out_c = in_h, in_w, out_h, out_w, conv.out_channels
        op.ksize, op.stride, op.num_src = k, stride, len(srcs)
        for i, (f, r) in enumerate(srcs):
            fh = f.h * (2 if r != L.RESAMPLE_NONE else 1)
            fw = f.w * (2 if r != L.RESAMPLE_NONE else 1)
            if (fh, fw) != (in_h, in_w) or f.b != f0.b:
                raise ValueError(f"concat sources disagree on shape: {(fh, fw)} vs {(in_h, in_w)}")
            op.src[i], op.src_c[i], op.src_resample[i] = L.ptr(f.t), f.c, r
            total_c += f.c
        if total_c != conv.in_channels:
            raise ValueError(f"conv expects {conv.in_channels} input channels, sources provide {total_c}")
        op.weight = L.ptr(self._pack(conv.weight, [f.c for f, _ in srcs]))
        if conv.bias is not None:
            # snapshot, like the packed weights: a plan never mixes stale weights with live biases
            bias = L.f32(conv.bias.detach(), self.device).clone()
            self.keep.append(bias)
            self._tracked.append((conv.bias, conv.bias._version))
            op.bias = L.ptr(bias)
        if self.math == L.MATH_TCH:
            if any(f.fmt != "split16" for f, _ in srcs):
                raise ValueError("math='tch': every conv source must be a split16 map (a multiple of 8 channels)")
            out = self.new(f0.b, out_h, out_w, conv.out_channels, "split16" if conv.out_channels % 64 == 0 else "f32")
        else:
            out = self.new(f0.b, out_h, out_w, conv.out_channels)
        if residual is not None:
            if (residual.b, residual.h, residual.w, residual.c) != (out.b, out.h, out.w, out.c) or residual.fmt != out.fmt:
                raise ValueError("residual shape / layout mismatch")
            op.residual = L.ptr(residual.t)
        op.act, op.act_slope = act, slope
        op.dst = L.ptr(out.t)
        self.ops.append(op)
        return out

    def finalize(self):
        """Hand out split-K scratch and freeze the descriptor array.  Ops that need scratch rotate over a small pool of
        slots: two ops on the same slot are serialised by the graph's dependency analysis, ops on different slots may
        overlap (in "sequence" mode one slot serves everything)."""
        needs = [int(L.lib().dtb200_conv_workspace_bytes(C.byref(op))) for op in self.ops]
        need = (max(needs + [0]) + 255) // 256 * 256
        users = [i for i, n in enumerate(needs) if n]
        if users:
            slots = max(1, min(self.workspace_slots, len(users))) if self.mode == "graph" else 1
            self.workspace = torch.empty(need * slots, dtype=torch.uint8, device=self.device)
            base = L.ptr(self.workspace)
            for n, i in enumerate(users):
                self.ops[i].workspace, self.ops[i].workspace_bytes = base + (n % slots) * need, need
        self._array = (L.ConvParams * len(self.ops))(*self.ops)
        return self

    def stale(self):
        """True when a parameter this plan snapshot (packed weights, bias copies) was modified in place afterwards
        (``param.data.copy_``, EMA, a fine-tune step): the owner rebuilds the plan instead of running stale weights."""
        return any(t._version != v for t, v in self._tracked)

    def analyze(self):
        """Host-side dependency analysis of the plan (no launch): per-op lane, level and direct dependencies."""
        n = len(self.ops)
        lane, level, off = (C.c_int32 * n)(), (C.c_int32 * n)(), (C.c_int32 * (n + 1))()
        cap = n * 8
        deps = (C.c_int32 * cap)()
        L.check(L.lib().dtb200_conv_graph_analyze(self._array, n, self.max_lanes, lane, level, off, deps, cap))
        return [dict(lane=lane[i], level=level[i], deps=list(deps[off[i]:off[i + 1]])) for i in range(n)]

    def graph_info(self):
        if self._graph is None:
            return None
        v = [C.c_int32() for _ in range(5)]
        L.check(L.lib().dtb200_conv_graph_info(self._graph, *[C.byref(x) for x in v]))
        return dict(zip(("ops", "kernels", "edges", "lanes", "depth"), (x.value for x in v)))

    def __del__(self):
        g, self._graph = getattr(self, "_graph", None), None
        if g is not None:
            try:
                L.lib().dtb200_conv_graph_destroy(g)
            except Exception:
                pass

    def flops(self):
        total = 0
        for op in self.ops:
            if op.ksize == 0:
                continue
            cin = sum(op.src_c[i] for i in range(op.num_src))
            total += 2 * op.batch * op.out_h * op.out_w * op.out_c * cin * op.ksize * op.ksize
        return total

    def load_inputs(self, tensors: dict):
        for name, x in tensors.items():
            f = self.inputs[name]
            if tuple(x.shape) != (f.b, f.c, f.h, f.w):
                raise ValueError(f"plan input {name}: expected {(f.b, f.c, f.h, f.w)}, got {tuple(x.shape)}")
            self.set_feature(f, x)

    def output_nchw(self, f):
        """A plan buffer as a fresh NCHW fp32 tensor (what the reference-facing ``forward`` methods return)."""
        return _nchw_out(f)

    def run(self):
        if self.mode != "graph":
            L.check(L.lib().dtb200_conv2d_sequence(self._array, len(self.ops), L.stream()))
            return
        if self._graph is None:
            # weights were packed on the current stream; the capture itself executes nothing
            g = L.fp()
            L.check(L.lib().dtb200_conv_graph_create(self._array, len(self.ops), self.max_lanes, C.byref(g)))
            self._graph = g
        L.check(L.lib().dtb200_conv_graph_launch(self._graph, L.stream()))


def _nchw_out(f: Feature):
    """Plan output -> fresh NCHW fp32 tensor (1-channel fp32 maps are a pure reshape)."""
    if f.fmt == "split16":
        return L.split16_to_nchw(f.t, f.b, f.c, f.h, f.w)
    if f.c == 1:
        return f.t.reshape(f.b, 1, f.h, f.w).clone()
    return L.nhwc_to_nchw(f.t)


# ----------------------------------------------------------------------------------------------------------------
# parameter-holding mirrors of the reference modules
# ----------------------------------------------------------------------------------------------------------------
class BasicBlock(nn.Module):
    """reference modules/layers.py:33-94 with norm_layer=Identity (bias on every conv, LeakyReLU(0.2), optional
    1x1 / strided-3x3 projection on the skip).  Keys: conv1, conv2, downsample.0."""

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=True)
        if inplanes == planes and stride == 1:
            self.downsample = None
        else:
            k = 1 if stride == 1 else 3
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes, k, stride=stride, padding=k // 2, bias=True),
                                            nn.Identity())
        self.stride = stride

    def emit(self, plan: ConvPlan, srcs):
        """layers.py:77-94 as 2-3 fused convs: t = lrelu(conv1(x)); out = lrelu(conv2(t) + (x | proj(x)))."""
        t = plan.conv(srcs, self.conv1, L.ACT_LEAKY, 0.2)
        if self.downsample is not None:
            skip = plan.conv(srcs, self.downsample[0])
        else:
            if len(srcs) != 1 or srcs[0][1] != L.RESAMPLE_NONE:
                raise ValueError("identity skip needs a single, non-resampled source")
            skip = srcs[0][0]
        return plan.conv([(t, L.RESAMPLE_NONE)], self.conv2, L.ACT_LEAKY, 0.2, residual=skip)

    def forward(self, x):
        """Standalone evaluation (reference modules/layers.py:77-94): NCHW in, NCHW out, through a plan of this block's 2-3
        descriptors.  Inside CVEncoder / DepthDecoderPP the block is emitted into the parent's plan instead."""
        math = getattr(self, "math", "exact")
        key = (tuple(x.shape), str(x.device), math, tuple(p._version for p in self.parameters()))
        cache = self.__dict__.setdefault("_plan", {})
        if key not in cache:
            cache.clear()
            plan = ConvPlan(x.device, math)
            fx = plan.input("x", *x.shape)
            out = self.emit(plan, [(fx, L.RESAMPLE_NONE)])
            cache[key] = (plan.finalize(), out)
        plan, out = cache[key]
        plan.load_inputs({"x": x})
        plan.run()
        return plan.output_nchw(out)


class _PlannedModule(nn.Module):
    """Caches compiled plans per (input shapes, device); invalidated when parameters are reloaded or moved."""

    def __init__(self, math="exact"):
        super().__init__()
        self.math = math
        self._plans = {}
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._plans.clear())

    def _apply(self, fn, recurse=True):
        self._plans = {}
        return super()._apply(fn, recurse)

    def set_math(self, math):
        self.math = math
        self._plans = {}
        return self

    def _plan_for(self, key, build):
        key = (key, self.math)
        if key not in self._plans or self._plans[key].stale():
            self._plans[key] = build()
        return self._plans[key]


class CVEncoder(_PlannedModule):
    """reference modules/networks.py:88-117."""

    def __init__(self, num_ch_cv, num_ch_enc, num_ch_outs, math="exact"):
        super().__init__(math)
        self.convs = nn.ModuleDict()
        self.num_ch_enc = []
        self.num_blocks = len(num_ch_outs)
        for i in range(self.num_blocks):
            cin = num_ch_cv if i == 0 else num_ch_outs[i - 1]
            cout = num_ch_outs[i]
            self.convs[f"ds_conv_{i}"] = BasicBlock(cin, cout, stride=1 if i == 0 else 2)
            self.convs[f"conv_{i}"] = nn.Sequential(BasicBlock(num_ch_enc[i] + cout, cout), BasicBlock(cout, cout))
            self.num_ch_enc.append(cout)

    def emit(self, plan, x: Feature, img_feats):
        outs = []
        for i in range(self.num_blocks):
            x = self.convs[f"ds_conv_{i}"].emit(plan, [(x, L.RESAMPLE_NONE)])
            # torch.cat([x, img_feats[i]]) (networks.py:114) is folded into the next block's loaders
            x = self.convs[f"conv_{i}"][0].emit(plan, [(x, L.RESAMPLE_NONE), (img_feats[i], L.RESAMPLE_NONE)])
            x = self.convs[f"conv_{i}"][1].emit(plan, [(x, L.RESAMPLE_NONE)])
            outs.append(x)
        return outs

    def forward(self, x, img_feats):
        shapes = (tuple(x.shape),) + tuple(tuple(f.shape) for f in img_feats)

        def build():
            plan = ConvPlan(x.device, self.math)
            fx = plan.input("x", *x.shape)
            fi = [plan.input(f"img{i}", *f.shape) for i, f in enumerate(img_feats)]
            plan.outputs = self.emit(plan, fx, fi)
            return plan.finalize()

        plan = self._plan_for((shapes, str(x.device)), build)
        plan.load_inputs({"x": x, **{f"img{i}": f for i, f in enumerate(img_feats)}})
        plan.run()
        return [_nchw_out(f) for f in plan.outputs]


def _double_basic_block(cin, cout):
    layers = nn.Sequential(BasicBlock(cin, cout))
    layers.add_module("conv_0", BasicBlock(cout, cout))  # key name of the reference (networks.py:13-17)
    return layers


class DepthDecoderPP(_PlannedModule):
    """reference modules/networks.py:20-85 (UNet++ decoder, 4 log-depth heads)."""

    def __init__(self, num_ch_enc, scales=range(4), num_output_channels=1, use_skips=True, math="exact"):
        super().__init__(math)
        self.num_output_channels = num_output_channels
        self.num_ch_enc = list(num_ch_enc)
        self.num_ch_dec = np.array([64, 64, 128, 256])
        self.convs = nn.ModuleDict()
        for j in range(1, 5):
            for i in range(4 - j, -1, -1):
                cout = int(self.num_ch_dec[i])
                total = 0
                cin = self.num_ch_enc[i + 1] if j == 1 else int(self.num_ch_dec[i + 1])
                self.convs[f"diag_conv_{i + 1}{j - 1}"] = BasicBlock(cin, cout)
                total += cout
                cin = self.num_ch_enc[i] if j == 1 else int(self.num_ch_dec[i])
                self.convs[f"right_conv_{i}{j - 1}"] = BasicBlock(cin, cout)
                total += cout
                if i + j != 4:
                    self.convs[f"up_conv_{i + 1}{j}"] = BasicBlock(int(self.num_ch_dec[i + 1]), cout)
                    total += cout
                self.convs[f"in_conv_{i}{j}"] = _double_basic_block(total, cout)
                self.convs[f"output_{i}"] = nn.Sequential(
                    BasicBlock(cout, cout) if i != 0 else nn.Identity(), nn.Conv2d(cout, num_output_channels, 1))

    def emit(self, plan, feats):
        """networks.py:65-85.  ``upsample`` (bilinear x2) and ``torch.cat`` are folded into in_conv's loaders; only the
        head evaluation that survives per scale (last write to the dict key, networks.py:81) is emitted."""
        prev = list(feats)
        heads = {}
        UP = L.RESAMPLE_BILINEAR_UP2
        for j in range(1, 5):
            col = []
            for i in range(4 - j, -1, -1):
                parts = [(self.convs[f"right_conv_{i}{j - 1}"].emit(plan, [(prev[i], L.RESAMPLE_NONE)]), L.RESAMPLE_NONE)]
                parts.append((self.convs[f"diag_conv_{i + 1}{j - 1}"].emit(plan, [(prev[i + 1], L.RESAMPLE_NONE)]), UP))
                if i + j != 4:
                    parts.append((self.convs[f"up_conv_{i + 1}{j}"].emit(plan, [(col[-1], L.RESAMPLE_NONE)]), UP))
                block = self.convs[f"in_conv_{i}{j}"]
                x = block[0].emit(plan, parts)
                x = block.conv_0.emit(plan, [(x, L.RESAMPLE_NONE)])
                col.append(x)
                if i + j == 4:
                    head = self.convs[f"output_{i}"]
                    h = x
                    if i != 0:
                        h = head[0].emit(plan, [(h, L.RESAMPLE_NONE)])
                    heads[f"log_depth_pred_s{i}_b1hw"] = plan.conv([(h, L.RESAMPLE_NONE)], head[1])
            prev = col[::-1]
        return heads

    def forward(self, input_features):
        shapes = tuple(tuple(f.shape) for f in input_features)
        dev = input_features[0].device

        def build():
            plan = ConvPlan(dev, self.math)
            fi = [plan.input(f"f{i}", *f.shape) for i, f in enumerate(input_features)]
            plan.outputs = self.emit(plan, fi)
            return plan.finalize()

        plan = self._plan_for((shapes, str(dev)), build)
        plan.load_inputs({f"f{i}": f for i, f in enumerate(input_features)})
        plan.run()
        return {k: _nchw_out(v) for k, v in sorted(plan.outputs.items(), reverse=True)}


class ConvBlock(nn.Module):
    """reference modules/networks_fast.py:6-27 (two 3x3 convs, ELU)."""

    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv1 = nn.Conv2d(in_ch, out_ch, 3, padding=1)
        self.conv2 = nn.Conv2d(out_ch, out_ch, 3, padding=1)

    def emit(self, plan, srcs):
        x = plan.conv(srcs, self.conv1, L.ACT_ELU)
        return plan.conv([(x, L.RESAMPLE_NONE)], self.conv2, L.ACT_ELU)


class ConvUpsampleAndConcatBlock(nn.Module):
    """reference modules/networks_fast.py:29-44."""

    def __init__(self, in_ch, out_ch, skip_chns):
        super().__init__()
        self.pre_concat_conv = ConvBlock(in_ch, out_ch)
        self.post_concat_conv = ConvBlock(out_ch + skip_chns, out_ch)

    def emit(self, plan, x, skip):
        x = self.pre_concat_conv.emit(plan, [(x, L.RESAMPLE_NONE)])
        # F.interpolate(nearest x2) + torch.cat (networks_fast.py:38-39) folded into the loader
        return self.post_concat_conv.emit(plan, [(x, L.RESAMPLE_NEAREST_UP2), (skip, L.RESAMPLE_NONE)])


class SkipDecoderRegression(_PlannedModule):
    """reference modules/networks_fast.py:47-141 (SkipDecoder + 1x1-conv regression heads)."""

    def __init__(self, input_channels, use_bn=False, math="exact"):
        super().__init__(math)
        ic = list(input_channels)[::-1]
        self.input_channels = ic
        self.output_channels = [256, 128, 64, 64]
        self.num_ch_dec = self.output_channels[::-1]
        for n in range(4):
            setattr(self, f"block{n + 1}", ConvUpsampleAndConcatBlock(ic[n], self.output_channels[n], ic[n + 1]))
        for n in range(4):
            setattr(self, f"out{n + 1}", nn.Sequential(
                nn.Conv2d(self.output_channels[n], 128, 1), nn.ELU(inplace=True), nn.Conv2d(128, 128, 1),
                nn.ELU(inplace=True), nn.Conv2d(128, 1, 1)))

    def emit(self, plan, feats):
        outs = {}
        x = feats[-1]
        for n in range(1, 5):
            x = getattr(self, f"block{n}").emit(plan, x, feats[-1 - n])
            s = 4 - n
            outs[f"feature_s{s}_b1hw"] = x
            head = getattr(self, f"out{n}")
            h = plan.conv([(x, L.RESAMPLE_NONE)], head[0], L.ACT_ELU)
            h = plan.conv([(h, L.RESAMPLE_NONE)], head[2], L.ACT_ELU)
            outs[f"log_depth_pred_s{s}_b1hw"] = plan.conv([(h, L.RESAMPLE_NONE)], head[4])
        return outs

    def forward(self, features):
        shapes = tuple(tuple(f.shape) for f in features)
        dev = features[0].device

        def build():
            plan = ConvPlan(dev, self.math)
            fi = [plan.input(f"f{i}", *f.shape) for i, f in enumerate(features)]
            plan.outputs = self.emit(plan, fi)
            return plan.finalize()

        plan = self._plan_for((shapes, str(dev)), build)
        plan.load_inputs({f"f{i}": f for i, f in enumerate(features)})
        plan.run()
        return {k: _nchw_out(v) for k, v in plan.outputs.items()}
