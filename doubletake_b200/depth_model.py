"""B200 mirror of the reference's depth models for the hot path.

``DepthModelCVHint.forward`` (reference experiment_modules/doubletake_model.py:265-425) and ``DepthModel.forward``
(experiment_modules/sr_depth_model.py:275-435) keep their signature
``forward(phase, cur_data, src_data, unbatched_matching_encoder_forward=False, return_mask=False)`` and return the same
dict (``log_depth_pred_s{0..3}_b1hw``, ``depth_pred_s{0..3}_b1hw``, ``lowest_cost_bhw``, ``overall_mask_bhw``).

The two image encoders (timm EfficientNetV2-S / ResNet18d priors, ResNet18-stem matching encoder) are UPSTREAM of the
boundary: they are injected as callables, or their outputs are supplied in the data dicts under
``image_prior_feats`` / ``matching_feats_bchw`` / ``matching_feats_bkchw``.  Everything between the encoders and the return
statement runs in hand-written CUDA through the C ABI: relative poses, fused cost volume, and ONE conv plan holding the
cost-volume encoder and the depth decoder back to back (no NCHW round trip between them), then exp.
Inference only: ``phase == "train"`` (flip augmentation + losses) is out of scope and raises.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn

from . import _lib as L
from .cost_volume import CostVolumeManager, FeatureMeshHintVolumeManager, FeatureVolumeManager
from .networks import ConvPlan, CVEncoder, DepthDecoderPP, SkipDecoderRegression, _nchw_out

PRIOR_CHANNELS = {"efficientnet": [24, 48, 64, 160, 256], "resnet18d": [64, 64, 128, 256, 512]}


@dataclass
class HotPathOptions:
    """The fields of the reference's ``Options`` (options.py:9-230) that the hot path reads, with its defaults."""

    image_encoder_name: str = "efficientnet"
    cv_encoder_type: str = "multi_scale_encoder"
    depth_decoder_name: str = "unet_pp"
    feature_volume_type: str = "mlp_mesh_hint_feature_volume"
    matching_num_depth_bins: int = 64
    matching_scale: int = 1
    matching_feature_dims: int = 16
    model_num_views: int = 8
    image_width: int = 512
    image_height: int = 384
    min_matching_depth: float = 0.25
    max_matching_depth: float = 5.0


class DepthModelCVHint(nn.Module):
    """Hot-path mirror of reference ``DepthModelCVHint`` (experiment_modules/doubletake_model.py:32-425)."""

    volume_types = {"mlp_mesh_hint_feature_volume": FeatureMeshHintVolumeManager}

    def __init__(self, opts, encoder=None, matching_model=None, math="exact", volume_math=None):
        """``math`` selects the conv-stack arithmetic ("exact": fp32 CUDA cores; "tc3x": tcgen05 3xTF32);
        ``volume_math`` the cost-volume MLP arithmetic (defaults to "exact")."""
        super().__init__()
        self.run_opts = opts
        self.math = math
        self.volume_math = volume_math or "exact"
        self.encoder = encoder  # image-prior encoder: image -> list of 5 maps (upstream of the boundary)
        self.matching_model = matching_model  # matching encoder: image -> (B,16,H/4,W/4) (upstream of the boundary)
        if encoder is not None and hasattr(encoder, "num_ch_enc"):
            num_ch_enc = list(encoder.num_ch_enc)
        else:
            fam = "efficientnet" if "efficientnet" in opts.image_encoder_name else "resnet18d"
            if fam not in opts.image_encoder_name:
                raise ValueError("Unrecognized option for image encoder type!")
            num_ch_enc = PRIOR_CHANNELS[fam]
        self.num_ch_enc = num_ch_enc
        ms = opts.matching_scale
        if opts.cv_encoder_type != "multi_scale_encoder":
            raise ValueError("Unrecognized option for cost volume encoder type!")
        self.cost_volume_net = CVEncoder(num_ch_cv=opts.matching_num_depth_bins, num_ch_enc=num_ch_enc[ms:],
                                         num_ch_outs=[64, 128, 256, 384], math=math)
        dec_in = num_ch_enc[:ms] + self.cost_volume_net.num_ch_enc
        if opts.depth_decoder_name == "unet_pp":
            self.depth_decoder = DepthDecoderPP(dec_in, math=math)
        elif opts.depth_decoder_name == "skip":
            self.depth_decoder = SkipDecoderRegression(dec_in, math=math)
        else:
            raise ValueError("Unrecognized option for depth decoder name!")
        if opts.feature_volume_type not in self.volume_types:
            raise ValueError(f"unsupported feature_volume_type {opts.feature_volume_type} for {type(self).__name__}")
        self.cost_volume = self.volume_types[opts.feature_volume_type](
            matching_height=opts.image_height // (2 ** (ms + 1)), matching_width=opts.image_width // (2 ** (ms + 1)),
            num_depth_bins=opts.matching_num_depth_bins, matching_dim_size=opts.matching_feature_dims,
            num_source_views=opts.model_num_views - 1, math=self.volume_math)
        self._plans = {}
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._plans.clear())

    def _apply(self, fn, recurse=True):
        self._plans = {}
        return super()._apply(fn, recurse)

    # -------------------------------------------------------------------------------------- upstream encoders
    def compute_matching_feats(self, cur_image, src_image, unbatched_matching_encoder_forward):
        """reference doubletake_model.py:210-263, using the injected matching encoder (not part of this engine)."""
        if self.matching_model is None:
            raise RuntimeError("no matching encoder injected and no precomputed matching features in the data dicts")
        B, K = src_image.shape[:2]
        if hasattr(self.matching_model, "forward_views"):
            # the B200 encoder: every view in one pass, features already in the cost-volume kernels' layouts (its per-image
            # InstanceNorm statistics make batching exact, so the reference's unbatched workaround is moot)
            return self.matching_model.forward_views(cur_image, src_image)
        if unbatched_matching_encoder_forward:
            cur = self.matching_model(cur_image)
            src = torch.stack([self.matching_model(src_image[:, k]) for k in range(K)], 1)
        else:
            allf = self.matching_model(torch.cat([cur_image[:, None], src_image], 1).flatten(0, 1))
            allf = allf.view(B, K + 1, *allf.shape[1:])
            cur, src = allf[:, 0], allf[:, 1:].contiguous()
        return cur, src

    # -------------------------------------------------------------------------------------- host -> device staging
    def _upload(self, cur_data, src_data, dev):
        """Inputs that still live in HOST memory (the reference moves the whole batch with ``to_gpu`` before forward,
        utils/generic_utils.py) are copied on a dedicated copy stream in the order the kernels need them: poses,
        intrinsics, matching features and hint first (the cost volume waits on event 0; EVERY non-list tensor is in this
        group, so no kernel can read an input whose copy it did not wait for), the list-valued image-prior feature maps
        last (only the conv plan waits on event 1).  With pinned buffers the copies of frame i+1 overlap the kernels of
        frame i, and a frame's prior maps travel while its cost volume is being computed.
        Returns (cur_data, src_data, (event_early, event_late)) -- events are None when nothing had to move."""
        def on_host(v):
            return torch.is_tensor(v) and not v.is_cuda

        trees = (cur_data, src_data)
        if not any(on_host(x) for d in trees for v in d.values() for x in (v if isinstance(v, (list, tuple)) else [v])):
            return cur_data, src_data, (None, None)
        main = torch.cuda.current_stream(dev)
        # bound the host's run-ahead to two frames: staged inputs are recycled by the caching allocator instead of growing
        # with every queued frame (a cudaMalloc inside the loop costs more than the copy it serves)
        inflight = self.__dict__.setdefault("_inflight", [])
        while len(inflight) >= 2:
            inflight.pop(0).synchronize()
        copy = self.__dict__.get("_copy_stream")
        if copy is None or copy.device != dev:
            copy = self.__dict__["_copy_stream"] = torch.cuda.Stream(dev)
        out = ({}, {})

        def move(v):
            if not on_host(v):
                return v
            t = v.to(dev, non_blocking=True)
            t.record_stream(main)
            return t

        with torch.cuda.stream(copy):
            late = []
            for d, o in zip(trees, out):
                for k, v in d.items():
                    if isinstance(v, (list, tuple)):
                        late.append((o, k, v))
                    else:
                        o[k] = move(v)  # every plain tensor is "early": only the list-valued prior maps travel late
            ev_early = copy.record_event()
            for o, k, v in late:
                o[k] = [move(x) for x in v] if isinstance(v, (list, tuple)) else move(v)
            ev_late = copy.record_event()
        return out[0], out[1], (ev_early, ev_late)

    # -------------------------------------------------------------------------------------- hot path
    def _relative_poses(self, cur_data, src_data, dev):
        """doubletake_model.py:341-349 on the device, one kernel."""
        scw = L.f32(src_data["cam_T_world_b44"], dev)
        swc = L.f32(src_data["world_T_cam_b44"], dev)
        ccw = L.f32(cur_data["cam_T_world_b44"], dev)
        cwc = L.f32(cur_data["world_T_cam_b44"], dev)
        B, K = scw.shape[:2]
        ext = torch.empty((B, K, 4, 4), dtype=torch.float32, device=dev)
        pose = torch.empty_like(ext)
        L.check(L.lib().dtb200_relative_poses(L.ptr(scw), L.ptr(swc), L.ptr(ccw), L.ptr(cwc), L.ptr(ext), L.ptr(pose),
                                              B, K, L.stream()))
        return ext, pose

    def _run_cost_volume(self, mcur, msrc, ext, pose, src_K, cur_invK, min_depth, max_depth, cur_data, return_mask):
        return self.cost_volume._run(mcur, msrc, ext, pose, src_K, cur_invK, min_depth, max_depth, cur_data, None,
                                     return_mask)

    def _network_plan(self, cv_shape, prior_feats):
        key = (tuple(cv_shape), tuple(tuple(f.shape) for f in prior_feats), self.math)
        if key not in self._plans or self._plans[key].stale():
            ms = self.run_opts.matching_scale
            plan = ConvPlan(prior_feats[0].device, self.math)
            fcv = plan.input("cv", *cv_shape)
            fp = [plan.input(f"prior{i}", *f.shape) for i, f in enumerate(prior_feats)]
            enc = self.cost_volume_net.emit(plan, fcv, fp[ms:])
            plan.outputs = self.depth_decoder.emit(plan, fp[:ms] + enc)
            self._plans[key] = plan.finalize()
        return self._plans[key]

    def forward(self, phase, cur_data, src_data, unbatched_matching_encoder_forward=False, return_mask=False):
        if phase == "train":
            raise NotImplementedError("doubletake_b200 is an inference engine: the training phase is out of scope")
        ms = self.run_opts.matching_scale
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("doubletake_b200 runs on CUDA only (no CPU fallback): move the model to a CUDA device")
        cur_data, src_data, (ev_early, ev_late) = self._upload(cur_data, src_data, dev)
        if ev_early is not None:
            torch.cuda.current_stream(dev).wait_event(ev_early)
        if "image_prior_feats" in cur_data:
            prior_feats = list(cur_data["image_prior_feats"])
            mcur, msrc = cur_data["matching_feats_bchw"], src_data["matching_feats_bkchw"]
        else:
            if self.encoder is None:
                raise RuntimeError("no image-prior encoder injected and no precomputed 'image_prior_feats'")
            prior_feats = list(self.encoder(cur_data["image_b3hw"]))
            mcur, msrc = self.compute_matching_feats(cur_data["image_b3hw"], src_data["image_b3hw"],
                                                     unbatched_matching_encoder_forward)
        src_K = src_data[f"K_s{ms}_b44"]
        cur_invK = cur_data[f"invK_s{ms}_b44"]
        ext, pose = self._relative_poses(cur_data, src_data, dev)
        # doubletake_model.py:374-376: depth bounds from the options (host floats -> planes bit-identical to torch CPU)
        min_depth = torch.tensor(self.run_opts.min_matching_depth).view(1, 1, 1, 1)
        max_depth = torch.tensor(self.run_opts.max_matching_depth).view(1, 1, 1, 1)
        cv = self._run_cost_volume(mcur, msrc, ext, pose, src_K, cur_invK, min_depth, max_depth, cur_data, return_mask)

        if ev_late is not None:
            torch.cuda.current_stream(dev).wait_event(ev_late)
        plan = self._network_plan(cv["volume"].shape, prior_feats)
        plan.load_inputs({"cv": cv["volume"], **{f"prior{i}": f for i, f in enumerate(prior_feats)}})
        plan.run()
        depth_outputs = {}
        for k, f in plan.outputs.items():
            log_depth = _nchw_out(f)
            depth_outputs[k] = log_depth
            lin = torch.empty_like(log_depth)
            L.check(L.lib().dtb200_exp(L.ptr(log_depth), L.ptr(lin), log_depth.numel(), L.stream()))
            depth_outputs[k.replace("log_", "")] = lin  # doubletake_model.py:410-418 (incl. the feature_s* quirk)
        if ev_late is not None:
            self.__dict__["_inflight"].append(torch.cuda.current_stream(dev).record_event())
        depth_outputs["lowest_cost_bhw"] = cv["lowest_cost"]
        depth_outputs["overall_mask_bhw"] = cv["mask"]
        return depth_outputs


class DepthModel(DepthModelCVHint):
    """Hot-path mirror of reference SimpleRecon ``DepthModel`` (experiment_modules/sr_depth_model.py:32-435): no hint;
    ``feature_volume_type`` selects the dot-product or the metadata-MLP volume (:186-194)."""

    volume_types = {"simple_cost_volume": CostVolumeManager, "mlp_feature_volume": FeatureVolumeManager}

    def _run_cost_volume(self, mcur, msrc, ext, pose, src_K, cur_invK, min_depth, max_depth, cur_data, return_mask):
        return self.cost_volume._run(mcur, msrc, ext, pose, src_K, cur_invK, min_depth, max_depth, None, None,
                                     return_mask)


def install(model, math="exact"):
    """The reference's own plug-in idiom (utils/model_utils.py:30-34), for a loaded REFERENCE LightningModule:
    replace ``cost_volume``, ``cost_volume_net`` and ``depth_decoder`` by B200 modules carrying the same weights.
    The reference's ``forward`` then runs unchanged on top of the CUDA kernels."""
    from .cost_volume import to_b200

    model.cost_volume = to_b200(model.cost_volume, math)
    dev = next(model.cost_volume_net.parameters()).device
    old_enc = model.cost_volume_net
    first = old_enc.convs["ds_conv_0"].conv1
    prior_ch = [old_enc.convs[f"conv_{i}"][0].conv1.in_channels - old_enc.num_ch_enc[i] for i in range(old_enc.num_blocks)]
    enc = CVEncoder(first.in_channels, prior_ch, list(old_enc.num_ch_enc), math=math)
    enc.load_state_dict(old_enc.state_dict())
    model.cost_volume_net = enc.to(dev)
    old_dec = model.depth_decoder
    if hasattr(old_dec, "convs"):
        dec = DepthDecoderPP(list(old_dec.num_ch_enc), math=math)
    else:
        dec = SkipDecoderRegression(list(old_dec.input_channels)[::-1], math=math)
    dec.load_state_dict(old_dec.state_dict())
    model.depth_decoder = dec.to(dev)
    return model
