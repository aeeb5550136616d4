"""B200 drop-ins for the reference's cost-volume managers.

Same constructor / ``build_cost_volume`` / ``forward`` signatures, return tuples and ``state_dict`` keys as
  * ``CostVolumeManager``            reference modules/cost_volume.py:9-363
  * ``FeatureVolumeManager``         reference modules/feature_volume.py:12-365
  * ``FeatureMeshHintVolumeManager`` reference modules/mesh_hint_volume.py:12-449
so they are swapped in the way the reference swaps in its own fast variant (utils/model_utils.py:30-34):
``model.cost_volume = doubletake_b200.to_b200(model.cost_volume)``.

The per-plane Python loop of the reference does not exist here: one call of ``dtb200_cost_volume`` (C ABI,
include/doubletake_b200.h) warps, matches, runs both MLPs, reduces the arg-max plane and writes the source-view mask.
There is no PyTorch fallback; CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib as L


class MLP(nn.Module):
    """Parameter container with the reference's key names (modules/networks.py:120-135:
    ``net.{0,2,4}.{weight,bias}``, LeakyReLU(0.01) in between, no final activation).  The layers are evaluated inside
    the fused cost-volume kernel; calling this module directly is not part of the hot path."""

    def __init__(self, channel_list, disable_final_activation=False):
        super().__init__()
        layers = []
        for i in range(len(channel_list) - 1):
            layers.append(nn.Linear(channel_list[i], channel_list[i + 1]))
            layers.append(nn.LeakyReLU(inplace=True))
        if disable_final_activation:
            layers = layers[:-1]
        self.net = nn.Sequential(*layers)

    def forward(self, x):
        """Standalone evaluation (reference modules/networks.py:133-135): ``x`` of shape (..., F) -> (..., out).  On the hot
        path these layers live inside the fused cost-volume kernel; called directly, the module runs them as 1x1 fused-conv
        descriptors (exact fp32 math) over the flattened rows -- CUDA only, like everything here."""
        from .networks import ConvPlan

        if not x.is_cuda:
            raise RuntimeError("doubletake_b200.MLP runs on CUDA only (no CPU fallback)")
        linears = [m for m in self.net if isinstance(m, nn.Linear)]
        acts = [isinstance(self.net[i + 1], nn.LeakyReLU) if i + 1 < len(self.net) else False
                for i, m in enumerate(self.net) if isinstance(m, nn.Linear)]
        lead, F = x.shape[:-1], x.shape[-1]
        rows = x.reshape(-1, F).float()
        n = rows.shape[0]
        key = (n, str(x.device), tuple((l.weight._version, l.bias._version) for l in linears))
        cache = self.__dict__.setdefault("_plan", {})
        if key not in cache:
            cache.clear()
            plan = ConvPlan(x.device, "exact")
            fpad = (F + 7) // 8 * 8                      # the SIMT conv reads input channels in groups of 8
            f = plan.input("x", 1, fpad, 1, n)
            holders = []
            cin = fpad
            for lin, act in zip(linears, acts):
                out_c = lin.out_features
                oc = out_c if (out_c % 64 == 0 or out_c < 64) else (out_c + 63) // 64 * 64
                conv = nn.Conv2d(cin, oc, 1).to(x.device)
                w = torch.zeros((oc, cin), device=x.device)
                w[:out_c, : lin.in_features] = lin.weight.detach()
                b = torch.zeros(oc, device=x.device)
                b[:out_c] = lin.bias.detach()
                conv.weight = nn.Parameter(w.view(oc, cin, 1, 1), requires_grad=False)
                conv.bias = nn.Parameter(b, requires_grad=False)
                holders.append(conv)
                f = plan.conv([(f, L.RESAMPLE_NONE)], conv, L.ACT_LEAKY if act else L.ACT_NONE, 0.01)
                cin = oc
            plan.finalize()
            cache[key] = (plan, f, holders, fpad, linears[-1].out_features)
        plan, out, _, fpad, out_features = cache[key]
        xin = torch.zeros((1, fpad, 1, n), device=x.device)
        xin[0, :F, 0] = rows.t()
        plan.load_inputs({"x": xin})
        plan.run()
        return plan.output_nchw(out)[0, :out_features, 0].t().reshape(*lead, out_features)


class _PixelGrid(nn.Module):
    """Holds the ``backprojector.pix_coords_13N`` buffer so reference checkpoints load with strict=True
    (utils/geometry_utils.py:34-52).  The kernel regenerates pixel centres from thread indices."""

    def __init__(self, height, width):
        super().__init__()
        ys, xs = torch.meshgrid(torch.arange(height), torch.arange(width), indexing="ij")
        pix = torch.stack([xs.flatten() + 0.5, ys.flatten() + 0.5, torch.ones(height * width)], 0)
        self.register_buffer("pix_coords_13N", pix.float().unsqueeze(0))


class _Eps(nn.Module):
    """Holds ``projector.eps`` (utils/geometry_utils.py:71-74)."""

    def __init__(self):
        super().__init__()
        self.register_buffer("eps", torch.tensor(1e-8).view(1, 1, 1))


class CostVolumeManager(nn.Module):
    """Dot-product plane-sweep cost volume (reference modules/cost_volume.py:9-363)."""

    kind = L.VOLUME_DOT
    per_view_mask = False

    def __init__(self, matching_height, matching_width, num_depth_bins=64, matching_dim_size=None,
                 num_source_views=None, math="exact"):
        super().__init__()
        self.num_depth_bins = num_depth_bins
        self.matching_height = matching_height
        self.matching_width = matching_width
        self.math = math
        self.initialise_for_projection()

    # ------------------------------------------------------------------------------------------ reference API
    def initialise_for_projection(self, device=None):
        """reference cost_volume.py:51-71 (buffers only; no projection modules are needed on this path)."""
        ramp = torch.linspace(0, 1, self.num_depth_bins).view(1, self.num_depth_bins, 1, 1)
        self.register_buffer("linear_ramp_1d11", ramp)
        self.backprojector = _PixelGrid(self.matching_height, self.matching_width)
        self.projector = _Eps()
        if device is not None:
            self.to(device)

    def generate_depth_planes(self, batch_size, min_depth, max_depth):
        """reference cost_volume.py:96-130: exp(log(min) + log(max/min) * linspace(0,1,D)), evaluated with the same
        torch ops on the device the depth bounds live on, expanded (view) to (B,D,H,W)."""
        planes_bd11 = self._planes_bd11(batch_size, min_depth, max_depth)
        return planes_bd11.expand(batch_size, self.num_depth_bins, self.matching_height, self.matching_width)

    def _planes_bd11(self, batch_size, min_depth, max_depth):
        min_depth = torch.as_tensor(min_depth, dtype=torch.float32)
        max_depth = torch.as_tensor(max_depth, dtype=torch.float32)
        ramp = self.linear_ramp_1d11.to(min_depth.device).expand(batch_size, self.num_depth_bins, 1, 1)
        return torch.exp(torch.log(min_depth) + torch.log(max_depth / min_depth) * ramp)

    def _cached_planes(self, batch_size, min_depth, max_depth, dev):
        """(B,D) plane depths on `dev`.  Host-side bounds (the DepthModel case: floats from the options) are evaluated
        once with torch CPU ops and cached, so steady-state steps issue no H2D copy and planes are bit-identical to
        the reference's CPU path; device-side bounds are evaluated on the device with the same ops."""
        min_t, max_t = torch.as_tensor(min_depth), torch.as_tensor(max_depth)
        if min_t.is_cuda or max_t.is_cuda:
            return L.f32(self._planes_bd11(batch_size, min_t.to(dev), max_t.to(dev)).reshape(batch_size, -1), dev)
        key = (batch_size, self.num_depth_bins, float(min_t), float(max_t), str(dev))
        cache = self.__dict__.setdefault("_plane_cache", {})
        if key not in cache:
            cache[key] = L.f32(self._planes_bd11(batch_size, min_t, max_t).reshape(batch_size, -1), dev)
        return cache[key]

    def indices_to_disparity(self, indices, depth_planes_bdhw):
        return torch.gather(depth_planes_bdhw, dim=1, index=indices.unsqueeze(1)).squeeze(1)

    def build_cost_volume(self, cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth,
                          max_depth, depth_planes_bdhw=None, return_mask=False):
        out = self._run(cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth,
                        None, depth_planes_bdhw, return_mask)
        return out["volume"], out["planes"], out["mask"]

    def forward(self, cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth,
                depth_planes_bdhw=None, return_mask=False):
        out = self._run(cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth,
                        None, depth_planes_bdhw, return_mask)
        return out["volume"], out["lowest_cost"], out["planes"], out["mask"]

    # ------------------------------------------------------------------------------------------ kernel launch
    def _mlp_weights(self):
        return {}

    def _run(self, cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth,
             hint, depth_planes_bdhw, return_mask):
        if not cur_feats.is_cuda:
            raise RuntimeError("doubletake_b200 cost volumes run on CUDA only (no CPU fallback)")
        dev = cur_feats.device
        B, K, Cf, H, W = src_feats.shape
        self.matching_height, self.matching_width = H, W  # reference cost_volume.py:335-341 (portrait re-init is moot)
        D = self.num_depth_bins
        cur = L.f32(cur_feats)
        # channels-last staging of the source maps: one 64-byte texel per bilinear tap
        if src_feats.dtype == torch.float32 and not src_feats.is_contiguous() and \
                src_feats.permute(0, 1, 3, 4, 2).is_contiguous():
            src_nhwc = src_feats.permute(0, 1, 3, 4, 2)  # caller already staged the maps channels-last
        else:
            src_nhwc = L.nchw_to_nhwc(L.f32(src_feats).reshape(B * K, Cf, H, W)).view(B, K, H, W, Cf)

        if depth_planes_bdhw is None:
            planes_dev = self._cached_planes(B, min_depth, max_depth, dev)
            planes_out = planes_dev.view(B, D, 1, 1).expand(B, D, H, W)
            per_pixel = 0
        else:
            planes_dev = L.f32(depth_planes_bdhw, dev)
            planes_out = depth_planes_bdhw
            per_pixel = 1

        volume = torch.empty((B, D, H, W), dtype=torch.float32, device=dev)
        lowest = torch.empty((B, H, W), dtype=torch.float32, device=dev)
        index = torch.empty((B, H, W), dtype=torch.int32, device=dev)
        want_mask = return_mask and self.kind != L.VOLUME_DOT
        mask_views = torch.empty((B, K, H, W), dtype=torch.uint8, device=dev) if want_mask else None
        mask_any = torch.empty((B, H, W), dtype=torch.uint8, device=dev) if want_mask else None

        keep = [cur, src_nhwc, planes_dev]
        p = L.CostVolumeParams()
        p.kind, p.math = self.kind, L.MATH_NAMES[self.math]
        p.batch, p.views, p.channels, p.height, p.width, p.planes = B, K, Cf, H, W, D
        p.cur_feats, p.src_feats_nhwc = L.ptr(cur), L.ptr(src_nhwc)
        for name, t in (("src_extrinsics", src_extrinsics), ("src_poses", src_poses), ("src_Ks", src_Ks),
                        ("cur_invK", cur_invK)):
            t = L.f32(t, dev)
            keep.append(t)
            setattr(p, name, L.ptr(t))
        p.plane_depths, p.planes_per_pixel = L.ptr(planes_dev), per_pixel
        if self.kind == L.VOLUME_MLP_HINT:
            hd = L.f32(hint["depth_hint_b1hw"], dev)
            hw = L.f32(hint["sampled_weights_b1hw"], dev)
            hm = L.f32(hint["depth_hint_mask_b1hw"], dev)
            keep += [hd, hw, hm]
            p.depth_hint, p.hint_weights, p.hint_mask = L.ptr(hd), L.ptr(hw), L.ptr(hm)
            p.hint_height, p.hint_width = hd.shape[-2], hd.shape[-1]
        for name, t in self._mlp_weights().items():
            t = L.f32(t.detach(), dev)
            keep.append(t)
            setattr(p, name, L.ptr(t))
        p.volume, p.lowest_cost, p.best_index = L.ptr(volume), L.ptr(lowest), L.ptr(index)
        p.mask_views, p.mask_any = L.ptr(mask_views), L.ptr(mask_any)
        ws = int(L.lib().dtb200_cost_volume_workspace_bytes(C.byref(p)))
        if ws:
            # tensor-core weight tiles: re-tiled once per weight version, then reused
            w = self._mlp_weights()
            key = (str(dev), ws) + tuple((t.data_ptr(), t._version) for t in (w["w1"], w["w2"]))
            cache = self.__dict__.setdefault("_tc_workspace", {})
            if key not in cache:
                cache.clear()
                work = torch.empty(ws, dtype=torch.uint8, device=dev)
                p.workspace, p.workspace_bytes = L.ptr(work), ws
                L.check(L.lib().dtb200_cost_volume_prepare(C.byref(p), L.stream()))
                cache[key] = work
            work = cache[key]
            p.workspace, p.workspace_bytes, p.workspace_prepared = L.ptr(work), ws, 1
        L.check(L.lib().dtb200_cost_volume(C.byref(p), L.stream()))

        mask = None
        if want_mask:
            mask = (mask_views if self.per_view_mask else mask_any).bool()
        return dict(volume=volume, lowest_cost=lowest, index=index, planes=planes_out, mask=mask)


class FeatureVolumeManager(CostVolumeManager):
    """Metadata-MLP feature volume (reference modules/feature_volume.py:12-365)."""

    kind = L.VOLUME_MLP

    def __init__(self, matching_height, matching_width, num_depth_bins=64, mlp_channels=None, matching_dim_size=16,
                 num_source_views=7, math="exact"):
        super().__init__(matching_height, matching_width, num_depth_bins, math=math)
        mlp_channels = list(mlp_channels) if mlp_channels is not None else [202, 128, 128, 1]
        # feature_volume.py:48-70: 16(K+1) visual + (K+1) depth + 3(K+1) rays + K angles + K masks + K dots + 3K pose
        mlp_channels[0] = (matching_dim_size + 10) * num_source_views + matching_dim_size + 4
        if mlp_channels[1:] != [128, 128, 1]:
            raise ValueError("doubletake_b200: the fused kernel implements the reference's [F,128,128,1] MLP only")
        self.matching_dim_size = matching_dim_size
        self.num_source_views = num_source_views
        self.mlp = MLP(channel_list=mlp_channels, disable_final_activation=True)

    def _mlp_weights(self):
        n = self.mlp.net
        return {"w1": n[0].weight, "b1": n[0].bias, "w2": n[2].weight, "b2": n[2].bias, "w3": n[4].weight,
                "b3": n[4].bias}

    def _check_views(self, src_feats):
        if src_feats.shape[1] != self.num_source_views:
            raise ValueError(f"expected {self.num_source_views} source views, got {src_feats.shape[1]}")

    def build_cost_volume(self, cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth,
                          max_depth, depth_planes_bdhw=None, return_mask=False):
        self._check_views(src_feats)
        return super().build_cost_volume(cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK,
                                         min_depth, max_depth, depth_planes_bdhw, return_mask)

    def forward(self, cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth,
                depth_planes_bdhw=None, return_mask=False):
        self._check_views(src_feats)
        return super().forward(cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth,
                               max_depth, depth_planes_bdhw, return_mask)

    def to_fast(self):
        """reference feature_volume.py:358-365: the fused kernel already is the fast path."""
        return self


class FeatureMeshHintVolumeManager(FeatureVolumeManager):
    """Metadata-MLP feature volume + rendered-depth hint MLP (reference modules/mesh_hint_volume.py:12-449).
    ``return_mask=True`` yields the per-view (B,K,H,W) mask of the last plane like the reference's slow manager
    (mesh_hint_volume.py:273-287); ``to_fast()`` yields the (B,H,W) any-view mask of its fast manager (:818-822)."""

    kind = L.VOLUME_MLP_HINT
    per_view_mask = True

    def __init__(self, matching_height, matching_width, num_depth_bins=64, mlp_channels=None, matching_dim_size=16,
                 num_source_views=7, math="exact"):
        super().__init__(matching_height, matching_width, num_depth_bins, mlp_channels, matching_dim_size,
                         num_source_views, math=math)
        self.hint_mlp = MLP(channel_list=[3, 12, 12, 1], disable_final_activation=True)

    def _mlp_weights(self):
        w = super()._mlp_weights()
        n = self.hint_mlp.net
        w.update({"hw1": n[0].weight, "hb1": n[0].bias, "hw2": n[2].weight, "hb2": n[2].bias, "hw3": n[4].weight,
                  "hb3": n[4].bias})
        return w

    def build_cost_volume(self, cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth,
                          max_depth, cv_depth_hint_dict, depth_planes_bdhw=None, return_mask=False):
        self._check_views(src_feats)
        out = self._run(cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth,
                        cv_depth_hint_dict, depth_planes_bdhw, return_mask)
        return out["volume"], out["planes"], out["mask"]

    def forward(self, cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth,
                cv_depth_hint_dict, depth_planes_bdhw=None, return_mask=False):
        self._check_views(src_feats)
        out = self._run(cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth,
                        cv_depth_hint_dict, depth_planes_bdhw, return_mask)
        return out["volume"], out["lowest_cost"], out["planes"], out["mask"]

    def to_fast(self):
        fast = FastFeatureMeshHintVolumeManager(self.matching_height, self.matching_width, self.num_depth_bins,
                                                matching_dim_size=self.matching_dim_size,
                                                num_source_views=self.num_source_views, math=self.math)
        fast.mlp = self.mlp
        fast.hint_mlp = self.hint_mlp
        return fast.to(self.linear_ramp_1d11.device)


class FastFeatureMeshHintVolumeManager(FeatureMeshHintVolumeManager):
    """Same kernel; (B,H,W) any-view mask contract of the reference's fast manager (mesh_hint_volume.py:818-822)."""

    per_view_mask = False


def to_b200(manager, math="exact"):
    """Convert a REFERENCE manager instance (duck-typed: ``num_depth_bins``, ``matching_height/width``, optional
    ``mlp`` / ``hint_mlp`` modules with ``net.{0,2,4}``) into its B200 drop-in, sharing the checkpoint weights."""
    name = type(manager).__name__
    args = (manager.matching_height, manager.matching_width, manager.num_depth_bins)
    if hasattr(manager, "mlp"):
        # F = (C + 10) K + C + 4 (feature_volume.py:48-70).  The reference managers do not store C or K; the fused kernels
        # implement the reference's C = 16 (options.py matching_feature_dims), so K follows from F -- anything else is refused
        in_features = manager.mlp.net[0].weight.shape[1]
        K = (in_features - 20) // 26
        if K < 1 or 26 * K + 20 != in_features:
            raise ValueError(f"to_b200: an MLP input width of {in_features} is not (16 + 10) K + 20 for any K: the B200 kernels "
                             "implement 16-channel matching features")
        cls = FeatureVolumeManager
        if hasattr(manager, "hint_mlp"):
            cls = FastFeatureMeshHintVolumeManager if name.startswith("Fast") else FeatureMeshHintVolumeManager
        new = cls(*args, matching_dim_size=16, num_source_views=K, math=math)
        new.mlp.load_state_dict(manager.mlp.state_dict())
        if hasattr(manager, "hint_mlp"):
            new.hint_mlp.load_state_dict(manager.hint_mlp.state_dict())
    else:
        new = CostVolumeManager(*args, math=math)
    device = next(iter(manager.buffers())).device if any(True for _ in manager.buffers()) else "cuda"
    return new.to(device)
