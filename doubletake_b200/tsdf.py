"""B200 mirror of the reference's TSDF volume and depth fuser (reference ``tools/tsdf.py``) -- SURVEY.md §8f row N2, the
hint-production step on the other side of the hot path: the predicted ``depth_pred_s0_b1hw`` is fused into a TSDF
(``TSDFFuser.integrate_depth``, tools/tsdf.py:414-558) and the fused confidence is read back as ``sampled_weights_b1hw``
for the next keyframe's hint (``TSDF.sample_tsdf``, :277-337; test_incremental.py:220-252).

Same class names, constructor arguments, method signatures and ``.npz`` format as the reference.  The volume lives in HBM
as the reference's fp16 tensors; ``integrate_depth`` is ONE CUDA kernel per batch of up to 8 frames
(``csrc/tsdf.cu``), ``sample_tsdf`` one kernel.  No CPU fallback: ``use_gpu=False`` raises.
Row N3 (hint rendering) is served WITHOUT a mesh: ``TSDF.render_depth_hint`` ray-casts the volume in one kernel and emits
the three hint tensors of the incremental loop; ``to_mesh*`` (marching cubes for mesh export) is not part of the engine.

Numerics: the kernels reproduce the reference's fp16 torch ops rounding for rounding.  fp16 ``grid_sample`` behaves
differently in ATen's CPU and CUDA builds (index arithmetic precision, non-finite indices); ``semantics="aten_cuda"``
(default: the reference fuser runs on CUDA) follows GridSampler.cuh and is pinned against torch's own CUDA ops on the GPU
box (tests/test_gpu_tsdf.py), ``"aten_cpu"`` is what the CPU-generated golden fixtures pin.
``sample_tsdf`` follows the reference's fp32 branch (its CUDA branch rounds the coordinates to fp16 first).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _matmul_h(a, b):
    """fp16 @ fp16 -> fp16 as ATen's generic CPU gemm evaluates it (fp32 accumulation in k order, one rounding), spelled
    out so the few 4x4 products below do not depend on which fp16 GEMM path the host CPU offers."""
    a32, b32 = a.float(), b.float()
    acc = torch.zeros(a32.shape[:-1] + b32.shape[-1:], dtype=torch.float32)
    for k in range(a32.shape[-1]):
        acc = acc + a32[..., :, k:k + 1] * b32[..., k:k + 1, :]
    return acc.half()


def get_frustum_bounds(invK_144, world_T_cam_144, min_depth=0.1, max_depth=10.0, img_h=480, img_w=640):
    """reference tools/tsdf.py:15-50 on fp16 CPU matrices ((4,4), as the reference's caller passes them): world-space
    bounding box of the view frustum."""
    corners = torch.tensor([[0, 0, 1, 1], [img_w, 0, 1, 1], [0, img_h, 1, 1], [img_w, img_h, 1, 1]],
                           dtype=invK_144.dtype).T
    pts = _matmul_h(invK_144, corners)
    near, far = pts.clone(), pts.clone()
    near[:3] *= min_depth
    far[:3] *= max_depth
    world = _matmul_h(world_T_cam_144, torch.cat((near, far), dim=1))
    return world.amin(dim=1)[:3], world.amax(dim=1)[:3]


def _matmul_h_np(a, b):
    """numpy twin of ``_matmul_h`` for the per-frame constants."""
    a32, b32 = a.astype(np.float32), b.astype(np.float32)
    acc = np.zeros(a32.shape[:-1] + b32.shape[-1:], dtype=np.float32)
    for k in range(a32.shape[-1]):
        acc = (acc + a32[..., :, k:k + 1] * b32[..., k:k + 1, :]).astype(np.float32)
    return acc.astype(np.float16)


def _frustum_bounds_np(invK, world_T_cam, min_depth, max_depth, img_h, img_w):
    """numpy twin of ``get_frustum_bounds`` (fp16 in, fp16 out)."""
    corners = np.array([[0, 0, 1, 1], [img_w, 0, 1, 1], [0, img_h, 1, 1], [img_w, img_h, 1, 1]], dtype=np.float16).T
    pts = _matmul_h_np(invK, corners)
    near, far = pts.copy(), pts.copy()
    near[:3] = (near[:3].astype(np.float32) * np.float32(min_depth)).astype(np.float16)
    far[:3] = (far[:3].astype(np.float32) * np.float32(max_depth)).astype(np.float16)
    world = _matmul_h_np(world_T_cam, np.concatenate([near, far], 1))
    return world.min(1)[:3], world.max(1)[:3]


class TSDF:
    """reference ``TSDF`` (tools/tsdf.py:53-337): fp16 voxel grid, values, weights."""

    VOX_MOD = 8  # volume dimensions are multiples of 8 (the kernels rely on it: one 16-byte vector per 8 voxels)

    def __init__(self, voxel_coords_3hwd, tsdf_values, tsdf_weights, voxel_size, origin, _origin_f32=None):
        self.voxel_coords_3hwd = None if voxel_coords_3hwd is None else voxel_coords_3hwd.half()
        self.tsdf_values = tsdf_values.half()
        self.tsdf_weights = tsdf_weights.half()
        self.voxel_size = voxel_size
        self.origin = origin.half()
        # set by from_bounds: the grid is origin + index * voxel_size, so the kernel regenerates it in registers
        self._origin_f32 = _origin_f32

    @classmethod
    def from_file(cls, tsdf_file):
        data = np.load(tsdf_file)
        return cls(torch.from_numpy(data["voxel_coords_3hwd"]), torch.from_numpy(data["tsdf_values"]),
                   torch.from_numpy(data["tsdf_weights"]), data["voxel_size"].item(), torch.from_numpy(data["origin"]))

    @classmethod
    def from_mesh(cls, mesh, voxel_size):
        """Bounds of ``mesh.vertices`` (any object with an (N,3) ``vertices`` array) plus 3 voxels (:102-121)."""
        vmax, vmin = np.asarray(mesh.vertices).max(0), np.asarray(mesh.vertices).min(0)
        bounds = {f"{a}min": vmin[i] - 3 * voxel_size for i, a in enumerate("xyz")}
        bounds.update({f"{a}max": vmax[i] + 3 * voxel_size for i, a in enumerate("xyz")})
        return cls.from_bounds(bounds, voxel_size)

    @classmethod
    def from_bounds(cls, bounds, voxel_size, lazy_grid=False):
        """reference :122-151.  ``lazy_grid=True`` (extension) skips materialising the (3,X,Y,Z) coordinate grid -- the
        kernels regenerate it bit-identically -- and allocates the volume directly in HBM: the reference's default
        20 m cube at 4 cm is 128 M voxels, i.e. 768 MB of coordinates that are never read here."""
        for key in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax"):
            if key not in bounds:
                raise KeyError("Provided bounds dict need to have keys'xmin', 'xmax', 'ymin', 'ymax', 'zmin', 'zmax'!")
        dims = tuple(int(np.ceil((bounds[a + "max"] - bounds[a + "min"]) / voxel_size / cls.VOX_MOD)) * cls.VOX_MOD
                     for a in "xyz")
        origin = torch.FloatTensor([bounds["xmin"], bounds["ymin"], bounds["zmin"]])
        if lazy_grid:
            values = torch.full(dims, -1.0, dtype=torch.float16, device="cuda")
            return cls(None, values, torch.zeros_like(values), voxel_size, origin, _origin_f32=origin.clone())
        coords = cls.generate_voxel_coords(origin, dims, voxel_size).half()
        return cls(coords, -torch.ones_like(coords[0]), torch.zeros_like(coords[0]), voxel_size, origin,
                   _origin_f32=origin.clone())

    @classmethod
    def generate_voxel_coords(cls, origin, volume_dims, voxel_size):
        grid = torch.meshgrid([torch.arange(vd) for vd in volume_dims], indexing="ij")
        return origin.view(3, 1, 1, 1) + torch.stack(grid, 0) * voxel_size

    def cuda(self):
        if self.voxel_coords_3hwd is not None:
            self.voxel_coords_3hwd = self.voxel_coords_3hwd.cuda()
        self.tsdf_values = self.tsdf_values.cuda()
        if self.tsdf_weights is not None:
            self.tsdf_weights = self.tsdf_weights.cuda()

    def cpu(self):
        if self.voxel_coords_3hwd is not None:
            self.voxel_coords_3hwd = self.voxel_coords_3hwd.cpu()
        self.tsdf_values = self.tsdf_values.cpu()
        if self.tsdf_weights is not None:
            self.tsdf_weights = self.tsdf_weights.cpu()

    def save_tsdf(self, filepath):
        if self.voxel_coords_3hwd is None:  # lazy grid: materialise it for the file format
            self.voxel_coords_3hwd = self.generate_voxel_coords(self._origin_f32, tuple(self.tsdf_values.shape),
                                                                self.voxel_size).half()
        np.savez_compressed(
            filepath, tsdf_values=self.tsdf_values.cpu().numpy().astype(np.float16),
            tsdf_weights=self.tsdf_weights.cpu().numpy().astype(np.float16),
            origin=self.origin.cpu().numpy().astype(np.float16),
            voxel_coords_3hwd=self.voxel_coords_3hwd.cpu().numpy().astype(np.float16), voxel_size=self.voxel_size)

    def to_mesh(self, *a, **k):
        raise NotImplementedError("doubletake_b200 renders the depth hint by ray casting the TSDF (render_depth_hint); "
                                  "mesh extraction for export (marching cubes) is not part of the engine")

    to_mesh_pytorch3d = save_mesh = to_mesh

    def render_depth_hint(self, world_T_cam_b44, invK_b44, height, width, z_near=0.05, z_far=10.0, weight_threshold=0.025,
                          max_steps=4096):
        """The rendered-depth hint of the incremental loop (reference test_incremental.py:186-252) without a mesh: ONE
        kernel ray-casts the fused TSDF from the current camera and returns the three tensors ``DepthModelCVHint.forward``
        consumes -- ``depth_hint_b1hw`` (NaN where no surface is seen or the fused confidence is below
        ``weight_threshold``), ``depth_hint_mask_b1hw`` (float 0/1) and ``sampled_weights_b1hw`` (0 where invalid) --
        plus the boolean ``depth_hint_mask_b_b1hw`` the reference keeps next to them.  It replaces
        ``fuser.get_mesh_pytorch3d`` (marching cubes), ``PyTorch3DMeshDepthRenderer.render``, ``BackprojectDepth``,
        ``fuser.sample_tsdf(..., "weights")`` and the masking lines :238-252.  ``invK_b44``: inverse intrinsics at the hint
        resolution (``cur_data["invK_s0_b44"]``); poses may live on the host or on the device."""
        self.cuda()
        dev = self.tsdf_values.device
        invK = L.f32(invK_b44, dev)
        pose = L.f32(world_T_cam_b44, dev)
        B = pose.shape[0]
        out = [torch.empty((B, 1, height, width), dtype=torch.float32, device=dev) for _ in range(3)]
        p = L.TsdfRaycastParams()
        p.values, p.weights = L.ptr(self.tsdf_values.contiguous()), L.ptr(self.tsdf_weights.contiguous())
        p.dims = (C.c_int32 * 3)(*self.tsdf_values.shape)
        p.origin_h = (C.c_float * 3)(*[float(v) for v in self.origin.float().cpu()])
        p.voxel_size = float(self.voxel_size)
        p.invK, p.world_T_cam = L.ptr(invK), L.ptr(pose)
        p.batch, p.height, p.width = B, height, width
        p.z_near, p.z_far, p.max_steps, p.weight_threshold = z_near, z_far, max_steps, weight_threshold
        p.depth_hint, p.hint_mask, p.sampled_weights = (L.ptr(t) for t in out)
        L.check(L.lib().dtb200_tsdf_raycast(C.byref(p), L.stream()))
        return {"depth_hint_b1hw": out[0], "depth_hint_mask_b1hw": out[1], "depth_hint_mask_b_b1hw": out[1] > 0,
                "sampled_weights_b1hw": out[2]}

    def sample_tsdf(self, world_points_N3, what_to_sample="tsdf", sampling_method="bilinear"):
        """(N,3) world points -> (N,) fp32 samples of the TSDF or the weights (tools/tsdf.py:277-337)."""
        if not (world_points_N3.ndim == 2 and world_points_N3.shape[1] == 3):
            raise ValueError("world_points_N3 must have shape (N, 3)! Instead got shape {}".format(world_points_N3.shape))
        if what_to_sample not in ("tsdf", "weights"):
            raise ValueError(f"what_to_sample must be 'tsdf' or 'weights', got {what_to_sample}")
        modes = {"bilinear": 0, "trilinear": 0, "nearest": 1}
        if sampling_method not in modes:
            raise ValueError(f"unknown sampling_method {sampling_method}")
        self.cuda()
        volume = (self.tsdf_values if what_to_sample == "tsdf" else self.tsdf_weights).contiguous()
        pts = L.f32(world_points_N3, volume.device)
        out = torch.empty(pts.shape[0], dtype=torch.float32, device=volume.device)
        dims = (C.c_int32 * 3)(*volume.shape)
        origin_h = (C.c_float * 3)(*[float(v) for v in self.origin.float().cpu()])
        L.check(L.lib().dtb200_tsdf_sample(L.ptr(volume), dims, origin_h, float(self.voxel_size), L.ptr(pts), L.ptr(out),
                                           pts.shape[0], modes[sampling_method], L.stream()))
        return out


def _host_half(x):
    """(B,4,4) poses / intrinsics as a host fp16 array.  numpy arrays and CPU tensors are used as they are (no device
    synchronisation: the incremental loop keeps poses on the host); a CUDA tensor costs one blocking copy."""
    if isinstance(x, np.ndarray):
        return x.astype(np.float16)
    return x.detach().to("cpu").half().numpy()


class TSDFFuser:
    """reference ``TSDFFuser`` (tools/tsdf.py:340-558)."""

    def __init__(self, tsdf, min_depth=0.5, max_depth=5.0, use_gpu=True, semantics="aten_cuda"):
        """``semantics`` selects which ATen build's fp16 ``grid_sample`` is reproduced: "aten_cuda" (default -- the
        reference fuser always runs with use_gpu=True, tools/fusers_helper.py) or "aten_cpu" (what fixtures generated by
        executing the reference on a CPU-only machine pin; differs only for non-finite / overflowing pixel coordinates)."""
        if not use_gpu:
            raise RuntimeError("doubletake_b200 runs on CUDA only (no CPU fallback): use_gpu must be True")
        if semantics not in L.TSDF_SEMANTICS:
            raise ValueError(f"semantics must be one of {sorted(L.TSDF_SEMANTICS)}")
        self.tsdf = tsdf
        self.min_depth = min_depth
        self.max_depth = max_depth
        self.use_gpu = use_gpu
        self.semantics = semantics
        self.truncation_size = 3.0
        self.maxW = 100.0

    voxel_coords_3hwd = property(lambda self: self.tsdf.voxel_coords_3hwd)
    def render_depth_hint(self, *args, **kwargs):
        """See ``TSDF.render_depth_hint``."""
        return self.tsdf.render_depth_hint(*args, **kwargs)

    tsdf_values = property(lambda self: self.tsdf.tsdf_values)
    tsdf_weights = property(lambda self: self.tsdf.tsdf_weights)
    voxel_size = property(lambda self: self.tsdf.voxel_size)
    shape = property(lambda self: self.tsdf.tsdf_values.shape)
    truncation = property(lambda self: self.truncation_size * self.tsdf.voxel_size)

    def _frame_constants(self, cam_T_world_44, K_44, img_h, img_w):
        """Per-frame host constants with the reference's own small fp16 ops (tools/tsdf.py:444-450,398-407): fp32 inverses
        rounded to fp16, the frustum box, P = (K @ cam_T_world)[:3].  numpy on 4x4 arrays (a few tens of microseconds)."""
        K = np.asarray(K_44, dtype=np.float16)
        T = np.asarray(cam_T_world_44, dtype=np.float16)
        depth_max = self.max_depth + self.truncation + 0.1
        invK = np.linalg.inv(K.astype(np.float32)).astype(np.float32).astype(np.float16)
        world_T_cam = np.linalg.inv(T.astype(np.float32)).astype(np.float32).astype(np.float16)
        lo, hi = _frustum_bounds_np(invK, world_T_cam, 0.01, depth_max, img_h, img_w)
        P = _matmul_h_np(K, T)[:3]
        return P.astype(np.float32).ravel().tolist(), lo.astype(np.float32).tolist(), hi.astype(np.float32).tolist()

    def _index_box(self, lo, hi, dims):
        """Conservative index cover [begin, end) of the voxels whose fp16 coordinate can lie strictly inside (lo, hi).
        Coordinates grow monotonically with the index, but the stored coordinate is ROUNDED to fp16: far from the origin one
        fp16 ulp spans several voxels (0.125 m at |x| >= 128 with 4 cm voxels), so the margin is derived from the ulp at the
        largest coordinate magnitude of the axis.  Caller-supplied grids (TSDF.from_file), non-finite boxes and volumes
        whose ulp exceeds 16 voxels scan the whole volume; z bounds are multiples of 8."""
        full = [0, 0, 0], list(dims)
        if self.tsdf._origin_f32 is None or not all(np.isfinite(lo + hi)):
            return full
        begin, end = [], []
        for a in range(3):
            o, vs = float(self.tsdf._origin_f32[a]), float(self.voxel_size)
            reach = max(abs(o), abs(o + dims[a] * vs), abs(lo[a]), abs(hi[a]), 6.2e-5)
            ulp = 2.0 ** (int(np.floor(np.log2(reach))) - 10)  # fp16 spacing at that magnitude
            margin = int(np.ceil(ulp / vs)) + 2
            if margin > 18:
                return full
            b = int(np.floor((lo[a] - o) / vs)) - margin
            e = int(np.ceil((hi[a] - o) / vs)) + margin + 1
            begin.append(min(max(b, 0), dims[a]))
            end.append(min(max(e, 0), dims[a]))
        begin[2] = begin[2] // 8 * 8
        end[2] = min((end[2] + 7) // 8 * 8, dims[2])
        return begin, end

    def integrate_depth(self, depth_b1hw, cam_T_world_T_b44, K_b44, depth_mask_b1hw=None, extended_neg_truncation=False):
        """Integrates a batch of depth maps into the volume, in order (tools/tsdf.py:414-558)."""
        self.tsdf.cuda()
        values, weights = self.tsdf.tsdf_values, self.tsdf.tsdf_weights
        if not (values.is_contiguous() and weights.is_contiguous()):
            raise RuntimeError("TSDF volumes must be contiguous")
        dev = values.device
        depth = depth_b1hw.to(dev).half().contiguous()
        mask = None if depth_mask_b1hw is None else depth_mask_b1hw.to(dev).to(torch.uint8).contiguous()
        B, _, img_h, img_w = depth.shape
        T_cpu, K_cpu = _host_half(cam_T_world_T_b44), _host_half(K_b44)
        p = L.TsdfIntegrateParams()
        p.values, p.weights = L.ptr(values), L.ptr(weights)
        if self.tsdf._origin_f32 is not None:
            p.voxel_coords = None
            p.origin = (C.c_float * 3)(*[float(v) for v in self.tsdf._origin_f32])
        else:
            p.voxel_coords = L.ptr(self.tsdf.voxel_coords_3hwd.contiguous())
        p.voxel_size = float(self.voxel_size)
        p.dims = (C.c_int32 * 3)(*values.shape)
        p.img_h, p.img_w = img_h, img_w
        p.semantics = L.TSDF_SEMANTICS[self.semantics]
        p.min_depth = float(self.min_depth)
        p.depth_range = float(self.max_depth - self.min_depth)
        p.max_depth_h = float(torch.tensor(self.max_depth).half())
        p.truncation = float(self.truncation)
        p.trunc_check_h = float(torch.tensor(-self.truncation * (1.5 if extended_neg_truncation else 1.0)).half())
        dims = tuple(values.shape)
        for s in range(0, B, L.TSDF_MAX_FRAMES):
            n = min(L.TSDF_MAX_FRAMES, B - s)
            p.num_frames = n
            begin, end = list(dims), [0, 0, 0]
            for i in range(n):
                P, lo, hi = self._frame_constants(T_cpu[s + i], K_cpu[s + i], img_h, img_w)
                b_i, e_i = self._index_box(lo, hi, dims)
                begin = [min(a, b) for a, b in zip(begin, b_i)]
                end = [max(a, b) for a, b in zip(end, e_i)]
                fr = p.frames[i]
                fr.depth = depth[s + i].data_ptr()
                fr.mask = None if mask is None else mask[s + i].data_ptr()
                fr.P = (C.c_float * 12)(*P)
                fr.box_min = (C.c_float * 3)(*lo)
                fr.box_max = (C.c_float * 3)(*hi)
            end = [max(b, e) for b, e in zip(begin, end)]  # empty cover -> empty box
            p.vox_begin, p.vox_end = (C.c_int32 * 3)(*begin), (C.c_int32 * 3)(*end)
            L.check(L.lib().dtb200_tsdf_integrate(C.byref(p), L.stream()))
        # `depth` / `mask` stay referenced until the launches are enqueued; the caching allocator keeps them valid on this stream
