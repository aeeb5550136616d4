"""ctypes binding of include/doubletake_b200.h -- the only way the Python host layer reaches the CUDA kernels.

There is NO CPU or PyTorch fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdoubletake_b200.so")

VOLUME_DOT, VOLUME_MLP, VOLUME_MLP_HINT = 0, 1, 2
MATH_EXACT, MATH_TC3X, MATH_TCH = 0, 1, 2
RESAMPLE_NONE, RESAMPLE_BILINEAR_UP2, RESAMPLE_NEAREST_UP2 = 0, 1, 2
ACT_NONE, ACT_LEAKY, ACT_ELU = 0, 1, 2
CONV_MAX_SRC = 3
MATH_NAMES = {"exact": MATH_EXACT, "tc3x": MATH_TC3X, "tch": MATH_TCH}

fp = C.c_void_p


class CostVolumeParams(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("math", C.c_int32),
        ("batch", C.c_int32), ("views", C.c_int32), ("channels", C.c_int32),
        ("height", C.c_int32), ("width", C.c_int32), ("planes", C.c_int32),
        ("cur_feats", fp), ("src_feats_nhwc", fp), ("src_extrinsics", fp), ("src_poses", fp),
        ("src_Ks", fp), ("cur_invK", fp), ("plane_depths", fp), ("planes_per_pixel", C.c_int32),
        ("depth_hint", fp), ("hint_weights", fp), ("hint_mask", fp),
        ("hint_height", C.c_int32), ("hint_width", C.c_int32),
        ("w1", fp), ("b1", fp), ("w2", fp), ("b2", fp), ("w3", fp), ("b3", fp),
        ("hw1", fp), ("hb1", fp), ("hw2", fp), ("hb2", fp), ("hw3", fp), ("hb3", fp),
        ("volume", fp), ("lowest_cost", fp), ("best_index", fp), ("mask_views", fp), ("mask_any", fp),
        ("workspace", fp), ("workspace_bytes", C.c_uint64), ("workspace_prepared", C.c_int32),
    ]


class ConvParams(C.Structure):
    _fields_ = [
        ("math", C.c_int32), ("batch", C.c_int32),
        ("in_h", C.c_int32), ("in_w", C.c_int32),
        ("out_h", C.c_int32), ("out_w", C.c_int32), ("out_c", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32), ("num_src", C.c_int32),
        ("src", fp * CONV_MAX_SRC), ("src_c", C.c_int32 * CONV_MAX_SRC), ("src_resample", C.c_int32 * CONV_MAX_SRC),
        ("weight", fp), ("bias", fp), ("residual", fp),
        ("act", C.c_int32), ("act_slope", C.c_float),
        ("dst", fp),
        ("workspace", fp), ("workspace_bytes", C.c_uint64),
    ]


TSDF_MAX_FRAMES = 8
TSDF_SEMANTICS = {"aten_cpu": 0, "aten_cuda": 1, "aten_cuda_half_index": 2}


class TsdfFrame(C.Structure):
    _fields_ = [("depth", fp), ("mask", fp), ("P", C.c_float * 12), ("box_min", C.c_float * 3), ("box_max", C.c_float * 3)]


class TsdfIntegrateParams(C.Structure):
    _fields_ = [
        ("values", fp), ("weights", fp), ("voxel_coords", fp),
        ("origin", C.c_float * 3), ("voxel_size", C.c_float), ("dims", C.c_int32 * 3),
        ("vox_begin", C.c_int32 * 3), ("vox_end", C.c_int32 * 3),
        ("img_h", C.c_int32), ("img_w", C.c_int32), ("num_frames", C.c_int32), ("semantics", C.c_int32),
        ("min_depth", C.c_float), ("depth_range", C.c_float), ("max_depth_h", C.c_float), ("truncation", C.c_float),
        ("trunc_check_h", C.c_float),
        ("frames", TsdfFrame * TSDF_MAX_FRAMES),
    ]


LAYOUT_F32, LAYOUT_SPLIT16 = 0, 1


class InstanceNormParams(C.Structure):
    _fields_ = [
        ("src", fp), ("src_layout", C.c_int32), ("src_channels", C.c_int32), ("src_border", C.c_int32),
        ("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("channels", C.c_int32),
        ("eps", C.c_float), ("act", C.c_int32), ("act_slope", C.c_float),
        ("dst", fp), ("dst_layout", C.c_int32), ("dst_border", C.c_int32), ("dst_nchw", fp), ("stats", fp),
    ]


class TsdfRaycastParams(C.Structure):
    _fields_ = [
        ("values", fp), ("weights", fp), ("dims", C.c_int32 * 3), ("origin_h", C.c_float * 3), ("voxel_size", C.c_float),
        ("invK", fp), ("world_T_cam", fp), ("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("z_near", C.c_float), ("z_far", C.c_float), ("max_steps", C.c_int32), ("weight_threshold", C.c_float),
        ("depth_hint", fp), ("hint_mask", fp), ("sampled_weights", fp),
    ]


# every symbol include/doubletake_b200.h declares (tests/test_capi_symbols.py checks the header against this list)
SYMBOLS = {
    "dtb200_abi_version": (C.c_int, []),
    "dtb200_debug_set": (C.c_int, [C.c_int]),
    "dtb200_debug_trace": (C.c_int, [C.c_void_p, C.c_int32]),
    "dtb200_last_error": (C.c_char_p, []),
    "dtb200_launch_count": (C.c_uint64, []),
    "dtb200_nchw_to_nhwc": (C.c_int, [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, fp]),
    "dtb200_nhwc_to_nchw": (C.c_int, [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, fp]),
    "dtb200_nchw_to_split16": (C.c_int, [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, fp]),
    "dtb200_split16_to_nchw": (C.c_int, [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, fp]),
    "dtb200_cost_volume_workspace_bytes": (C.c_uint64, [C.POINTER(CostVolumeParams)]),
    "dtb200_cost_volume_prepare": (C.c_int, [C.POINTER(CostVolumeParams), fp]),
    "dtb200_cost_volume": (C.c_int, [C.POINTER(CostVolumeParams), fp]),
    "dtb200_packed_conv_weight_floats": (C.c_uint64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "dtb200_pack_conv_weight": (C.c_int, [C.c_int32, fp, fp, C.c_int32, C.c_int32, C.c_int32, fp]),
    "dtb200_packed_conv_weight_floats_srcs": (C.c_uint64, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32]),
    "dtb200_pack_conv_weight_srcs": (C.c_int, [C.c_int32, fp, fp, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32, fp]),
    "dtb200_conv_workspace_bytes": (C.c_uint64, [C.POINTER(ConvParams)]),
    "dtb200_conv2d": (C.c_int, [C.POINTER(ConvParams), fp]),
    "dtb200_conv2d_sequence": (C.c_int, [C.POINTER(ConvParams), C.c_int32, fp]),
    "dtb200_conv_graph_create": (C.c_int, [C.POINTER(ConvParams), C.c_int32, C.c_int32, C.POINTER(fp)]),
    "dtb200_conv_graph_launch": (C.c_int, [fp, fp]),
    "dtb200_conv_graph_info": (C.c_int, [fp] + [C.POINTER(C.c_int32)] * 5),
    "dtb200_conv_graph_destroy": (None, [fp]),
    "dtb200_conv_graph_analyze": (C.c_int, [C.POINTER(ConvParams), C.c_int32, C.c_int32] + [C.POINTER(C.c_int32)] * 4
                                  + [C.c_int32]),
    "dtb200_relative_poses": (C.c_int, [fp, fp, fp, fp, fp, fp, C.c_int, C.c_int, fp]),
    "dtb200_exp": (C.c_int, [fp, fp, C.c_uint64, fp]),
    "dtb200_tsdf_integrate": (C.c_int, [C.POINTER(TsdfIntegrateParams), fp]),
    "dtb200_encoder_stem": (C.c_int, [fp, fp, fp, fp, fp, C.c_int, C.c_int, C.c_int, fp]),
    "dtb200_encoder_pool": (C.c_int, [fp, fp, C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int32, fp]),
    "dtb200_instance_norm": (C.c_int, [C.POINTER(InstanceNormParams), fp]),
    "dtb200_tsdf_raycast": (C.c_int, [C.POINTER(TsdfRaycastParams), fp]),
    "dtb200_tsdf_sample": (C.c_int, [fp, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.c_float, fp, fp, C.c_int64, C.c_int32, fp]),
}

_lib = None


def lib():
    """Load (once) and return the C-ABI library.  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m doubletake_b200.build` (nvcc, sm_100a). "
                "doubletake_b200 has no CPU / PyTorch fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.dtb200_abi_version() != 1:
            raise RuntimeError("libdoubletake_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc: int):
    if rc != 0:
        raise RuntimeError(f"doubletake_b200 C-ABI error {rc}: {lib().dtb200_last_error().decode()}")


def ptr(t):
    """Device pointer of a dense CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("doubletake_b200 kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("doubletake_b200 kernels need contiguous tensors")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def f32(t, device=None):
    """Dense fp32 view/copy of `t` (optionally moved to `device`)."""
    if t.dtype != torch.float32:
        t = t.float()
    if device is not None and t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def nchw_to_nhwc(x, out=None):
    """(N,C,H,W) fp32 CUDA -> (N,H,W,C) via the library's transpose kernel."""
    x = f32(x)
    n, c, h, w = x.shape
    if out is None:
        out = torch.empty((n, h, w, c), dtype=torch.float32, device=x.device)
    check(lib().dtb200_nchw_to_nhwc(ptr(x), ptr(out), n, c, h, w, stream()))
    return out


def nhwc_to_nchw(x, out=None):
    x = f32(x)
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    check(lib().dtb200_nhwc_to_nchw(ptr(x), ptr(out), n, c, h, w, stream()))
    return out


def nchw_to_split16(x, out):
    """(N,C,H,W) fp32 CUDA -> split16 (N,H,W,2,C) fp16 written into `out` (any dense buffer of N*H*W*C*4 bytes)."""
    x = f32(x)
    n, c, h, w = x.shape
    check(lib().dtb200_nchw_to_split16(ptr(x), ptr(out), n, c, h, w, stream()))
    return out


def split16_to_nchw(buf, n, c, h, w):
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=buf.device)
    check(lib().dtb200_split16_to_nchw(ptr(buf), ptr(out), n, c, h, w, stream()))
    return out


def launch_count():
    return int(lib().dtb200_launch_count())
