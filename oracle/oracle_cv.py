"""ctypes loader of the plain-C cost-volume oracle (oracle/oracle_cv.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle_cv.so")
fp = C.c_void_p


class Params(C.Structure):
    _fields_ = [("kind", C.c_int32), ("B", C.c_int32), ("K", C.c_int32), ("C", C.c_int32), ("H", C.c_int32),
                ("W", C.c_int32), ("D", C.c_int32),
                ("cur_feats", fp), ("src_feats", fp), ("src_extrinsics", fp), ("src_poses", fp), ("src_Ks", fp),
                ("cur_invK", fp), ("planes", fp),
                ("w1", fp), ("b1", fp), ("w2", fp), ("b2", fp), ("w3", fp), ("b3", fp),
                ("hw1", fp), ("hb1", fp), ("hw2", fp), ("hb2", fp), ("hw3", fp), ("hb3", fp),
                ("hint_depth", fp), ("hint_weights", fp), ("hint_mask", fp), ("hint_h", C.c_int32), ("hint_w", C.c_int32),
                ("volume", fp), ("index", fp), ("lowest", fp), ("mask_views", fp), ("mask_any", fp)]


def lib():
    src = os.path.join(HERE, "oracle_cv.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    h = C.CDLL(LIB)
    h.orc_cost_volume.restype = C.c_int
    h.orc_cost_volume.argtypes = [C.POINTER(Params)]
    return h


def cost_volume(kind, cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, planes, weights=None, hint=None):
    """kind: "dot" | "mlp" | "hint".  Arrays are numpy / torch fp32; ``weights`` uses the reference's state_dict keys;
    ``hint``: dict with depth_hint_b1hw / sampled_weights_b1hw / depth_hint_mask_b1hw.  Returns a dict of numpy arrays."""
    keep = []

    def arr(x, dtype=np.float32):
        a = np.ascontiguousarray(np.asarray(x.detach().cpu().numpy() if hasattr(x, "detach") else x, dtype=dtype))
        keep.append(a)
        return a

    p = Params()
    src = arr(src_feats)
    p.B, p.K, p.C, p.H, p.W = src.shape
    pl = arr(planes).reshape(-1)
    p.D = pl.shape[0]
    p.kind = {"dot": 0, "mlp": 1, "hint": 2}[kind]
    for name, val in (("cur_feats", cur_feats), ("src_extrinsics", src_extrinsics), ("src_poses", src_poses),
                      ("src_Ks", src_Ks), ("cur_invK", cur_invK)):
        setattr(p, name, arr(val).ctypes.data)
    p.src_feats, p.planes = src.ctypes.data, pl.ctypes.data
    if kind != "dot":
        for f, key in (("w1", "mlp.net.0.weight"), ("b1", "mlp.net.0.bias"), ("w2", "mlp.net.2.weight"),
                       ("b2", "mlp.net.2.bias"), ("w3", "mlp.net.4.weight"), ("b3", "mlp.net.4.bias")):
            setattr(p, f, arr(weights[key]).ctypes.data)
    if kind == "hint":
        for f, key in (("hw1", "hint_mlp.net.0.weight"), ("hb1", "hint_mlp.net.0.bias"), ("hw2", "hint_mlp.net.2.weight"),
                       ("hb2", "hint_mlp.net.2.bias"), ("hw3", "hint_mlp.net.4.weight"), ("hb3", "hint_mlp.net.4.bias")):
            setattr(p, f, arr(weights[key]).ctypes.data)
        hd = arr(hint["depth_hint_b1hw"])
        p.hint_h, p.hint_w = hd.shape[-2:]
        p.hint_depth = hd.ctypes.data
        p.hint_weights = arr(hint["sampled_weights_b1hw"]).ctypes.data
        p.hint_mask = arr(hint["depth_hint_mask_b1hw"]).ctypes.data
    out = dict(volume=np.empty((p.B, p.D, p.H, p.W), np.float32), index=np.empty((p.B, p.H, p.W), np.int32),
               lowest_cost=np.empty((p.B, p.H, p.W), np.float32), mask_views=np.zeros((p.B, p.K, p.H, p.W), np.uint8),
               mask_any=np.zeros((p.B, p.H, p.W), np.uint8))
    p.volume, p.index, p.lowest = out["volume"].ctypes.data, out["index"].ctypes.data, out["lowest_cost"].ctypes.data
    p.mask_views, p.mask_any = out["mask_views"].ctypes.data, out["mask_any"].ctypes.data
    rc = lib().orc_cost_volume(C.byref(p))
    if rc != 0:
        raise RuntimeError(f"orc_cost_volume failed with {rc}")
    out["planes"] = pl
    return out
