"""On-disk formats on either side of the hot path, byte-compatible with the reference so its dataset
classes, evaluation and fusion scripts interoperate with this engine:

* rendered depth-hint PNG pairs (``rendered_depth_<frame>.png`` 16-bit, depth x 2048, 0 = no hint;
  ``sampled_weights_<frame>.png`` 16-bit, weight x 8192) as ``ScannetDataset.load_depth_hint`` reads them
  (reference datasets/scannet_dataset.py:577-630, utils/generic_utils.py:221-268) -- the bridge between the first and the
  second pass of the offline two-pass mode (test_offline_two_pass.py);
* cached model outputs, one pickle per keyframe, as ``cache_model_outputs`` writes them (utils/generic_utils.py:304-352);
* keyframe tuple files and the TSDF ``.npz`` live next to their users (``sharding.read_frame_tuples``, ``tsdf.TSDF``).

Host-side I/O only (PIL / pickle): nothing here touches the GPU path.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch
from PIL import Image

DEPTH_HINT_SCALE = 2048.0     # scannet_dataset.py:608  read_image_file(..., value_scale_factor=1 / 2048)
HINT_WEIGHT_SCALE = 8192.0    # scannet_dataset.py:616  read_image_file(..., value_scale_factor=1 / 8192)


def _read_u16_png(path, scale):
    """``read_image_file`` (generic_utils.py:221-268) for a single-channel 16-bit PNG: PIL -> float tensor (1,H,W) / scale.
    (torchvision's ``to_tensor`` turns an ``I;16`` image into the raw integer values as float; so does this.)"""
    img = np.asarray(Image.open(path))
    return torch.from_numpy(img.astype(np.float32))[None] * (1.0 / scale)


def _write_u16_png(path, x_hw, scale):
    v = np.nan_to_num(np.asarray(x_hw, dtype=np.float64), nan=0.0, posinf=0.0, neginf=0.0) * scale
    Image.fromarray(np.clip(np.rint(v), 0, 65535).astype(np.uint16)).save(path)


def hint_paths(depth_hint_dir, scan_id, frame_id):
    d = os.path.join(depth_hint_dir, scan_id)
    return (os.path.join(d, f"rendered_depth_{int(frame_id)}.png"), os.path.join(d, f"sampled_weights_{int(frame_id)}.png"))


def write_depth_hint(depth_hint_dir, scan_id, frame_id, hint):
    """Store one keyframe's hint (the dict ``TSDF.render_depth_hint`` / the reference's incremental loop produce; tensors of
    shape (1,1,H,W) or (1,H,W)) as the PNG pair the reference dataset loads.  NaN / masked pixels become 0 = "no hint"."""
    os.makedirs(os.path.join(depth_hint_dir, scan_id), exist_ok=True)
    depth = torch.as_tensor(hint["depth_hint_b1hw"]).detach().float().cpu().reshape(hint["depth_hint_b1hw"].shape[-2:])
    weights = torch.as_tensor(hint["sampled_weights_b1hw"]).detach().float().cpu().reshape(depth.shape)
    if "depth_hint_mask_b1hw" in hint:
        mask = torch.as_tensor(hint["depth_hint_mask_b1hw"]).detach().float().cpu().reshape(depth.shape) > 0
        depth = torch.where(mask, depth, torch.zeros_like(depth))
    dpath, wpath = hint_paths(depth_hint_dir, scan_id, frame_id)
    _write_u16_png(dpath, depth.numpy(), DEPTH_HINT_SCALE)
    _write_u16_png(wpath, weights.numpy(), HINT_WEIGHT_SCALE)
    return dpath, wpath


def load_depth_hint(depth_hint_dir, scan_id, frame_id, depth_height=None, depth_width=None, mark_all_empty=False, flip=False):
    """``ScannetDataset.load_depth_hint`` (scannet_dataset.py:577-630) without the training-only partial-render coin flip:
    returns ``depth_hint_b1hw`` (NaN where empty), ``depth_hint_mask_b1hw`` (float), ``depth_hint_mask_b_b1hw`` (bool),
    ``sampled_weights_b1hw``, each (1,H,W) like a dataset item before collation."""
    if mark_all_empty:
        if depth_height is None or depth_width is None:
            raise ValueError("mark_all_empty needs depth_height and depth_width")
        depth = torch.full((1, depth_height, depth_width), float("nan"))
        mask = torch.zeros_like(depth)
        mask_b = torch.zeros_like(depth).bool()
        weights = torch.zeros_like(depth)
    else:
        dpath, wpath = hint_paths(depth_hint_dir, scan_id, frame_id)
        depth = _read_u16_png(dpath, DEPTH_HINT_SCALE)
        mask = (depth > 0).float()
        mask_b = depth > 0
        depth[~mask_b] = float("nan")
        weights = _read_u16_png(wpath, HINT_WEIGHT_SCALE)
        if flip:
            depth, mask, mask_b, weights = (torch.flip(t, (-1,)) for t in (depth, mask, mask_b, weights))
    return {"depth_hint_b1hw": depth, "depth_hint_mask_b1hw": mask, "depth_hint_mask_b_b1hw": mask_b,
            "sampled_weights_b1hw": weights}


def cache_model_outputs(output_path, outputs, cur_data, src_data, batch_ind, batch_size):
    """``cache_model_outputs`` (generic_utils.py:304-352): one ``<frame_id>.pickle`` per batch element holding
    ``depth_pred_s0_b1hw``, ``overall_mask_bhw``, (``cv_confidence_b1hw``), ``K_full_depth_b44``, ``K_s0_b44``,
    ``cam_T_world_b44``, ``frame_id`` and ``src_ids`` -- what the reference's fusion / evaluation scripts read back.
    Tensors are stored on the CPU."""
    os.makedirs(output_path, exist_ok=True)
    paths = []
    for i in range(outputs["depth_pred_s0_b1hw"].shape[0]):
        if "frame_id_string" in cur_data:
            frame_id = cur_data["frame_id_string"][i]
        else:
            frame_id = f"{batch_ind * batch_size + i:6d}"   # the reference's f"{str(frame_id):6d}" raises; this is its intent
        elem = {"depth_pred_s0_b1hw": outputs["depth_pred_s0_b1hw"][i].unsqueeze(0).detach().cpu(),
                "overall_mask_bhw": None if outputs.get("overall_mask_bhw") is None
                else outputs["overall_mask_bhw"][i].unsqueeze(0).detach().cpu()}
        if "cv_confidence_b1hw" in outputs:
            elem["cv_confidence_b1hw"] = outputs["cv_confidence_b1hw"][i].unsqueeze(0).detach().cpu()
        for key in ("K_full_depth_b44", "K_s0_b44", "cam_T_world_b44"):
            elem[key] = cur_data[key][i].unsqueeze(0).detach().cpu()
        elem["frame_id"] = frame_id
        elem["src_ids"] = [ids[i] for ids in src_data.get("frame_id_string", [])]
        path = os.path.join(output_path, f"{frame_id}.pickle")
        with open(path, "wb") as handle:
            pickle.dump(elem, handle)
        paths.append(path)
    return paths


def load_cached_outputs(path):
    with open(path, "rb") as handle:
        return pickle.load(handle)
