"""On-disk formats: hint PNG pairs and cached-output pickles, against the reference's own readers where
the reference tree is present (build container)."""
import os
import sys

import numpy as np
import pytest
import torch

from doubletake_b200 import formats as fm

REF = "/root/reference/src"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_hint(h=24, w=32, seed=0):
    g = torch.Generator().manual_seed(seed)
    depth = torch.rand(1, 1, h, w, generator=g) * 4 + 0.5
    mask = (torch.rand(1, 1, h, w, generator=g) > 0.3).float()
    depth[mask == 0] = float("nan")
    weights = torch.rand(1, 1, h, w, generator=g) * mask
    return {"depth_hint_b1hw": depth, "depth_hint_mask_b1hw": mask, "sampled_weights_b1hw": weights}


def test_hint_png_round_trip(tmp_path):
    hint = make_hint()
    dpath, wpath = fm.write_depth_hint(str(tmp_path), "scene0707_00", "000120", hint)
    assert dpath.endswith("scene0707_00/rendered_depth_120.png") and wpath.endswith("sampled_weights_120.png")
    back = fm.load_depth_hint(str(tmp_path), "scene0707_00", 120)
    assert back["depth_hint_b1hw"].shape == (1, 24, 32)
    mask = hint["depth_hint_mask_b1hw"][0] > 0
    assert torch.equal(back["depth_hint_mask_b_b1hw"], mask) and torch.equal(back["depth_hint_mask_b1hw"], mask.float())
    assert torch.isnan(back["depth_hint_b1hw"][~mask]).all()
    # 16-bit quantisation: depth to 1/2048 m, weights to 1/8192
    assert float((back["depth_hint_b1hw"][mask] - hint["depth_hint_b1hw"][0][mask]).abs().max()) <= 0.5 / 2048 + 1e-7
    assert float((back["sampled_weights_b1hw"] - hint["sampled_weights_b1hw"][0]).abs().max()) <= 0.5 / 8192 + 1e-7
    flipped = fm.load_depth_hint(str(tmp_path), "scene0707_00", 120, flip=True)
    assert torch.equal(flipped["sampled_weights_b1hw"], torch.flip(back["sampled_weights_b1hw"], (-1,)))
    empty = fm.load_depth_hint(str(tmp_path), "x", 0, 24, 32, mark_all_empty=True)
    assert torch.isnan(empty["depth_hint_b1hw"]).all() and not empty["depth_hint_mask_b_b1hw"].any()
    with pytest.raises(ValueError):
        fm.load_depth_hint(str(tmp_path), "x", 0, mark_all_empty=True)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_hint_pngs_read_back_identically_through_the_reference_reader(tmp_path):
    """The reference's own ``read_image_file`` (utils/generic_utils.py:221-268) on the files this module writes."""
    for p in (os.path.join(ROOT, "oracle", "ref_stubs"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    from doubletake.utils.generic_utils import read_image_file

    hint = make_hint(seed=3)
    dpath, wpath = fm.write_depth_hint(str(tmp_path), "s", 7, hint)
    ours = fm.load_depth_hint(str(tmp_path), "s", 7)
    ref_depth = read_image_file(dpath, value_scale_factor=1 / 2048)
    ref_w = read_image_file(wpath, value_scale_factor=1 / 8192)
    assert ref_depth.shape == (1, 24, 32)
    assert torch.equal(ref_w, ours["sampled_weights_b1hw"])
    m = ref_depth > 0
    assert torch.equal(m, ours["depth_hint_mask_b_b1hw"]) and torch.equal(ref_depth[m], ours["depth_hint_b1hw"][m])


def test_cached_output_pickles(tmp_path):
    B = 2
    outputs = {"depth_pred_s0_b1hw": torch.rand(B, 1, 6, 8), "overall_mask_bhw": torch.rand(B, 6, 8) > 0.5}
    cur = {"frame_id_string": ["000010", "000020"], "K_full_depth_b44": torch.eye(4).expand(B, 4, 4),
           "K_s0_b44": torch.eye(4).expand(B, 4, 4) * 2, "cam_T_world_b44": torch.eye(4).expand(B, 4, 4)}
    src = {"frame_id_string": [["000001", "000002"], ["000003", "000004"], ["000005", "000006"]]}
    paths = fm.cache_model_outputs(str(tmp_path), outputs, cur, src, batch_ind=0, batch_size=B)
    assert [os.path.basename(p) for p in paths] == ["000010.pickle", "000020.pickle"]
    e = fm.load_cached_outputs(paths[1])
    assert set(e) == {"depth_pred_s0_b1hw", "overall_mask_bhw", "K_full_depth_b44", "K_s0_b44", "cam_T_world_b44", "frame_id", "src_ids"}
    assert e["frame_id"] == "000020" and e["src_ids"] == ["000002", "000004", "000006"]
    assert torch.equal(e["depth_pred_s0_b1hw"], outputs["depth_pred_s0_b1hw"][1:2]) and e["K_s0_b44"].shape == (1, 4, 4)
    outputs["cv_confidence_b1hw"] = torch.rand(B, 1, 6, 8)
    cur2 = {k: v for k, v in cur.items() if k != "frame_id_string"}
    paths = fm.cache_model_outputs(str(tmp_path / "n"), outputs, cur2, {}, batch_ind=3, batch_size=B)
    assert os.path.basename(paths[0]) == "     6.pickle" and "cv_confidence_b1hw" in fm.load_cached_outputs(paths[0])
