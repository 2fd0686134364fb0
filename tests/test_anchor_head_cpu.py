"""AnchorHeadSingleV2 mirror (cpd_b200/anchor_head.py) against the REFERENCE's own class, imported unmodified through
cpd_b200.compat.reference: anchors, the point-occupancy anchor mask, AxisAlignedTargetAssigner targets, the three
losses and the decoded boxes, all on the same conv outputs (pure torch on both sides => CPU; the convolutions themselves are
DenseConv2d, covered by the -m gpu tests).  Needs the reference checkout (skipped on the GPU box)."""
import importlib
import os

import numpy as np
import pytest
import torch

from cpd_b200.synth import synth_gt_boxes, synth_scan

REF = os.environ.get("CPD_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "cpd")), reason="reference checkout not present")

RANGE = [-75.2, -75.2, -2.0, 75.2, 75.2, 4.0]
GRID = [1504, 1504, 40]
NAMES = ["Vehicle", "Pedestrian", "Cyclist"]


def _cfg(match_height=False):
    from cpd_b200 import anchor_head
    return anchor_head.default_cfg(match_height)


@pytest.fixture(scope="module")
def ref_mod():
    from cpd_b200.compat import reference
    reference.install_reference(REF)
    try:
        import cv2  # noqa: F401
    except ImportError:                     # anchor_head_single.py imports cv2 at module level and never uses it
        import sys
        import types
        sys.modules.setdefault("cv2", types.ModuleType("cv2"))
    mod = importlib.import_module("cpd.models.dense_heads.anchor_head_single")
    yield mod
    reference.uninstall_reference()


def test_anchor_head_v2_matches_reference(ref_mod, monkeypatch):
    from cpd_b200 import anchor_head
    from cpd_b200.compat.reference import EasyDict
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    torch.manual_seed(0)
    cfg = _cfg()
    ref = ref_mod.AnchorHeadSingleV2(EasyDict(cfg), 1, 64, 3, NAMES, np.array(GRID), np.array(RANGE), predict_boxes_when_training=True).train()
    mine = anchor_head.AnchorHeadSingleV2(cfg, 1, 64, 3, NAMES, GRID, RANGE, predict_boxes_when_training=True).train()
    # same parameter names and shapes => reference checkpoints load
    rs, ms = ref.state_dict(), mine.state_dict()
    assert set(rs) == set(ms) and all(tuple(rs[k].shape) == tuple(ms[k].shape) for k in rs)
    assert float(ms["conv_cls.3.bias"][0]) == np.float32(-4.59)
    for a, b in zip(ref.anchors_root, mine.anchors_root):
        assert torch.equal(a, b)
    B, H, W = 2, 188, 188
    pts = [torch.from_numpy(synth_scan(6000, 3 + i)) for i in range(B)]
    pts[1][:200, 0] = -74.9                                       # points in the first block column / row: the wrap-around of the mask
    pts[1][:200, 1] = -74.9
    points = torch.cat([torch.cat([torch.full((p.shape[0], 1), float(i)), p], 1) for i, p in enumerate(pts)], 0)
    gt = torch.from_numpy(np.stack([synth_gt_boxes(30, 11 + i) for i in range(B)])).float()
    gt[1, 20:] = 0                                                # trailing padding rows
    gt[0, 5] = 0                                                  # a padding row in the middle
    n = mine.num_anchors_per_location
    feat = torch.randn(B, 64, H, W)
    # drive the reference's forward with fixed conv outputs (its nn.Conv2d layers are not what is under test)
    outs = dict(cls=torch.randn(B, n * 3, H, W) * 2 - 3, reg=torch.randn(B, n * 2, H, W) * 0.3, height=torch.randn(B, n, H, W) * 0.3,
                dim=torch.randn(B, n * 3, H, W) * 0.2, ang=torch.randn(B, n, H, W) * 0.5, dir=torch.randn(B, n * 2, H, W))
    for name, key in (("conv_cls", "cls"), ("conv_reg", "reg"), ("conv_height", "height"), ("conv_dim", "dim"), ("conv_ang", "ang"), ("conv_dir_cls", "dir")):
        monkeypatch.setattr(getattr(ref, name), "forward", lambda x, k=key: outs[k])
    bd_ref = dict(points=points, st_features_2d=feat, gt_boxes=gt.clone(), batch_size=B)
    out_ref = ref(bd_ref)
    loss_ref, tb_ref = ref.get_loss()
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
    bd = dict(gt_boxes=gt.clone(), batch_size=B)
    out = mine.head_post(bd, nhwc(outs["cls"]), nhwc(torch.cat([outs["reg"], outs["height"], outs["dim"], outs["ang"]], 1)), nhwc(outs["dir"]),
                         points[:, 1:3])
    mask_ref = ref.get_anchor_mask(bd_ref, feat.shape)
    assert torch.equal(mine.get_anchor_mask(points[:, 1:3], (H, W)), mask_ref) and 0 < int(mask_ref.sum()) < H * W
    assert bool(mask_ref[-10:, -10:].any())                       # the negative-index wrap is exercised
    for key in ("cls_preds", "box_preds", "dir_cls_preds"):
        assert torch.equal(mine.forward_ret_dict[key], ref.forward_ret_dict[key]), key
    assert torch.equal(mine.forward_ret_dict["box_cls_labels"], ref.forward_ret_dict["box_cls_labels"])
    assert int((ref.forward_ret_dict["box_cls_labels"] > 0).sum()) > 20
    for key in ("box_reg_targets", "reg_weights", "gt_ious"):
        a, b = mine.forward_ret_dict[key], ref.forward_ret_dict[key]
        assert a.shape == b.shape and float((a - b).abs().max()) <= 1e-6, key
    assert torch.equal(out["gt_ious"], out_ref["gt_ious"])
    assert float((out["batch_cls_preds"] - out_ref["batch_cls_preds"]).abs().max()) == 0.0
    assert float((out["batch_box_preds"] - out_ref["batch_box_preds"]).abs().max()) <= 1e-5
    loss, tb = mine.get_loss()
    assert abs(float(loss) - float(loss_ref)) <= 1e-5 * max(1.0, abs(float(loss_ref)))
    for key in ("rpn_loss_cls", "rpn_loss_loc", "rpn_loss_dir"):
        assert abs(float(tb[key]) - tb_ref[key]) <= 1e-5 * max(1.0, abs(tb_ref[key])), key
