"""Golden-vector checks that need no reference checkout (fixtures written by the REFERENCE's own Python,
tests/golden/make_golden_heads.py): AnchorHeadSingleV2's anchor mask, target assigner, decoded boxes and losses, and
HeightCompression.bev_align.  Pure torch on the CPU."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_anchor_head_against_reference_golden():
    from cpd_b200 import anchor_head
    G = np.load(os.path.join(GOLD, "anchor_head_ref.npz"))
    grid, rng = [int(v) for v in G["grid"]], [float(v) for v in G["range"]]
    head = anchor_head.AnchorHeadSingleV2(anchor_head.default_cfg(), 1, 64, 3, ["Vehicle", "Pedestrian", "Cyclist"], grid, rng).train()
    t = lambda k: torch.from_numpy(G[k].astype(np.float32))
    nhwc = lambda x: x.permute(0, 2, 3, 1).contiguous()
    pts, gt = torch.from_numpy(G["points"]), torch.from_numpy(G["gt"])
    bd = dict(gt_boxes=gt.clone(), batch_size=gt.shape[0])
    box = torch.cat([t("out_reg"), t("out_height"), t("out_dim"), t("out_ang")], 1)
    out = head.head_post(bd, nhwc(t("out_cls")), nhwc(box), nhwc(t("out_dir")), pts[:, 1:3])
    H, W = G["mask"].shape
    mask = head.get_anchor_mask(pts[:, 1:3], (H, W))
    assert np.array_equal(mask.numpy(), G["mask"]) and 0 < int(mask.sum()) < H * W
    fr = head.forward_ret_dict
    assert np.array_equal(fr["box_cls_labels"].numpy(), G["box_cls_labels"])
    for key in ("box_reg_targets", "reg_weights", "gt_ious"):
        assert fr[key].shape == G[key].shape and float(np.abs(fr[key].numpy() - G[key]).max()) <= 1e-6, key
    assert float(np.abs(out["batch_cls_preds"].numpy() - G["batch_cls_preds"]).max()) == 0.0
    assert float(np.abs(out["batch_box_preds"].numpy() - G["batch_box_preds"]).max()) <= 1e-5
    loss, tb = head.get_loss()
    for got, key in ((loss, "loss"), (tb["rpn_loss_cls"], "loss_cls"), (tb["rpn_loss_loc"], "loss_loc"), (tb["rpn_loss_dir"], "loss_dir")):
        assert abs(float(got) - float(G[key])) <= 1e-5 * max(1.0, abs(float(G[key]))), key


def test_bev_align_against_reference_golden():
    from cpd_b200 import backbone
    G = np.load(os.path.join(GOLD, "bev_align_ref.npz"))
    pcr, vs, stride = [float(v) for v in G["pcr"]], [float(v) for v in G["vs"]], int(G["stride"])
    feat = torch.from_numpy(G["feat"].astype(np.float32))
    tp = torch.from_numpy(G["transform_param"])
    for nhwc in (True, False):
        m = backbone.HeightCompression(dict(NUM_BEV_FEATURES=8), nhwc=nhwc, voxel_size=vs, point_cloud_range=pcr)
        x = feat.contiguous(memory_format=torch.channels_last) if nhwc else feat
        for j, stage in enumerate((1, 2)):
            got = m.bev_align(x, tp, stride, stage)
            assert got.shape == G["aligned"][j].shape
            assert float(np.abs(got.numpy() - G["aligned"][j]).max()) <= 3e-5 * max(1.0, float(np.abs(G["aligned"][j]).max()))   # (pixel coordinates up to 80 carry 4e-6 of fp32 rounding into the bilinear weights)
