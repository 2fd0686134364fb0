"""Golden fixtures produced by the REFERENCE's own Python (build container only: needs /root/reference), so that the checks in
tests/test_anchor_head_cpu.py / tests/test_bev_align_cpu.py also run where the reference checkout is absent:

  anchor_head_ref.npz  cpd/models/dense_heads/anchor_head_single.py::AnchorHeadSingleV2 (imported unmodified through
                       cpd_b200.compat.reference) on a 40 x 60 map: inputs (points, gt boxes, conv outputs) and its anchor mask,
                       target-assigner outputs, decoded boxes and losses.
  bev_align_ref.npz    cpd/models/backbones_2d/map_to_bev/height_compression.py::HeightCompression.bev_align for three stages.

    python tests/golden/make_golden_heads.py
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CPD_REFERENCE_ROOT", "/root/reference")

GRID = [480, 320, 40]                                  # -> 60 x 40 feature map at stride 8
RANGE = [0.0, -16.0, -2.0, 48.0, 16.0, 4.0]
NAMES = ["Vehicle", "Pedestrian", "Cyclist"]


def main():
    from cpd_b200 import anchor_head
    from cpd_b200.compat import reference
    from cpd_b200.compat.reference import EasyDict
    reference.install_reference(REF)
    sys.modules.setdefault("cv2", types.ModuleType("cv2"))            # imported at module level, never used
    torch.Tensor.cuda = lambda self, *a, **k: self                     # the reference calls .cuda() on its anchors / code weights
    mod = importlib.import_module("cpd.models.dense_heads.anchor_head_single")
    torch.manual_seed(1)
    g = np.random.default_rng(1)
    cfg = anchor_head.default_cfg()
    ref = mod.AnchorHeadSingleV2(EasyDict(cfg), 1, 64, 3, NAMES, np.array(GRID), np.array(RANGE), predict_boxes_when_training=True).train()
    B, H, W = 2, GRID[1] // 8, GRID[0] // 8
    n = sum(ref.num_anchors_per_location) if isinstance(ref.num_anchors_per_location, (list, tuple)) else ref.num_anchors_per_location
    pts = np.concatenate([np.concatenate([np.full((3000, 1), b, np.float32), g.uniform(0, 20, (3000, 1)).astype(np.float32),
                                          g.uniform(-15.9, -0.5, (3000, 1)).astype(np.float32), g.uniform(-2, 2, (3000, 2)).astype(np.float32)], 1)
                          for b in range(B)], 0)                      # (the far third of the map holds no points: its anchors are masked out)
    pts[:50, 1:3] = [0.3, -15.7]                                       # first block row / column: the negative-index wrap of the mask
    gt = np.zeros((B, 24, 8), np.float32)
    for b in range(B):
        for k in range(18):
            cls = 1 + k % 3
            size = [[4.7, 2.1, 1.7], [0.91, 0.86, 1.73], [1.78, 0.84, 1.78]][cls - 1]
            src = pts[pts[:, 0] == b][g.integers(0, 3000)]
            gt[b, k] = [src[1], src[2], g.uniform(-1, 1), *(np.array(size) * g.uniform(0.85, 1.15, 3)), g.uniform(-3.1, 3.1), cls]
    gt[0, 4] = 0                                                       # a padding row in the middle
    h16 = lambda t: t.half().float()                                   # inputs are fp16-representable: the fixture stores them as fp16
    outs = dict(cls=h16(torch.randn(B, n * 3, H, W) * 2 - 3), reg=h16(torch.randn(B, n * 2, H, W) * 0.3), height=h16(torch.randn(B, n, H, W) * 0.3),
                dim=h16(torch.randn(B, n * 3, H, W) * 0.2), ang=h16(torch.randn(B, n, H, W) * 0.5), dir=h16(torch.randn(B, n * 2, H, W)))
    for name, key in (("conv_cls", "cls"), ("conv_reg", "reg"), ("conv_height", "height"), ("conv_dim", "dim"), ("conv_ang", "ang"), ("conv_dir_cls", "dir")):
        getattr(ref, name).forward = (lambda x, k=key: outs[k])
    bd = dict(points=torch.from_numpy(pts), st_features_2d=torch.randn(B, 64, H, W), gt_boxes=torch.from_numpy(gt).clone(), batch_size=B)
    out = ref(bd)
    loss, tb = ref.get_loss()
    mask = ref.get_anchor_mask(bd, bd["st_features_2d"].shape)
    fr = ref.forward_ret_dict
    assert int((fr["box_cls_labels"] > 0).sum()) > 10 and bool(mask[-10:, -10:].any()) and int(mask.sum()) < H * W
    np.savez_compressed(os.path.join(HERE, "anchor_head_ref.npz"), grid=np.array(GRID), range=np.array(RANGE, np.float32), points=pts, gt=gt,
                        **{"out_" + k: v.numpy().astype(np.float16) for k, v in outs.items()}, mask=mask.numpy(), box_cls_labels=fr["box_cls_labels"].numpy(),
                        box_reg_targets=fr["box_reg_targets"].numpy(), reg_weights=fr["reg_weights"].numpy(), gt_ious=fr["gt_ious"].numpy(),
                        batch_cls_preds=out["batch_cls_preds"].detach().numpy(), batch_box_preds=out["batch_box_preds"].detach().numpy(),
                        loss=np.float32(float(loss)), loss_cls=np.float32(tb["rpn_loss_cls"]), loss_loc=np.float32(tb["rpn_loss_loc"]),
                        loss_dir=np.float32(tb["rpn_loss_dir"]))
    print("anchor_head_ref.npz: positives", int((fr["box_cls_labels"] > 0).sum()), "masked locations", int(mask.sum()), "loss", float(loss))

    hc = importlib.import_module("cpd.models.backbones_2d.map_to_bev.height_compression")
    pcr, vs, stride = [0.0, -16.0, -3.0, 25.6, 16.0, 1.0], [0.05, 0.05, 0.1], 8
    m = hc.HeightCompression(EasyDict(dict(NUM_BEV_FEATURES=8)), 1, voxel_size=vs, point_cloud_range=pcr)
    h, w = int(round(32 / 0.05 / stride)), int(round(25.6 / 0.05 / stride))
    feat = (torch.randn(2, 8, h, w) * (torch.rand(2, 1, h, w) < 0.3)).half().float()
    tp = np.stack([np.stack([[g.uniform(-0.78, 0.78), float(s % 2 if b == 0 else (s + 1) % 2), g.uniform(0.95, 1.05)] for s in range(3)]) for b in range(2)])
    tp_t = torch.from_numpy(tp).float()
    aligned = np.stack([m.bev_align(feat.clone(), tp_t, stride, i).numpy() for i in (1, 2)])
    np.savez_compressed(os.path.join(HERE, "bev_align_ref.npz"), pcr=np.array(pcr, np.float32), vs=np.array(vs, np.float32), stride=stride,
                        feat=feat.numpy().astype(np.float16), transform_param=tp.astype(np.float32), aligned=aligned.astype(np.float32))
    print("bev_align_ref.npz:", aligned.shape, float(np.abs(aligned).max()))


if __name__ == "__main__":
    main()
