"""AnchorHeadSingleV2 on the dense tcgen05 convolutions.

Host-side mirror of
  * cpd/models/dense_heads/anchor_head_single.py:9-192          (get_layer, AnchorHeadSingleV2: same module / parameter names)
  * cpd/models/dense_heads/anchor_head_template.py:13-384       (anchors, losses, generate_predicted_boxes)
  * cpd/models/dense_heads/target_assigner/anchor_generator.py  (AnchorGenerator)
  * cpd/models/dense_heads/target_assigner/axis_aligned_target_assigner.py:48-243
  * cpd/utils/box_coder_utils.py:5-79 (ResidualCoder), cpd/utils/box_utils.py:238-287 (nearest-BEV IoU),
    cpd/utils/loss_utils.py:10-207 (focal / smooth-L1 / weighted CE), cpd/utils/common_utils.py:17-20 (limit_period)
used by tools/cfgs/models/waymo_unsupervised/voxel_rcnn_{oyster,dbscan}_single_train.yaml.

The convolutions run on cpd_b200.bev.DenseConv2d (TMA-fed tcgen05 kernel, fused BatchNorm); everything after them is torch,
as in the reference, but written without the reference's host round trips: the anchor mask is two small matrix products
instead of a numpy loop over occupied cells, and the target assigner works on masked IoU columns instead of slicing the
boxes of each class out (no `.cpu()`, no data-dependent shapes apart from the masked-anchor count itself).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import iou3d_nms_utils
from .backbone import _cfg
from .bev import DenseConv2d, DenseMap, DenseSequential


def default_cfg(match_height=False):
    """DENSE_HEAD of tools/cfgs/models/waymo_unsupervised/voxel_rcnn_oyster_single_train.yaml:36-92."""
    gen = [dict(class_name=n, anchor_sizes=[s], anchor_rotations=[0, 1.57], anchor_bottom_heights=[0], align_center=False,
                feature_map_stride=8, matched_threshold=0.55, unmatched_threshold=u)
           for n, s, u in (("Vehicle", [4.7, 2.1, 1.7], 0.5), ("Pedestrian", [0.91, 0.86, 1.73], 0.4), ("Cyclist", [1.78, 0.84, 1.78], 0.4))]
    return dict(NAME="AnchorHeadSingleV2", CLASS_AGNOSTIC=False, USE_DIRECTION_CLASSIFIER=True, DIR_OFFSET=0.78539, DIR_LIMIT_OFFSET=0.0,
                NUM_DIR_BINS=2, ANCHOR_GENERATOR_CONFIG=gen,
                TARGET_ASSIGNER_CONFIG=dict(NAME="AxisAlignedTargetAssigner", POS_FRACTION=-1.0, SAMPLE_SIZE=512, NORM_BY_NUM_EXAMPLES=False,
                                            MATCH_HEIGHT=match_height, BOX_CODER="ResidualCoder"),
                LOSS_CONFIG=dict(LOSS_WEIGHTS=dict(cls_weight=1.0, loc_weight=2.0, dir_weight=0.2, code_weights=[1.0] * 7)))


def limit_period(val, offset=0.5, period=np.pi):
    return val - torch.floor(val / period + offset) * period


class ResidualCoder:
    """box_coder_utils.py:5-79."""

    def __init__(self, code_size=7, encode_angle_by_sincos=False, **kwargs):
        self.code_size = code_size + (1 if encode_angle_by_sincos else 0)
        self.encode_angle_by_sincos = encode_angle_by_sincos

    def encode_torch(self, boxes, anchors):
        anchors = torch.cat([anchors[..., :3], anchors[..., 3:6].clamp_min(1e-5), anchors[..., 6:]], -1)
        boxes = torch.cat([boxes[..., :3], boxes[..., 3:6].clamp_min(1e-5), boxes[..., 6:]], -1)
        xa, ya, za, dxa, dya, dza, ra, *cas = torch.split(anchors, 1, dim=-1)
        xg, yg, zg, dxg, dyg, dzg, rg, *cgs = torch.split(boxes, 1, dim=-1)
        diagonal = torch.sqrt(dxa ** 2 + dya ** 2)
        rts = [torch.cos(rg) - torch.cos(ra), torch.sin(rg) - torch.sin(ra)] if self.encode_angle_by_sincos else [rg - ra]
        cts = [g - a for g, a in zip(cgs, cas)]
        return torch.cat([(xg - xa) / diagonal, (yg - ya) / diagonal, (zg - za) / dza, torch.log(dxg / dxa), torch.log(dyg / dya),
                          torch.log(dzg / dza), *rts, *cts], dim=-1)

    def decode_torch(self, enc, anchors):
        xa, ya, za, dxa, dya, dza, ra, *cas = torch.split(anchors, 1, dim=-1)
        if self.encode_angle_by_sincos:
            xt, yt, zt, dxt, dyt, dzt, cost, sint, *cts = torch.split(enc, 1, dim=-1)
        else:
            xt, yt, zt, dxt, dyt, dzt, rt, *cts = torch.split(enc, 1, dim=-1)
        diagonal = torch.sqrt(dxa ** 2 + dya ** 2)
        rg = torch.atan2(sint + torch.sin(ra), cost + torch.cos(ra)) if self.encode_angle_by_sincos else rt + ra
        cgs = [t + a for t, a in zip(cts, cas)]
        return torch.cat([xt * diagonal + xa, yt * diagonal + ya, zt * dza + za, torch.exp(dxt) * dxa, torch.exp(dyt) * dya,
                          torch.exp(dzt) * dza, rg, *cgs], dim=-1)


def generate_anchors(anchor_generator_cfg, grid_size, point_cloud_range, anchor_ndim=7, device=None):
    """anchor_generator.py:17-61 + anchor_head_template.py:46-61 -> ([ (z, y, x, num_size, num_rot, ndim) ], [anchors per location])."""
    rng = [float(v) for v in point_cloud_range]
    out, per_loc = [], []
    for cfg in anchor_generator_cfg:
        fm = [int(g) // int(cfg["feature_map_stride"]) for g in grid_size[:2]]
        sizes, rots, heights = cfg["anchor_sizes"], cfg["anchor_rotations"], cfg["anchor_bottom_heights"]
        per_loc.append(len(rots) * len(sizes) * len(heights))
        if cfg.get("align_center", False):
            xs, ys = (rng[3] - rng[0]) / fm[0], (rng[4] - rng[1]) / fm[1]
            xo, yo = xs / 2, ys / 2
        else:
            xs, ys = (rng[3] - rng[0]) / (fm[0] - 1), (rng[4] - rng[1]) / (fm[1] - 1)
            xo, yo = 0, 0
        x = torch.arange(rng[0] + xo, rng[3] + 1e-5, step=xs, dtype=torch.float32)
        y = torch.arange(rng[1] + yo, rng[4] + 1e-5, step=ys, dtype=torch.float32)
        z = x.new_tensor(heights)
        size, rot = x.new_tensor(sizes), x.new_tensor(rots)
        gx, gy, gz = torch.meshgrid([x, y, z], indexing="ij")
        a = torch.stack((gx, gy, gz), dim=-1)[:, :, :, None, :].repeat(1, 1, 1, size.shape[0], 1)
        a = torch.cat((a, size.view(1, 1, 1, -1, 3).repeat([*a.shape[0:3], 1, 1])), dim=-1)
        a = a[:, :, :, :, None, :].repeat(1, 1, 1, 1, rot.shape[0], 1)
        a = torch.cat((a, rot.view(1, 1, 1, 1, -1, 1).repeat([*a.shape[0:3], size.shape[0], 1, 1])), dim=-1)
        a = a.permute(2, 1, 0, 3, 4, 5).contiguous()                                  # [z, y, x, num_size, num_rot, 7]
        a[..., 2] += a[..., 5] / 2                                                     # bottom height -> box centre
        if anchor_ndim != 7:
            a = torch.cat((a, a.new_zeros([*a.shape[:-1], anchor_ndim - 7])), dim=-1)
        out.append(a.to(device) if device is not None else a)
    return out, per_loc


def boxes3d_nearest_bev_iou(boxes_a, boxes_b):
    """box_utils.py:238-287: IoU of the axis-aligned BEV boxes nearest to the rotated ones."""
    def aligned(b):
        rot = limit_period(b[:, 6], offset=0.5, period=np.pi).abs()
        dims = torch.where(rot[:, None] < np.pi / 4, b[:, [3, 4]], b[:, [4, 3]])
        return torch.cat((b[:, 0:2] - dims / 2, b[:, 0:2] + dims / 2), dim=1)
    a, b = aligned(boxes_a), aligned(boxes_b)
    x_len = torch.clamp_min(torch.min(a[:, 2, None], b[None, :, 2]) - torch.max(a[:, 0, None], b[None, :, 0]), min=0)
    y_len = torch.clamp_min(torch.min(a[:, 3, None], b[None, :, 3]) - torch.max(a[:, 1, None], b[None, :, 1]), min=0)
    area_a, area_b = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]), (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    inter = x_len * y_len
    return inter / torch.clamp_min(area_a[:, None] + area_b[None, :] - inter, min=1e-6)


class AxisAlignedTargetAssigner:
    """axis_aligned_target_assigner.py:9-243 (single-head layout; POS_FRACTION sampling kept with the reference's RNG calls)."""

    def __init__(self, model_cfg, class_names, box_coder, match_height=False):
        gen, tgt = model_cfg["ANCHOR_GENERATOR_CONFIG"], _cfg(model_cfg["TARGET_ASSIGNER_CONFIG"])
        self.box_coder, self.match_height = box_coder, match_height
        self.class_names = list(class_names)
        self.anchor_class_names = [c["class_name"] for c in gen]
        self.pos_fraction = tgt["POS_FRACTION"] if tgt["POS_FRACTION"] >= 0 else None
        self.sample_size = tgt["SAMPLE_SIZE"]
        self.norm_by_num_examples = tgt["NORM_BY_NUM_EXAMPLES"]
        self.matched = {c["class_name"]: c["matched_threshold"] for c in gen}
        self.unmatched = {c["class_name"]: c["unmatched_threshold"] for c in gen}

    def assign_targets(self, all_anchors, gt_boxes_with_classes):
        """all_anchors: per anchor class (z, n_loc, num_size, num_rot, 7) (already masked); gt (B, M, 7 + 1) -> dict of
        box_cls_labels (B, A) int32, box_reg_targets (B, A, code), reg_weights (B, A), gt_ious (B, A); A ordered
        location-major, then anchor class, then size / rotation -- the reference's `torch.cat(..., dim=-1)` of per-class maps."""
        B = gt_boxes_with_classes.shape[0]
        gt_classes = gt_boxes_with_classes[:, :, -1]
        gt_boxes = gt_boxes_with_classes[:, :, :-1]
        name_id = {n: i + 1 for i, n in enumerate(self.class_names)}
        out = dict(box_cls_labels=[], box_reg_targets=[], reg_weights=[], gt_ious=[])
        for k in range(B):
            cur_gt, cur_cls = gt_boxes[k], gt_classes[k].int()
            # the reference drops the trailing all-zero rows (:67-71); rows in between stay.  Zero boxes have IoU 0 with every
            # anchor and can never match, so masking them out instead of slicing gives identical targets without a host sync
            nonzero = cur_gt.sum(1) != 0
            per_class = []
            for cname, anchors in zip(self.anchor_class_names, all_anchors):
                fm = anchors.shape[:2]
                col_ok = nonzero & (cur_cls == name_id[cname])
                t = self.assign_targets_single(anchors.reshape(-1, anchors.shape[-1]), cur_gt, cur_cls, col_ok,
                                               self.matched[cname], self.unmatched[cname])
                per_class.append({key: (v.view(*fm, -1, self.box_coder.code_size) if key == "box_reg_targets" else v.view(*fm, -1))
                                  for key, v in t.items()})
            out["box_reg_targets"].append(torch.cat([t["box_reg_targets"] for t in per_class], dim=-2).view(-1, self.box_coder.code_size))
            for key in ("box_cls_labels", "gt_ious", "reg_weights"):
                out[key].append(torch.cat([t[key] for t in per_class], dim=-1).view(-1))
        return {key: torch.stack(v, dim=0) for key, v in out.items()}

    def assign_targets_single(self, anchors, gt_boxes, gt_classes, col_ok, matched_threshold=0.6, unmatched_threshold=0.45):
        """:159-243 with the class' boxes given as a column mask over all M rows of the frame."""
        A, G, dev = anchors.shape[0], gt_boxes.shape[0], anchors.device
        labels = torch.full((A,), -1, dtype=torch.int32, device=dev)
        if G == 0 or A == 0:
            labels[:] = 0
            return dict(box_cls_labels=labels, box_reg_targets=anchors.new_zeros((A, self.box_coder.code_size)),
                        reg_weights=anchors.new_zeros((A,)), gt_ious=anchors.new_zeros((A,)))
        ov = iou3d_nms_utils.boxes_iou3d_gpu(anchors[:, 0:7].contiguous(), gt_boxes[:, 0:7].contiguous()) if self.match_height \
            else boxes3d_nearest_bev_iou(anchors[:, 0:7], gt_boxes[:, 0:7])
        ov = torch.where(col_ok[None, :], ov, ov.new_full((), -1.0))
        a_max = ov.max(dim=1)[0]
        col = torch.arange(G, device=dev)
        a_arg = torch.where(ov == a_max[:, None], col[None, :], G).min(dim=1)[0].clamp(max=G - 1)   # first maximum, like numpy's argmax
        any_gt = col_ok.any()
        g_max = ov.max(dim=0)[0]
        g_max = torch.where(g_max == 0, g_max.new_full((), -1.0), g_max)                            # gts no anchor overlaps never force a match
        force = ((ov == g_max[None, :]) & col_ok[None, :] & (g_max[None, :] > 0)).any(dim=1)
        cls_of = gt_classes[a_arg].to(torch.int32)
        labels = torch.where(force, cls_of, labels)
        pos = a_max >= matched_threshold
        labels = torch.where(pos, cls_of, labels)
        bg = a_max < unmatched_threshold
        if self.pos_fraction is not None:                      # (never set in the shipped configs; keeps the reference's RNG calls)
            fg_inds = (labels > 0).nonzero()[:, 0]
            num_fg = int(self.pos_fraction * self.sample_size)
            if len(fg_inds) > num_fg:
                disable = torch.randperm(len(fg_inds))[:len(fg_inds) - num_fg]
                labels[disable.to(dev)] = -1                   # (sic: the reference indexes `labels` with positions in fg_inds)
            bg_inds = bg.nonzero()[:, 0]
            num_bg = self.sample_size - int((labels > 0).sum())
            if len(bg_inds) > num_bg:
                labels[bg_inds[torch.randint(0, len(bg_inds), size=(num_bg,)).to(dev)]] = 0
        else:
            labels = torch.where(bg, torch.zeros_like(labels), labels)
            labels = torch.where(force, cls_of, labels)
        labels = torch.where(any_gt, labels, torch.zeros_like(labels))                              # no box of this class: all background
        fg = labels > 0
        enc = self.box_coder.encode_torch(gt_boxes[a_arg], anchors)
        bbox_targets = torch.where(fg[:, None], enc, torch.zeros_like(enc))
        reg_weights = fg.to(anchors.dtype)
        if self.norm_by_num_examples:
            reg_weights = reg_weights / (labels >= 0).sum().clamp(min=1).to(anchors.dtype)
        ious = torch.where(any_gt, a_max.clamp(min=0), torch.zeros_like(a_max))
        return dict(box_cls_labels=labels, box_reg_targets=bbox_targets, reg_weights=reg_weights, gt_ious=ious)


def sigmoid_focal_loss(logits, target, weights, alpha=0.25, gamma=2.0):
    """loss_utils.py:10-73."""
    p = torch.sigmoid(logits)
    alpha_w = target * alpha + (1 - target) * (1 - alpha)
    pt = target * (1.0 - p) + (1.0 - target) * p
    bce = torch.clamp(logits, min=0) - logits * target + torch.log1p(torch.exp(-torch.abs(logits)))
    return alpha_w * torch.pow(pt, gamma) * bce * weights.unsqueeze(-1)


def weighted_smooth_l1(pred, target, weights, code_weights, beta=1.0 / 9.0):
    """loss_utils.py:76-135."""
    target = torch.where(torch.isnan(target), pred, target)
    diff = (pred - target) * code_weights.view(1, 1, -1)
    n = torch.abs(diff)
    loss = n if beta < 1e-5 else torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta)
    return loss * weights.unsqueeze(-1)


def get_layer(dim, out_dim, init=None):
    """anchor_head_single.py:9-29: conv3x3(bias) + BN + ReLU + conv1x1(bias), same Sequential indices (0, 1, 2, 3)."""
    conv = DenseConv2d(dim, dim, 3, padding=1, bias=True)
    nn.init.normal_(conv.weight, mean=0, std=0.001)
    conv2 = DenseConv2d(dim, out_dim, 1, bias=True)
    if init is None:
        nn.init.normal_(conv2.weight, mean=0, std=0.001)
    else:
        conv2.bias.data.fill_(init)
    return DenseSequential(conv, nn.BatchNorm2d(dim), nn.ReLU(), conv2)


class AnchorHeadSingleV2(nn.Module):
    """anchor_head_single.py:31-192 on AnchorHeadTemplate (anchor_head_template.py:13-384)."""

    def __init__(self, model_cfg, num_frames=1, input_channels=512, num_class=3, class_names=("Vehicle", "Pedestrian", "Cyclist"),
                 grid_size=(1504, 1504, 40), point_cloud_range=(-75.2, -75.2, -2, 75.2, 75.2, 4), predict_boxes_when_training=True, **kwargs):
        super().__init__()
        self.model_cfg = _cfg(model_cfg)
        self.num_frames, self.num_class, self.class_names = num_frames, num_class, list(class_names)
        self.predict_boxes_when_training = predict_boxes_when_training
        tcfg = _cfg(self.model_cfg["TARGET_ASSIGNER_CONFIG"])
        assert tcfg.get("NAME", "AxisAlignedTargetAssigner") == "AxisAlignedTargetAssigner" and not self.model_cfg.get("USE_MULTIHEAD", False)
        self.box_coder = ResidualCoder(**(tcfg.get("BOX_CODER_CONFIG") or {}))
        self.grid_size, self.range = [int(g) for g in grid_size], [float(v) for v in point_cloud_range]
        self.voxel_size = (self.range[3] - self.range[0]) / self.grid_size[0]
        anchors, per_loc = generate_anchors(self.model_cfg["ANCHOR_GENERATOR_CONFIG"], self.grid_size, self.range, self.box_coder.code_size)
        self.anchors_root = anchors                      # moved to the features' device on first use
        self.num_anchors_per_location = sum(per_loc)
        self.target_assigner = AxisAlignedTargetAssigner(self.model_cfg, self.class_names, self.box_coder, match_height=tcfg["MATCH_HEIGHT"])
        shard_c = 64
        self.shared_conv = DenseSequential(DenseConv2d(input_channels, shard_c, 3, padding=1, bias=True), nn.BatchNorm2d(shard_c), nn.ReLU())
        n = self.num_anchors_per_location
        self.conv_cls = get_layer(shard_c, n * self.num_class, -4.59)
        self.conv_reg = get_layer(shard_c, n * 2)
        self.conv_height = get_layer(shard_c, n * 1)
        self.conv_dim = get_layer(shard_c, n * 3)
        self.conv_ang = get_layer(shard_c, n * 1)
        self.conv_dir_cls = None
        if self.model_cfg.get("USE_DIRECTION_CLASSIFIER", None) is not None:
            self.conv_dir_cls = DenseConv2d(input_channels, n * self.model_cfg["NUM_DIR_BINS"], 1)
        lw = self.model_cfg["LOSS_CONFIG"]["LOSS_WEIGHTS"]
        self.register_buffer("code_weights", torch.tensor(lw["code_weights"], dtype=torch.float32), persistent=False)
        self.forward_ret_dict = {}

    # ---- anchors --------------------------------------------------------------------------
    def get_anchor_mask(self, points_xy, shape_hw):
        """:88-126: BEV cells within +-10 feature-map pixels of a 10x10-pixel block that holds a point.  points_xy (N, 2).
        The reference marks blocks on the host and loops over them in numpy; here block occupancy (hb, wb) is expanded with
        two 0/1 membership matrices: pixel y belongs to blocks {y // 10, y // 10 + 1}, and -- Python's negative indexing in
        the reference's `mask[inds] = 1` -- the last 10 rows / columns also belong to block 0."""
        H, W = int(shape_hw[0]), int(shape_hw[1])
        hb, wb = H // 10, W // 10
        dev = points_xy.device
        stride = float(np.round(self.voxel_size * 8.0 * 10.0))
        in_x = ((points_xy[:, 0] - self.range[0]) / stride).long().clamp(max=wb - 1)
        in_y = ((points_xy[:, 1] - self.range[1]) / stride).long().clamp(max=hb - 1)
        large = torch.zeros(hb * wb, dtype=torch.float32, device=dev)
        large.index_fill_(0, torch.remainder(in_y, hb) * wb + torch.remainder(in_x, wb), 1.0)     # (negative indices wrap, as in the reference)

        def membership(n, nb):
            p = torch.arange(n, device=dev)
            m = torch.zeros(n, nb, dtype=torch.float32, device=dev)
            q = p // 10
            m[p[q < nb], q[q < nb]] = 1.0
            m[p[q + 1 < nb], (q + 1)[q + 1 < nb]] = 1.0
            m[p >= n - 10, 0] = 1.0
            return m
        return (membership(H, hb) @ large.view(hb, wb) @ membership(W, wb).t()) > 0

    # ---- forward --------------------------------------------------------------------------
    def forward(self, data_dict):
        x = data_dict.get("st_features_2d_map")
        if x is None:
            x = DenseMap.from_nchw(data_dict["st_features_2d"])
        shard = self.shared_conv(x)
        rows = [self.conv_cls(shard).data, torch.cat([self.conv_reg(shard).data, self.conv_height(shard).data, self.conv_dim(shard).data,
                                                       self.conv_ang(shard).data], dim=1)]
        rows.append(self.conv_dir_cls(x).data if self.conv_dir_cls is not None else None)          # NHWC rows == permute(0, 2, 3, 1)
        pts = data_dict["points"]
        if isinstance(pts, (list, tuple)):
            pts_xy = torch.cat([p[:, 0:2] for p in pts], 0)
        else:
            pts_xy = pts[:, 1:3]                                                                    # [batch, x, y, z, ...] as in the reference
        return self.head_post(data_dict, *[None if r is None else r.view(x.n, x.h, x.w, -1) for r in rows], pts_xy)

    def head_post(self, data_dict, cls_map, box_map, dir_map, points_xy):
        """Everything after the convolutions (:128-188): anchor mask, masked predictions, targets, decoded boxes.
        cls_map / box_map / dir_map: (N, H, W, C) maps."""
        H, W = cls_map.shape[1], cls_map.shape[2]
        mask = self.get_anchor_mask(points_xy.to(cls_map.device), (H, W))
        self.anchors = [a.to(cls_map.device)[:, mask, ...] for a in self.anchors_root]
        cls_preds, box_preds = cls_map[:, mask, :], box_map[:, mask, :]
        dir_cls_preds = dir_map[:, mask, :] if dir_map is not None else None
        self.forward_ret_dict.update(cls_preds=cls_preds, box_preds=box_preds)
        if dir_cls_preds is not None:
            self.forward_ret_dict["dir_cls_preds"] = dir_cls_preds
        if self.training:
            targets = self.target_assigner.assign_targets(self.anchors, data_dict["gt_boxes"])
            self.forward_ret_dict.update(targets)
            data_dict["gt_ious"] = targets["gt_ious"]
        if not self.training or self.predict_boxes_when_training:
            cls_b, box_b = self.generate_predicted_boxes(data_dict["batch_size"], cls_preds, box_preds, dir_cls_preds)
            data_dict.update(batch_cls_preds=cls_b, batch_box_preds=box_b, cls_preds_normalized=False)
        return data_dict

    def _flat_anchors(self, batch_size):
        anchors = torch.cat(self.anchors, dim=-3)
        return anchors.view(1, -1, anchors.shape[-1]).repeat(batch_size, 1, 1)

    def generate_predicted_boxes(self, batch_size, cls_preds, box_preds, dir_cls_preds=None):
        """anchor_head_template.py:330-381."""
        anchors = self._flat_anchors(batch_size)
        A = anchors.shape[1]
        cls_b = cls_preds.reshape(batch_size, A, -1).float()
        box_b = self.box_coder.decode_torch(box_preds.reshape(batch_size, A, -1), anchors)
        if dir_cls_preds is not None:
            off, lim = self.model_cfg["DIR_OFFSET"], self.model_cfg["DIR_LIMIT_OFFSET"]
            labels = torch.max(dir_cls_preds.reshape(batch_size, A, -1), dim=-1)[1]
            period = 2 * np.pi / self.model_cfg["NUM_DIR_BINS"]
            rot = limit_period(box_b[..., 6] - off, lim, period)
            box_b = torch.cat([box_b[..., :6], (rot + off + period * labels.to(box_b.dtype)).unsqueeze(-1), box_b[..., 7:]], dim=-1)
        return cls_b, box_b

    # ---- losses ---------------------------------------------------------------------------
    def get_loss(self):
        """anchor_head_template.py:171-327 (cls + box regression + direction); tb_dict holds tensors (no .item() syncs)."""
        lw = self.model_cfg["LOSS_CONFIG"]["LOSS_WEIGHTS"]
        fr = self.forward_ret_dict
        cls_preds, labels = fr["cls_preds"], fr["box_cls_labels"]
        B = int(cls_preds.shape[0])
        cared, positives, negatives = labels >= 0, labels > 0, labels == 0
        cls_weights = (negatives * 1.0 + 1.0 * positives).float()
        reg_weights = positives.float()
        if self.num_class == 1:
            labels = torch.where(positives, torch.ones_like(labels), labels)
        norm = torch.clamp(positives.sum(1, keepdim=True).float(), min=1.0)
        cls_weights, reg_weights = cls_weights / norm, reg_weights / norm
        cls_targets = (labels * cared.type_as(labels)).long()
        one_hot = F.one_hot(cls_targets, self.num_class + 1).to(cls_preds.dtype)[..., 1:]
        cls_loss = sigmoid_focal_loss(cls_preds.reshape(B, -1, self.num_class), one_hot, cls_weights).sum() / B * lw["cls_weight"]
        tb = {"rpn_loss_cls": cls_loss.detach()}
        # box regression with the sin-difference encoding of the heading
        anchors = self._flat_anchors(B)
        box_preds = fr["box_preds"].reshape(B, -1, fr["box_preds"].shape[-1] // self.num_anchors_per_location)
        tgt = fr["box_reg_targets"]
        p_sin = torch.cat([box_preds[..., :6], torch.sin(box_preds[..., 6:7]) * torch.cos(tgt[..., 6:7]), box_preds[..., 7:]], dim=-1)
        t_sin = torch.cat([tgt[..., :6], torch.cos(box_preds[..., 6:7]) * torch.sin(tgt[..., 6:7]), tgt[..., 7:]], dim=-1)
        loc_loss = weighted_smooth_l1(p_sin, t_sin, reg_weights, self.code_weights.to(p_sin.device)).sum() / B * lw["loc_weight"]
        box_loss = loc_loss
        tb["rpn_loss_loc"] = loc_loss.detach()
        if fr.get("dir_cls_preds") is not None:
            nb = self.model_cfg["NUM_DIR_BINS"]
            rot_gt = tgt[..., 6] + anchors[..., 6]
            offset_rot = limit_period(rot_gt - self.model_cfg["DIR_OFFSET"], 0, 2 * np.pi)
            dir_t = torch.clamp(torch.floor(offset_rot / (2 * np.pi / nb)).long(), min=0, max=nb - 1)
            logits = fr["dir_cls_preds"].reshape(B, -1, nb)
            w = positives.type_as(logits)
            w = w / torch.clamp(w.sum(-1, keepdim=True), min=1.0)
            dir_loss = (F.cross_entropy(logits.permute(0, 2, 1), dir_t, reduction="none") * w).sum() / B * lw["dir_weight"]
            box_loss = box_loss + dir_loss
            tb["rpn_loss_dir"] = dir_loss.detach()
        loss = cls_loss + box_loss
        tb["rpn_loss"] = loss.detach()
        return loss, tb
