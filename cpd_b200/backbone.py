"""Sparse 3-D backbones of CPD rebuilt on cpd_b200.sparse.

Host-side mirror of cpd/models/backbones_3d/spconv_backbone.py: ``VoxelBackBone8x``
(:138-395) and ``VoxelResBackBone8x`` (:398-600) with the same constructor signature,
config keys (NUM_FILTERS, OUT_FEATURES, RETURN_NUM_FEATURES_AS_DICT, MM, last_pad), module
names (=> identical state_dict keys, so CPD checkpoints load), indice_keys and batch_dict
contract.  The reference classes themselves also run unchanged on the
``cpd_b200.compat`` spconv shim; these mirrors exist so that the hot path can be built and
benchmarked without the reference tree, and so that the residual blocks can use the fused
conv+BN+residual+ReLU epilogue in eval mode.
"""
from functools import partial

import torch
import torch.nn as nn

from . import sparse as sp
from .sparse import fold_bn


class _Cfg(dict):
    """dict with attribute access (stand-in for the reference's EasyDict model_cfg)."""
    __getattr__ = dict.get


def _cfg(c):
    return c if hasattr(c, "get") and not isinstance(c, dict) else _Cfg(c or {})


def conv_bn_relu(cin, cout, ksize, norm_fn, indice_key=None, stride=1, padding=0, conv_type="subm"):
    """post_act_block (spconv_backbone.py:13-35): conv(bias=False) + BN + ReLU."""
    if conv_type == "subm":
        conv = sp.SubMConv3d(cin, cout, ksize, bias=False, indice_key=indice_key)
    elif conv_type == "spconv":
        conv = sp.SparseConv3d(cin, cout, ksize, stride=stride, padding=padding, bias=False, indice_key=indice_key)
    elif conv_type == "inverseconv":
        conv = sp.SparseInverseConv3d(cin, cout, ksize, indice_key=indice_key, bias=False)
    else:
        raise NotImplementedError(conv_type)
    return sp.SparseSequential(conv, norm_fn(cout), nn.ReLU())


class SparseBasicBlock(sp.SparseModule):
    """Residual block of spconv_backbone.py:100-136: two SubM 3x3x3 convs (bias=True), BN after
    each, ReLU after the first and after the residual add."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, norm_fn=None, downsample=None, indice_key=None):
        super().__init__()
        assert norm_fn is not None
        self.conv1 = sp.SubMConv3d(inplanes, planes, 3, stride=stride, padding=1, bias=True, indice_key=indice_key)
        self.bn1 = norm_fn(planes)
        self.relu = nn.ReLU()
        self.conv2 = sp.SubMConv3d(planes, planes, 3, stride=stride, padding=1, bias=True, indice_key=indice_key)
        self.bn2 = norm_fn(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        if not self.training and not torch.is_grad_enabled() and self.downsample is None:
            s1, h1 = fold_bn(self.bn1)
            s2, h2 = fold_bn(self.bn2)
            out = self.conv1.forward_fused(x, s1, h1, relu=True)
            return self.conv2.forward_fused(out, s2, h2, relu=True, residual=x.features)
        if self.downsample is None and sp.bn_fusable(self.bn1) and sp.bn_fusable(self.bn2):
            out = self.conv1.forward_bn_train(x, self.bn1, relu=True)
            return self.conv2.forward_bn_train(out, self.bn2, relu=True, residual=x.features)
        identity = x
        out = self.conv1(x)
        out = out.replace_feature(self.relu(self.bn1(out.features)))
        out = self.conv2(out)
        out = out.replace_feature(self.bn2(out.features))
        if self.downsample is not None:
            identity = self.downsample(x)
        return out.replace_feature(self.relu(out.features + identity.features))


class _Backbone8xBase(nn.Module):
    RES = False

    def __init__(self, model_cfg, input_channels, grid_size, num_frames=1, **kwargs):
        super().__init__()
        cfg = _cfg(model_cfg)
        self.model_cfg = cfg
        self.num_frames = num_frames
        self.return_num_features_as_dict = cfg.get("RETURN_NUM_FEATURES_AS_DICT", False)
        self.out_features = cfg.get("OUT_FEATURES", 128)
        nf = cfg.get("NUM_FILTERS", [16, 32, 64, 128])
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        gs = [int(g) for g in grid_size]
        self.sparse_shape = [gs[2] + 1, gs[1], gs[0]]          # grid_size[::-1] + [1, 0, 0]
        self._build_tower("", input_channels, nf, norm_fn, full=True)
        last_pad = cfg.get("last_pad", 0)
        self.conv_out = sp.SparseSequential(
            sp.SparseConv3d(nf[3], self.out_features, (3, 1, 1), stride=(2, 1, 1), padding=last_pad, bias=False,
                            indice_key="spconv_down2"),
            norm_fn(self.out_features), nn.ReLU())
        if cfg.get("MM", False):
            self._build_tower("_2", input_channels, nf, norm_fn, full=not self.RES)
        self.num_point_features = self.out_features
        if self.return_num_features_as_dict:
            self.num_point_features = {"x_conv1": nf[0], "x_conv2": nf[1], "x_conv3": nf[2], "x_conv4": nf[3]}

    def _stage_blocks(self, c, norm_fn, key, n):
        if self.RES:
            return [SparseBasicBlock(c, c, norm_fn=norm_fn, indice_key=key) for _ in range(n)]
        return [conv_bn_relu(c, c, 3, norm_fn, indice_key=key, padding=1) for _ in range(n)]

    def _build_tower(self, sfx, cin, nf, norm_fn, full):
        k = (lambda name: name + sfx)
        sub = "res" if self.RES else "subm"
        # Res tower indexes its stage-1 convs 'subm1' (input) and 'res1' (blocks); plain tower shares 'subm1'
        setattr(self, "conv_input" + sfx, sp.SparseSequential(
            sp.SubMConv3d(cin, nf[0], 3, padding=1, bias=False, indice_key=k("subm1")), norm_fn(nf[0]), nn.ReLU()))
        n1 = 2 if self.RES else 1
        setattr(self, "conv1" + sfx, sp.SparseSequential(*self._stage_blocks(nf[0], norm_fn, k(sub + "1"), n1)))
        pads = {2: 1, 3: 1, 4: (0, 1, 1)}
        for s in (2, 3, 4):
            n = 2 if full else 1
            down = conv_bn_relu(nf[s - 2], nf[s - 1], 3, norm_fn, stride=2, padding=pads[s], indice_key=k(f"spconv{s}"),
                                conv_type="spconv")
            setattr(self, f"conv{s}" + sfx, sp.SparseSequential(down, *self._stage_blocks(nf[s - 1], norm_fn, k(f"{sub}{s}"), n)))

    SORT_INPUT = True   # process the tower in ascending linear-key order (see _run_tower)

    def _sorted_order(self, coords):
        # The voxelizer emits rows in first-appearance (i.e. shuffled) order.  A sparse tensor is a set, so the
        # tower may visit it in any order: ascending ((b*D+z)*H+y)*W+x -- the order every strided conv already
        # produces -- makes each 128-row tile spatially compact, so gathers hit L2/L1 and empty taps are skipped.
        d, h, w = self.sparse_shape
        c64 = coords.long()
        key = ((c64[:, 0] * d + c64[:, 1]) * h + c64[:, 2]) * w + c64[:, 3]
        perm = torch.argsort(key)
        return perm, coords.index_select(0, perm).contiguous()

    def plan_tower(self, sfx, coords, batch_size, with_out):
        """Everything of a tower that depends on the voxel COORDINATES only: the visiting order and the whole rulebook
        chain (spconv's indice_dict), built by walking the tower's convolutions without features.  All data-dependent
        sizes -- hence all host syncs of the sparse path -- live here, so a plan can be made one step ahead on a side
        stream (CPDHotPathDetector.prepare) and the forward pass itself never waits for the device."""
        coords = coords.int() if coords.dtype != torch.int32 else coords
        perm = None
        if self.SORT_INPUT and coords.shape[0] > 0:
            perm, coords = self._sorted_order(coords)
        x = sp.SparseConvTensor(None, coords, self.sparse_shape, batch_size)
        names = ["conv_input", "conv1", "conv2", "conv3", "conv4"]
        seqs = [getattr(self, n + sfx) for n in names] + ([self.conv_out] if with_out else [])
        for seq in seqs:
            for m in seq.modules():                   # registration order == execution order in these towers
                if isinstance(m, sp.SparseConvolution):
                    rb, out_hash = m.get_rulebook(x)
                    if rb.nbr_fwd.shape[1] > 1:
                        rb.nbr_fwd_t                  # transposed table for the weight-gradient kernel
                    cin = m.in_channels + (-m.in_channels) % 8
                    if rb.sorted_table("fwd", cin) is None:
                        rb.masks_fwd                  # per-tile tap masks (k-block skipping) of the unsorted table
                    if self.training and torch.is_grad_enabled():
                        if rb.kind == "strided":
                            rb.bwd_sorted(m.out_channels)                 # input-gradient table in tap-pattern order
                        elif cin == m.in_channels:                        # (the 5-channel input needs no gradient)
                            rb.sorted_table("fwd", m.out_channels)
                    x = m._wrap_output(x, rb, out_hash, None)
        return dict(perm=perm, coords=coords, indice_dict=x.indice_dict)

    def _run_tower(self, sfx, feats, coords, batch_size, with_out, plan=None):
        coords = coords.int() if coords.dtype != torch.int32 else coords
        indice_dict = None
        if plan is not None:
            coords, indice_dict = plan["coords"], plan["indice_dict"]
            if plan["perm"] is not None:
                feats = feats.index_select(0, plan["perm"])
        elif self.SORT_INPUT and coords.shape[0] > 0:
            perm, coords = self._sorted_order(coords)
            feats = feats.index_select(0, perm)
        x = sp.SparseConvTensor(feats, coords, self.sparse_shape, batch_size, indice_dict=indice_dict)
        x = getattr(self, "conv_input" + sfx)(x)
        c1 = getattr(self, "conv1" + sfx)(x)
        c2 = getattr(self, "conv2" + sfx)(c1)
        c3 = getattr(self, "conv3" + sfx)(c2)
        c4 = getattr(self, "conv4" + sfx)(c3)
        out = self.conv_out(c4) if with_out else None
        return out, {"x_conv1": c1, "x_conv2": c2, "x_conv3": c3, "x_conv4": c4}

    _STRIDES = {"x_conv1": 1, "x_conv2": 2, "x_conv3": 4, "x_conv4": 8}


class VoxelResBackBone8x(_Backbone8xBase):
    """spconv_backbone.py:398-600: residual tower; in training with MM a second tower
    (``*_2`` modules, one block per stage from stage 2 on, no conv_out) runs on
    ``voxel_features1`` / ``voxel_coords1`` (:560-598)."""
    RES = True

    def forward(self, batch_dict):
        bs = batch_dict["batch_size"]
        out, ms = self._run_tower("", batch_dict["voxel_features"], batch_dict["voxel_coords"], bs, True,
                                  plan=batch_dict.get("tower_plan"))
        batch_dict.update({"encoded_spconv_tensor": out, "encoded_spconv_tensor_stride": 8,
                           "multi_scale_3d_features": ms, "multi_scale_3d_strides": dict(self._STRIDES)})
        if self.training and self.model_cfg.get("MM", False):
            _, ms2 = self._run_tower("_2", batch_dict["voxel_features1"], batch_dict["voxel_coords1"], bs, False,
                                     plan=batch_dict.get("tower_plan1"))
            batch_dict.update({"encoded_spconv_tensor_stride_mm": 8, "multi_scale_3d_features_mm": ms2,
                               "multi_scale_3d_strides": dict(self._STRIDES)})
        return batch_dict


class VoxelBackBone8x(_Backbone8xBase):
    """spconv_backbone.py:138-395: plain tower.  Training loops the stages separately
    (:285-331); eval concatenates the stages along X into one [D, H, 4*W] tensor and splits
    the results with ``decompose_tensor`` whose strict ``<`` bounds are kept (:241-260,332-393)."""
    RES = False

    def decompose_tensor(self, tensor, i, batch_size):
        w4 = tensor.spatial_shape[2] // 4
        lo, hi = i * w4, (i + 1) * w4
        x = tensor.indices[:, 3]
        mask = (lo < x) & (x < hi)
        coords = tensor.indices[mask].clone()
        coords[:, 3] -= lo
        return sp.SparseConvTensor(tensor.features[mask], coords.int().contiguous(),
                                   [tensor.spatial_shape[0], tensor.spatial_shape[1], w4], batch_size)

    def forward(self, batch_dict):
        stages = batch_dict["transform_param"].shape[1] if "transform_param" in batch_dict else 1
        bs = batch_dict["batch_size"]
        sid = lambda i: "" if i == 0 else str(i)
        if self.training:
            for i in range(stages):
                out, ms = self._run_tower("", batch_dict["voxel_features" + sid(i)], batch_dict["voxel_coords" + sid(i)], bs, True)
                batch_dict.update({"encoded_spconv_tensor" + sid(i): out, "encoded_spconv_tensor_stride" + sid(i): 8,
                                   "multi_scale_3d_features" + sid(i): ms,
                                   "multi_scale_3d_strides" + sid(i): dict(self._STRIDES)})
            return batch_dict
        feats, coords = [], []
        for i in range(stages):
            feats.append(batch_dict["voxel_features" + sid(i)])
            c = batch_dict["voxel_coords" + sid(i)].clone()
            c[:, 3] += i * self.sparse_shape[2]
            coords.append(c)
        wide = [self.sparse_shape[0], self.sparse_shape[1], self.sparse_shape[2] * 4]
        keep = self.sparse_shape
        self.sparse_shape = wide
        try:
            out, ms = self._run_tower("", torch.cat(feats, 0), torch.cat(coords, 0), bs, True)
        finally:
            self.sparse_shape = keep
        for i in range(stages):
            batch_dict.update({
                "encoded_spconv_tensor" + sid(i): self.decompose_tensor(out, i, bs),
                "encoded_spconv_tensor_stride" + sid(i): 8,
                "multi_scale_3d_features" + sid(i): {"x_conv1": None, "x_conv2": None,
                                                    "x_conv3": self.decompose_tensor(ms["x_conv3"], i, bs),
                                                    "x_conv4": self.decompose_tensor(ms["x_conv4"], i, bs)},
                "multi_scale_3d_strides" + sid(i): dict(self._STRIDES)})
        return batch_dict


def _rotate_z(points, angle):
    """common_utils.rotate_points_along_z for one cloud: points (N, 3), angle scalar tensor."""
    c, s_ = torch.cos(angle), torch.sin(angle)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    rot = torch.stack((c, s_, z, -s_, c, z, z, z, o)).view(3, 3).to(points.dtype)
    return points @ rot


def x_transform_points(points, param, backward=False):
    """X_TRAIN.forward_with_param / backward_with_param on points (cpd/datasets/augmentor/X_transform.py:53-160) with the
    default queue world_rotation -> world_flip (x axis: y := -y) -> world_scaling; param = (angle, flip, scale).
    backward runs the queue reversed with the inverse rotation / scale (the flip is its own inverse)."""
    pts = points.clone()
    if not backward:
        pts[:, 0:3] = _rotate_z(pts[:, 0:3], param[0])
        if bool(param[1] != 0):
            pts[:, 1] = -pts[:, 1]
        pts[:, 0:3] = pts[:, 0:3] * param[2]
    else:
        pts[:, 0:3] = pts[:, 0:3] / param[2]
        if bool(param[1] != 0):
            pts[:, 1] = -pts[:, 1]
        pts[:, 0:3] = _rotate_z(pts[:, 0:3], -param[0])
    return pts


def bilinear_sample_rows(rows, h, w, x, y):
    """height_compression.py:5-36 (bilinear_interpolate_torch) on an NHWC row matrix: rows (h*w, C) of ONE image, x / y (P,)
    pixel coordinates -> (P, C).  Corner indices are clamped like the reference's (weights are not)."""
    x0 = torch.floor(x).long()
    y0 = torch.floor(y).long()
    x1, y1 = x0 + 1, y0 + 1
    x0, x1 = x0.clamp(0, w - 1), x1.clamp(0, w - 1)
    y0, y1 = y0.clamp(0, h - 1), y1.clamp(0, h - 1)
    wa = (x1.type_as(x) - x) * (y1.type_as(y) - y)
    wb = (x1.type_as(x) - x) * (y - y0.type_as(y))
    wc = (x - x0.type_as(x)) * (y1.type_as(y) - y)
    wd = (x - x0.type_as(x)) * (y - y0.type_as(y))
    at = lambda yy, xx: rows.index_select(0, yy * w + xx)
    return at(y0, x0) * wa[:, None] + at(y1, x0) * wb[:, None] + at(y0, x1) * wc[:, None] + at(y1, x1) * wd[:, None]


class HeightCompression(nn.Module):
    """map_to_bev/height_compression.py:38-175: dense() + view, plus the test-time-augmentation alignment
    (ALIGN / ALIGN_METHOD in {first, max, mean, weighted_max}, bev_align :81-105).  ``nhwc=True`` writes the BEV map
    channels-last in one pass and returns it as a (B, C*D, H, W) tensor in torch.channels_last memory format (same
    logical values), which is also the layout bev_align samples from (a pixel is one contiguous row)."""

    def __init__(self, model_cfg=None, nhwc=True, num_frames=1, voxel_size=None, point_cloud_range=None, **kwargs):
        super().__init__()
        self.model_cfg = _cfg(model_cfg)
        self.num_bev_features = self.model_cfg.get("NUM_BEV_FEATURES", 256)
        self.nhwc = nhwc
        self.num_frames = num_frames
        self.voxel_size, self.point_cloud_range = voxel_size, point_cloud_range
        self._grids = {}

    def get_pseudo_points(self, pts_range, voxel_size, stride, device):
        """:48-66: centres of the stride-downsampled BEV pixels as (H, W, 3) points, z = 0 (float64 arange like numpy's)."""
        key = (tuple(float(v) for v in pts_range), tuple(float(v) for v in voxel_size), int(stride), str(device))
        if key not in self._grids:
            import numpy as np
            xs, ys = voxel_size[0] * stride, voxel_size[1] * stride
            x = np.arange(pts_range[0] + xs / 2, pts_range[3], xs)
            y = np.arange(pts_range[1] + ys / 2, pts_range[4] + ys / 2, ys)
            x, y = np.meshgrid(x, y)
            g = np.stack([x, y, np.zeros_like(x)]).astype(np.float32)
            self._grids[key] = torch.from_numpy(g).permute(1, 2, 0).contiguous().to(device)
        return self._grids[key]

    def bev_align(self, bev_feat, transform_param, stride, stage_i):
        """:81-105: resample stage i's BEV map onto stage 0's frame.  bev_feat (B, C, H, W) (any memory format);
        transform_param (B, stages, 3)."""
        n, c, h, w = bev_feat.shape
        rows = bev_feat.permute(0, 2, 3, 1).reshape(n, h * w, c)              # a view for channels_last maps
        pr, vs = self.point_cloud_range, self.voxel_size
        out = []
        for b in range(n):
            grid = self.get_pseudo_points(pr, vs, stride, bev_feat.device).reshape(-1, 3)
            pts = x_transform_points(grid, transform_param[b][stage_i])
            pts = x_transform_points(pts, transform_param[b][0], backward=True)
            x = (pts[:, 0] - pr[0]) / vs[0] / stride
            y = (pts[:, 1] - pr[1]) / vs[1] / stride
            out.append(bilinear_sample_rows(rows[b], h, w, x, y).reshape(h, w, c))
        return torch.stack(out).permute(0, 3, 1, 2)                           # logical NCHW over NHWC memory

    def forward(self, batch_dict):
        stages = batch_dict["transform_param"].shape[1] if "transform_param" in batch_dict else 1
        batch_dict["spatial_features_stride"] = batch_dict["encoded_spconv_tensor_stride"]
        align = "transform_param" in batch_dict and self.model_cfg.get("ALIGN", False)
        all_feat = []
        for i in range(stages):
            sid = "" if i == 0 else str(i)
            t = batch_dict["encoded_spconv_tensor" + sid]
            if self.nhwc:
                sf = t.dense_bev_nhwc().permute(0, 3, 1, 2)           # logical NCHW, channels_last strides
            else:
                d = t.dense()
                n, c, dd, h, w = d.shape
                sf = d.view(n, c * dd, h, w)
            batch_dict["spatial_features" + sid] = sf
            if i == 0:
                all_feat.append(sf)
            elif align:
                all_feat.append(self.bev_align(sf, batch_dict["transform_param"], batch_dict["spatial_features_stride"], i))
        if align:
            method = self.model_cfg.get("ALIGN_METHOD", "first")
            stack = torch.stack(all_feat)
            if method == "max":
                batch_dict["spatial_features"] = stack.max(0)[0]
            elif method == "mean":
                batch_dict["spatial_features"] = stack.mean(0)
            elif method == "weighted_max":
                w1, w2 = self.model_cfg.get("W1", 0.9), self.model_cfg.get("W2", 0.1)
                batch_dict["spatial_features"] = w1 * batch_dict["spatial_features"] + w2 * stack.max(0)[0]
        return batch_dict
