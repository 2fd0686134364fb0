"""RoI grid pooling of VoxelRCNN[Proto]Head on libcpd_b200.so.

Host-side mirrors of
  * cpd/models/roi_heads/voxel_rcnn_head.py:186-273,365-386   (roi_grid_pool, get_global_grid_points_of_roi, get_dense_grid_points)
  * cpd/ops/pointnet2/pointnet2_stack/voxel_pool_modules.py:8-130   (NeighborVoxelSAModuleMSG: same sub-module / parameter names)
  * cpd/ops/pointnet2/pointnet2_stack/voxel_query_utils.py:12-110    (VoxelQuery, VoxelQueryAndGrouping)
  * cpd/utils/common_utils.py:35-57,66-82                            (rotate_points_along_z, get_voxel_centers)
The voxel -> row lookup goes through the coordinate hash of the level's SparseConvTensor (the one its rulebooks already
built) instead of the dense (B, Z, Y, X) map of cpd/utils/spconv_utils.py:4-21; query and grouping run in cpd_voxel_query /
cpd_group_points[_bwd] and use GLOBAL row indices, so the reference's per-batch rebasing loop (and its batch_cnt tensors)
disappears.  The small pointwise MLPs stay torch, as in the reference.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


def rotate_points_along_z(points, angle):
    """common_utils.py:35-57: points (B, N, 3+C), angle (B,)."""
    cosa, sina = torch.cos(angle), torch.sin(angle)
    zeros, ones = angle.new_zeros(points.shape[0]), angle.new_ones(points.shape[0])
    rot = torch.stack((cosa, sina, zeros, -sina, cosa, zeros, zeros, zeros, ones), dim=1).view(-1, 3, 3).float()
    return torch.cat((torch.matmul(points[:, :, 0:3], rot), points[:, :, 3:]), dim=-1)


def get_voxel_centers(voxel_coords, downsample_times, voxel_size, point_cloud_range):
    """common_utils.py:66-82: (N, 3) zyx cells -> xyz centres."""
    centers = voxel_coords[:, [2, 1, 0]].float()
    vs = torch.tensor(voxel_size, device=centers.device).float() * downsample_times
    return (centers + 0.5) * vs + torch.tensor(point_cloud_range[0:3], device=centers.device).float()


def get_dense_grid_points(rois, batch_size_rcnn, grid_size):
    """voxel_rcnn_head.py:377-386."""
    dense_idx = rois.new_ones((grid_size, grid_size, grid_size)).nonzero().repeat(batch_size_rcnn, 1, 1).float()
    size = rois.reshape(batch_size_rcnn, -1)[:, 3:6]
    return (dense_idx + 0.5) / grid_size * size.unsqueeze(1) - size.unsqueeze(1) / 2


def get_global_grid_points_of_roi(rois, grid_size):
    """voxel_rcnn_head.py:365-375."""
    rois = rois.reshape(-1, rois.shape[-1])
    local = get_dense_grid_points(rois, rois.shape[0], grid_size)
    glob = rotate_points_along_z(local.clone(), rois[:, 6]).squeeze(1) + rois[:, 0:3].unsqueeze(1)
    return glob, local


class _GroupPoints(torch.autograd.Function):
    """pointnet2_utils.GroupingOperation with global row indices: (n, c), (m, ns) -> (m, c, ns)."""

    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n = features.shape[0]
        return ops.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return ops.group_points_bwd(grad_out.contiguous(), idx, ctx.n), None


def group_points(features, idx):
    return _GroupPoints.apply(features, idx)


class VoxelQueryAndGrouping(nn.Module):
    """voxel_query_utils.py:54-110."""

    def __init__(self, max_range, radius, nsample):
        super().__init__()
        self.max_range, self.radius, self.nsample = max_range, radius, nsample

    def forward(self, new_coords, xyz, new_xyz, features, sp_tensor):
        """new_coords (m, 4) [b, z, y, x]; xyz (n, 3) voxel centres; features (n, c) -> (m, c, ns), (m, 3, ns), (m,) bool."""
        with torch.no_grad():
            idx, empty = ops.voxel_query(new_xyz, new_coords, xyz, sp_tensor.spatial_shape, sp_tensor.batch_size, self.max_range, self.radius,
                                         self.nsample, hash_buf=sp_tensor.coord_hash())
        return group_points(features, idx), group_points(xyz, idx), empty


class NeighborVoxelSAModuleMSG(nn.Module):
    """voxel_pool_modules.py:8-130 -- same constructor arguments, same parameter names."""

    def __init__(self, *, query_ranges, radii, nsamples, mlps, use_xyz=True, pool_method="max_pool"):
        super().__init__()
        assert len(query_ranges) == len(nsamples) == len(mlps)
        self.groupers, self.mlps_in, self.mlps_pos, self.mlps_out = nn.ModuleList(), nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for rng, ns, rad, spec in zip(query_ranges, nsamples, radii, mlps):
            self.groupers.append(VoxelQueryAndGrouping(rng, rad, ns))
            self.mlps_in.append(nn.Sequential(nn.Conv1d(spec[0], spec[1], kernel_size=1, bias=False), nn.BatchNorm1d(spec[1])))
            self.mlps_pos.append(nn.Sequential(nn.Conv2d(3, spec[1], kernel_size=1, bias=False), nn.BatchNorm2d(spec[1])))
            self.mlps_out.append(nn.Sequential(nn.Conv1d(spec[1], spec[2], kernel_size=1, bias=False), nn.BatchNorm1d(spec[2]), nn.ReLU()))
        self.relu = nn.ReLU()
        self.pool_method = pool_method
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Conv1d)):
                nn.init.kaiming_normal_(m.weight)
            if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                nn.init.constant_(m.weight, 1.0)
                nn.init.constant_(m.bias, 0)

    def forward(self, xyz, new_xyz, new_coords, features, sp_tensor):
        """xyz (n, 3) voxel centres, new_xyz (m, 3), new_coords (m, 4) [b, x, y, z] (the reference's order), features (n, c)
        -> (m, sum_k mlps[k][-1])."""
        new_coords = new_coords[:, [0, 3, 2, 1]].contiguous()                   # -> [b, z, y, x]
        outs = []
        for k, grouper in enumerate(self.groupers):
            f_in = self.mlps_in[k](features.permute(1, 0).unsqueeze(0)).permute(0, 2, 1).contiguous().view(features.shape[0], -1)
            gf, gx, empty = grouper(new_coords, xyz, new_xyz, f_in, sp_tensor)
            gf = gf.masked_fill(empty[:, None, None], 0.0)
            gx = (gx - new_xyz.unsqueeze(-1)).masked_fill(empty[:, None, None], 0.0)
            pos = self.mlps_pos[k](gx.permute(1, 0, 2).unsqueeze(0))            # (1, C, m, ns)
            nf = self.relu(gf.permute(1, 0, 2).unsqueeze(0) + pos)
            if self.pool_method == "max_pool":
                nf = F.max_pool2d(nf, kernel_size=[1, nf.size(3)]).squeeze(-1)
            elif self.pool_method == "avg_pool":
                nf = F.avg_pool2d(nf, kernel_size=[1, nf.size(3)]).squeeze(-1)
            else:
                raise NotImplementedError(self.pool_method)
            outs.append(self.mlps_out[k](nf).squeeze(0).permute(1, 0))
        return torch.cat(outs, dim=1)


DEFAULT_POOL_CFG = dict(
    FEATURES_SOURCE=["x_conv3", "x_conv4"], GRID_SIZE=6,
    POOL_LAYERS=dict(x_conv3=dict(MLPS=[[32, 32], [32, 32]], QUERY_RANGES=[[2, 2, 2], [4, 4, 4]], POOL_RADIUS=[0.4, 0.8], NSAMPLE=[16, 16],
                                  POOL_METHOD="max_pool"),
                     x_conv4=dict(MLPS=[[32, 32], [32, 32]], QUERY_RANGES=[[2, 2, 2], [4, 4, 4]], POOL_RADIUS=[0.8, 1.6], NSAMPLE=[16, 16],
                                  POOL_METHOD="max_pool")))


class RoIGridPool(nn.Module):
    """The pooling stage of VoxelRCNNProtoHead (voxel_rcnn_head.py:29-44,186-273) with the config of
    tools/cfgs/models/waymo_unsupervised/voxel_rcnn_cproto_center.yaml:108-124: per RoI a GRID_SIZE^3 lattice of query
    points, per feature source a NeighborVoxelSAModuleMSG over the backbone's multi-scale sparse tensors."""

    def __init__(self, input_channels, voxel_size, point_cloud_range, pool_cfg=None):
        super().__init__()
        cfg = pool_cfg or DEFAULT_POOL_CFG
        self.pool_cfg, self.voxel_size, self.point_cloud_range = cfg, [float(v) for v in voxel_size], [float(v) for v in point_cloud_range]
        self.roi_grid_pool_layers = nn.ModuleList()
        self.num_features = 0
        for src in cfg["FEATURES_SOURCE"]:
            lc = cfg["POOL_LAYERS"][src]
            mlps = [[input_channels[src]] + list(m) for m in lc["MLPS"]]
            self.roi_grid_pool_layers.append(NeighborVoxelSAModuleMSG(query_ranges=lc["QUERY_RANGES"], nsamples=lc["NSAMPLE"], radii=lc["POOL_RADIUS"],
                                                                      mlps=mlps, pool_method=lc["POOL_METHOD"]))
            self.num_features += sum(m[-1] for m in mlps)

    def forward(self, rois, multi_scale_3d_features, multi_scale_3d_strides):
        """rois (B, N, 7+) -> (B*N, GRID_SIZE^3, C)."""
        B = rois.shape[0]
        gs = self.pool_cfg["GRID_SIZE"]
        grid_xyz, _ = get_global_grid_points_of_roi(rois, gs)
        grid_xyz = grid_xyz.view(B, -1, 3)
        pr, vs = self.point_cloud_range, self.voxel_size
        gc = torch.cat([(grid_xyz[:, :, 0:1] - pr[0]) // vs[0], (grid_xyz[:, :, 1:2] - pr[1]) // vs[1], (grid_xyz[:, :, 2:3] - pr[2]) // vs[2]], dim=-1)
        bidx = torch.arange(B, device=rois.device, dtype=gc.dtype).view(B, 1, 1).expand(B, gc.shape[1], 1)
        pooled = []
        for k, src in enumerate(self.pool_cfg["FEATURES_SOURCE"]):
            t = multi_scale_3d_features[src]
            stride = multi_scale_3d_strides[src]
            xyz = get_voxel_centers(t.indices[:, 1:4], stride, vs, pr)
            coords = torch.cat([bidx, gc // stride], dim=-1).int()                 # [b, x, y, z], as the reference builds it
            f = self.roi_grid_pool_layers[k](xyz.contiguous(), grid_xyz.contiguous().view(-1, 3), coords.contiguous().view(-1, 4),
                                             t.features.contiguous(), t)
            pooled.append(f.view(-1, gs ** 3, f.shape[-1]))
        return torch.cat(pooled, dim=-1)
