"""Module with the entry points of the reference's pybind extension ``cpd.ops.pointnet2.pointnet2_stack.pointnet2_stack_cuda``
that the RoI grid pooling uses (src/pointnet2_api.cpp:14,20-21), same names, argument order and in-place outputs, backed by
libcpd_b200.so -- so the reference's own voxel_query_utils.py / pointnet2_utils.py / voxel_pool_modules.py run on the shim
unchanged.  The other entries of that extension (ball query, FPS, interpolation, vector pool) are outside the hot path."""
import torch

from . import ops


def voxel_query_wrapper(M, R1, R2, R3, nsample, radius, z_range, y_range, x_range, new_xyz, xyz, new_coords, point_indices, idx):
    """voxel_query.cpp:25-42: point_indices is the reference's dense (B, Z, Y, X) voxel -> row map; idx (M, nsample) int32 is
    filled in place; empty balls get idx[:, 0] = -1 like upstream (the caller zeroes them, voxel_query_utils.py:41-42)."""
    out, empty = ops.voxel_query(new_xyz, new_coords, xyz, [R1, R2, R3], point_indices.shape[0], [z_range, y_range, x_range], radius, nsample,
                                 dense_map=point_indices)
    idx.copy_(out)
    idx[empty, 0] = -1
    return 1


def _global_rows(idx, idx_batch_cnt, features_batch_cnt):
    """per-batch local indices -> global rows (what the reference kernels do with their batch_cnt prefix sums)."""
    starts = torch.cumsum(features_batch_cnt, 0) - features_batch_cnt
    return (idx + torch.repeat_interleave(starts.to(idx.dtype), idx_batch_cnt.long()).view(-1, 1)).int().contiguous()


def group_points_wrapper(B, M, C, nsample, features, features_batch_cnt, idx, idx_batch_cnt, out):
    out.copy_(ops.group_points(features, _global_rows(idx, idx_batch_cnt, features_batch_cnt)))
    return 1


def group_points_grad_wrapper(B, M, C, N, nsample, grad_out, idx, idx_batch_cnt, features_batch_cnt, grad_features):
    grad_features.copy_(ops.group_points_bwd(grad_out, _global_rows(idx, idx_batch_cnt, features_batch_cnt), N))
    return 1
