"""Voxelizer front-ends over cpd_voxelize.

* ``Point2VoxelCPU3d`` -- same constructor/method names as the spconv class the reference
  wraps (cpd/datasets/processor/data_processor.py:24,35-41,53-58); despite the inherited name
  the arithmetic runs on the GPU (host numpy in, host numpy out: H2D + kernels + D2H).
* ``voxelize_batch`` -- device-resident batched form used by the detector/bench path.
* ``MeanVFE`` -- the reference's VoxelFeatureExtractor (vfe/mean_vfe.py:6-61); when the batch
  came from ``voxelize_batch`` the mean is already there (fused into the scatter kernel).
"""
import numpy as np
import torch
import torch.nn as nn

from . import ops


class TvTensor:
    """Stand-in for cumm.tensorview.Tensor: wraps a numpy array, ``.numpy()`` returns a copy."""

    def __init__(self, arr):
        self._arr = arr

    def numpy(self):
        return np.array(self._arr, copy=True)

    def numpy_view(self):
        return self._arr

    @property
    def shape(self):
        return self._arr.shape


def tv_from_numpy(arr):
    return TvTensor(arr)


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("cpd_b200 voxelizer needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


class Point2VoxelCPU3d:
    """spconv.utils.Point2VoxelCPU3d of the shim.  Like upstream it is a HOST routine (cpd_voxelize_cpu: numpy in, numpy
    out, no CUDA): the reference calls it inside Dataset.__getitem__, i.e. in forked DataLoader worker processes
    (data_processor.py:133-144), where a CUDA context must not be created.  The hot path does not go through here:
    voxelize_batch / cpd_voxelize take the raw device points in the main process."""

    def __init__(self, vsize_xyz, coors_range_xyz, num_point_features, max_num_points_per_voxel, max_num_voxels):
        self.vsize = [float(v) for v in vsize_xyz]
        self.coors_range = [float(v) for v in coors_range_xyz]
        self.num_point_features = int(num_point_features)
        self.max_pts = int(max_num_points_per_voxel)
        self.max_voxels = int(max_num_voxels)
        self.grid_size = [int(round((self.coors_range[3 + i] - self.coors_range[i]) / self.vsize[i])) for i in range(3)]

    def point_to_voxel(self, pc):
        import ctypes as C

        from . import _lib
        arr = pc.numpy_view() if isinstance(pc, TvTensor) else np.asarray(pc)
        arr = np.ascontiguousarray(arr, dtype=np.float32)
        assert arr.ndim == 2 and arr.shape[1] == self.num_point_features
        n, c = arr.shape
        cap = max(1, min(n, self.max_voxels))
        voxels = np.empty((cap, self.max_pts, c), np.float32)
        coords = np.empty((cap, 3), np.int32)
        num = np.empty((cap,), np.int32)
        rng = np.asarray(self.coors_range, np.float32)
        vs = np.asarray(self.vsize, np.float32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        m = _lib.lib().cpd_voxelize_cpu(vp(arr), n, c, vp(rng), vp(vs), self.max_pts, self.max_voxels, vp(voxels), vp(coords), vp(num))
        if m < 0:
            _lib.check(int(m), "cpd_voxelize_cpu")
        return TvTensor(voxels[:m]), TvTensor(coords[:m]), TvTensor(num[:m])


class PointToVoxel:
    """spconv.pytorch.utils.PointToVoxel look-alike (imported, never called, by
    spconv_backbone.py:10): device tensors in, device tensors out."""

    def __init__(self, vsize_xyz, coors_range_xyz, num_point_features, max_num_voxels, max_num_points_per_voxel,
                 device=None):
        self.vsize, self.coors_range = list(vsize_xyz), list(coors_range_xyz)
        self.max_voxels, self.max_pts = int(max_num_voxels), int(max_num_points_per_voxel)

    def __call__(self, pc, clear_voxels=True, empty_mean=False):
        out = ops.voxelize(pc, [0, pc.shape[0]], self.coors_range, self.vsize, self.max_pts, self.max_voxels,
                           want_voxels=True, want_mean=False)
        return out["voxels"], out["coords"][:, 1:].contiguous(), out["num"]


def gather_features_by_pc_voxel_id(seg_res_features, pc_voxel_id, invalid_value=0):
    res = torch.full((pc_voxel_id.shape[0], seg_res_features.shape[1]), invalid_value, dtype=seg_res_features.dtype,
                     device=seg_res_features.device)
    valid = pc_voxel_id >= 0
    res[valid] = seg_res_features[pc_voxel_id[valid].long()]
    return res


def mask_and_shuffle_points(points, pc_range, generator=None, shuffle=True):
    """DataProcessor.mask_points_and_boxes_outside_range + shuffle_points (cpd/datasets/processor/data_processor.py:77-126,
    cpd/utils/common_utils.py:60-63) for one raw sweep ALREADY ON THE DEVICE: keep the points whose x, y lie inside the
    range (inclusive at both ends, z unchecked -- exactly the reference's mask), then apply a random permutation.
    No host sync: the mask is applied by a stable partition (rejected points are moved behind the kept ones and carry NaN
    coordinates, which the voxelizer drops like any out-of-range point), so the tensor keeps its static size.
    Returns (points (n, C), number of kept points as a 0-d device tensor)."""
    x, y = points[:, 0], points[:, 1]
    keep = (x >= pc_range[0]) & (x <= pc_range[3]) & (y >= pc_range[1]) & (y <= pc_range[4])
    n = points.shape[0]
    if shuffle:
        perm = torch.randperm(n, device=points.device, generator=generator)
        points, keep = points.index_select(0, perm), keep.index_select(0, perm)
    order = torch.sort((~keep).to(torch.uint8), stable=True)[1]          # kept points first, order preserved
    out = points.index_select(0, order)
    n_keep = keep.sum()
    tail = torch.arange(n, device=points.device) >= n_keep
    out = torch.where(tail[:, None], torch.full_like(out, float("nan")), out)
    return out, n_keep


def voxelize_batch(points_list, pc_range, voxel_size, max_pts=5, max_voxels=1000000, want_voxels=False):
    """list of (n_i, C) CUDA tensors (or one concatenated tensor + offsets) -> batch_dict entries
    ``voxels``, ``voxel_coords`` (float like load_data_to_gpu would make them is NOT done: int32
    [b,z,y,x]), ``voxel_num_points``, ``voxel_features`` (MeanVFE already applied)."""
    if isinstance(points_list, (list, tuple)):
        offs = [0]
        for p in points_list:
            offs.append(offs[-1] + p.shape[0])
        pts = torch.cat(list(points_list), 0) if len(points_list) > 1 else points_list[0]
    else:
        pts, offs = points_list
    out = ops.voxelize(pts, offs, pc_range, voxel_size, max_pts, max_voxels, want_voxels=want_voxels, want_mean=True)
    return dict(voxels=out["voxels"], voxel_coords=out["coords"], voxel_num_points=out["num"],
                voxel_features=out["mean"], voxel_counts=out["counts"])


class MeanVFE(nn.Module):
    """VoxelFeatureExtractor of the CPD configs (vfe/mean_vfe.py:6-61).  Same batch_dict keys;
    frames/stages loop like the reference.  If ``voxel_features<id>`` is already present
    (produced fused by ``voxelize_batch``) it is kept."""

    def __init__(self, model_cfg=None, num_point_features=5, num_frames=1, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg or {}
        self.num_point_features = num_point_features
        self.num_frames = num_frames
        self.model = self.model_cfg.get("MODEL", None) if hasattr(self.model_cfg, "get") else None

    def get_output_feature_dim(self):
        return self.num_point_features

    @staticmethod
    def _mean(voxels, num):
        s = voxels.sum(dim=1)
        return (s / torch.clamp_min(num.view(-1, 1), 1.0).type_as(voxels)).contiguous()

    def forward(self, batch_dict, **kwargs):
        for i in range(self.num_frames):
            fid = "" if i == 0 else str(i)
            for mm in ("", "_mm"):
                vk, nk, ok = "voxels" + mm + fid, "voxel_num_points" + mm + fid, "voxel_features" + mm + fid
                if mm and "mm" not in batch_dict:
                    continue
                if batch_dict.get(vk) is None:      # absent, or the fused voxelizer already produced voxel_features
                    continue
                feat = self._mean(batch_dict[vk], batch_dict[nk])
                if self.model == "max" and not mm:
                    feat[:, -1] = batch_dict[vk].max(dim=1)[0][:, -1]
                batch_dict[ok] = feat
        return batch_dict


VoxelFeatureExtractor = MeanVFE
