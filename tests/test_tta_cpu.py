"""VoxelBackBone8x's multi-stage (TTA) eval path -- stages concatenated along X into one [D, H, 4 W] tensor, one tower pass,
strict-`<` decompose_tensor per stage (spconv_backbone.py:241-260, 332-393) -- for 2 and 3 stages, against the
oracle pipeline on the same concatenated input.  Runs on the CPU: the kernels behind cpd_b200.ops are the oracle-backed test
backend (tests/cpu_backend.py), what is exercised is the mirror's host logic; tests/test_gpu_parity2.py runs the same check on
the CUDA kernels at the full 1504 x 6016 grid."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

RANGE = [0.0, -3.2, -1.0, 6.4, 3.2, 1.0]
VS = [0.1, 0.1, 0.05]
GRID = [64, 64, 40]


def _cloud(seed, n=4000):
    g = np.random.default_rng(seed)
    return np.concatenate([g.uniform(0, 6.4, (n, 1)), g.uniform(-3.2, 3.2, (n, 1)), g.uniform(-1, 1, (n, 1)), g.uniform(0, 1, (n, 2))], 1).astype(np.float32)


def make_net(seed=0):
    from cpd_b200 import backbone
    torch.manual_seed(seed)
    net = backbone.VoxelBackBone8x(dict(NUM_FILTERS=[8, 8, 16, 16], OUT_FEATURES=16), 5, GRID)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.uniform_(-0.2, 0.2); m.running_var.uniform_(0.3, 1.0); m.weight.data.uniform_(0.7, 1.3); m.bias.data.uniform_(-0.2, 0.2)
    return net.eval()


def check_tta(net, batch_dict, stage_frames, pc_range, voxel_size, tol=1e-4):
    from oracle import pipeline
    want = pipeline.backbone_forward_tta(net, stage_frames, pc_range, voxel_size)
    for i, w in enumerate(want):
        sid = "" if i == 0 else str(i)
        got = {"out": batch_dict["encoded_spconv_tensor" + sid], "x_conv3": batch_dict["multi_scale_3d_features" + sid]["x_conv3"],
               "x_conv4": batch_dict["multi_scale_3d_features" + sid]["x_conv4"]}
        for key, (f, c, shape) in w.items():
            t = got[key]
            assert t.spatial_shape == shape, (i, key, t.spatial_shape, shape)
            assert np.array_equal(t.indices.cpu().numpy(), c), (i, key)
            assert len(c) > 0 and int((c[:, 3] == 0).sum()) == 0                       # strict `<`: column 0 of every slab is dropped
            assert float(np.abs(t.features.cpu().numpy() - f).max()) <= tol * max(1.0, float(np.abs(f).max())), (i, key)
        assert batch_dict["multi_scale_3d_features" + sid]["x_conv1"] is None and batch_dict["encoded_spconv_tensor_stride" + sid] == 8


@pytest.mark.parametrize("stages", [2, 3])
def test_multi_stage_eval_matches_oracle(oracle, stages):
    from cpu_backend import cpu_ops
    net = make_net()
    stage_frames = [[_cloud(10 * i + b) for b in range(2)] for i in range(stages)]
    bd = dict(batch_size=2, transform_param=torch.zeros(2, stages, 3))
    for i, frames in enumerate(stage_frames):
        sid = "" if i == 0 else str(i)
        feats, coords = [], []
        for b, pts in enumerate(frames):
            v, c, n = oracle.voxelize(pts, RANGE, VS)
            feats.append(torch.from_numpy(oracle.mean_vfe(v, n)))
            coords.append(torch.from_numpy(np.concatenate([np.full((len(c), 1), b, np.int32), c], 1)))
        bd["voxel_features" + sid], bd["voxel_coords" + sid] = torch.cat(feats), torch.cat(coords)
    with cpu_ops(), torch.no_grad():
        out = net(bd)
    check_tta(net, out, stage_frames, RANGE, VS)
