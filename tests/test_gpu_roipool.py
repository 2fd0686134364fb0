"""GPU parity of the RoI grid pooling primitives: cpd_voxel_query / cpd_group_points[_bwd] against the
reference's OWN kernels -- cpd/ops/pointnet2/pointnet2_stack/src/voxel_query_gpu.cu and group_points_gpu.cu compiled
unmodified for sm_100a into oracle/_ref/libpointnet2_ref_gpu.so -- bit for bit, and the pooling modules
(NeighborVoxelSAModuleMSG, RoIGridPool: voxel_pool_modules.py:8-130, voxel_rcnn_head.py:186-273) against a plain torch
evaluation of the same formulas."""
import ctypes as C

import numpy as np
import pytest
import torch

from cpd_b200.synth import PC_RANGE, VOXEL_SIZE, synth_scan

pytestmark = pytest.mark.gpu


def vp(t):
    return C.c_void_p(t.data_ptr())


def _level(cuda, stride, seeds=(31, 32, 33)):
    """A sparse level like x_conv3 / x_conv4: distinct cells of `stride`-downsampled voxels of a few sweeps."""
    from cpd_b200 import sparse as sp, voxel
    frames = [torch.from_numpy(synth_scan(30000, s)).to(cuda) for s in seeds]
    bd = voxel.voxelize_batch(frames, PC_RANGE, VOXEL_SIZE)
    c = bd["voxel_coords"].clone()
    c[:, 1:] //= stride
    c = torch.unique(c, dim=0).int().contiguous()
    shape = [(41 + stride - 1) // stride, 1504 // stride, 1504 // stride]
    return sp.SparseConvTensor(torch.randn(c.shape[0], 64, device=cuda), c, shape, len(seeds))


def _queries(t, stride, cuda, n_rois=40, grid=6):
    from cpd_b200 import roipool
    g = torch.Generator().manual_seed(5)
    centers = roipool.get_voxel_centers(t.indices[:, 1:4], stride, list(VOXEL_SIZE), list(PC_RANGE))
    B = t.batch_size
    rois = torch.zeros(B, n_rois, 7, device=cuda)
    for b in range(B):
        rows = torch.nonzero(t.indices[:, 0] == b).view(-1)
        pick = rows[torch.randint(0, rows.numel(), (n_rois,), generator=g).to(cuda)]
        rois[b, :, 0:3] = centers[pick] + torch.randn(n_rois, 3, generator=g).to(cuda) * 0.3
        rois[b, :, 3:6] = torch.tensor([4.5, 2.0, 1.7], device=cuda) * (0.7 + 0.6 * torch.rand(n_rois, 3, generator=g).to(cuda))
        rois[b, :, 6] = (torch.rand(n_rois, generator=g).to(cuda) * 2 - 1) * 3.1
    rois[0, 0, 0:3] = torch.tensor([500.0, 500.0, 50.0], device=cuda)        # far outside: empty balls, out-of-grid cells
    xyz, _ = roipool.get_global_grid_points_of_roi(rois, grid)
    xyz = xyz.view(B, -1, 3)
    pr, vs = list(PC_RANGE), list(VOXEL_SIZE)
    gc = torch.cat([(xyz[..., 0:1] - pr[0]) // vs[0], (xyz[..., 1:2] - pr[1]) // vs[1], (xyz[..., 2:3] - pr[2]) // vs[2]], -1) // stride
    bidx = torch.arange(B, device=cuda, dtype=gc.dtype).view(B, 1, 1).expand(B, gc.shape[1], 1)
    coords_bzyx = torch.cat([bidx, gc], -1)[..., [0, 3, 2, 1]].int().contiguous().view(-1, 4)
    return rois, xyz.contiguous().view(-1, 3), coords_bzyx, centers.contiguous()


@pytest.mark.parametrize("stride,rng,radius,nsample", [(4, (2, 2, 2), 0.4, 16), (4, (4, 4, 4), 0.8, 16), (8, (2, 2, 2), 0.8, 16), (8, (4, 4, 4), 1.6, 7)])
def test_voxel_query_and_grouping_bit_exact_vs_reference_kernels(oracle, cuda, stride, rng, radius, nsample):
    from cpd_b200 import ops
    lib = oracle.ref_pointnet2_lib()
    if lib is None:
        pytest.skip("oracle/_ref/libpointnet2_ref_gpu.so not built")
    t = _level(cuda, stride)
    rois, new_xyz, new_coords, xyz = _queries(t, stride, cuda)
    M, N, B = new_xyz.shape[0], xyz.shape[0], t.batch_size
    Z, Y, X = t.spatial_shape
    # the reference's dense voxel -> row map (cpd/utils/spconv_utils.py:4-21)
    dense = torch.full((B, Z, Y, X), -1, dtype=torch.int32, device=cuda)
    ind = t.indices.long()
    dense[ind[:, 0], ind[:, 1], ind[:, 2], ind[:, 3]] = torch.arange(N, dtype=torch.int32, device=cuda)
    ref_idx = torch.zeros(M, nsample, dtype=torch.int32, device=cuda)                   # voxel_query_utils.py:33 zero-initialises
    assert lib.ref_voxel_query(M, Z, Y, X, nsample, C.c_float(radius), rng[0], rng[1], rng[2], vp(new_xyz), vp(xyz), vp(new_coords), vp(dense),
                               vp(ref_idx)) == 0
    ref_empty = ref_idx[:, 0] == -1
    ref_idx[ref_empty] = 0                                                              # voxel_query_utils.py:41-42
    idx_h, empty_h = ops.voxel_query(new_xyz, new_coords, xyz, t.spatial_shape, B, rng, radius, nsample, hash_buf=t.coord_hash())
    idx_d, empty_d = ops.voxel_query(new_xyz, new_coords, xyz, t.spatial_shape, B, rng, radius, nsample, dense_map=dense)
    assert torch.equal(idx_h, ref_idx) and torch.equal(empty_h, ref_empty)              # hash lookup == dense map == reference kernel
    assert torch.equal(idx_d, ref_idx) and torch.equal(empty_d, ref_empty)
    assert 0 < int(ref_empty.sum()) < M and int((ref_idx != ref_idx[:, :1]).any(1).sum()) > 0
    # grouping: the reference takes per-batch local indices + batch counts
    feats = t.features.contiguous()
    cnt_f = torch.bincount(t.indices[:, 0].long(), minlength=B).int()
    cnt_q = torch.full((B,), M // B, dtype=torch.int32, device=cuda)
    starts = (torch.cumsum(cnt_f, 0) - cnt_f).int()
    local = (ref_idx - torch.repeat_interleave(starts, cnt_q.long()).view(-1, 1)).contiguous()
    local[ref_empty] = 0                                                                # what voxel_query_utils.py:85-91 leaves behind
    glob = (local + torch.repeat_interleave(starts, cnt_q.long()).view(-1, 1)).int().contiguous()
    c = feats.shape[1]
    ref_out = torch.empty(M, c, nsample, device=cuda)
    assert lib.ref_group_points(B, M, c, nsample, vp(feats), vp(cnt_f), vp(local), vp(cnt_q), vp(ref_out)) == 0
    out = ops.group_points(feats, glob)
    assert torch.equal(out, ref_out)
    gout = torch.randn(M, c, nsample, device=cuda)
    ref_g = torch.zeros(N, c, device=cuda)
    assert lib.ref_group_points_grad(B, M, c, N, nsample, vp(gout), vp(local), vp(cnt_q), vp(cnt_f), vp(ref_g)) == 0
    g = ops.group_points_bwd(gout, glob, N)
    assert float((g - ref_g).abs().max()) <= 1e-5 * max(1.0, float(ref_g.abs().max()))  # fp32 atomics: order differs
    # the pybind-surface shim the reference's own Python calls
    from cpd_b200 import pointnet2_stack_cuda as shim
    idx2 = torch.zeros(M, nsample, dtype=torch.int32, device=cuda)
    shim.voxel_query_wrapper(M, Z, Y, X, nsample, radius, rng[0], rng[1], rng[2], new_xyz, xyz, new_coords, dense, idx2)
    e2 = idx2[:, 0] == -1
    idx2[e2] = 0
    assert torch.equal(idx2, ref_idx) and torch.equal(e2, ref_empty)
    out2 = torch.empty(M, c, nsample, device=cuda)
    shim.group_points_wrapper(B, M, c, nsample, feats, cnt_f, local, cnt_q, out2)
    assert torch.equal(out2, ref_out)


def _torch_sa_module(mod, xyz, new_xyz, coords_bzyx, feats, t):
    """NeighborVoxelSAModuleMSG.forward (voxel_pool_modules.py:73-130) in plain torch indexing, float64."""
    from cpd_b200 import ops
    outs = []
    for k, grouper in enumerate(mod.groupers):
        idx, empty = ops.voxel_query(new_xyz, coords_bzyx, xyz, t.spatial_shape, t.batch_size, grouper.max_range, grouper.radius, grouper.nsample,
                                     hash_buf=t.coord_hash())
        idx = idx.long()
        f_in = mod.mlps_in[k](feats.permute(1, 0).unsqueeze(0)).squeeze(0).permute(1, 0)           # (n, c)
        gf = f_in[idx].permute(0, 2, 1)                                                            # (m, c, ns)
        gx = xyz.to(feats.dtype)[idx].permute(0, 2, 1) - new_xyz.to(feats.dtype).unsqueeze(-1)
        gf = gf.masked_fill(empty[:, None, None], 0.0)
        gx = gx.masked_fill(empty[:, None, None], 0.0)
        nf = torch.relu(gf.permute(1, 0, 2).unsqueeze(0) + mod.mlps_pos[k](gx.permute(1, 0, 2).unsqueeze(0)))
        nf = nf.max(dim=3)[0]
        outs.append(mod.mlps_out[k](nf).squeeze(0).permute(1, 0))
    return torch.cat(outs, 1)


def test_neighbor_voxel_sa_module_and_roi_grid_pool(cuda):
    from cpd_b200 import roipool
    torch.manual_seed(4)
    # the pointwise MLPs are torch Conv1d / Conv2d (as in the reference): cuDNN would run them in TF32 by default, which is
    # 1e-3, not the 1e-4 this comparison with the float64 evaluation holds the grouping kernels to
    torch.backends.cudnn.allow_tf32, tf32_was = False, torch.backends.cudnn.allow_tf32
    try:
        _sa_module_and_pool(cuda, roipool)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32_was


def _sa_module_and_pool(cuda, roipool):
    import copy
    t = _level(cuda, 4)
    rois, new_xyz, coords_bzyx, xyz = _queries(t, 4, cuda, n_rois=16)
    mod = roipool.NeighborVoxelSAModuleMSG(query_ranges=[[2, 2, 2], [4, 4, 4]], radii=[0.4, 0.8], nsamples=[16, 16],
                                           mlps=[[64, 32, 32], [64, 32, 32]]).to(cuda).eval()
    for m in mod.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.uniform_(-0.2, 0.2); m.running_var.uniform_(0.5, 1.5)
    feats = t.features.clone().requires_grad_(True)
    y = mod(xyz, new_xyz, coords_bzyx[:, [0, 3, 2, 1]].contiguous(), feats, t)            # the module takes [b, x, y, z] like the reference's
    ref_mod = copy.deepcopy(mod).double()
    fr = t.features.detach().double().requires_grad_(True)
    yr = _torch_sa_module(ref_mod, xyz, new_xyz, coords_bzyx, fr, t)
    assert y.shape == (new_xyz.shape[0], 64)
    assert float((y.double() - yr).abs().max()) <= 1e-4 * max(1.0, float(yr.abs().max()))
    dy = torch.randn_like(y)
    y.backward(dy)
    yr.backward(dy.double())
    assert float((feats.grad.double() - fr.grad).abs().max()) <= 1e-4 * max(1.0, float(fr.grad.abs().max()))
    for (n, p), (_, q) in zip(mod.named_parameters(), ref_mod.named_parameters()):
        assert float((p.grad.double() - q.grad).abs().max()) <= 1e-4 * max(1.0, float(q.grad.abs().max())), n
    # the whole pooling stage on two scales
    t4 = _level(cuda, 8)
    t4 = t4.replace_feature(torch.randn(t4.indices.shape[0], 128, device=cuda, requires_grad=True))
    t3 = t.replace_feature(feats.detach().clone().requires_grad_(True))
    pool = roipool.RoIGridPool(dict(x_conv3=64, x_conv4=128), list(VOXEL_SIZE), list(PC_RANGE)).to(cuda).train()
    pooled = pool(rois, dict(x_conv3=t3, x_conv4=t4), dict(x_conv3=4, x_conv4=8))
    assert pooled.shape == (rois.shape[0] * rois.shape[1], 216, pool.num_features) and pool.num_features == 128
    pooled.square().mean().backward()
    assert t3.features.grad is not None and t4.features.grad is not None and float(t3.features.grad.abs().sum()) > 0
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in pool.parameters())


def test_detector_train_step_with_roi_grid_pool(cuda):
    """CPDHotPathDetector(roi_grid_pool=True): the CenterHead's proposals are pooled from x_conv3 / x_conv4 of both towers
    (voxel_rcnn_head.py:186-343) instead of the stand-in regulariser; every pooling and tower parameter receives a gradient."""
    import numpy as np

    from cpd_b200 import detector
    from cpd_b200.synth import synth_gt_boxes
    torch.manual_seed(0)
    det = detector.CPDHotPathDetector(roi_grid_pool=True, rois_per_image=16).to(cuda).train()
    bs = 2
    batch = dict(points=[torch.from_numpy(synth_scan(20000, 90 + i)).to(cuda) for i in range(bs)],
                 points1=[torch.from_numpy(synth_scan(20000, 590 + i)).to(cuda) for i in range(bs)],
                 gt_boxes=torch.from_numpy(np.stack([synth_gt_boxes(30, 90 + i) for i in range(bs)])).to(cuda))
    loss, tb = det(batch)
    assert torch.isfinite(loss) and "roi_pooled_abs_mean" in tb
    loss.backward()
    for name, p in det.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
    for pool in (det.roi_pool, det.roi_pool_mm):
        assert any(float(p.grad.abs().sum()) > 0 for p in pool.parameters())
    assert float(det.backbone_3d.conv4_2[1].conv2.weight.grad.abs().sum()) > 0      # the MM tower trains through the pooling stage
