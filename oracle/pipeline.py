"""Oracle-side interpreter: runs a cpd_b200 backbone module tree with the CPU oracle
(eval-mode BatchNorm folded to its affine form).  TEST INFRASTRUCTURE ONLY -- used by
tests/ and by bench.py's cpu_baseline / `--impl reference` legs, never by the product."""
import numpy as np
import torch
import torch.nn as nn

from . import oracle as O


def _bn_affine(m):
    s = (m.weight / torch.sqrt(m.running_var + m.eps)).detach().cpu().numpy()
    return s, m.bias.detach().cpu().numpy() - m.running_mean.detach().cpu().numpy() * s


def _conv(m, feats, coords, shape, cache):
    key = (m.indice_key, len(coords), tuple(shape))
    if key not in cache:
        cache[key] = (O.rulebook_subm(coords, shape, m.kernel_size) if m.subm else
                      O.rulebook_strided(coords, shape, m.kernel_size, m.stride, m.padding))
    rb = cache[key]
    y = O.spconv_fwd(feats, m.weight.detach().cpu().numpy(), None if m.bias is None else m.bias.detach().cpu().numpy(), rb)
    return y, rb.out_coords, rb.out_shape


def run_sequential(seq, feats, coords, shape, cache):
    from cpd_b200 import sparse as sp
    from cpd_b200.backbone import SparseBasicBlock
    for m in seq._modules.values():
        if isinstance(m, sp.SparseSequential):
            feats, coords, shape = run_sequential(m, feats, coords, shape, cache)
        elif isinstance(m, SparseBasicBlock):
            ident = feats
            y, _, _ = _conv(m.conv1, feats, coords, shape, cache)
            s, b = _bn_affine(m.bn1)
            y = np.maximum(y * s + b, 0)
            y, _, _ = _conv(m.conv2, y, coords, shape, cache)
            s, b = _bn_affine(m.bn2)
            feats = np.maximum(y * s + b + ident, 0)
        elif isinstance(m, sp.SparseConvolution):
            feats, coords, shape = _conv(m, feats, coords, shape, cache)
        elif isinstance(m, nn.BatchNorm1d):
            s, b = _bn_affine(m)
            feats = feats * s + b
        elif isinstance(m, nn.ReLU):
            feats = np.maximum(feats, 0)
        else:
            raise TypeError(type(m))
    return feats, coords, shape


def decompose(feats, coords, shape, i=0):
    """VoxelBackBone8x.decompose_tensor (spconv_backbone.py:241-260): strict `<` on both slab
    bounds, so x == i*W/4 is dropped -- a reference quirk that is reproduced, not fixed."""
    w4 = shape[2] // 4
    keep = (i * w4 < coords[:, 3]) & (coords[:, 3] < (i + 1) * w4)
    c = coords[keep].copy()
    c[:, 3] -= i * w4
    return feats[keep], c, [shape[0], shape[1], w4]


def backbone_forward(net, frames, pc_range, voxel_size, max_pts=5, max_voxels=1000000, sfx="", eval_wide=False):
    """frames: list of (n_i, C) numpy clouds -> (features, coords (M,4), shape, stage dict).
    eval_wide: VoxelBackBone8x's eval path (spconv_backbone.py:332-393): the tower runs on a
    [D, H, 4*W] grid and the outputs are cut back with decompose()."""
    feats, coords = [], []
    for b, pts in enumerate(frames):
        v, c, n = O.voxelize(pts, pc_range, voxel_size, max_pts, max_voxels)
        feats.append(O.mean_vfe(v, n))
        coords.append(np.concatenate([np.full((len(c), 1), b, np.int32), c], 1))
    feats, coords = np.concatenate(feats, 0), np.concatenate(coords, 0)
    shape, cache, stages = list(net.sparse_shape), {}, {}
    if eval_wide:
        shape[2] *= 4
    names = ["conv_input", "conv1", "conv2", "conv3", "conv4"] + (["conv_out"] if sfx == "" else [])
    for name in names:
        feats, coords, shape = run_sequential(getattr(net, name + sfx if name != "conv_out" else name), feats, coords, shape, cache)
        stages[name] = (feats, coords, list(shape))
    if eval_wide:
        feats, coords, shape = decompose(feats, coords, shape, 0)
    return feats, coords, shape, stages


def bev_dense(feats, coords, batch, shape):
    d = O.dense(feats, coords, batch, shape)
    n, c, dd, h, w = d.shape
    return d.reshape(n, c * dd, h, w)
