"""The size-independent properties of tests/test_properties_cpu.py on the CUDA kernels, at the benchmark's sizes (160 k-point
sweeps, bs = 2-4): where the oracle would take minutes, the domain's own invariants are the check (task brief section 3)."""
import numpy as np
import pytest
import torch

from cpd_b200.synth import PC_RANGE, VOXEL_SIZE, synth_scan

pytestmark = pytest.mark.gpu
SHAPE = [41, 1504, 1504]


def _key(c, shape):
    c = c.long()
    return ((c[:, 0] * shape[0] + c[:, 1]) * shape[1] + c[:, 2]) * shape[2] + c[:, 3]


def test_voxel_set_permutation_invariant_at_full_size(cuda):
    from cpd_b200 import voxel
    frames = [torch.from_numpy(synth_scan(160000, 40 + i)).to(cuda) for i in range(2)]
    a = voxel.voxelize_batch(frames, PC_RANGE, VOXEL_SIZE, want_voxels=True)
    g = torch.Generator(device="cpu").manual_seed(1)
    shuffled = [f[torch.randperm(f.shape[0], generator=g).to(cuda)] for f in frames]
    b = voxel.voxelize_batch(shuffled, PC_RANGE, VOXEL_SIZE, want_voxels=True)
    ka, kb = _key(a["voxel_coords"], SHAPE), _key(b["voxel_coords"], SHAPE)
    oa, ob = torch.argsort(ka), torch.argsort(kb)
    assert torch.equal(ka[oa], kb[ob]) and ka.unique().numel() == ka.numel()          # same SET of voxels, no duplicates
    assert torch.equal(a["voxel_num_points"][oa], b["voxel_num_points"][ob])
    # per-voxel point count = min(points in the cell, 5), recomputed with torch
    pts = torch.cat([torch.cat([torch.full((f.shape[0], 1), i, device=cuda), f[:, :3]], 1) for i, f in enumerate(frames)])
    lo = torch.tensor(PC_RANGE[:3], device=cuda)
    vs = torch.tensor(VOXEL_SIZE, device=cuda)
    ijk = torch.floor((pts[:, 1:] - lo) / vs).long()
    grid = torch.tensor([1504, 1504, 40], device=cuda)
    ok = ((ijk >= 0) & (ijk < grid)).all(1)
    cell = torch.stack([pts[:, 0].long(), ijk[:, 2], ijk[:, 1], ijk[:, 0]], 1)[ok]
    uk, cnt = torch.unique(_key(cell, [40, 1504, 1504]), return_counts=True)
    ka40 = _key(a["voxel_coords"], [40, 1504, 1504])
    o40 = torch.argsort(ka40)
    assert torch.equal(ka40[o40], uk) and torch.equal(a["voxel_num_points"][o40].long(), cnt.clamp(max=5))


def _level(cuda, n_frames=2):
    from cpd_b200 import sparse as sp, voxel
    frames = [torch.from_numpy(synth_scan(160000, 60 + i)).to(cuda) for i in range(n_frames)]
    bd = voxel.voxelize_batch(frames, PC_RANGE, VOXEL_SIZE)
    c = bd["voxel_coords"]
    c = c[torch.argsort(_key(c, SHAPE))].contiguous()
    return sp.SparseConvTensor(None, c, SHAPE, n_frames)


def test_subm_table_is_mirror_symmetric_and_strided_outputs_sorted(cuda):
    from cpd_b200 import ops
    t = _level(cuda)
    m = t.indices.shape[0]
    nbr = ops.subm_table(t.indices, t.spatial_shape, t.batch_size, 3, t.coord_hash())
    rows = torch.arange(m, device=cuda, dtype=torch.int32)
    assert torch.equal(nbr[:, 13], rows)                                               # centre tap: the site itself
    for k in range(13):                                                                # nbr[o, k] = i  <=>  nbr[i, 26 - k] = o
        o = torch.nonzero(nbr[:, k] >= 0).view(-1)
        i = nbr[o, k].long()
        assert torch.equal(nbr[i, 26 - k].long(), o)
        assert int((nbr[:, k] >= 0).sum()) == int((nbr[:, 26 - k] >= 0).sum())
    # neighbours really are at the tap's offset
    k = 5
    o = torch.nonzero(nbr[:, k] >= 0).view(-1)
    d = t.indices[nbr[o, k].long()] - t.indices[o]
    assert torch.equal(d, torch.tensor([0, k // 9 - 1, (k // 3) % 3 - 1, k % 3 - 1], device=cuda, dtype=d.dtype).expand_as(d))
    # strided 3x3x3 / 2: outputs = exactly the cells reached, in ascending linear-key order
    oc, oshape = ops.strided_outputs(t.indices, t.spatial_shape, t.batch_size, 3, 2, 1)
    ko = _key(oc, oshape)
    assert bool((ko[1:] > ko[:-1]).all())
    c = t.indices.long()
    want = []
    for kz in range(3):
        for ky in range(3):
            for kx in range(3):
                n = torch.stack([c[:, 1] + 1 - kz, c[:, 2] + 1 - ky, c[:, 3] + 1 - kx], 1)
                ok = ((n >= 0) & (n % 2 == 0)).all(1) & (n[:, 0] // 2 < oshape[0]) & (n[:, 1] // 2 < oshape[1]) & (n[:, 2] // 2 < oshape[2])
                want.append(torch.cat([c[ok, :1], n[ok] // 2], 1))
    want = torch.unique(_key(torch.cat(want), oshape))
    assert torch.equal(ko, want)


@pytest.mark.parametrize("c", [16, 64])
def test_gather_gemm_linear_and_backward_is_its_adjoint_at_full_size(cuda, c):
    from cpd_b200 import ops
    t = _level(cuda)
    m = t.indices.shape[0]
    nbr = ops.subm_table(t.indices, t.spatial_shape, t.batch_size, 3, t.coord_hash())
    g = torch.Generator(device="cpu").manual_seed(c)
    w = (torch.randn(c, 27, c, generator=g) * (27 * c) ** -0.5).to(cuda)
    x1, x2, dy = (torch.randn(m, c, generator=g).to(cuda) for _ in range(3))
    y1, y2, y12 = ops.gather_gemm(x1, w, nbr), ops.gather_gemm(x2, w, nbr), ops.gather_gemm(x1 + 2 * x2, w, nbr)
    assert float((y12 - (y1 + 2 * y2)).abs().max()) <= 1e-4 * max(1.0, float(y12.abs().max()))
    wt = ops.weight_transpose(w, flip_taps=True)
    dx = ops.gather_gemm(dy, wt, nbr)                                                  # SubM input-gradient: same table, flipped taps
    dw, _ = ops.gather_wgrad(x1, dy, nbr.t().contiguous(), tap_major=True)
    lhs = float((dy.double() * y1.double()).sum())
    scale = float((dy.double().abs() * y1.double().abs()).sum())
    assert abs(lhs - float((dx.double() * x1.double()).sum())) <= 1e-5 * scale        # <dy, A x> == <A^T dy, x>
    assert abs(lhs - float((dw.double() * w.double()).sum())) <= 1e-5 * scale         # == <dW, W>


def test_dense_round_trip_at_full_size(cuda):
    from cpd_b200 import ops
    t = _level(cuda, 3)
    c = t.indices.clone()
    c[:, 1] //= 21
    c[:, 2:] //= 8
    c = torch.unique(c, dim=0).int().contiguous()
    feat = torch.randn(c.shape[0], 128, device=cuda)
    for cl in (False, True):
        d = ops.sparse_to_dense(feat, c, 3, [2, 188, 188], channels_last=cl)
        idx = c.long()
        if cl:                                                                         # (B, H, W, C*D) with channel index ch*D + z
            back = d.view(3, 188, 188, 128, 2)[idx[:, 0], idx[:, 2], idx[:, 3], :, idx[:, 1]]
        else:
            back = d[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]]
        assert torch.equal(back, feat) and int(torch.count_nonzero(d)) == int(torch.count_nonzero(feat))
