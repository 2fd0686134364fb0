"""CPU suite: pins the oracle (oracle/cpd_oracle.c) against
  * the reference's own iou3d_cpu.cpp (golden iou_ref.npz; live oracle/_ref when present),
  * torch.nn.functional.conv3d / autograd on densified inputs (independent pin of the
    sparse-conv restatement: the reference ships no spconv source or vectors),
  * a pure-Python restatement of the sequential voxelizer (small case), and the committed
    golden fixtures.
"""
import os

import numpy as np
import pytest
import torch

from cpd_b200.synth import PC_RANGE, VOXEL_SIZE, synth_nms_boxes, synth_scan

G = os.path.join(os.path.dirname(__file__), "golden")


def py_voxelize(pts, rng, vs, max_pts, max_voxels):
    """SURVEY.md Appendix A.1, literally, in Python (fp32 arithmetic via numpy scalars)."""
    grid = [int(np.round((np.float32(rng[3 + j]) - np.float32(rng[j])) / np.float32(vs[j]))) for j in range(3)]
    lut, vox, co, num = {}, [], [], []
    for p in pts:
        c = []
        for j in range(3):
            q = np.floor((np.float32(p[j]) - np.float32(rng[j])) / np.float32(vs[j]))
            if q < 0 or q >= grid[j]:
                c = None
                break
            c.append(int(q))
        if c is None:
            continue
        key = (c[2], c[1], c[0])
        if key not in lut:
            if len(vox) >= max_voxels:
                continue
            lut[key] = len(vox)
            vox.append(np.zeros((max_pts, pts.shape[1]), np.float32))
            co.append(key)
            num.append(0)
        v = lut[key]
        if num[v] < max_pts:
            vox[v][num[v]] = p
            num[v] += 1
    return np.array(vox, np.float32).reshape(-1, max_pts, pts.shape[1]), np.array(co, np.int32).reshape(-1, 3), np.array(num, np.int32)


def test_voxelizer_matches_python_restatement(oracle):
    pts = synth_scan(2500, 3)
    pts[10] = [75.2, 0, 0, 0.1, 0]
    pts[11] = [-75.2, -75.2, -2.0, 0.1, 0]
    for mv in (1000000, 300):
        v, c, n = oracle.voxelize(pts, PC_RANGE, VOXEL_SIZE, 5, mv)
        pv, pc, pn = py_voxelize(pts, PC_RANGE, VOXEL_SIZE, 5, mv)
        assert np.array_equal(c, pc) and np.array_equal(n, pn) and np.array_equal(v, pv)
    assert n.max() <= 5 and len(np.unique(c, axis=0)) == len(c)


def test_voxelizer_golden_and_edges(oracle):
    g = np.load(os.path.join(G, "voxel_small.npz"))
    v, c, n = oracle.voxelize(g["points"], PC_RANGE, VOXEL_SIZE, 5, 1000000)
    assert np.array_equal(v, g["voxels"]) and np.array_equal(c, g["coords"]) and np.array_equal(n, g["num"])
    # empty input / everything out of range
    v, c, n = oracle.voxelize(np.zeros((0, 5), np.float32), PC_RANGE, VOXEL_SIZE)
    assert v.shape == (0, 5, 5) and c.shape == (0, 3)
    far = np.full((7, 5), 500.0, np.float32)
    assert oracle.voxelize(far, PC_RANGE, VOXEL_SIZE)[0].shape[0] == 0
    # mean VFE = sum / clamp(num, 1)
    v, c, n = oracle.voxelize(g["points"], PC_RANGE, VOXEL_SIZE)
    m = oracle.mean_vfe(v, n)
    ref = torch.from_numpy(v).sum(1) / torch.clamp_min(torch.from_numpy(n).view(-1, 1).float(), 1.0)
    assert np.allclose(m, ref.numpy(), atol=1e-6)


def _rand_sparse(seed, batch, shape, m, c):
    rng = np.random.default_rng(seed)
    vol = shape[0] * shape[1] * shape[2]
    cells = rng.choice(batch * vol, m, replace=False)
    coords = np.stack([cells // vol, (cells // (shape[1] * shape[2])) % shape[0], (cells // shape[2]) % shape[1],
                       cells % shape[2]], 1).astype(np.int32)
    return coords, rng.normal(0, 1, (m, c)).astype(np.float32), rng


def _densify(coords, x, batch, shape):
    d = torch.zeros(batch, x.shape[1], *shape, dtype=torch.float64)
    d[coords[:, 0], :, coords[:, 1], coords[:, 2], coords[:, 3]] = torch.from_numpy(x).double()
    return d


@pytest.mark.parametrize("ksize,stride,pad", [(3, 1, 1), (3, 2, 1), ((3, 1, 1), (2, 1, 1), 0), (3, 2, (0, 1, 1))])
def test_sparse_conv_matches_dense_conv3d(oracle, ksize, stride, pad):
    batch, shape = 2, [9, 12, 10]
    coords, x, rng = _rand_sparse(1, batch, shape, 400, 6)
    ks = [ksize] * 3 if np.isscalar(ksize) else list(ksize)
    w = rng.normal(0, 0.3, (7, *ks, 6)).astype(np.float32)
    b = rng.normal(0, 0.1, 7).astype(np.float32)
    subm = stride == 1
    rb = oracle.rulebook_subm(coords, shape, ksize) if subm else oracle.rulebook_strided(coords, shape, ksize, stride, pad)
    y = oracle.spconv_fwd(x, w, b, rb)
    xd = _densify(coords, x, batch, shape).requires_grad_(True)
    wt = torch.from_numpy(w).double().permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
    bt = torch.from_numpy(b).double().requires_grad_(True)
    st = [stride] * 3 if np.isscalar(stride) else list(stride)
    pd = [pad] * 3 if np.isscalar(pad) else list(pad)
    yd = torch.nn.functional.conv3d(xd, wt, bt, stride=st, padding=pd)
    oc = rb.out_coords
    assert list(yd.shape[2:]) == list(rb.out_shape)
    got = yd[oc[:, 0], :, oc[:, 1], oc[:, 2], oc[:, 3]]
    assert np.abs(got.detach().numpy() - y).max() < 1e-4
    if not subm:
        # active output set == every output with a non-empty receptive field, ascending linear key
        occ = torch.nn.functional.conv3d((_densify(coords, np.ones((len(coords), 1), np.float32), batch, shape)),
                                         torch.ones(1, 1, *ks, dtype=torch.float64), stride=st, padding=pd)
        assert int((occ > 0).sum()) == rb.m_out
        key = ((oc[:, 0].astype(np.int64) * rb.out_shape[0] + oc[:, 1]) * rb.out_shape[1] + oc[:, 2]) * rb.out_shape[2] + oc[:, 3]
        assert np.all(np.diff(key) > 0)
    # backward (A.5) against autograd of the dense conv restricted to the active outputs
    dy = rng.normal(0, 1, y.shape).astype(np.float32)
    dx, dw, db = oracle.spconv_bwd(x, w, dy, rb)
    (got * torch.from_numpy(dy).double()).sum().backward()
    gx = xd.grad[coords[:, 0], :, coords[:, 1], coords[:, 2], coords[:, 3]].numpy()
    assert np.abs(gx - dx).max() < 1e-4
    assert np.abs(wt.grad.permute(0, 2, 3, 4, 1).numpy() - dw).max() < 1e-3
    assert np.abs(bt.grad.numpy() - db).max() < 1e-4


def test_sparse_conv_golden(oracle):
    g = np.load(os.path.join(G, "spconv_small.npz"))
    shape = [int(s) for s in g["shape"]]
    rb = oracle.rulebook_subm(g["coords"], shape, 3)
    assert np.allclose(oracle.spconv_fwd(g["x"], g["w"], g["bias"], rb), g["y_subm"], atol=1e-6)
    rs = oracle.rulebook_strided(g["coords"], shape, 3, 2, 1)
    assert np.array_equal(rs.out_coords, g["out_coords"]) and rs.out_shape == [int(s) for s in g["out_shape"]]
    assert np.allclose(oracle.spconv_fwd(g["x"], g["w"], None, rs), g["y_strided"], atol=1e-6)


def test_subm_rulebook_invariants(oracle):
    coords, x, _ = _rand_sparse(4, 3, [6, 20, 20], 900, 4)
    rb = oracle.rulebook_subm(coords, [6, 20, 20], 3)
    assert rb.pair_cnt[13] == len(coords)                       # centre tap pairs every site with itself
    assert np.array_equal(rb.pair_cnt, rb.pair_cnt[::-1])       # offsets are symmetric
    d = oracle.dense(x, coords, 3, [6, 20, 20])
    assert np.array_equal(d[coords[:, 0], :, coords[:, 1], coords[:, 2], coords[:, 3]], x)
    assert np.count_nonzero(d) == np.count_nonzero(x)


def test_iou_pinned_to_reference_golden(oracle):
    g = np.load(os.path.join(G, "iou_ref.npz"))
    got = oracle.iou_bev(g["a"], g["b"])
    assert np.array_equal(got.view(np.uint32), g["iou"].view(np.uint32)), "oracle IoU is not bit-identical to iou3d_cpu.cpp"
    assert (g["iou"] > 0.3).sum() > 30


def test_iou_pinned_to_reference_live(oracle):
    ref = oracle.ref_cpu_module()
    if ref is None:
        pytest.skip("oracle/_ref not built on this box")
    a, _ = synth_nms_boxes(300, 21)
    b, _ = synth_nms_boxes(260, 22)
    out = torch.zeros(300, 260)
    ref.boxes_iou_bev_cpu(torch.from_numpy(a), torch.from_numpy(b), out)
    got = oracle.iou_bev(a, b)
    assert np.array_equal(got.view(np.uint32), out.numpy().view(np.uint32))
    assert not np.array_equal(oracle.iou_bev(a, a), oracle.iou_bev(a, a).T)   # not bitwise symmetric


def test_nms_greedy_semantics(oracle):
    boxes, scores = synth_nms_boxes(700, 5)
    order = np.argsort(-scores, kind="stable")
    b = boxes[order]
    keep, mask = oracle.nms(b, 0.5, rotated=True, return_mask=True)
    iou = oracle.iou_bev(b, b)
    ref_keep, alive = [], np.ones(len(b), bool)
    for i in range(len(b)):
        if alive[i]:
            ref_keep.append(i)
            alive[i + 1:] &= ~(iou[i, i + 1:] > 0.5)
    assert keep.tolist() == ref_keep
    bits = ((mask[:, :, None] >> np.arange(64, dtype=np.uint64)) & 1).reshape(len(b), -1)[:, :len(b)].astype(bool)
    assert np.array_equal(np.triu(bits, 1), np.triu(iou > 0.5, 1))
    assert len(oracle.nms(b[:0], 0.5)) == 0 and oracle.nms(b[:1], 0.5).tolist() == [0]
    kn = oracle.nms(b, 0.5, rotated=False)
    assert 0 < len(kn) <= len(b)


def test_bf16x3_split_error_model():
    """The tensor-core kernels compute x.w as hi.hi + lo.hi + hi.lo with hi = RN_bf16(v), lo = RN_bf16(v - hi) (split.cu,
    spconv_tc.cu).  Restated here with torch CPU bfloat16 (same round-to-nearest-even): the representation keeps
    2^-16 of |v| and a 3456-term dot product (27 taps x 128 channels) stays within 1e-4 of its magnitude scale --
    the bound DESIGN.md quotes, independent of the GPU."""
    import torch
    torch.manual_seed(0)
    x = torch.randn(512, 3456) * torch.logspace(-2, 2, 3456)
    w = torch.randn(3456, 64) / 3456 ** 0.5
    split = lambda v: (v.to(torch.bfloat16), (v - v.to(torch.bfloat16).float()).to(torch.bfloat16))
    xh, xl = split(x)
    wh, wl = split(w)
    rel = ((xh.float() + xl.float() - x).abs() / x.abs().clamp_min(1e-30)).max()
    assert float(rel) <= 2.0 ** -16                                   # hi + lo carries 16+ mantissa bits
    d = lambda a, b: a.double() @ b.double()                          # products of bf16 values are exact in fp64
    got = d(xh, wh) + d(xl, wh) + d(xh, wl)                           # the lo.lo term is dropped, as in the kernels
    ref = x.double() @ w.double()
    scale = (x.abs().double() @ w.abs().double())                     # sum |x||w|: what the error is relative to
    assert float(((got - ref).abs() / scale).max()) <= 3.0 * 2.0 ** -16
    assert float((got - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
