"""Property tests of the oracle (the reference ships no vectors for the spconv part, so beyond the dense
conv3d cross-check the restatement is held to the size-independent properties the domain offers).  hypothesis drives small random
clouds / site sets; the same properties are checked on the CUDA kernels at full size in tests/test_gpu_properties.py."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

RANGE = [0.0, -4.0, -1.0, 8.0, 4.0, 1.0]
VS = [0.1, 0.1, 0.25]
GRID = [80, 80, 8]            # x, y, z cells


def _cloud(seed, n):
    g = np.random.default_rng(seed)
    pts = np.concatenate([g.uniform(-0.5, 8.5, (n, 1)), g.uniform(-4.5, 4.5, (n, 1)), g.uniform(-1.2, 1.2, (n, 1)), g.uniform(0, 1, (n, 2))], 1)
    return pts.astype(np.float32)


def _sites(seed, m, shape, batch=2):
    g = np.random.default_rng(seed)
    cells = g.choice(batch * shape[0] * shape[1] * shape[2], size=min(m, batch * shape[0] * shape[1] * shape[2] // 2), replace=False)
    cells.sort()
    b, r = np.divmod(cells, shape[0] * shape[1] * shape[2])
    z, r = np.divmod(r, shape[1] * shape[2])
    y, x = np.divmod(r, shape[2])
    return np.stack([b, z, y, x], 1).astype(np.int32)


FAST = dict(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck))


@settings(**FAST)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(0, 3000))
def test_voxel_set_is_permutation_invariant_and_slots_are_first_come(oracle, seed, n):
    pts = _cloud(seed, n)
    v, c, num = oracle.voxelize(pts, RANGE, VS, 5, 100000)
    # every in-range point's cell is a voxel, every voxel holds min(count, 5) points, rows are first-appearance ordered
    ijk = np.floor((pts[:, :3] - np.array(RANGE[:3], np.float32)) / np.array(VS, np.float32)).astype(np.int64)
    ok = np.all((ijk >= 0) & (ijk < np.array(GRID)), 1)
    keys = (ijk[:, 2] * GRID[1] + ijk[:, 1]) * GRID[0] + ijk[:, 0]
    uniq, first, counts = np.unique(keys[ok], return_index=True, return_counts=True)
    vk = (c[:, 0].astype(np.int64) * GRID[1] + c[:, 1]) * GRID[0] + c[:, 2]
    assert len(vk) == len(uniq) and np.array_equal(np.sort(vk), uniq)
    order = np.argsort(first, kind="stable")
    assert np.array_equal(vk, uniq[order])                                   # voxel id = rank of the cell's first point
    assert np.array_equal(num, np.minimum(counts[order], 5))
    # the SET of voxels (and the per-voxel point counts) does not depend on the order of the points
    perm = np.random.default_rng(seed + 1).permutation(n)
    v2, c2, num2 = oracle.voxelize(pts[perm], RANGE, VS, 5, 100000)
    k2 = (c2[:, 0].astype(np.int64) * GRID[1] + c2[:, 1]) * GRID[0] + c2[:, 2]
    a, b = np.argsort(vk), np.argsort(k2)
    assert np.array_equal(vk[a], k2[b]) and np.array_equal(num[a], num2[b])
    # slots hold the FIRST points of the cell in input order; the rest of a voxel's rows are zero
    if len(vk):
        j = int(np.argmax(num))
        members = np.nonzero(ok & (keys == vk[j]))[0][:5]
        assert np.array_equal(v[j, :len(members)], pts[members]) and not v[j, len(members):].any()


@settings(**FAST)
@given(seed=st.integers(0, 10 ** 6), m=st.integers(1, 600), cin=st.sampled_from([1, 3, 8]), cout=st.sampled_from([1, 4]))
def test_subm_keeps_the_site_set_and_is_linear(oracle, seed, m, cin, cout):
    shape = [5, 12, 9]
    coords = _sites(seed, m, shape)
    rb = oracle.rulebook_subm(coords, shape, 3)
    assert rb.m_out == rb.m_in == len(coords)                                # SubM: output sites == input sites
    centre = rb.K // 2
    assert rb.pair_cnt[centre] == len(coords)
    assert np.array_equal(rb.pair_in[centre, :len(coords)], rb.pair_out[centre, :len(coords)])
    for k in range(rb.K):                                                    # tap k and its mirror hold the same pairs, swapped
        n = rb.pair_cnt[k]
        assert n == rb.pair_cnt[rb.K - 1 - k]
        a = set(zip(rb.pair_in[k, :n].tolist(), rb.pair_out[k, :n].tolist()))
        b = set(zip(rb.pair_out[rb.K - 1 - k, :n].tolist(), rb.pair_in[rb.K - 1 - k, :n].tolist()))
        assert a == b
    g = np.random.default_rng(seed)
    w = g.normal(0, 1, (cout, 3, 3, 3, cin)).astype(np.float32)
    x1, x2 = g.normal(0, 1, (len(coords), cin)).astype(np.float32), g.normal(0, 1, (len(coords), cin)).astype(np.float32)
    y1, y2, y12 = oracle.spconv_fwd(x1, w, None, rb), oracle.spconv_fwd(x2, w, None, rb), oracle.spconv_fwd(x1 + 2 * x2, w, None, rb)
    assert np.allclose(y12, y1 + 2 * y2, atol=1e-4 * max(1.0, float(np.abs(y12).max())))
    # <dy, conv(x)> == <conv^T(dy), x> == <dW, w>: the backward is the adjoint of the forward
    dy = g.normal(0, 1, y1.shape).astype(np.float32)
    dx, dw, _ = oracle.spconv_bwd(x1, w, dy, rb, need_bias=False)
    lhs = float((dy.astype(np.float64) * y1).sum())
    assert abs(lhs - float((dx.astype(np.float64) * x1).sum())) <= 1e-3 * max(1.0, abs(lhs))
    assert abs(lhs - float((dw.astype(np.float64) * w).sum())) <= 1e-3 * max(1.0, abs(lhs))


@settings(**FAST)
@given(seed=st.integers(0, 10 ** 6), m=st.integers(1, 500), stride=st.sampled_from([(2, 2, 2), (1, 2, 2), (2, 1, 1)]),
       pad=st.sampled_from([(1, 1, 1), (0, 1, 1), (0, 0, 0)]))
def test_strided_outputs_are_exactly_the_cells_an_input_reaches(oracle, seed, m, stride, pad):
    shape = [6, 11, 10]
    ks = (3, 3, 3) if stride != (2, 1, 1) else (3, 1, 1)
    pad = pad if stride != (2, 1, 1) else (pad[0], 0, 0)
    coords = _sites(seed, m, shape)
    rb = oracle.rulebook_strided(coords, shape, ks, stride, pad)
    osh = [(shape[i] + 2 * pad[i] - ks[i]) // stride[i] + 1 for i in range(3)]
    assert rb.out_shape == osh
    want = set()
    for b, z, y, x in coords.tolist():
        for kz in range(ks[0]):
            for ky in range(ks[1]):
                for kx in range(ks[2]):
                    n = (z + pad[0] - kz, y + pad[1] - ky, x + pad[2] - kx)
                    if all(v >= 0 and v % s == 0 and v // s < o for v, s, o in zip(n, stride, osh)):
                        want.add((b, n[0] // stride[0], n[1] // stride[1], n[2] // stride[2]))
    got = [tuple(r) for r in rb.out_coords.tolist()]
    assert len(got) == len(set(got)) and set(got) == want
    key = [((b * osh[0] + z) * osh[1] + y) * osh[2] + x for b, z, y, x in got]
    assert key == sorted(key)                                                # deterministic, batch-major order


@settings(**FAST)
@given(seed=st.integers(0, 10 ** 6), m=st.integers(0, 400), c=st.sampled_from([1, 5, 16]), batch=st.integers(1, 3))
def test_dense_round_trip(oracle, seed, m, c, batch):
    shape = [3, 7, 9]
    coords = _sites(seed, m, shape, batch) if m else np.zeros((0, 4), np.int32)
    feat = np.random.default_rng(seed).normal(0, 1, (len(coords), c)).astype(np.float32)
    d = oracle.dense(feat, coords, batch, shape)
    assert d.shape == (batch, c, *shape)
    if len(coords):
        back = d[coords[:, 0], :, coords[:, 1], coords[:, 2], coords[:, 3]]
        assert np.array_equal(back, feat)                                    # gather(dense(x)) == x
    assert np.count_nonzero(d) == np.count_nonzero(feat)                     # nothing else is written


def _boxes(seed, n):
    g = np.random.default_rng(seed)
    c = g.uniform(-8, 8, (max(n // 4, 1), 2))
    xy = c[g.integers(0, len(c), n)] + g.normal(0, 0.7, (n, 2))
    return np.concatenate([xy, g.uniform(-1, 1, (n, 1)), g.uniform(0.6, 5.0, (n, 2)), g.uniform(1.0, 2.0, (n, 1)), g.uniform(-3.2, 3.2, (n, 1))],
                          1).astype(np.float32)


@settings(**FAST)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 120), thresh=st.sampled_from([0.1, 0.5, 0.8]))
def test_rotated_iou_and_nms_invariants(oracle, seed, n, thresh):
    b = _boxes(seed, n)
    iou = oracle.iou_bev(b, b)
    assert iou.shape == (n, n) and np.all(iou >= 0) and np.all(iou <= 1 + 1e-4)
    assert np.allclose(np.diag(iou), 1.0, atol=2e-3)                          # a box overlaps itself completely (fp32 clipping slack)
    assert np.allclose(iou, iou.T, atol=2e-3)                                 # symmetric up to the clipping order
    far = b.copy()
    far[:, 0] += 100.0
    assert not oracle.iou_bev(b, far).any()                                   # disjoint boxes
    keep = oracle.nms(b, thresh)                                              # boxes are taken as already score-sorted
    assert keep[0] == 0 and np.all(np.diff(keep) > 0)                         # the best box always survives; order is kept
    kept = b[keep]
    sub = np.triu(oracle.iou_bev(kept, kept), 1)
    assert sub.max(initial=0.0) <= thresh + 1e-5                              # survivors do not overlap above the threshold
    removed = np.setdiff1d(np.arange(n), keep)
    if len(removed):                                                          # every removed box is covered by an earlier survivor
        cover = oracle.iou_bev(b[removed], kept)
        earlier = np.asarray(keep)[None, :] < removed[:, None]
        assert np.all((cover * earlier).max(1) > thresh)
    assert np.array_equal(oracle.nms(kept, thresh), np.arange(len(kept)))     # idempotent


@settings(max_examples=20, deadline=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 10 ** 6), dims=st.tuples(st.integers(3, 9), st.integers(4, 13), st.integers(4, 11)),
       geom=st.sampled_from([((3, 3, 3), (1, 1, 1), (1, 1, 1)), ((3, 3, 3), (2, 2, 2), (1, 1, 1)), ((3, 3, 3), (2, 2, 2), (0, 1, 1)),
                             ((3, 1, 1), (2, 1, 1), (0, 0, 0)), ((3, 3, 3), (1, 2, 2), (1, 1, 1)), ((1, 3, 3), (1, 1, 1), (0, 1, 1))]),
       cin=st.integers(1, 6), cout=st.integers(1, 5), fill=st.floats(0.05, 0.6))
def test_sparse_conv_equals_dense_conv3d_on_random_geometry(oracle, seed, dims, geom, cin, cout, fill):
    """The independent pin of the spconv restatement (tests/test_oracle_cpu.py) over random grids, kernel / stride / padding
    combinations, channel counts and occupancies: forward and all three gradients == torch conv3d on the densified input,
    read at the active outputs; the active output set == the cells with a non-empty receptive field."""
    import torch
    ks, stv, pd = geom
    subm = stv == (1, 1, 1) and all(k % 2 == 1 for k in ks) and pd == tuple(k // 2 for k in ks)
    shape, batch = list(dims), 2
    g = np.random.default_rng(seed)
    ncell = batch * shape[0] * shape[1] * shape[2]
    cells = np.sort(g.choice(ncell, size=max(1, int(fill * ncell)), replace=False))
    b, r = np.divmod(cells, shape[0] * shape[1] * shape[2])
    z, r = np.divmod(r, shape[1] * shape[2])
    y, x = np.divmod(r, shape[2])
    coords = np.stack([b, z, y, x], 1).astype(np.int32)
    xf = g.normal(0, 1, (len(coords), cin)).astype(np.float32)
    w = g.normal(0, 0.4, (cout, *ks, cin)).astype(np.float32)
    bias = g.normal(0, 0.2, cout).astype(np.float32)
    osh = [(shape[i] + 2 * pd[i] - ks[i]) // stv[i] + 1 for i in range(3)]
    if min(osh) <= 0:
        return
    rb = oracle.rulebook_subm(coords, shape, list(ks)) if subm else oracle.rulebook_strided(coords, shape, list(ks), list(stv), list(pd))
    yv = oracle.spconv_fwd(xf, w, bias, rb)
    xd = torch.zeros(batch, cin, *shape, dtype=torch.float64)
    xd[coords[:, 0], :, coords[:, 1], coords[:, 2], coords[:, 3]] = torch.from_numpy(xf).double()
    xd.requires_grad_(True)
    wt = torch.from_numpy(w).double().permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
    bt = torch.from_numpy(bias).double().requires_grad_(True)
    yd = torch.nn.functional.conv3d(xd, wt, bt, stride=list(stv), padding=list(pd))
    oc = rb.out_coords
    assert list(yd.shape[2:]) == list(rb.out_shape)
    if not subm:
        occ_in = torch.zeros(batch, 1, *shape, dtype=torch.float64)
        occ_in[coords[:, 0], 0, coords[:, 1], coords[:, 2], coords[:, 3]] = 1.0
        occ = torch.nn.functional.conv3d(occ_in, torch.ones(1, 1, *ks, dtype=torch.float64), stride=list(stv), padding=list(pd))
        assert int((occ > 0).sum()) == rb.m_out
    got = yd[oc[:, 0], :, oc[:, 1], oc[:, 2], oc[:, 3]]
    scale = max(1.0, float(got.detach().abs().max()))
    assert float(np.abs(got.detach().numpy() - yv).max()) <= 1e-4 * scale
    dy = g.normal(0, 1, yv.shape).astype(np.float32)
    dx, dw, db = oracle.spconv_bwd(xf, w, dy, rb)
    (got * torch.from_numpy(dy).double()).sum().backward()
    gx = xd.grad[coords[:, 0], :, coords[:, 1], coords[:, 2], coords[:, 3]].numpy()
    assert float(np.abs(gx - dx).max()) <= 1e-4 * max(1.0, float(np.abs(gx).max()))
    gw = wt.grad.permute(0, 2, 3, 4, 1).numpy()
    assert float(np.abs(gw - dw).max()) <= 1e-4 * max(1.0, float(np.abs(gw).max()))
    assert float(np.abs(bt.grad.numpy() - db).max()) <= 1e-4 * max(1.0, float(np.abs(db).max()))
