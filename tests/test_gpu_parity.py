"""GPU parity suite (`-m gpu`): every CUDA entry of the C ABI against the CPU oracle on the same
seeded inputs -- bit-exact for voxel indices / point slots / rulebooks / NMS keep lists and
masks, <= 1e-4 for fp32 features -- plus the reference's own CUDA kernels (oracle/_ref,
compiled unmodified for sm_100a) for the IoU/NMS bit patterns."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from cpd_b200.synth import PC_RANGE, VOXEL_SIZE, synth_nms_boxes, synth_scan

pytestmark = pytest.mark.gpu
TOL = 1e-4   # north_star: features within 1e-4 fp32


def dev(a, d):
    return torch.from_numpy(np.ascontiguousarray(a)).to(d)


# ------------------------------------------------------------------------------- voxelizer
def _check_vox(oracle, ops, cuda, frames, max_voxels=1000000):
    offs = np.cumsum([0] + [len(f) for f in frames]).tolist()
    pts = np.concatenate(frames, 0) if frames else np.zeros((0, 5), np.float32)
    out = ops.voxelize(dev(pts, cuda), offs, PC_RANGE, VOXEL_SIZE, 5, max_voxels)
    counts = out["counts"].cpu().numpy()
    row = 0
    for b, f in enumerate(frames):
        v, c, n = oracle.voxelize(f, PC_RANGE, VOXEL_SIZE, 5, max_voxels)
        m = len(c)
        assert counts[b] == m
        assert np.array_equal(out["coords"][row:row + m].cpu().numpy(), np.concatenate([np.full((m, 1), b, np.int32), c], 1))
        assert np.array_equal(out["num"][row:row + m].cpu().numpy(), n)
        assert np.array_equal(out["voxels"][row:row + m].cpu().numpy(), v)          # bit-exact point slots
        assert np.allclose(out["mean"][row:row + m].cpu().numpy(), oracle.mean_vfe(v, n), atol=1e-6)
        row += m
    assert counts[-1] == row == out["coords"].shape[0]


def test_voxelize_single_frames(oracle, cuda):
    from cpd_b200 import ops
    _check_vox(oracle, ops, cuda, [synth_scan(16000, 0)])
    _check_vox(oracle, ops, cuda, [synth_scan(160000, 1)])


def test_voxelize_ragged_batch_and_edges(oracle, cuda):
    from cpd_b200 import ops
    a, b = synth_scan(30000, 2), synth_scan(9000, 3)
    b[100] = [75.2, 0, 0, 0.5, 0]
    b[101] = [0, 0, 4.0, 0.5, 0]
    b[102] = [-75.2, -75.2, -2.0, 0.5, 0]
    unshuffled = a[np.lexsort((a[:, 1], a[:, 0]))]                     # many same-cell neighbours in one warp
    empty = np.zeros((0, 5), np.float32)
    far = np.full((50, 5), 400.0, np.float32)
    _check_vox(oracle, ops, cuda, [a, empty, b, far, unshuffled])
    _check_vox(oracle, ops, cuda, [a, b], max_voxels=2000)              # MAX_NUMBER_OF_VOXELS clamp
    dup = np.repeat(a[:40], 20, axis=0)                                 # 20 points per cell: only first 5 kept
    _check_vox(oracle, ops, cuda, [dup])


# ------------------------------------------------------------------------------- rulebooks
def _table_from_pairs(rb, m_out):
    t = np.full((m_out, rb.K), -1, np.int32)
    for k in range(rb.K):
        n = rb.pair_cnt[k]
        t[rb.pair_out[k, :n], k] = rb.pair_in[k, :n]
    return t


def _scene(oracle, n_pts, seeds):
    coords = []
    for b, s in enumerate(seeds):
        _, c, _ = oracle.voxelize(synth_scan(n_pts, s), PC_RANGE, VOXEL_SIZE)
        coords.append(np.concatenate([np.full((len(c), 1), b, np.int32), c], 1))
    return np.concatenate(coords, 0)


def test_rulebooks_bit_exact(oracle, cuda):
    from cpd_b200 import ops
    shape = [41, 1504, 1504]
    coords = _scene(oracle, 20000, [4, 5])
    dc = dev(coords, cuda)
    h = ops.build_hash(dc, shape, 2)
    nbr = ops.subm_table(dc, shape, 2, 3, h).cpu().numpy()
    assert np.array_equal(nbr, _table_from_pairs(oracle.rulebook_subm(coords, shape, 3), len(coords)))
    cur, cur_shape, cur_h = coords, shape, h
    for ks, st, pd in [(3, 2, 1), (3, 2, 1), (3, 2, (0, 1, 1)), ((3, 1, 1), (2, 1, 1), 0)]:
        rb = oracle.rulebook_strided(cur, cur_shape, ks, st, pd)
        oc, oshape = ops.strided_outputs(dev(cur, cuda), cur_shape, 2, ks, st, pd)
        assert oshape == rb.out_shape
        assert np.array_equal(oc.cpu().numpy(), rb.out_coords)          # same set, same (sorted) order
        oh = ops.build_hash(oc, oshape, 2)
        fwd, bwd = ops.strided_tables(dev(cur, cuda), cur_shape, cur_h, oc, oshape, oh, 2, ks, st, pd)
        assert np.array_equal(fwd.cpu().numpy(), _table_from_pairs(rb, rb.m_out))
        tb = np.full((len(cur), rb.K), -1, np.int32)
        for k in range(rb.K):
            n = rb.pair_cnt[k]
            tb[rb.pair_in[k, :n], k] = rb.pair_out[k, :n]
        assert np.array_equal(bwd.cpu().numpy(), tb)
        cur, cur_shape, cur_h = rb.out_coords, rb.out_shape, oh
    assert cur_shape == [2, 188, 188]


# ------------------------------------------------------------------------------- sparse conv
@pytest.mark.parametrize("cin,cout,kind", [(5, 16, "subm"), (16, 16, "subm"), (32, 64, "s2"), (64, 64, "subm"),
                                           (128, 128, "down"), (64, 128, "s2p0")])
def test_gather_gemm_fwd_bwd(oracle, cuda, cin, cout, kind):
    from cpd_b200 import ops
    shape = [41, 1504, 1504] if cin <= 16 else [11, 376, 376]
    coords = _scene(oracle, 12000, [6])
    if cin > 16:
        coords = np.unique(np.concatenate([coords[:, :1], coords[:, 1:] // 4], 1), axis=0).astype(np.int32)
    cfg = {"subm": (3, 1, 1), "s2": (3, 2, 1), "down": ((3, 1, 1), (2, 1, 1), 0), "s2p0": (3, 2, (0, 1, 1))}[kind]
    rng = np.random.default_rng(cin * 1000 + cout)
    m = len(coords)
    x = rng.normal(0, 1, (m, cin)).astype(np.float32)
    rb = oracle.rulebook_subm(coords, shape, cfg[0]) if kind == "subm" else oracle.rulebook_strided(coords, shape, *cfg)
    w = (rng.normal(0, 1, (cout, rb.K, cin)) / np.sqrt(cin * 4)).astype(np.float32)
    bias = rng.normal(0, 0.1, cout).astype(np.float32)
    y_ref = oracle.spconv_fwd(x, w, bias, rb)
    nbr = dev(_table_from_pairs(rb, rb.m_out), cuda)
    dx_, dw_ = dev(x, cuda), dev(w, cuda)
    y = ops.gather_gemm(dx_, dw_, nbr, bias=dev(bias, cuda), algo=ops.ALGO_SIMT)
    assert np.abs(y.cpu().numpy() - y_ref).max() <= TOL
    # fused epilogue: affine + residual + relu, and BN statistics
    scale, shift = rng.uniform(0.5, 1.5, cout).astype(np.float32), rng.normal(0, 0.2, cout).astype(np.float32)
    res = rng.normal(0, 1, y_ref.shape).astype(np.float32)
    stats = torch.zeros(2, cout, device=cuda)
    yf = ops.gather_gemm(dx_, dw_, nbr, bias=dev(bias, cuda), scale=dev(scale, cuda), shift=dev(shift, cuda),
                         residual=dev(res, cuda), relu=True, stats=stats, algo=ops.ALGO_SIMT)
    assert np.abs(yf.cpu().numpy() - np.maximum(y_ref * scale + shift + res, 0)).max() <= TOL
    assert np.allclose(stats[0].cpu().numpy(), y_ref.sum(0), rtol=1e-4, atol=1e-2)
    assert np.allclose(stats[1].cpu().numpy(), (y_ref.astype(np.float64) ** 2).sum(0), rtol=1e-4, atol=1e-2)
    # backward
    dy = rng.normal(0, 1, y_ref.shape).astype(np.float32)
    gx, gw, gb = oracle.spconv_bwd(x, w, dy, rb)
    ddy = dev(dy, cuda)
    if kind == "subm":
        dxg = ops.gather_gemm(ddy, ops.weight_transpose(dw_, flip_taps=True), nbr, algo=ops.ALGO_SIMT)
    else:
        tb = np.full((m, rb.K), -1, np.int32)
        for k in range(rb.K):
            n = rb.pair_cnt[k]
            tb[rb.pair_in[k, :n], k] = rb.pair_out[k, :n]
        dxg = ops.gather_gemm(ddy, ops.weight_transpose(dw_, flip_taps=False), dev(tb, cuda), algo=ops.ALGO_SIMT)
    assert np.abs(dxg.cpu().numpy() - gx).max() <= TOL
    scale_w = max(1.0, np.abs(gw).max())
    algos = [ops.ALGO_SIMT] + ([ops.ALGO_TCGEN05] if cin >= 8 and cin % 8 == 0 and cout >= 8 else [])
    nbr_t = nbr.t().contiguous()
    for algo in algos:                                             # fp32 SIMT and tcgen05 (3xTF32, MN-major) weight gradients
        dwg, dbg = ops.gather_wgrad(dx_, ddy, nbr_t, want_bias=True, algo=algo, tap_major=True)
        assert np.abs(dwg.cpu().numpy() - gw).max() <= TOL * scale_w, algo
        assert np.abs(dbg.cpu().numpy() - gb).max() <= TOL * max(1.0, np.abs(gb).max())
    if cin % 8 == 0 and cout in (16, 32, 64, 128, 256):            # tcgen05 forward / input-gradient
        yt = ops.gather_gemm(dx_, dw_, nbr, bias=dev(bias, cuda), algo=ops.ALGO_TCGEN05)
        assert np.abs(yt.cpu().numpy() - y_ref).max() <= TOL


def test_sparse_to_dense(oracle, cuda):
    from cpd_b200 import ops
    rng = np.random.default_rng(9)
    shape, batch, c = [2, 188, 188], 3, 128
    cells = rng.choice(batch * 2 * 188 * 188, 9000, replace=False)
    coords = np.stack([cells // (2 * 188 * 188), (cells // (188 * 188)) % 2, (cells // 188) % 188, cells % 188], 1).astype(np.int32)
    f = rng.normal(0, 1, (9000, c)).astype(np.float32)
    ref = oracle.dense(f, coords, batch, shape)
    got = ops.sparse_to_dense(dev(f, cuda), dev(coords, cuda), batch, shape, channels_last=False)
    assert np.array_equal(got.cpu().numpy(), ref)
    nhwc = ops.sparse_to_dense(dev(f, cuda), dev(coords, cuda), batch, shape, channels_last=True)
    assert np.array_equal(nhwc.permute(0, 3, 1, 2).cpu().numpy(), ref.reshape(batch, c * 2, 188, 188))
    back = ops.sparse_to_dense_bwd(nhwc, dev(coords, cuda), c, batch, shape, channels_last=True)
    assert np.array_equal(back.cpu().numpy(), f)


# ------------------------------------------------------------------------------- iou3d / nms
def _ref_gpu(oracle):
    lib = oracle.ref_gpu_lib()
    if lib is None:
        pytest.skip("oracle/_ref/libiou3d_ref_gpu.so not built")
    return lib


def test_iou_matrices(oracle, cuda):
    from cpd_b200 import ops
    a, _ = synth_nms_boxes(700, 31)
    b, _ = synth_nms_boxes(500, 32)
    b[:200] = a[:200] + np.random.default_rng(1).normal(0, 0.05, (200, 7)).astype(np.float32)
    da, db = dev(a, cuda), dev(b, cuda)
    iou = ops.iou_bev(da, db).cpu().numpy()
    ovl = ops.iou_bev(da, db, overlap=True).cpu().numpy()
    assert np.abs(iou - oracle.iou_bev(a, b)).max() < 1e-5          # CPU libm vs libdevice: not bitwise
    assert np.abs(ovl - oracle.overlap_bev(a, b)).max() < 1e-4
    lib = oracle.ref_gpu_lib()
    if lib is not None:                                             # the reference's own kernels: bitwise
        r_iou, r_ovl = torch.zeros(700, 500, device=cuda), torch.zeros(700, 500, device=cuda)
        vp = C.c_void_p
        assert lib.ref_iou_bev(vp(da.data_ptr()), 700, vp(db.data_ptr()), 500, vp(r_iou.data_ptr())) == 0
        assert lib.ref_overlap_bev(vp(da.data_ptr()), 700, vp(db.data_ptr()), 500, vp(r_ovl.data_ptr())) == 0
        assert np.array_equal(iou.view(np.uint32), r_iou.cpu().numpy().view(np.uint32))
        assert np.array_equal(ovl.view(np.uint32), r_ovl.cpu().numpy().view(np.uint32))


@pytest.mark.parametrize("n,thresh,rotated", [(1, 0.8, True), (63, 0.8, True), (500, 0.8, True), (500, 0.3, True),
                                              (4096, 0.3, True), (1000, 0.5, False)])
def test_nms_keep_bit_exact(oracle, cuda, n, thresh, rotated):
    from cpd_b200 import ops
    boxes, scores = synth_nms_boxes(n, 40 + n)
    b = boxes[np.argsort(-scores, kind="stable")]
    keep_ref = oracle.nms(b, thresh, rotated=rotated)
    keep, nk = ops.nms(dev(b, cuda), thresh, rotated=rotated)
    k = int(nk.item())
    assert k == len(keep_ref) and np.array_equal(keep[:k].cpu().numpy(), keep_ref)
    lib = oracle.ref_gpu_lib()
    if lib is not None:                                             # bit-matrix vs the reference kernel
        cb = (n + 63) // 64
        ref_mask = torch.zeros(n, cb, dtype=torch.int64, device=cuda)
        fn = lib.ref_nms_mask if rotated else lib.ref_nms_normal_mask
        assert fn(C.c_void_p(dev(b, cuda).data_ptr()), n, C.c_float(thresh), C.c_void_p(ref_mask.data_ptr())) == 0
        mine = ops.nms_mask(dev(b, cuda), thresh, rotated=rotated).cpu().numpy()
        refm = ref_mask.cpu().numpy()
        for r in range(n):                                          # tiles left of the diagonal are never read upstream
            assert np.array_equal(mine[r, r // 64:], refm[r, r // 64:])


def test_nms_module_surface(oracle, cuda):
    from cpd_b200 import iou3d_nms_cuda, iou3d_nms_utils
    boxes, scores = synth_nms_boxes(300, 77)
    db, ds = dev(boxes, cuda), dev(scores, cuda)
    order = np.argsort(-scores, kind="stable")
    keep_ref = order[oracle.nms(boxes[order], 0.8)]
    keep_cpu = torch.LongTensor(300)                                # the reference passes a CPU LongTensor
    num = iou3d_nms_cuda.nms_gpu(db[torch.from_numpy(order).to(cuda)].contiguous(), keep_cpu, 0.8)
    assert np.array_equal(order[keep_cpu[:num].numpy()], keep_ref)
    sel, _ = iou3d_nms_utils.nms_gpu(db, ds, 0.8)
    assert np.array_equal(sel.cpu().numpy(), keep_ref)
    out = torch.zeros(300, 300, device=cuda)
    assert iou3d_nms_cuda.boxes_iou_bev_gpu(db, db, out) == 1
    assert np.abs(out.cpu().numpy() - oracle.iou_bev(boxes, boxes)).max() < 1e-5
    i3 = iou3d_nms_utils.boxes_iou3d_gpu(db[:50], db[:60])
    assert i3.shape == (50, 60) and float(i3.max()) <= 1.0 + 1e-5


# ------------------------------------------------------------------------------- backbone end to end
def test_voxel_backbone8x_forward_matches_oracle(oracle, cuda):
    """BASELINE config 1 geometry: 16 k-point cloud through VoxelBackBone8x, eval mode."""
    from cpd_b200 import backbone, voxel
    torch.manual_seed(0)
    net = backbone.VoxelBackBone8x(dict(NUM_FILTERS=[16, 32, 64, 128], OUT_FEATURES=128), 5, [1504, 1504, 40]).to(cuda)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.uniform_(-0.2, 0.2); m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.7, 1.3); m.bias.data.uniform_(-0.2, 0.2)
    net.eval()
    pts = synth_scan(16000, 0)
    bd = voxel.voxelize_batch([dev(pts, cuda)], PC_RANGE, VOXEL_SIZE)
    bd["batch_size"] = 1
    with torch.no_grad():
        out = net(dict(bd))
    t = out["encoded_spconv_tensor"]
    from oracle import pipeline
    feats, coords, shape, _ = pipeline.backbone_forward(net, [pts], PC_RANGE, VOXEL_SIZE, eval_wide=True)
    assert t.spatial_shape == shape == [2, 188, 188]
    assert np.array_equal(t.indices.cpu().numpy(), coords)
    ref_scale = max(1.0, float(np.abs(feats).max()))
    assert np.abs(t.features.cpu().numpy() - feats).max() <= TOL * ref_scale
    # unfused (module-by-module) path gives the same result as the fused eval path
    with torch.enable_grad():
        out2 = net(dict(bd))
    assert np.abs(out2["encoded_spconv_tensor"].features.detach().cpu().numpy() - feats).max() <= TOL * ref_scale


def test_res_backbone_train_step_gradients(oracle, cuda):
    """VoxelResBackBone8x (MM tower on) forward+backward: gradients reach every parameter and the
    input-gradient of the first residual block matches the oracle's A.5 restatement."""
    from cpd_b200 import backbone, sparse as sp, voxel
    torch.manual_seed(1)
    net = backbone.VoxelResBackBone8x(dict(NUM_FILTERS=[16, 32, 64, 128], OUT_FEATURES=128, MM=True,
                                           RETURN_NUM_FEATURES_AS_DICT=True), 5, [1504, 1504, 40]).to(cuda).train()
    frames = [dev(synth_scan(8000, s), cuda) for s in (1, 2)]
    frames1 = [dev(synth_scan(8000, s), cuda) for s in (3, 4)]
    bd = voxel.voxelize_batch(frames, PC_RANGE, VOXEL_SIZE)
    bd1 = voxel.voxelize_batch(frames1, PC_RANGE, VOXEL_SIZE)
    bd.update(batch_size=2, voxel_features1=bd1["voxel_features"], voxel_coords1=bd1["voxel_coords"])
    out = net(bd)
    loss = out["encoded_spconv_tensor"].features.square().mean()
    for t in out["multi_scale_3d_features_mm"].values():
        loss = loss + t.features.square().mean()
    loss.backward()
    missing = [n for n, p in net.named_parameters() if p.grad is None or not torch.isfinite(p.grad).all()]
    assert not missing, missing
    # one SubM layer's autograd against the oracle
    conv = sp.SubMConv3d(16, 16, 3, bias=True, indice_key="t").to(cuda)
    x = sp.SparseConvTensor(torch.randn(bd["voxel_coords"].shape[0], 16, device=cuda, requires_grad=True),
                            bd["voxel_coords"], [41, 1504, 1504], 2)
    y = conv(x)
    dy = torch.randn_like(y.features)
    y.features.backward(dy)
    rb = oracle.rulebook_subm(bd["voxel_coords"].cpu().numpy(), [41, 1504, 1504], 3)
    gx, gw, gb = oracle.spconv_bwd(x.features.detach().cpu().numpy(), conv.weight.detach().cpu().numpy(),
                                   dy.cpu().numpy(), rb)
    assert np.abs(x.features.grad.cpu().numpy() - gx).max() <= TOL
    assert np.abs(conv.weight.grad.cpu().numpy() - gw).max() <= TOL * max(1.0, np.abs(gw).max())
    assert np.abs(conv.bias.grad.cpu().numpy() - gb).max() <= TOL * max(1.0, np.abs(gb).max())


# ------------------------------------------------------------------------------- fused training BatchNorm
@pytest.mark.parametrize("m,c,relu,res", [(5000, 16, True, False), (70001, 128, True, True), (333, 64, False, False), (20000, 256, True, False)])
def test_fused_bn_train_matches_torch(cuda, m, c, relu, res):
    """conv-epilogue statistics + cpd_bn_train_fwd/bwd against nn.BatchNorm1d (+ residual) (+ ReLU) in float64."""
    from cpd_b200 import sparse as sp
    torch.manual_seed(m + c)
    x = (torch.randn(m, c, device=cuda) * 1.7 + 0.3).requires_grad_(True)
    r = torch.randn(m, c, device=cuda, requires_grad=True) if res else None
    bn = torch.nn.BatchNorm1d(c, eps=1e-3, momentum=0.01).to(cuda).train()
    bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.uniform_(-0.3, 0.3)
    ref = torch.nn.BatchNorm1d(c, eps=1e-3, momentum=0.01).double().train()
    ref.load_state_dict({k: v.double().cpu() if v.is_floating_point() else v.cpu() for k, v in bn.state_dict().items()})
    stats = torch.stack([x.detach().sum(0), (x.detach() ** 2).sum(0)])            # what the conv epilogue would emit
    y = sp.bn_train(x, stats, bn, relu, r, dx_split=True)
    ys = sp._carried_split(y)                                                     # the same pass also wrote the bf16 hi | lo image of y
    assert ys is not None
    hi = y.detach().to(torch.bfloat16)
    assert torch.equal(ys[:, 0, :].view(torch.int16), hi.view(torch.int16))
    assert torch.equal(ys[:, 1, :].view(torch.int16), (y.detach() - hi.float()).to(torch.bfloat16).view(torch.int16))
    xr = x.detach().double().cpu().requires_grad_(True)
    rr = r.detach().double().cpu().requires_grad_(True) if res else None
    yr = ref(xr)
    if res:
        yr = yr + rr
    if relu:
        yr = torch.relu(yr)
    assert float((y.detach().double().cpu() - yr.detach()).abs().max()) <= TOL
    dy = torch.randn_like(y)
    y.backward(dy)
    yr.backward(dy.double().cpu())
    sc = lambda t: max(1.0, float(t.abs().max()))
    assert float((x.grad.double().cpu() - xr.grad).abs().max()) <= TOL * sc(xr.grad)
    dxs = sp._carried_split(x.grad)                                               # ... and the backward the image of dx
    if dxs is not None:
        assert torch.equal(dxs[:, 0, :].view(torch.int16), x.grad.to(torch.bfloat16).view(torch.int16))
    assert float((bn.weight.grad.double().cpu() - ref.weight.grad).abs().max()) <= TOL * sc(ref.weight.grad)
    assert float((bn.bias.grad.double().cpu() - ref.bias.grad).abs().max()) <= TOL * sc(ref.bias.grad)
    if res:
        assert float((r.grad.double().cpu() - rr.grad).abs().max()) <= TOL
    assert float((bn.running_mean.double().cpu() - ref.running_mean).abs().max()) <= 1e-5
    assert float((bn.running_var.double().cpu() - ref.running_var).abs().max()) <= 1e-5
    assert int(bn.num_batches_tracked) == 1


# ------------------------------------------------------------------------------- tcgen05 operand format + scheduling aids
def test_split_rows_image_is_the_rn_bf16_split(cuda):
    """cpd_split_rows: hi = RN_bf16(x), lo = RN_bf16(x - hi), bit-exact; column sums from the same pass."""
    from cpd_b200 import ops
    torch.manual_seed(5)
    for m, c in ((1, 8), (777, 16), (4099, 32), (3000, 128), (513, 40)):
        x = torch.randn(m, c, device=cuda) * torch.logspace(-3, 3, c, device=cuda)
        x[0, 0] = 0.0
        xs, cs = ops.split_rows(x, colsum=True)
        hi = x.to(torch.bfloat16)
        lo = (x - hi.float()).to(torch.bfloat16)
        assert xs.shape == (m, 2, c)
        assert torch.equal(xs[:, 0, :].view(torch.int16), hi.view(torch.int16))
        assert torch.equal(xs[:, 1, :].view(torch.int16), lo.view(torch.int16))
        rec = xs[:, 0, :].float() + xs[:, 1, :].float()
        assert float(((rec - x).abs() / x.abs().clamp_min(1e-30)).max()) <= 2.0 ** -15      # x = hi + lo to ~2^-16
        if c & (c - 1) == 0:
            ref = x.double().sum(0)
            assert float((cs.double() - ref).abs().max()) <= 1e-4 * max(1.0, float(x.abs().sum(0).max()))
        else:
            assert cs is None
    assert ops.split_rows(torch.randn(5, 5, device=cuda)) is None         # needs c % 8 == 0


def test_tile_tap_masks_and_kblock_skipping_are_exact(cuda):
    """Tile masks only skip work that multiplies zero rows: results are bit-identical with and without them,
    also for a strided input-gradient table whose rows were grouped by tap pattern."""
    from cpd_b200 import ops
    torch.manual_seed(6)
    m_in, m_out, cin, cout, K = 5000, 3333, 32, 64, 27
    nbr = torch.randint(0, m_in, (m_out, K), device=cuda, dtype=torch.int32)
    nbr[torch.rand(m_out, K, device=cuda) > 0.3] = -1
    nbr[:512, 9:] = -1                       # whole tiles without taps 9..26
    nbr[1024:1152] = -1                      # an empty tile
    masks = ops.tile_tap_masks(nbr)
    valid = (nbr >= 0)
    for t in range((m_out + 127) // 128):
        bits = valid[128 * t:128 * (t + 1)].any(0).cpu().numpy()
        assert int(masks[t].item()) & 0xffffffff == sum(1 << k for k in range(K) if bits[k])
    x = torch.randn(m_in, cin, device=cuda)
    w = torch.randn(cout, K, cin, device=cuda) * 0.05
    y0 = ops.gather_gemm(x, w, nbr, algo=ops.ALGO_TCGEN05)
    y1 = ops.gather_gemm(x, w, nbr, algo=ops.ALGO_TCGEN05, tile_masks=masks)
    assert torch.equal(y0, y1)
    assert float(y1[1024:1152].abs().max()) == 0.0
    ref = ops.gather_gemm(x, w, nbr, algo=ops.ALGO_SIMT)
    assert float((y1 - ref).abs().max()) <= TOL * max(1.0, float(ref.abs().max()))
    # rows grouped by tap pattern, results scattered back
    key = (valid.long() << torch.arange(K, device=cuda)).sum(1)
    perm = torch.argsort(key, stable=True)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(m_out, device=cuda)
    nbs = nbr[perm].contiguous()
    y2 = ops.gather_gemm(x, w, nbs, algo=ops.ALGO_TCGEN05, tile_masks=ops.tile_tap_masks(nbs))[inv]
    assert float((y2 - y0).abs().max()) <= 2e-6 * max(1.0, float(y0.abs().max()))
    # ... or scattered by the kernel's epilogue itself
    y3 = ops.gather_gemm(x, w, nbs, algo=ops.ALGO_TCGEN05, tile_masks=ops.tile_tap_masks(nbs), out_rows=perm.to(torch.int32))
    assert torch.equal(y3, y2)


def test_shared_split_images_give_the_same_results(cuda):
    from cpd_b200 import ops
    torch.manual_seed(7)
    m, cin, cout, K = 4000, 64, 32, 27
    nbr = torch.randint(-m, m, (m, K), device=cuda, dtype=torch.int32).clamp_(min=-1)
    x, dy = torch.randn(m, cin, device=cuda), torch.randn(m, cout, device=cuda)
    w = torch.randn(cout, K, cin, device=cuda) * 0.05
    xs, (dys, db) = ops.split_rows(x), ops.split_rows(dy, colsum=True)
    assert torch.equal(ops.gather_gemm(x, w, nbr), ops.gather_gemm(x, w, nbr, x_split=xs))
    nbr_t = nbr.t().contiguous()
    dw0, db0 = ops.gather_wgrad(x, dy, nbr_t, want_bias=True, tap_major=True)
    dw1, _ = ops.gather_wgrad(x, dy, nbr_t, tap_major=True, x_split=xs, dy_split=dys)
    scale = max(1.0, float(dw0.abs().max()))
    assert float((dw0 - dw1).abs().max()) <= 1e-5 * scale                 # split-K atomics: order differs run to run
    assert float((db0 - db).abs().max()) <= 1e-4 * max(1.0, float(db0.abs().max()))
