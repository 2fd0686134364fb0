"""GPU parity, second tier (`-m gpu`): the production paths the first suite only reached through the SIMT kernels.

* tcgen05 input-gradient (SubM flipped taps, strided plain table, strided table grouped by tap pattern with the
  epilogue row scatter `out_rows`) and the tcgen05 fused epilogue (bias / folded-BN affine / residual / ReLU / BatchNorm
  statistics) against the CPU oracle's A.3-A.5 restatement (oracle.spconv_fwd / spconv_bwd);
* the weight gradient held to a bound stated against sum |x| |dy| (bf16x3 products carry 2^-16 of each |x||dy| term);
* one case at the BENCHMARKED size (4 x 160 k-point frames: stage-2 SubM 32->32 over the real rulebook; a 4 x 188 x 188
  256->256 3x3 BEV convolution) -- 32-bit index arithmetic, multi-wave scheduling, N-tiling;
* the whole train step of CPDHotPathDetector against the CPU port (oracle.pipeline.CpuDetector -- the arm
  `bench.py --impl reference` times): same weights, same batch, loss and every parameter gradient.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from cpd_b200.synth import PC_RANGE, VOXEL_SIZE, synth_gt_boxes, synth_scan

pytestmark = pytest.mark.gpu
TOL = 1e-4


def dev(a, d):
    return torch.from_numpy(np.ascontiguousarray(a)).to(d)


def _table_from_pairs(rb, m_out):
    t = np.full((m_out, rb.K), -1, np.int32)
    for k in range(rb.K):
        n = rb.pair_cnt[k]
        t[rb.pair_out[k, :n], k] = rb.pair_in[k, :n]
    return t


def _bwd_table(rb, m_in):
    t = np.full((m_in, rb.K), -1, np.int32)
    for k in range(rb.K):
        n = rb.pair_cnt[k]
        t[rb.pair_in[k, :n], k] = rb.pair_out[k, :n]
    return t


def _scene(oracle, n_pts, seeds, div=1):
    coords = []
    for b, s in enumerate(seeds):
        _, c, _ = oracle.voxelize(synth_scan(n_pts, s), PC_RANGE, VOXEL_SIZE)
        coords.append(np.concatenate([np.full((len(c), 1), b, np.int32), c], 1))
    coords = np.concatenate(coords, 0)
    if div > 1:
        coords = np.unique(np.concatenate([coords[:, :1], coords[:, 1:] // div], 1), axis=0).astype(np.int32)
    return coords


CASES = [(16, 16, "subm"), (32, 32, "subm"), (32, 64, "s2"), (64, 64, "subm"), (128, 128, "subm"), (128, 128, "down"),
         (64, 128, "s2p0"), (16, 32, "s2")]


@pytest.mark.parametrize("cin,cout,kind", CASES)
def test_tcgen05_fwd_epilogue_and_dgrad_match_oracle(oracle, cuda, cin, cout, kind):
    """Everything the training / inference step runs on the tensor cores, against the oracle (not against SIMT)."""
    from cpd_b200 import ops
    from cpd_b200.sparse import Rulebook
    shape = [41, 1504, 1504] if cin <= 16 else [11, 376, 376]
    coords = _scene(oracle, 14000, [11, 12], div=1 if cin <= 16 else 4)
    cfg = {"subm": (3, 1, 1), "s2": (3, 2, 1), "down": ((3, 1, 1), (2, 1, 1), 0), "s2p0": (3, 2, (0, 1, 1))}[kind]
    rng = np.random.default_rng(cin * 977 + cout)
    m = len(coords)
    x = rng.normal(0, 1, (m, cin)).astype(np.float32)
    rb = oracle.rulebook_subm(coords, shape, cfg[0]) if kind == "subm" else oracle.rulebook_strided(coords, shape, *cfg)
    w = (rng.normal(0, 1, (cout, rb.K, cin)) / np.sqrt(cin * 4)).astype(np.float32)
    bias = rng.normal(0, 0.1, cout).astype(np.float32)
    y_ref = oracle.spconv_fwd(x, w, bias, rb)
    nbr = dev(_table_from_pairs(rb, rb.m_out), cuda)
    dx_, dw_, db_ = dev(x, cuda), dev(w, cuda), dev(bias, cuda)
    masks = ops.tile_tap_masks(nbr) if rb.K <= 32 else None
    # ---- forward with the fused epilogue, tcgen05 ----
    scale, shift = rng.uniform(0.5, 1.5, cout).astype(np.float32), rng.normal(0, 0.2, cout).astype(np.float32)
    res = rng.normal(0, 1, y_ref.shape).astype(np.float32)
    stats = torch.full((2, cout), 7.0, device=cuda)                     # the call must overwrite, not accumulate
    yf = ops.gather_gemm(dx_, dw_, nbr, bias=db_, scale=dev(scale, cuda), shift=dev(shift, cuda), residual=dev(res, cuda),
                         relu=True, stats=stats, algo=ops.ALGO_TCGEN05, tile_masks=masks)
    ys = max(1.0, float(np.abs(y_ref).max()))
    assert np.abs(yf.cpu().numpy() - np.maximum(y_ref * scale + shift + res, 0)).max() <= TOL * ys
    assert np.allclose(stats[0].cpu().numpy(), y_ref.astype(np.float64).sum(0), rtol=1e-4, atol=1e-4 * rb.m_out ** 0.5 * ys)
    assert np.allclose(stats[1].cpu().numpy(), (y_ref.astype(np.float64) ** 2).sum(0), rtol=2e-4, atol=1e-2)
    # ---- input gradient on tcgen05 ----
    dy = rng.normal(0, 1, y_ref.shape).astype(np.float32)
    gx, gw, gb = oracle.spconv_bwd(x, w, dy, rb)
    ddy = dev(dy, cuda)
    gs = max(1.0, float(np.abs(gx).max()))
    if kind == "subm":
        dxg = ops.gather_gemm(ddy, ops.weight_transpose(dw_, flip_taps=True), nbr, algo=ops.ALGO_TCGEN05, tile_masks=masks)
        assert np.abs(dxg.cpu().numpy() - gx).max() <= TOL * gs
    else:
        tb = dev(_bwd_table(rb, m), cuda)
        wt = ops.weight_transpose(dw_, flip_taps=False)
        dxg = ops.gather_gemm(ddy, wt, tb, algo=ops.ALGO_TCGEN05)
        assert np.abs(dxg.cpu().numpy() - gx).max() <= TOL * gs
        # the production path: rows grouped by tap pattern (Rulebook.bwd_sorted) + epilogue row scatter (out_rows)
        book = Rulebook("strided", nbr, tb, None, None, shape, rb.out_shape, list(oracle._tri(cfg[0])), list(oracle._tri(cfg[1])),
                        list(oracle._tri(cfg[2])))
        grouped = book.bwd_sorted()
        assert (grouped is not None) == (rb.K >= 4)                     # the (3,1,1) conv_out kernel is not worth regrouping
        if grouped is not None:
            nbs, out_rows, gmasks = grouped
            assert sorted(out_rows.cpu().tolist()) == list(range(m))    # a permutation of the input rows
            dxs = ops.gather_gemm(ddy, wt, nbs, algo=ops.ALGO_TCGEN05, x_split=ops.split_rows(ddy), tile_masks=gmasks, out_rows=out_rows)
            assert np.abs(dxs.cpu().numpy() - gx).max() <= TOL * gs
    # ---- weight / bias gradient on tcgen05, image of dy shared with the input gradient ----
    dys, dbias = ops.split_rows(ddy, colsum=True)
    dwg, _ = ops.gather_wgrad(dx_, ddy, nbr.t().contiguous(), algo=ops.ALGO_TCGEN05, tap_major=True, x_split=ops.split_rows(dx_), dy_split=dys)
    assert np.abs(dwg.cpu().numpy() - gw).max() <= TOL * max(1.0, float(np.abs(gw).max()))
    assert np.abs(dbias.cpu().numpy() - gb).max() <= TOL * max(1.0, float(np.abs(gb).max()))


def test_wgrad_error_is_bounded_by_sum_abs_products(cuda):
    """bf16x3 drops the lo.lo term and rounds hi / lo to bf16: every product x*dy carries <= ~2^-16 |x||dy| of error, so
    |dW - dW_exact| <= c * 2^-16 * sum_o |dy[o,co]| |x[nbr[o,k],ci]| ELEMENTWISE -- also when the sum itself cancels
    (BatchNorm backward makes dy zero-mean per channel), where a tolerance relative to max|dW| says little."""
    from cpd_b200 import ops
    torch.manual_seed(21)
    m, cin, cout, K = 60000, 64, 64, 27
    nbr = torch.randint(0, m, (m, K), device=cuda, dtype=torch.int32)
    nbr[torch.rand(m, K, device=cuda) > 0.5] = -1
    x = torch.randn(m, cin, device=cuda).abs_() + 0.5                   # post-ReLU-like, positive
    dy = torch.randn(m, cout, device=cuda)
    dy -= dy.mean(0, keepdim=True)                                      # zero-mean per channel: sums cancel heavily
    nbr_t = nbr.t().contiguous()
    dw, _ = ops.gather_wgrad(x, dy, nbr_t, algo=ops.ALGO_TCGEN05, tap_major=True)
    bound, _ = ops.gather_wgrad(x.abs(), dy.abs(), nbr_t, algo=ops.ALGO_SIMT, tap_major=True)        # sum |x||dy| in fp32
    # exact reference in float64, tap by tap
    ref = torch.zeros(cout, K, cin, dtype=torch.float64, device=cuda)
    xd, dyd = x.double(), dy.double()
    for k in range(K):
        idx = nbr[:, k].long()
        ok = idx >= 0
        ref[:, k, :] = dyd[ok].t() @ xd[idx[ok]]
    err = (dw.double() - ref).abs()
    ratio = float((err / bound.double().clamp_min(1e-30)).max())
    print(f"wgrad: max |err| / sum|x||dy| = {ratio:.3e} (2^-16 = {2.0 ** -16:.3e}); max |err| {float(err.max()):.3e}, max |dW| {float(ref.abs().max()):.3e}")
    assert ratio <= 2.0 ** -15                                          # measured ~2^-18: fp32 accumulation included
    # same statement for the forward GEMM
    w = torch.randn(cout, K, cin, device=cuda) * 0.05
    y = ops.gather_gemm(x, w, nbr, algo=ops.ALGO_TCGEN05)
    yb = ops.gather_gemm(x.abs(), w.abs(), nbr, algo=ops.ALGO_SIMT)
    yref = torch.zeros(m, cout, dtype=torch.float64, device=cuda)
    for k in range(K):
        idx = nbr[:, k].long()
        ok = idx >= 0
        yref[ok] += xd[idx[ok]] @ w[:, k, :].double().t()
    ratio_y = float(((y.double() - yref).abs() / yb.double().clamp_min(1e-30)).max())
    print(f"fwd: max |err| / sum|x||w| = {ratio_y:.3e}")
    assert ratio_y <= 2.0 ** -15


def test_bench_scale_sparse_and_dense(oracle, cuda):
    """The sizes bench.py runs: bs=4 x 160 k points.  Stage-2 SubM 32->32 over the real stage-2 rulebook (hundreds of
    thousands of rows, thousands of tiles, several waves of the persistent scheduler) against the oracle; a
    4 x 188 x 188 256->256 3x3 convolution and the 512->64 shared head conv (cin > 256: k-block bitmap words > 1)
    against float64 torch on the first and the last frame."""
    from cpd_b200 import ops, voxel
    from cpd_b200.bev import DenseConv2d, DenseMap
    frames = [synth_scan(160000, 900 + i) for i in range(4)]
    bd = voxel.voxelize_batch([dev(f, cuda) for f in frames], PC_RANGE, VOXEL_SIZE)
    c1 = bd["voxel_coords"]
    # visit order of the backbone: ascending linear key
    key = ((c1[:, 0].long() * 41 + c1[:, 1]) * 1504 + c1[:, 2]) * 1504 + c1[:, 3]
    c1 = c1[torch.argsort(key)].contiguous()
    c2, shape2 = ops.strided_outputs(c1, [41, 1504, 1504], 4, 3, 2, 1)
    c2 = c2.clone()
    assert shape2 == [21, 752, 752]
    h2 = ops.build_hash(c2, shape2, 4)
    nbr = ops.subm_table(c2, shape2, 4, 3, h2)
    m = c2.shape[0]
    assert m > 150000, m
    rng = np.random.default_rng(5)
    x = rng.normal(0, 1, (m, 32)).astype(np.float32)
    w = (rng.normal(0, 1, (32, 27, 32)) / np.sqrt(32 * 14)).astype(np.float32)
    b = rng.normal(0, 0.1, 32).astype(np.float32)
    rb = oracle.rulebook_subm(c2.cpu().numpy(), shape2, 3)
    assert np.array_equal(nbr.cpu().numpy(), _table_from_pairs(rb, m))           # the rulebook itself at this size
    y_ref = oracle.spconv_fwd(x, w, b, rb)
    stats = torch.empty(2, 32, device=cuda)
    y = ops.gather_gemm(dev(x, cuda), dev(w, cuda), nbr, bias=dev(b, cuda), stats=stats, algo=ops.ALGO_TCGEN05,
                        tile_masks=ops.tile_tap_masks(nbr))
    err = float(np.abs(y.cpu().numpy() - y_ref).max())
    print(f"stage-2 SubM 32->32: {m} rows, {rb.n_pairs} pairs, max |err| {err:.2e}")
    assert err <= TOL * max(1.0, float(np.abs(y_ref).max()))
    assert np.allclose(stats[0].cpu().numpy(), y_ref.astype(np.float64).sum(0), rtol=1e-4, atol=0.5)
    dy = rng.normal(0, 1, y_ref.shape).astype(np.float32)
    gx, gw, gb = oracle.spconv_bwd(x, w, dy, rb)
    ddy = dev(dy, cuda)
    dxg = ops.gather_gemm(ddy, ops.weight_transpose(dev(w, cuda), flip_taps=True), nbr, algo=ops.ALGO_TCGEN05)
    assert np.abs(dxg.cpu().numpy() - gx).max() <= TOL * max(1.0, float(np.abs(gx).max()))
    dwg, _ = ops.gather_wgrad(dev(x, cuda), ddy, nbr.t().contiguous(), algo=ops.ALGO_TCGEN05, tap_major=True)
    assert np.abs(dwg.cpu().numpy() - gw).max() <= TOL * max(1.0, float(np.abs(gw).max()))
    # ---- dense: 4 x 188 x 188 ----
    torch.manual_seed(8)
    for cin, cout in ((256, 256), (512, 64)):
        conv = DenseConv2d(cin, cout, 3, padding=1, bias=True).to(cuda)
        xi = torch.randn(4, cin, 188, 188, device=cuda) * (torch.rand(4, 1, 188, 188, device=cuda) < 0.3)
        xi.requires_grad_(True)
        yo = conv(DenseMap.from_nchw(xi)).nchw()
        go = torch.randn_like(yo)
        yo.backward(go)
        wd, bdb = conv.weight.detach().double().cpu(), conv.bias.detach().double().cpu()
        for f in (0, 3):
            xr = xi[f:f + 1].detach().double().cpu().requires_grad_(True)
            yr = F.conv2d(xr, wd, bdb, padding=1)
            assert float((yo[f:f + 1].detach().double().cpu() - yr).abs().max()) <= TOL * max(1.0, float(yr.abs().max())), (cin, cout, f)
            yr.backward(go[f:f + 1].double().cpu())
            assert float((xi.grad[f:f + 1].double().cpu() - xr.grad).abs().max()) <= TOL * max(1.0, float(xr.grad.abs().max())), (cin, cout, f)


def test_detector_train_step_matches_cpu_port(cuda):
    """The two arms of the headline ratio, on the same batch with the same weights: CPDHotPathDetector (CUDA) against
    oracle.pipeline.CpuDetector (oracle C for voxelizer / sparse convs, torch CPU for the dense head).

    Loss: 1e-4.  Gradients: ~50 conv + training-mode BatchNorm + ReLU stages are ill-conditioned -- plain fp32 torch
    differs from float64 torch by ~1e-3 in the BEV weight gradients already (measured) -- so an absolute
    tolerance says nothing.  The bound is stated against the reference's OWN sensitivity: the CPU port is run a second
    time with every weight perturbed by a relative 2^-17 (the size of one bf16x3 operand rounding); the CUDA arm must
    sit within a small multiple of the gradient change that perturbation causes, group by group."""
    from cpd_b200 import detector
    from oracle import pipeline
    torch.manual_seed(0)
    det = detector.CPDHotPathDetector().to(cuda).train()
    frames = [synth_scan(20000, 300 + i) for i in range(2)]
    frames1 = [synth_scan(20000, 400 + i) for i in range(2)]
    gt = np.stack([synth_gt_boxes(30, 300 + i) for i in range(2)])
    cpu = pipeline.CpuDetector(det)
    loss_ref = cpu.train_step(frames, frames1, gt)
    ref = cpu.named_grads()
    # the same CPU port with weights perturbed at the 2^-17 level: how far does the REFERENCE move?
    pert = pipeline.CpuDetector(det)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for prm in list(pert.det.parameters()) + list(pert.dense_mods.parameters()):
            prm.mul_(1.0 + (torch.rand(prm.shape, generator=g) * 2 - 1) * 2.0 ** -17)
    loss_pert = pert.train_step(frames, frames1, gt)
    refp = pert.named_grads()
    loss, _ = det(dict(points=[dev(f, cuda) for f in frames], points1=[dev(f, cuda) for f in frames1], gt_boxes=dev(gt, cuda)))
    loss.backward()
    print(f"loss: cuda {float(loss):.6f}  cpu port {loss_ref:.6f}  cpu port with 2^-17 weight noise {loss_pert:.6f}")
    assert abs(float(loss) - loss_ref) <= 1e-4 * max(1.0, abs(loss_ref))
    groups = {}
    for name, prm in det.named_parameters():
        assert prm.grad is not None and name in ref, name
        r = ref[name].double()
        if float(r.norm()) / r.numel() ** 0.5 < 1e-9:          # mathematically zero (a conv bias in front of BatchNorm): noise vs noise
            continue
        key = ".".join(name.split(".")[:2])
        e = groups.setdefault(key, [0.0, 0.0, 0.0])
        e[0] += float((prm.grad.detach().double().cpu() - r).square().sum())
        e[1] += float((refp[name].double() - r).square().sum())
        e[2] += float(r.square().sum())
    worst = 0.0
    for key, (d_cuda, d_pert, nrm) in groups.items():
        rc, rp = (d_cuda / nrm) ** 0.5, (d_pert / nrm) ** 0.5
        worst = max(worst, rc / max(rp, 1e-7))
        print(f"  {key:34s} rel L2: cuda vs port {rc:.2e}   port(2^-17 noise) vs port {rp:.2e}   ratio {rc / max(rp, 1e-7):.2f}")
    assert worst <= 8.0, worst


@pytest.mark.parametrize("m,K", [(1, 27), (127, 27), (128, 27), (5000, 27), (70001, 27), (300, 9), (513, 3), (1000, 32), (777, 1)])
def test_table_permute_and_transpose_match_torch(cuda, m, K):
    """cpd_table_permute == (index_select by the permutation, int32 rows, cpd_tile_tap_masks of the result) and
    cpd_table_transpose == .t().contiguous(): integer work, bit-exact."""
    from cpd_b200 import ops
    g = torch.Generator().manual_seed(m * 31 + K)
    nbr = torch.randint(-1, max(m, 2), (m, K), generator=g, dtype=torch.int32)
    nbr[torch.rand(m, K, generator=g) < 0.5] = -1
    if m > 256:
        nbr[128:256, : max(1, K // 2)] = -1                     # a tile that lacks whole taps
    nbr = nbr.to(cuda)
    perm = torch.randperm(m, generator=g).to(cuda)
    out, rows, masks = ops.table_permute(nbr, perm)
    ref = nbr.index_select(0, perm)
    assert torch.equal(out, ref) and torch.equal(rows, perm.to(torch.int32))
    assert torch.equal(masks, ops.tile_tap_masks(ref))
    t = ops.table_transpose(nbr)
    assert t.shape == (K, m) and torch.equal(t, nbr.t().contiguous())


def test_multi_stage_tta_eval_at_full_width(cuda, oracle):
    """VoxelBackBone8x eval with 3 TTA stages on the real grid: one tower pass over [41, 1504, 6016] (X-concatenated stages),
    strict-`<` decompose per stage (spconv_backbone.py:241-260,332-393) == the oracle pipeline on the same concatenated input."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_tta_cpu import check_tta
    from cpd_b200 import backbone, voxel
    torch.manual_seed(0)
    net = backbone.VoxelBackBone8x(dict(NUM_FILTERS=[16, 32, 64, 128], OUT_FEATURES=128), 5, [1504, 1504, 40])
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.uniform_(-0.2, 0.2); m.running_var.uniform_(0.5, 1.5); m.weight.data.uniform_(0.7, 1.3); m.bias.data.uniform_(-0.2, 0.2)
    net = net.to(cuda).eval()
    stage_frames = [[synth_scan(12000, 30 + i)] for i in range(3)]
    bd = dict(batch_size=1, transform_param=torch.zeros(1, 3, 3, device=cuda))
    for i, frames in enumerate(stage_frames):
        sid = "" if i == 0 else str(i)
        v = voxel.voxelize_batch([torch.from_numpy(f).to(cuda) for f in frames], PC_RANGE, VOXEL_SIZE)
        bd["voxel_features" + sid], bd["voxel_coords" + sid] = v["voxel_features"], v["voxel_coords"]
    with torch.no_grad():
        out = net(bd)
    check_tta(net, out, stage_frames, PC_RANGE, VOXEL_SIZE)
