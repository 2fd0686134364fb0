"""GPU parity of the table-free dense convolutions (cpd_conv2d_fwd / _dgrad / _wgrad, cpd_convt2d_fwd: tiled TMA loads of
shifted pixel patches feeding tcgen05) against float64 torch -- the reference runs these layers as nn.Conv2d /
nn.ConvTranspose2d (base_bev_backbone.py:31-59, center_head.py:11-45,73-80).  Tolerance 1e-4 of the output scale."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _close(a, b, tol=TOL):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    err = float((a - b).abs().max())
    assert err <= tol * max(1.0, float(b.abs().max())), (err, float(b.abs().max()))


def _rows(x):            # NCHW -> NHWC rows
    n, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(n * h * w, c).contiguous()


def _nchw(rows, n, h, w):
    return rows.view(n, h, w, -1).permute(0, 3, 1, 2)


@pytest.mark.parametrize("cin,cout,k,n,h,w", [(64, 64, 3, 2, 47, 50), (256, 128, 3, 1, 33, 21), (128, 256, 3, 3, 16, 8), (512, 64, 3, 1, 40, 37),
                                              (64, 16, 3, 2, 9, 130), (128, 256, 1, 2, 31, 45), (256, 256, 3, 1, 94, 94)])
def test_conv2d_fwd_epilogue_dgrad_wgrad(cuda, cin, cout, k, n, h, w):
    from cpd_b200 import ops
    torch.manual_seed(cin + cout + k)
    pad = k // 2
    assert ops.conv2d_ok(cin, k, cout)
    x = torch.randn(n, cin, h, w, device=cuda) * (torch.rand(n, 1, h, w, device=cuda) < 0.6)
    wt = torch.randn(cout, cin, k, k, device=cuda) / (cin * k * k) ** 0.5
    bias = torch.randn(cout, device=cuda) * 0.1
    w_kc = wt.permute(0, 2, 3, 1).reshape(cout, k * k, cin).contiguous()
    xr = _rows(x)
    xs = ops.split_rows(xr)
    ref = F.conv2d(x.double().cpu(), wt.double().cpu(), bias.double().cpu(), padding=pad)
    y = ops.conv2d_fwd(xs, n, h, w, w_kc, k, pad, bias=bias)
    _close(_nchw(y, n, h, w), ref)
    # fused epilogue: folded-BN affine, residual, ReLU, BatchNorm statistics of the pre-affine output
    scale, shift = torch.rand(cout, device=cuda) + 0.5, torch.randn(cout, device=cuda) * 0.2
    res = torch.randn(n * h * w, cout, device=cuda)
    stats = torch.full((2, cout), 3.0, device=cuda)
    yf = ops.conv2d_fwd(xs, n, h, w, w_kc, k, pad, bias=bias, scale=scale, shift=shift, residual=res, relu=True, stats=stats)
    ref_rows = _rows(ref)
    _close(yf, torch.relu(ref_rows * scale.double().cpu() + shift.double().cpu() + res.double().cpu()))
    assert torch.allclose(stats[0].double().cpu(), ref_rows.sum(0), rtol=1e-4, atol=1e-3 * (n * h * w) ** 0.5)
    assert torch.allclose(stats[1].double().cpu(), ref_rows.square().sum(0), rtol=2e-4, atol=1e-2)
    # input / weight gradients
    dy = torch.randn(n, cout, h, w, device=cuda)
    xg = x.double().cpu().requires_grad_(True)
    wg = wt.double().cpu().requires_grad_(True)
    F.conv2d(xg, wg, None, padding=pad).backward(dy.double().cpu())
    dys = ops.split_rows(_rows(dy))
    if ops.conv2d_ok(cout, k, cin):
        dx = ops.conv2d_dgrad(dys, n, h, w, cin, w_kc, k, pad)
        _close(_nchw(dx, n, h, w), xg.grad)
    dw = ops.conv2d_wgrad(xs, dys, n, h, w, k, pad)                   # (cout, k*k, cin)
    _close(dw.view(cout, k, k, cin).permute(0, 3, 1, 2), wg.grad)


def test_conv2d_unpadded_and_rectangular_tiles(cuda):
    """pad = 0 (output smaller than the input) and images smaller than one 16 x 8 tile."""
    from cpd_b200 import ops
    torch.manual_seed(2)
    for (n, h, w, k, pad) in ((2, 20, 19, 3, 0), (1, 5, 3, 3, 1), (4, 8, 16, 1, 0)):
        cin, cout = 64, 32
        x = torch.randn(n, cin, h, w, device=cuda)
        wt = torch.randn(cout, cin, k, k, device=cuda) * 0.05
        y = ops.conv2d_fwd(ops.split_rows(_rows(x)), n, h, w, wt.permute(0, 2, 3, 1).reshape(cout, k * k, cin).contiguous(), k, pad)
        ref = F.conv2d(x.double().cpu(), wt.double().cpu(), None, padding=pad)
        _close(_nchw(y, n, ref.shape[2], ref.shape[3]), ref)


@pytest.mark.parametrize("s,cin,cout", [(2, 256, 256), (2, 64, 128), (1, 128, 256), (4, 64, 16)])
def test_convt2d_fwd_and_module_grads(cuda, s, cin, cout):
    from cpd_b200 import bev, ops
    torch.manual_seed(s + cin)
    n, h, w = 2, 13, 22
    m = bev.DenseConvTranspose2d(cin, cout, s, stride=s).to(cuda)
    x = torch.randn(n, cin, h, w, device=cuda, requires_grad=True)
    ref_x = x.detach().double().cpu().requires_grad_(True)
    ref_w = m.weight.detach().double().cpu().requires_grad_(True)
    ref = F.conv_transpose2d(ref_x, ref_w, stride=s)
    stats = torch.empty(2, cout, device=cuda)
    y = ops.convt2d_fwd(ops.split_rows(_rows(x.detach())), n, h, w, m.weight.detach(), s, stats=stats)
    _close(_nchw(y, n, h * s, w * s), ref)
    assert torch.allclose(stats[0].double().cpu(), _rows(ref.detach()).sum(0), rtol=1e-4, atol=1e-3 * (n * h * w * s * s) ** 0.5)
    out = m(bev.DenseMap.from_nchw(x)).nchw()                       # module path with autograd (pixel-unshuffle backward)
    _close(out, ref)
    dy = torch.randn_like(out)
    out.backward(dy)
    ref.backward(dy.double().cpu())
    _close(x.grad, ref_x.grad)
    _close(m.weight.grad, ref_w.grad)


def test_deblock_convt_bn_train_fused(cuda):
    """ConvTranspose + BatchNorm2d(train) (+ ReLU) as DenseSequential runs it (statistics from the GEMM epilogues).
    Gradients are compared WITHOUT the ReLU: with it, the few outputs within rounding distance of zero flip their gate
    and each flip moves dx by O(|W|) -- a property of ReLU, not of the kernels (the ReLU forward is checked)."""
    from cpd_b200 import bev
    torch.manual_seed(9)
    for s in (1, 2):
        for relu in (True, False):
            mods = [bev.DenseConvTranspose2d(128, 256, s, stride=s), torch.nn.BatchNorm2d(256, eps=1e-3, momentum=0.01)] + ([torch.nn.ReLU()] if relu else [])
            seq = bev.DenseSequential(*mods).to(cuda).train()
            rmods = [torch.nn.ConvTranspose2d(128, 256, s, stride=s, bias=False), torch.nn.BatchNorm2d(256, eps=1e-3, momentum=0.01)] + \
                    ([torch.nn.ReLU()] if relu else [])
            ref = torch.nn.Sequential(*rmods).double().train()
            ref[0].weight.data.copy_(seq[0].weight.data.double().cpu())
            x = torch.randn(2, 128, 17, 12, device=cuda, requires_grad=True)
            xr = x.detach().double().cpu().requires_grad_(True)
            y, yr = seq(bev.DenseMap.from_nchw(x)).nchw(), ref(xr)
            _close(y, yr)
            _close(seq[1].running_var, ref[1].running_var, 1e-5)
            if relu:
                continue
            dy = torch.randn_like(y)
            y.backward(dy)
            yr.backward(dy.double().cpu())
            _close(x.grad, xr.grad)
            _close(seq[0].weight.grad, ref[0].weight.grad)
            _close(seq[1].weight.grad, ref[1].weight.grad)
            _close(seq[1].bias.grad, ref[1].bias.grad)
