"""TEST INFRASTRUCTURE ONLY: a CPU stand-in for the kernels behind cpd_b200.ops, built on the oracle, so that the HOST
logic of the package (spconv shim surface, rulebook caching by indice_key, SparseSequential fusion rules, the reference's
own module files running on the shim) can be exercised on a box without a GPU.  Never imported by cpd_b200/."""
import contextlib

import numpy as np
import torch

from cpd_b200 import ops, sparse
from oracle import oracle as O


def _fwd_table(rb, m_out):
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


def _gather_gemm(x, w, nbr, bias=None, scale=None, shift=None, residual=None, relu=False, stats=None, algo=0, out=None,
                 x_split=None, tile_masks=None, out_rows=None):
    cout, cin = w.shape[0], w.shape[-1]
    w = w.detach().reshape(cout, -1, cin).double()
    xd = x.detach().double()
    y = torch.zeros((nbr.shape[0], cout), dtype=torch.float64)
    for k in range(w.shape[1]):
        idx = nbr[:, k].long()
        ok = idx >= 0
        if bool(ok.any()):
            y[ok] += xd[idx[ok]] @ w[:, k, :].t()
    if bias is not None:
        y += bias.detach().double()
    if stats is not None:
        stats.copy_(torch.stack([y.sum(0), (y * y).sum(0)]).float())
    if scale is not None:
        y = y * scale.double() + shift.double()
    if out_rows is not None:
        z = torch.empty_like(y)
        z[out_rows.long()] = y
        y = z
    if residual is not None:
        y = y + residual.double()
    if relu:
        y = torch.relu(y)
    return y.float()


def _weight_transpose(w, flip_taps=False):
    cout, cin = w.shape[0], w.shape[-1]
    w = w.detach().reshape(cout, -1, cin)
    if flip_taps:
        w = w.flip(1)
    return w.permute(2, 1, 0).contiguous()


def _gather_wgrad(x, dy, nbr, want_bias=False, algo=0, tap_major=False, x_split=None, dy_split=None):
    if tap_major:
        nbr = nbr.t()
    xd, dyd = x.detach().double(), dy.detach().double()
    K = nbr.shape[1]
    dw = torch.zeros((dy.shape[1], K, x.shape[1]), dtype=torch.float64)
    for k in range(K):
        idx = nbr[:, k].long()
        ok = idx >= 0
        if bool(ok.any()):
            dw[:, k, :] = dyd[ok].t() @ xd[idx[ok]]
    return dw.float(), (dyd.sum(0).float() if want_bias else None)


def _dense(feat, coords, batch, shape, channels_last=False):
    d = torch.from_numpy(O.dense(feat.detach().numpy(), coords.numpy(), batch, list(shape)))
    if channels_last:
        n, c, dd, h, w = d.shape
        return d.reshape(n, c * dd, h, w).permute(0, 2, 3, 1).contiguous()
    return d


@contextlib.contextmanager
def cpu_ops():
    saved = {k: getattr(ops, k) for k in ("build_hash", "subm_table", "strided_outputs", "strided_tables", "gather_gemm", "tile_tap_masks",
                                          "split_rows", "sparse_to_dense", "gather_wgrad", "weight_transpose")}
    saved_min = sparse.Rulebook.SORT_MIN_ROWS

    def strided_outputs(coords, shape, batch, ksize, stride, padding):
        rb = O.rulebook_strided(coords.numpy(), list(shape), ksize, stride, padding)
        return torch.from_numpy(rb.out_coords), rb.out_shape

    def strided_tables(in_coords, in_shape, in_hash, out_coords, out_shape, out_hash, batch, ksize, stride, padding, want_bwd=True):
        rb = O.rulebook_strided(in_coords.numpy(), list(in_shape), ksize, stride, padding)
        assert np.array_equal(rb.out_coords, out_coords.numpy())
        return torch.from_numpy(_fwd_table(rb, rb.m_out)), (torch.from_numpy(_bwd_table(rb, rb.m_in)) if want_bwd else None)

    ops.build_hash = lambda coords, shape, batch: torch.zeros(1)
    ops.subm_table = lambda coords, shape, batch, ksize, h: torch.from_numpy(
        _fwd_table(O.rulebook_subm(coords.numpy(), list(shape), ksize), coords.shape[0]))
    ops.strided_outputs, ops.strided_tables = strided_outputs, strided_tables
    ops.gather_gemm = _gather_gemm
    ops.gather_wgrad, ops.weight_transpose = _gather_wgrad, _weight_transpose
    ops.tile_tap_masks = lambda nbr: torch.zeros(((nbr.shape[0] + 127) // 128,), dtype=torch.int32)
    ops.split_rows = lambda x, colsum=False: (None, None) if colsum else None
    ops.sparse_to_dense = _dense
    sparse.Rulebook.SORT_MIN_ROWS = 1 << 60
    try:
        yield
    finally:
        for k, v in saved.items():
            setattr(ops, k, v)
        sparse.Rulebook.SORT_MIN_ROWS = saved_min
