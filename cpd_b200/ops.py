"""Functional torch-level wrappers over the C ABI (device tensors in, device tensors out).

torch is plumbing here: it owns device memory and the stream; all arithmetic happens in
libcpd_b200.so.  Every function requires CUDA tensors and raises otherwise.
"""
import ctypes as C

import torch

from . import _lib


# bench.py sets this to a list to time launches with CUDA events on the launching stream: entries are
# (start_event, end_event, meta dict) -- gather-GEMM / wgrad / conv2d (meta holds the table or the pair count), voxelizer, rulebooks.
PROFILE = None


def _prof_begin():
    if PROFILE is None:
        return None
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    return e0


def _prof_end(e0, **meta):
    if e0 is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        nbr = meta.pop("nbr", None)
        if nbr is not None:                  # keep the PAIR COUNT (a device scalar, computed after the timed bracket), not the table:
            p = getattr(nbr, "_cpd_pairs", None)      # holding every table of every step would pin gigabytes in the allocator
            if p is None:
                p = (nbr >= 0).sum()
                nbr._cpd_pairs = p
            meta["pairs"] = p
        PROFILE.append((e0, e1, meta))


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _i32x3(v):
    if isinstance(v, int):
        v = (v, v, v)
    v = [int(x) for x in v]
    assert len(v) == 3
    return (C.c_int32 * 3)(*v)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.CpdError("cpd_b200 ops need CUDA tensors (there is no CPU fallback)")


def _f32c(t):
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.contiguous().float()


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------
# voxelizer
# ------------------------------------------------------------------------------------
def voxelize(points, frame_offsets, pc_range, voxel_size, max_pts=5, max_voxels=1000000,
             want_voxels=True, want_mean=True, cap_rows=None, sync=True):
    """Batched Point2VoxelCPU3d + collate + MeanVFE (see include/cpd_b200.h).

    points (N,C) float32 cuda, frames concatenated; frame_offsets: python ints, len batch+1.
    Returns dict(voxels (M,max_pts,C) | None, coords (M,4) int32 [b,z,y,x], num (M,) int32,
    mean (M,C) | None, counts (batch+1,) int32 device).  With sync=True tensors are sliced to
    the true M (one D2H read of the counter, as any data-dependent shape needs); with
    sync=False they keep cap_rows rows and `counts` stays on the device.
    """
    _need_cuda(points)
    L = _lib.lib()
    pts = _f32c(points)
    n, c = pts.shape
    batch = len(frame_offsets) - 1
    dev = pts.device
    if cap_rows is None:
        cap_rows = int(min(n, batch * int(max_voxels)))
    cap_rows = max(int(cap_rows), 1)
    offs = (C.c_int64 * (batch + 1))(*[int(o) for o in frame_offsets])
    rng = (C.c_float * 6)(*[float(v) for v in pc_range])
    vs = (C.c_float * 3)(*[float(v) for v in voxel_size])
    voxels = torch.empty((cap_rows, max_pts, c), dtype=torch.float32, device=dev) if want_voxels else None
    mean = torch.empty((cap_rows, c), dtype=torch.float32, device=dev) if want_mean else None
    coords = torch.empty((cap_rows, 4), dtype=torch.int32, device=dev)
    num = torch.empty((cap_rows,), dtype=torch.int32, device=dev)
    counts = torch.empty((batch + 1,), dtype=torch.int32, device=dev)
    wsb = L.cpd_voxelize_workspace_bytes(n, batch, max_pts, cap_rows)
    ws = _ws(wsb, dev)
    e0 = _prof_begin()
    _lib.check(L.cpd_voxelize(_ptr(pts), n, c, offs, batch, rng, vs, int(max_pts), int(max_voxels), cap_rows,
                              _ptr(voxels), _ptr(coords), _ptr(num), _ptr(mean), _ptr(counts), _ptr(ws), wsb,
                              _stream()), "cpd_voxelize")
    out = dict(voxels=voxels, coords=coords, num=num, mean=mean, counts=counts)
    _prof_end(e0, kind="voxelize", n=n, c=c, counts=counts, batch=batch, max_pts=max_pts, want_voxels=want_voxels, want_mean=want_mean)
    if sync:
        m = int(counts[batch].item())
        if m > cap_rows:
            raise _lib.CpdError(f"cpd_voxelize: {m} voxels exceed cap_rows={cap_rows}")
        for k in ("voxels", "coords", "num", "mean"):
            if out[k] is not None:
                out[k] = out[k][:m]
    return out


# ------------------------------------------------------------------------------------
# rulebooks
# ------------------------------------------------------------------------------------
def build_hash(coords, shape, batch):
    _need_cuda(coords)
    L = _lib.lib()
    m = coords.shape[0]
    nb = L.cpd_coord_hash_bytes(m)
    h = torch.empty(nb, dtype=torch.uint8, device=coords.device)
    _lib.check(L.cpd_coord_hash_build(_ptr(coords), m, _i32x3(shape), int(batch), _ptr(h), nb, _stream()),
               "cpd_coord_hash_build")
    return h


def subm_table(coords, shape, batch, ksize, hash_buf):
    _need_cuda(coords, hash_buf)
    L = _lib.lib()
    ks = _i32x3(ksize)
    K = ks[0] * ks[1] * ks[2]
    m = coords.shape[0]
    nbr = torch.empty((m, K), dtype=torch.int32, device=coords.device)
    e0 = _prof_begin()
    _lib.check(L.cpd_rulebook_subm(_ptr(coords), m, _i32x3(shape), int(batch), ks, _ptr(hash_buf), hash_buf.numel(),
                                   _ptr(nbr), _stream()), "cpd_rulebook_subm")
    _prof_end(e0, kind="rulebook_subm", m_in=m, m_out=m, K=K, nbr=nbr)
    return nbr


def strided_outputs(coords, shape, batch, ksize, stride, padding):
    """-> (out_coords (M_out,4) int32 sorted by linear key, out_shape [D,H,W])."""
    _need_cuda(coords)
    L = _lib.lib()
    ks, st, pd, sh = _i32x3(ksize), _i32x3(stride), _i32x3(padding), _i32x3(shape)
    K = ks[0] * ks[1] * ks[2]
    m = coords.shape[0]
    dev = coords.device
    wsb = L.cpd_rulebook_strided_workspace_bytes(sh, int(batch), ks, st, pd)
    if wsb == 0:
        raise _lib.CpdError("cpd_rulebook_strided: empty output shape")
    ws = _ws(wsb, dev)
    oshape = (C.c_int32 * 3)()
    # an input site feeds at most prod(ceil(k / s)) outputs (8 for the k=3, s=2 convs of the CPD towers, K for stride 1)
    fan = 1
    for k_, s_ in zip(ks, st):
        fan *= -(-k_ // s_)
    cap = max(m * min(K, fan), 1)
    ocoords = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    n_out = torch.empty((1,), dtype=torch.int32, device=dev)
    e0 = _prof_begin()
    _lib.check(L.cpd_rulebook_strided_outputs(_ptr(coords), m, sh, int(batch), ks, st, pd, oshape, cap, _ptr(ocoords),
                                              _ptr(n_out), _ptr(ws), wsb, _stream()), "cpd_rulebook_strided_outputs")
    _prof_end(e0, kind="rulebook_strided_outputs", m_in=m, n_out=n_out, K=K, cells=int(wsb) * 8)
    mo = int(n_out.item())
    if mo > cap:
        raise _lib.CpdError("cpd_rulebook_strided_outputs: output capacity exceeded")
    out = ocoords[:mo]
    if mo * 2 < cap:                 # do not pin the oversized buffer for the rulebook's lifetime
        out = out.clone()
    return out, [int(oshape[0]), int(oshape[1]), int(oshape[2])]


def strided_tables(in_coords, in_shape, in_hash, out_coords, out_shape, out_hash, batch, ksize, stride, padding,
                   want_bwd=True):
    _need_cuda(in_coords, out_coords)
    L = _lib.lib()
    ks, st, pd = _i32x3(ksize), _i32x3(stride), _i32x3(padding)
    K = ks[0] * ks[1] * ks[2]
    dev = in_coords.device
    m_in, m_out = in_coords.shape[0], out_coords.shape[0]
    fwd = torch.empty((m_out, K), dtype=torch.int32, device=dev)
    bwd = torch.empty((m_in, K), dtype=torch.int32, device=dev) if want_bwd else None
    e0 = _prof_begin()
    _lib.check(L.cpd_rulebook_strided_tables(_ptr(in_coords), m_in, _i32x3(in_shape), _ptr(in_hash), in_hash.numel(),
                                             _ptr(out_coords), m_out, _i32x3(out_shape),
                                             _ptr(out_hash) if want_bwd else None,
                                             out_hash.numel() if want_bwd else 0, int(batch), ks, st, pd,
                                             _ptr(fwd), _ptr(bwd), _stream()), "cpd_rulebook_strided_tables")
    _prof_end(e0, kind="rulebook_strided_tables", m_in=m_in, m_out=m_out, K=K, nbr=fwd, both=want_bwd)
    return fwd, bwd


def conv2d_table(n, h, w, kh, kw, stride, pad, transposed, ho, wo, device):
    L = _lib.lib()
    nbr = torch.empty((n * ho * wo, kh * kw), dtype=torch.int32, device=device)
    _lib.check(L.cpd_conv2d_table(n, h, w, kh, kw, stride, pad, int(bool(transposed)), ho, wo, _ptr(nbr), _stream()),
               "cpd_conv2d_table")
    return nbr


# ------------------------------------------------------------------------------------
# gather-GEMM family
# ------------------------------------------------------------------------------------
ALGO_AUTO, ALGO_SIMT, ALGO_TCGEN05 = 0, 1, 2



def tc_gemm_ok(cin, K, cout):
    """Shapes the tcgen05 gather-GEMM takes under ALGO_AUTO (mirrors cpd_gather_gemm's dispatch)."""
    return cin % 8 == 0 and cin >= 8 and K <= 27 and K * cin <= 8192 and (cout in (16, 32, 64, 128) or (cout % 256 == 0 and 0 < cout <= 2048))


def tc_wgrad_ok(cin, K, cout):
    """Shapes the row-stationary tcgen05 weight-gradient takes (mirrors gather_wgrad_rows_supported)."""
    return cin >= 8 and cin % 8 == 0 and ((cin <= 256 and 256 % cin == 0) or cin % 256 == 0) and cout >= 8 and cout % 8 == 0 and K <= 64


def split_rows(x, colsum=False):
    """Split-row image of a row matrix (m, c), c % 8 == 0: (m, 2, c) bf16, row = [hi(c) | lo(c)] -- the operand
    format of the tcgen05 kernels (include/cpd_b200.h).  Returns None when the layout does not apply.
    colsum=True (c a power of two): returns (image, per-channel column sums of x) from the same pass."""
    _need_cuda(x)
    if x.dim() != 2 or x.shape[1] % 8 or x.shape[1] < 8:
        return (None, None) if colsum else None
    L = _lib.lib()
    x = _f32c(x)
    c = x.shape[1]
    xs = torch.empty((x.shape[0], 2, c), dtype=torch.bfloat16, device=x.device)
    cs = torch.empty((c,), dtype=torch.float32, device=x.device) if colsum and c <= 2048 and c & (c - 1) == 0 else None
    _lib.check(L.cpd_split_rows(_ptr(x), x.shape[0], c, _ptr(xs), _ptr(cs), _stream()), "cpd_split_rows")
    return (xs, cs) if colsum else xs


def tile_tap_masks(nbr):
    """Per 128-row tile of a (m, K <= 32) neighbour table: bitmask of the taps that have a neighbour in the tile."""
    _need_cuda(nbr)
    L = _lib.lib()
    m, K = nbr.shape
    masks = torch.empty(((m + 127) // 128,), dtype=torch.int32, device=nbr.device)
    _lib.check(L.cpd_tile_tap_masks(_ptr(nbr), m, K, _ptr(masks), _stream()), "cpd_tile_tap_masks")
    return masks


def tap_block_keys(nbr, taps_per_block):
    """Per table row: bitmap of the k-blocks (groups of taps_per_block taps) it has neighbours in -- the sort key of
    Rulebook.sorted_table."""
    _need_cuda(nbr)
    L = _lib.lib()
    m, K = nbr.shape
    keys = torch.empty((m,), dtype=torch.int32, device=nbr.device)
    _lib.check(L.cpd_tap_block_keys(_ptr(nbr), m, K, int(taps_per_block), _ptr(keys), _stream()), "cpd_tap_block_keys")
    return keys


def table_permute(nbr, perm, want_masks=True):
    """(nbr[perm], perm as int32, tile tap masks of the permuted table) in one pass over the table."""
    _need_cuda(nbr, perm)
    L = _lib.lib()
    m, K = nbr.shape
    assert perm.dtype == torch.int64 and perm.numel() == m and perm.is_contiguous() and nbr.is_contiguous()
    out = torch.empty_like(nbr)
    rows = torch.empty((m,), dtype=torch.int32, device=nbr.device)
    masks = torch.empty(((m + 127) // 128,), dtype=torch.int32, device=nbr.device) if want_masks else None
    _lib.check(L.cpd_table_permute(_ptr(nbr), m, K, _ptr(perm), _ptr(out), _ptr(rows), _ptr(masks), _stream()), "cpd_table_permute")
    return out, rows, masks


def table_transpose(nbr):
    """(m, K) neighbour table -> tap-major (K, m) (the layout cpd_gather_wgrad reads)."""
    _need_cuda(nbr)
    L = _lib.lib()
    m, K = nbr.shape
    assert nbr.dtype == torch.int32 and nbr.is_contiguous()
    out = torch.empty((K, m), dtype=torch.int32, device=nbr.device)
    _lib.check(L.cpd_table_transpose(_ptr(nbr), m, K, _ptr(out), _stream()), "cpd_table_transpose")
    return out


def gather_gemm(x, w, nbr, bias=None, scale=None, shift=None, residual=None, relu=False, stats=None,
                algo=ALGO_AUTO, out=None, x_split=None, tile_masks=None, out_rows=None):
    """y[o] = epi(sum_k W[:,k,:] x[nbr[o,k]]);  w is (cout, K, cin) (any (cout, ..., cin) view).
    x_split: split_rows(x) if the caller already has it (shared between forward and weight-gradient)."""
    _need_cuda(x, w, nbr)
    L = _lib.lib()
    x = _f32c(x)
    cout, cin = w.shape[0], w.shape[-1]
    w = _f32c(w).view(cout, -1, cin)
    K = w.shape[1]
    m_out = nbr.shape[0]
    assert nbr.shape[1] == K and nbr.dtype == torch.int32 and nbr.is_contiguous()
    assert x.shape[1] == cin
    assert x_split is None or (x_split.shape[0] == x.shape[0] and x_split.shape[2] == cin and x_split.is_contiguous())
    y = out if out is not None else torch.empty((m_out, cout), dtype=torch.float32, device=x.device)
    residual = _f32c(residual) if residual is not None else None
    wsb = L.cpd_gather_gemm_workspace_bytes(x.shape[0], m_out, cin, K, cout, algo, int(x_split is not None))
    ws = _ws(wsb, x.device) if wsb else None
    e0 = _prof_begin()
    assert tile_masks is None or tile_masks.numel() == (m_out + 127) // 128
    assert out_rows is None or (out_rows.dtype == torch.int32 and out_rows.numel() == m_out and out_rows.is_contiguous())
    _lib.check(L.cpd_gather_gemm(_ptr(x), _ptr(x_split), x.shape[0], cin, _ptr(w), K, cout, _ptr(nbr), _ptr(tile_masks), _ptr(out_rows),
                                 m_out, _ptr(bias),
                                 _ptr(scale), _ptr(shift), _ptr(residual), int(bool(relu)), _ptr(stats), _ptr(y), int(algo),
                                 _ptr(ws), wsb, _stream()), "cpd_gather_gemm")
    _prof_end(e0, kind="gather_gemm", m_in=x.shape[0], m_out=m_out, cin=cin, cout=cout, K=K, nbr=nbr, residual=residual is not None)
    return y


def gather_wgrad(x, dy, nbr, want_bias=False, algo=ALGO_AUTO, tap_major=False, x_split=None, dy_split=None):
    """dw (cout, K, cin), dbias (cout,) | None.  tap_major: nbr is the transposed (K, m_out) table.
    x_split / dy_split: split_rows() images the caller already has."""
    _need_cuda(x, dy, nbr)
    L = _lib.lib()
    x, dy = _f32c(x), _f32c(dy)
    cin, cout, K = x.shape[1], dy.shape[1], nbr.shape[0 if tap_major else 1]
    assert nbr.is_contiguous() and nbr.shape[1 if tap_major else 0] == dy.shape[0]
    dw = torch.empty((cout, K, cin), dtype=torch.float32, device=x.device)
    db = torch.empty((cout,), dtype=torch.float32, device=x.device) if want_bias else None
    wsb = 0 if algo == ALGO_SIMT else L.cpd_gather_wgrad_workspace_bytes(x.shape[0], dy.shape[0], cin, K, cout,
                                                                          int(x_split is not None), int(dy_split is not None))
    ws = _ws(wsb, x.device) if wsb else None
    e0 = _prof_begin()
    _lib.check(L.cpd_gather_wgrad(_ptr(x), _ptr(x_split), x.shape[0], cin, _ptr(dy), _ptr(dy_split), dy.shape[0], cout, _ptr(nbr),
                                  int(bool(tap_major)), K, _ptr(dw), _ptr(db), int(algo), _ptr(ws), wsb, _stream()),
               "cpd_gather_wgrad")
    _prof_end(e0, kind="gather_wgrad", m_in=x.shape[0], m_out=dy.shape[0], cin=cin, cout=cout, K=K, nbr=nbr, residual=False)
    return dw, db


# ------------------------------------------------------------------------------------
# dense stride-1 convolutions on the TMA-fed tcgen05 kernel (no neighbour table)
# ------------------------------------------------------------------------------------
_CONV2D_OK = {}


def conv2d_ok(cin, k, cout):
    """Shapes cpd_conv2d_fwd takes (cin % 64 == 0, cout in {16, 32, 64, 128, 256 j})."""
    key = (int(cin), int(k), int(cout))
    if key not in _CONV2D_OK:
        _CONV2D_OK[key] = bool(_lib.lib().cpd_conv2d_supported(key[0], key[1], key[1], key[2]))
    return _CONV2D_OK[key]


def conv2d_fwd(x_split, n, h, w, weight, k, pad, bias=None, scale=None, shift=None, residual=None, relu=False, stats=None):
    """y (n*ho*wo, cout) = epi(conv(x, weight)); x_split: split-row image of the NHWC rows (n*h*w, 2, cin);
    weight (cout, k*k, cin)."""
    _need_cuda(x_split, weight)
    L = _lib.lib()
    cout, cin = weight.shape[0], weight.shape[-1]
    weight = _f32c(weight).view(cout, -1, cin)
    assert weight.shape[1] == k * k and x_split.shape == (n * h * w, 2, cin) and x_split.is_contiguous()
    ho, wo = h + 2 * pad - k + 1, w + 2 * pad - k + 1
    y = torch.empty((n * ho * wo, cout), dtype=torch.float32, device=weight.device)
    residual = _f32c(residual) if residual is not None else None
    wsb = L.cpd_conv2d_workspace_bytes(cin, k, k, cout)
    ws = _ws(wsb, weight.device)
    e0 = _prof_begin()
    _lib.check(L.cpd_conv2d_fwd(_ptr(x_split), n, h, w, cin, _ptr(weight), k, k, pad, cout, _ptr(bias), _ptr(scale), _ptr(shift),
                                _ptr(residual), int(bool(relu)), _ptr(stats), _ptr(y), _ptr(ws), wsb, _stream()), "cpd_conv2d_fwd")
    _prof_end(e0, kind="conv2d", m_in=n * h * w, m_out=n * ho * wo, cin=cin, cout=cout, K=k * k, P=n * ho * wo * k * k, residual=residual is not None)
    return y


def conv2d_dgrad(dy_split, n, h, w, cin, weight, k, pad):
    """dx (n*h*w, cin) of conv2d_fwd; dy_split: split-row image of dy (n*ho*wo, 2, cout); weight (cout, k*k, cin)."""
    _need_cuda(dy_split, weight)
    L = _lib.lib()
    cout = weight.shape[0]
    weight = _f32c(weight).view(cout, -1, cin)
    ho, wo = h + 2 * pad - k + 1, w + 2 * pad - k + 1
    assert dy_split.shape == (n * ho * wo, 2, cout) and dy_split.is_contiguous()
    dx = torch.empty((n * h * w, cin), dtype=torch.float32, device=weight.device)
    wsb = L.cpd_conv2d_dgrad_workspace_bytes(cin, k, k, cout)
    ws = _ws(wsb, weight.device)
    e0 = _prof_begin()
    _lib.check(L.cpd_conv2d_dgrad(_ptr(dy_split), n, h, w, cin, _ptr(weight), k, k, pad, cout, _ptr(dx), _ptr(ws), wsb, _stream()),
               "cpd_conv2d_dgrad")
    _prof_end(e0, kind="conv2d", m_in=n * ho * wo, m_out=n * h * w, cin=cout, cout=cin, K=k * k, P=n * h * w * k * k, residual=False)
    return dx


def conv2d_wgrad(x_split, dy_split, n, h, w, k, pad):
    """dw (cout, k*k, cin) of conv2d_fwd from the two split-row images (pixel table built in the call's workspace)."""
    _need_cuda(x_split, dy_split)
    L = _lib.lib()
    cin, cout = x_split.shape[2], dy_split.shape[2]
    dw = torch.empty((cout, k * k, cin), dtype=torch.float32, device=x_split.device)
    wsb = L.cpd_conv2d_wgrad_workspace_bytes(n, h, w, k, k, pad)
    ws = _ws(wsb, x_split.device)
    _lib.check(L.cpd_conv2d_wgrad(_ptr(x_split), _ptr(dy_split), n, h, w, cin, k, k, pad, cout, _ptr(dw), _ptr(ws), wsb, _stream()),
               "cpd_conv2d_wgrad")
    return dw


def convt2d_fwd(x_split, n, h, w, weight, s, scale=None, shift=None, relu=False, stats=None):
    """nn.ConvTranspose2d with kernel == stride == s: y (n*h*s*w*s, cout); weight in the torch layout (cin, cout, s, s)."""
    _need_cuda(x_split, weight)
    L = _lib.lib()
    cin, cout = weight.shape[0], weight.shape[1]
    weight = _f32c(weight)
    assert x_split.shape == (n * h * w, 2, cin) and x_split.is_contiguous() and weight.shape[2] == weight.shape[3] == s
    y = torch.empty((n * h * s * w * s, cout), dtype=torch.float32, device=weight.device)
    wsb = L.cpd_convt2d_workspace_bytes(cin, s, cout)
    ws = _ws(wsb, weight.device)
    e0 = _prof_begin()
    _lib.check(L.cpd_convt2d_fwd(_ptr(x_split), n, h, w, cin, _ptr(weight), s, cout, _ptr(scale), _ptr(shift), int(bool(relu)), _ptr(stats),
                                 _ptr(y), _ptr(ws), wsb, _stream()), "cpd_convt2d_fwd")
    _prof_end(e0, kind="convt2d", m_in=n * h * w, m_out=n * h * s * w * s, cin=cin, cout=cout, K=1, P=n * h * s * w * s, residual=False)
    return y


def weight_transpose(w, flip_taps=False):
    """(cout, K, cin) -> (cin, K, cout), optionally reversing taps."""
    _need_cuda(w)
    L = _lib.lib()
    cout, cin = w.shape[0], w.shape[-1]
    w = _f32c(w).view(cout, -1, cin)
    K = w.shape[1]
    wt = torch.empty((cin, K, cout), dtype=torch.float32, device=w.device)
    _lib.check(L.cpd_weight_transpose(_ptr(w), cout, K, cin, int(bool(flip_taps)), _ptr(wt), _stream()),
               "cpd_weight_transpose")
    return wt


def sparse_to_dense(feat, coords, batch, shape, channels_last=False):
    _need_cuda(feat, coords)
    L = _lib.lib()
    feat = _f32c(feat)
    m, c = feat.shape
    d, h, w = [int(v) for v in shape]
    if channels_last:
        out = torch.empty((batch, h, w, c * d), dtype=torch.float32, device=feat.device)
    else:
        out = torch.empty((batch, c, d, h, w), dtype=torch.float32, device=feat.device)
    _lib.check(L.cpd_sparse_to_dense(_ptr(feat), _ptr(coords), m, c, int(batch), _i32x3(shape), int(bool(channels_last)),
                                     _ptr(out), _stream()), "cpd_sparse_to_dense")
    return out


def sparse_to_dense_bwd(dout, coords, c, batch, shape, channels_last=False):
    _need_cuda(dout, coords)
    L = _lib.lib()
    dout = _f32c(dout)
    m = coords.shape[0]
    dfeat = torch.empty((m, c), dtype=torch.float32, device=dout.device)
    _lib.check(L.cpd_sparse_to_dense_bwd(_ptr(dout), _ptr(coords), m, c, int(batch), _i32x3(shape),
                                         int(bool(channels_last)), _ptr(dfeat), _stream()), "cpd_sparse_to_dense_bwd")
    return dfeat


# ------------------------------------------------------------------------------------
# iou3d_nms
# ------------------------------------------------------------------------------------
def iou_bev(a, b, out=None, overlap=False):
    _need_cuda(a, b)
    L = _lib.lib()
    a, b = _f32c(a), _f32c(b)
    if out is None:
        out = torch.zeros((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    fn = L.cpd_overlap_bev if overlap else L.cpd_iou_bev
    _lib.check(fn(_ptr(a), a.shape[0], _ptr(b), b.shape[0], _ptr(out), _stream()), "cpd_iou_bev")
    return out


def nms(boxes_sorted, thresh, rotated=True):
    """Greedy NMS over boxes sorted by descending score -> (keep int64 (n,) device, n_keep int32 (1,) device)."""
    _need_cuda(boxes_sorted)
    L = _lib.lib()
    b = _f32c(boxes_sorted)
    n = b.shape[0]
    keep = torch.empty((max(n, 1),), dtype=torch.int64, device=b.device)
    n_keep = torch.zeros((1,), dtype=torch.int32, device=b.device)
    wsb = L.cpd_nms_workspace_bytes(n)
    ws = _ws(wsb, b.device)
    fn = L.cpd_nms_rotated if rotated else L.cpd_nms_normal
    _lib.check(fn(_ptr(b), n, float(thresh), _ptr(keep), _ptr(n_keep), _ptr(ws), wsb, _stream()), "cpd_nms")
    return keep, n_keep


def nms_mask(boxes_sorted, thresh, rotated=True):
    _need_cuda(boxes_sorted)
    L = _lib.lib()
    b = _f32c(boxes_sorted)
    n = b.shape[0]
    cb = (n + 63) // 64
    mask = torch.zeros((n, cb), dtype=torch.int64, device=b.device)
    _lib.check(L.cpd_nms_mask(_ptr(b), n, float(thresh), int(bool(rotated)), _ptr(mask), _stream()), "cpd_nms_mask")
    return mask


# ------------------------------------------------------------------------------------
# fused training-mode BatchNorm (+ReLU, + residual) on row matrices
# ------------------------------------------------------------------------------------
def bn_train_fwd(x, stats, gamma, beta, residual, relu, eps, momentum, running_mean, running_var, want_split=False):
    """-> (y, mean_invstd (2,c), split-row image of y | None).  stats (2,c) = per-channel sum / sum of squares of x
    (from gather_gemm)."""
    _need_cuda(x, stats)
    L = _lib.lib()
    x = _f32c(x)
    m, c = x.shape
    y = torch.empty_like(x)
    ys = torch.empty((m, 2, c), dtype=torch.bfloat16, device=x.device) if want_split and c % 8 == 0 else None
    mi = torch.empty((2, c), dtype=torch.float32, device=x.device)
    residual = _f32c(residual) if residual is not None else None
    _lib.check(L.cpd_bn_train_fwd(_ptr(x), m, c, _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(residual), int(bool(relu)),
                                  float(eps), float(momentum), _ptr(running_mean), _ptr(running_var), _ptr(mi), _ptr(y),
                                  _ptr(ys), _stream()), "cpd_bn_train_fwd")
    return y, mi, ys


def bn_train_bwd(x, y, dy, mean_invstd, gamma, relu, want_residual, want_split=False):
    """-> (dx, dresidual | None, dgamma, dbeta, split-row image of dx | None)."""
    _need_cuda(x, dy)
    L = _lib.lib()
    dy = _f32c(dy)
    m, c = x.shape
    dx = torch.empty_like(x)
    dxs = torch.empty((m, 2, c), dtype=torch.bfloat16, device=x.device) if want_split and c % 8 == 0 else None
    dres = torch.empty_like(x) if want_residual else None
    gb = torch.empty((2, c), dtype=torch.float32, device=x.device)
    _lib.check(L.cpd_bn_train_bwd(_ptr(x), _ptr(y), _ptr(dy), m, c, _ptr(mean_invstd), _ptr(gamma), int(bool(relu)), _ptr(dx),
                                  _ptr(dxs), _ptr(dres), _ptr(gb), _stream()), "cpd_bn_train_bwd")
    return dx, dres, gb[1], gb[0], dxs


# ------------------------------------------------------------------------------------
# RoI grid pooling primitives (voxel query + point grouping)
# ------------------------------------------------------------------------------------
def voxel_query(new_xyz, new_coords, xyz, shape, batch, ranges, radius, nsample, hash_buf=None, dense_map=None):
    """-> (idx (m, nsample) int32 global rows, empty (m,) bool).  new_coords (m, 4) int32 [b, z, y, x].  Pass the level's
    coordinate hash (build_hash / SparseConvTensor.coord_hash()) or the reference's dense (B, Z, Y, X) int32 map."""
    _need_cuda(new_xyz, new_coords, xyz)
    L = _lib.lib()
    new_xyz, xyz = _f32c(new_xyz), _f32c(xyz)
    new_coords = new_coords if (new_coords.dtype == torch.int32 and new_coords.is_contiguous()) else new_coords.int().contiguous()
    m = new_xyz.shape[0]
    idx = torch.empty((m, nsample), dtype=torch.int32, device=xyz.device)
    empty = torch.empty((m,), dtype=torch.uint8, device=xyz.device)
    assert (hash_buf is None) != (dense_map is None)
    if dense_map is not None:
        assert dense_map.dtype == torch.int32 and dense_map.is_contiguous()
    _lib.check(L.cpd_voxel_query(_ptr(new_xyz), _ptr(new_coords), m, _ptr(xyz), _ptr(hash_buf), hash_buf.numel() if hash_buf is not None else 0,
                                 _ptr(dense_map), _i32x3(shape), int(batch), _i32x3(ranges), float(radius), int(nsample), _ptr(idx), _ptr(empty),
                                 _stream()), "cpd_voxel_query")
    return idx, empty.bool()


def group_points(features, idx):
    """(n, c), (m, nsample) int32 global rows -> (m, c, nsample)."""
    _need_cuda(features, idx)
    L = _lib.lib()
    features = _f32c(features)
    m, ns = idx.shape
    c = features.shape[1]
    out = torch.empty((m, c, ns), dtype=torch.float32, device=features.device)
    _lib.check(L.cpd_group_points(_ptr(features), _ptr(idx), m, c, ns, _ptr(out), _stream()), "cpd_group_points")
    return out


def group_points_bwd(grad_out, idx, n):
    _need_cuda(grad_out, idx)
    L = _lib.lib()
    grad_out = _f32c(grad_out)
    m, c, ns = grad_out.shape
    g = torch.empty((n, c), dtype=torch.float32, device=grad_out.device)
    _lib.check(L.cpd_group_points_bwd(_ptr(grad_out), _ptr(idx), m, c, ns, n, _ptr(g), _stream()), "cpd_group_points_bwd")
    return g
