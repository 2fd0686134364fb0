"""spconv-compatible sparse tensor and convolution modules on top of libcpd_b200.so.

Mirrors the third-party surface the reference consumes:
``SparseConvTensor``, ``SparseModule``, ``SparseSequential``, ``SubMConv3d``,
``SparseConv3d``, ``SparseInverseConv3d`` (constructible, no forward -- never requested by
the reference, cpd/models/backbones_3d/spconv_backbone.py:24) and
``conv.SparseConvolution`` (cpd/utils/spconv_utils.py:49).  Weights are exposed in the
spconv 2.x layout (cout, kz, ky, kx, cin) so CPD checkpoints load through
cpd/models/detectors/detector3d_template.py:388-419 unchanged.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import ops


def _triple(v):
    if isinstance(v, (list, tuple, np.ndarray)):
        assert len(v) == 3
        return [int(t) for t in v]
    return [int(v)] * 3


class Rulebook:
    """Cached per indice_key, like spconv's indice_dict entries."""

    def __init__(self, kind, nbr_fwd, nbr_bwd, in_coords, out_coords, in_shape, out_shape, ksize, stride, padding):
        self.kind = kind                    # 'subm' | 'strided'
        self.nbr_fwd = nbr_fwd              # (m_out, K) int32
        self.nbr_bwd = nbr_bwd              # strided: (m_in, K); subm: None (same table, flipped taps)
        self.in_coords, self.out_coords = in_coords, out_coords
        self.in_shape, self.out_shape = in_shape, out_shape
        self.ksize, self.stride, self.padding = ksize, stride, padding
        self.out_hash = None                # coordinate hash of out_coords (strided only)
        self._nbr_fwd_t = None              # (K, m_out) transposed table for the weight-gradient kernel
        self._masks_fwd = None              # per-tile tap masks of nbr_fwd (k-block skipping in the tcgen05 kernel)
        self._sorted = {}                   # (which table, taps per k-block) -> (rows in tap-pattern order, their rows, tile masks)

    @property
    def nbr_fwd_t(self):
        if self._nbr_fwd_t is None:
            nb = self.nbr_fwd
            self._nbr_fwd_t = ops.table_transpose(nb) if (nb.is_cuda and nb.shape[1] <= 64 and nb.shape[0] > 0) else nb.t().contiguous()
        return self._nbr_fwd_t

    @property
    def masks_fwd(self):
        """Tile tap masks of the forward table; only worth it for 3-D kernels (z-boundary tiles lose whole taps)."""
        if self.nbr_fwd.shape[1] != 27 or self.nbr_fwd.shape[0] == 0:
            return None
        if self._masks_fwd is None:
            self._masks_fwd = ops.tile_tap_masks(self.nbr_fwd)
        return self._masks_fwd

    SORT_MIN_ROWS = 2048

    def sorted_table(self, which, channels):
        """A 27-tap neighbour table visited in TAP-PATTERN order.  A 128-row tile in spatial order mixes rows with
        different neighbourhoods, so nearly every k-block of the tensor-core kernel is active for it although a row uses
        only 16-60 % of its taps (lidar surfaces are thin sheets; the input sites of a stride-2 conv feed <= 8 of 27
        taps).  Sorting the rows by the bitmap of the k-blocks they need (cpd_tap_block_keys; `channels` = contraction
        width, 64 / channels taps share a k-block) makes tiles whose rows agree: measured on the synthetic sweeps, active
        k-blocks drop 0.69 -> 0.37 (16 ch), 0.76 -> 0.57 (32 ch), 0.76 -> 0.64 (64 ch) and the slots of the active ones
        are ~90 % full.  The kernel skips the rest (tile masks) and scatters row r to y[rows[r]] in its epilogue.
        which: 'fwd' (forward of any conv; input-gradient of a SubM conv, same table with flipped taps) or 'bwd'
        (input-gradient of a strided conv).  Returns (table, rows int32, tile masks) or None when it does not apply."""
        nb = self.nbr_fwd if which == "fwd" else self.nbr_bwd
        if nb is None or nb.shape[1] != 27 or nb.shape[0] < self.SORT_MIN_ROWS:
            return None
        tpb = taps_per_block(channels)
        key = (which, tpb)
        if key not in self._sorted:
            keys = ops.tap_block_keys(nb, tpb)
            if (27 + tpb - 1) // tpb <= 15:                      # <= 15 key bits: half the radix passes on an int16 key
                keys = keys.to(torch.int16)
            perm = torch.sort(keys, stable=True)[1]              # stable: spatial locality survives inside a group
            self._sorted[key] = ops.table_permute(nb, perm)      # (sorted table, its rows as int32, tile masks) in one pass
        return self._sorted[key]

    def bwd_sorted(self, channels=64):
        """Input-gradient table of a strided conv in tap-pattern order (see sorted_table)."""
        if all(int(s) == 1 for s in self.stride):
            return None
        return self.sorted_table("bwd", channels)


def taps_per_block(channels):
    """Taps sharing one 64-element k-block of the tcgen05 gather-GEMM (csrc/spconv_tc.cu)."""
    return max(1, 64 // int(channels)) if channels < 64 else 1


class SparseConvTensor:
    """features (N,C) float32; indices (N,4) int32 [batch,z,y,x]; spatial_shape [D,H,W]."""

    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, voxel_num=None, indice_dict=None,
                 benchmark=False):
        self.features = features
        self.indices = indices if indices.dtype == torch.int32 else indices.int()
        if not self.indices.is_contiguous():
            self.indices = self.indices.contiguous()
        self.spatial_shape = [int(s) for s in spatial_shape]   # accepts a numpy int64 array (spconv_backbone.py:151)
        self.batch_size = int(batch_size)
        self.indice_dict = indice_dict if indice_dict is not None else {}
        self._hash = None                                      # coordinate hash of `indices` (lazily built)
        self.grid, self.voxel_num, self.benchmark = grid, voxel_num, benchmark

    # -- spconv 2.x API -------------------------------------------------------------
    def replace_feature(self, feature):
        t = SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.grid, self.voxel_num,
                             self.indice_dict, self.benchmark)
        t._hash = self._hash
        return t

    @property
    def spatial_size(self):
        return int(np.prod(self.spatial_shape))

    def find_indice_pair(self, key):
        return None if key is None else self.indice_dict.get(key)

    def coord_hash(self):
        if self._hash is None:
            self._hash = ops.build_hash(self.indices, self.spatial_shape, self.batch_size)
        return self._hash

    def dense(self, channels_first=True):
        """(B, C, D, H, W) like upstream; channels_first=False gives (B, D, H, W, C)."""
        out = _ToDense.apply(self.features, self.indices, self.batch_size, tuple(self.spatial_shape), False)
        return out if channels_first else out.permute(0, 2, 3, 4, 1).contiguous()

    def dense_bev_nhwc(self):
        """(B, H, W, C*D): the NHWC form of HeightCompression's (B, C*D, H, W) map
        (cpd/models/backbones_2d/map_to_bev/height_compression.py:136-138), written in one pass."""
        return _ToDense.apply(self.features, self.indices, self.batch_size, tuple(self.spatial_shape), True)


class _ToDense(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, coords, batch, shape, channels_last):
        ctx.save_for_backward(coords)
        ctx.meta = (feat.shape[1], batch, shape, channels_last)
        return ops.sparse_to_dense(feat, coords, batch, shape, channels_last)

    @staticmethod
    def backward(ctx, dout):
        (coords,) = ctx.saved_tensors
        c, batch, shape, cl = ctx.meta
        return ops.sparse_to_dense_bwd(dout.contiguous(), coords, c, batch, shape, cl), None, None, None, None


class _GatherConv(torch.autograd.Function):
    """y = gather-GEMM(x, W, nbr) (+bias); backward = dgrad gather-GEMM + wgrad (Appendix A.5).
    With want_stats the epilogue also returns the per-channel (sum, sum of squares) of y for a fused BatchNorm."""

    @staticmethod
    def forward(ctx, x, weight, bias, rb, algo, want_stats=False):
        ctx.set_materialize_grads(False)      # no zero-filled "gradient" tensor for the non-differentiable stats output
        ctx.rb, ctx.algo = rb, algo
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        stats = torch.empty((2, weight.shape[0]), dtype=torch.float32, device=x.device) if want_stats else None   # cleared by the call
        cout, cin, K = weight.shape[0], weight.shape[-1], rb.nbr_fwd.shape[1]
        # narrow outputs (the 1-3 channel CenterHead maps): zero-pad the output channels so the tensor-core kernels apply
        ctx.pad_out = algo != ops.ALGO_SIMT and cout < 8 and cin % 8 == 0 and not want_stats
        # dense stride-1 convs (rb.geo: a pixel table of the BEV maps) run table-free on the TMA-fed kernel
        geo = getattr(rb, "geo", None) if algo != ops.ALGO_SIMT else None
        if geo is not None and geo["stride"] != 1:
            geo = None
        if ctx.pad_out:
            wp = torch.nn.functional.pad(weight.reshape(cout, K, cin), (0, 0, 0, 0, 0, 16 - cout))
            bp = torch.nn.functional.pad(bias, (0, 16 - cout)) if bias is not None else None
            ctx.xs = None
            if geo is not None and ops.conv2d_ok(cin, geo["k"], 16):
                xs = _carried_split(x)
                if xs is None:
                    xs = ops.split_rows(x)
                return ops.conv2d_fwd(xs, geo["n"], geo["h"], geo["w"], wp, geo["k"], geo["pad"], bias=bp)[:, :cout].contiguous()
            return ops.gather_gemm(x, wp, rb.nbr_fwd, bias=bp, algo=algo)[:, :cout].contiguous()
        # the split-row image of x (bf16 hi | lo, the tcgen05 operand format) is built once and shared by the
        # forward GEMM and the weight-gradient
        needs_grad = bool(ctx.needs_input_grad[1])          # (grad mode is off inside forward: ask the context)
        xs = None
        if algo != ops.ALGO_SIMT and (ops.tc_gemm_ok(cin, K, cout) or (needs_grad and ops.tc_wgrad_ok(cin, K, cout))):
            xs = _carried_split(x)
            if xs is None:
                xs = ops.split_rows(x)
        ctx.xs = xs if needs_grad else None
        srt = rb.sorted_table("fwd", cin) if (xs is not None and ops.tc_gemm_ok(cin, K, cout)) else None
        if geo is not None and xs is not None and ops.conv2d_ok(cin, geo["k"], cout):
            y = ops.conv2d_fwd(xs, geo["n"], geo["h"], geo["w"], weight, geo["k"], geo["pad"], bias=bias, stats=stats)
        elif srt is not None:                                # rows in tap-pattern order, epilogue scatters them back
            y = ops.gather_gemm(x, weight, srt[0], bias=bias, stats=stats, algo=algo, x_split=xs, tile_masks=srt[2], out_rows=srt[1])
        else:
            y = ops.gather_gemm(x, weight, rb.nbr_fwd, bias=bias, stats=stats, algo=algo, x_split=xs, tile_masks=rb.masks_fwd)
        if not want_stats:
            return y
        ctx.mark_non_differentiable(stats)
        return y, stats

    @staticmethod
    def backward(ctx, dy, *unused):
        if dy is None:
            return None, None, None, None, None, None
        x, weight = ctx.saved_tensors
        rb = ctx.rb
        dy = dy.contiguous()
        dx = dw = db = None
        cout, cin, K = weight.shape[0], weight.shape[-1], rb.nbr_fwd.shape[1]
        want_w = ctx.needs_input_grad[1] or ctx.has_bias
        if ctx.pad_out:                                # dy (m, cout < 8) -> (m, 8): zero channels contribute nothing
            dyp = torch.nn.functional.pad(dy, (0, 8 - cout))
            dys = ops.split_rows(dyp)
            if ctx.needs_input_grad[0]:
                wt = torch.nn.functional.pad(ops.weight_transpose(weight, flip_taps=rb.kind == "subm"), (0, 8 - cout))
                dx = ops.gather_gemm(dyp, wt, rb.nbr_fwd if rb.kind == "subm" else rb.nbr_bwd, algo=ctx.algo, x_split=dys)
            if want_w:
                tm = K > 1
                dwp, dbp = ops.gather_wgrad(x, dyp, rb.nbr_fwd_t if tm else rb.nbr_fwd, want_bias=ctx.has_bias, tap_major=tm, dy_split=dys)
                dw, db = dwp[:cout].reshape(weight.shape), (dbp[:cout] if dbp is not None else None)
            return dx, dw, db, None, None, None
        # one split-row image of dy serves the input-gradient GEMM and the weight-gradient
        dys = db_fused = None
        if ctx.algo != ops.ALGO_SIMT and ((ctx.needs_input_grad[0] and ops.tc_gemm_ok(cout, K, cin)) or
                                          (want_w and ops.tc_wgrad_ok(cin, K, cout))):
            if ctx.has_bias:
                dys, db_fused = ops.split_rows(dy, colsum=True)          # the bias gradient falls out of the same pass
            else:
                dys = _carried_split(dy)                                 # written by the BatchNorm backward that produced dy
                if dys is None:
                    dys = ops.split_rows(dy)
        if ctx.needs_input_grad[0]:
            tc_dgrad = dys is not None and ops.tc_gemm_ok(cout, K, cin)
            geo = getattr(rb, "geo", None) if ctx.algo != ops.ALGO_SIMT else None
            if geo is not None and geo["stride"] == 1 and dys is not None and ops.conv2d_ok(cout, geo["k"], cin):
                dx = ops.conv2d_dgrad(dys, geo["n"], geo["h"], geo["w"], cin, weight, geo["k"], geo["pad"])      # table-free, TMA-fed
            elif rb.kind == "subm":
                wt = ops.weight_transpose(weight, flip_taps=True)
                srt = rb.sorted_table("fwd", cout) if tc_dgrad else None
                if srt is not None:
                    dx = ops.gather_gemm(dy, wt, srt[0], algo=ctx.algo, x_split=dys, tile_masks=srt[2], out_rows=srt[1])
                else:
                    dx = ops.gather_gemm(dy, wt, rb.nbr_fwd, algo=ctx.algo, x_split=dys, tile_masks=rb.masks_fwd)
            else:
                wt = ops.weight_transpose(weight, flip_taps=False)
                grouped = rb.bwd_sorted(cout) if tc_dgrad else None
                if grouped is not None:
                    nbs, out_rows, masks = grouped                      # the epilogue scatters row r to dx[out_rows[r]]
                    dx = ops.gather_gemm(dy, wt, nbs, algo=ctx.algo, x_split=dys, tile_masks=masks, out_rows=out_rows)
                else:
                    dx = ops.gather_gemm(dy, wt, rb.nbr_bwd, algo=ctx.algo, x_split=dys)
        if want_w:
            wb = ctx.has_bias and db_fused is None
            if rb.nbr_fwd.shape[1] > 1:
                dw, db = ops.gather_wgrad(x, dy, rb.nbr_fwd_t, want_bias=wb, tap_major=True, x_split=ctx.xs, dy_split=dys)
            else:
                dw, db = ops.gather_wgrad(x, dy, rb.nbr_fwd, want_bias=wb, x_split=ctx.xs, dy_split=dys)
            dw = dw.view_as(weight)
            if db_fused is not None:
                db = db_fused
        ctx.xs = None
        return dx, dw, db, None, None, None


class _BNTrain(torch.autograd.Function):
    """Training-mode BatchNorm (+residual) (+ReLU) on a row matrix, statistics supplied by the conv epilogue."""

    @staticmethod
    def forward(ctx, x, stats, gamma, beta, residual, bn, relu, dx_split):
        ctx.set_materialize_grads(False)      # the split-row image output is non-differentiable: autograd would otherwise
        y, mi, ys = ops.bn_train_fwd(x, stats, gamma, beta, residual, relu, bn.eps, bn.momentum if bn.momentum is not None else 0.1,
                                     bn.running_mean if bn.track_running_stats else None,
                                     bn.running_var if bn.track_running_stats else None, want_split=True)
        if bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        ctx.relu, ctx.has_res, ctx.dx_split = relu, residual is not None, dx_split
        ctx.save_for_backward(x, y if relu else None, mi, gamma)
        if ys is None:
            ys = y.new_empty(0)
        ctx.mark_non_differentiable(ys)
        return y, ys

    @staticmethod
    def backward(ctx, dy, _unused=None):          # zero-fill a tensor of its size in every backward (measured: 53 fills / step)
        if dy is None:
            return None, None, None, None, None, None, None, None
        x, y, mi, gamma = ctx.saved_tensors
        dx, dres, dgamma, dbeta, dxs = ops.bn_train_bwd(x, y, dy, mi, gamma, ctx.relu, ctx.has_res, want_split=ctx.dx_split)
        if dxs is not None:
            dx._cpd_split = dxs            # picked up by the producing convolution's backward (its dy)
        return dx, None, (dgamma if gamma is not None else None), (dbeta if gamma is not None else None), dres, None, None, None


def bn_train(y, stats, bn, relu, residual=None, dx_split=False):
    """Fused training BatchNorm (+residual)(+ReLU).  The output carries the split-row image of itself (`_cpd_split`,
    written by the same pass) for the convolution that consumes it; dx_split: the backward also emits the image of dx
    for the convolution that produced y (worth it when that convolution has no bias gradient to take from its own split)."""
    out, ys = _BNTrain.apply(y, stats, bn.weight, bn.bias, residual, bn, relu, dx_split)
    if ys.numel():
        out._cpd_split = ys
    return out


def _carried_split(t):
    """The split-row image a producer attached to tensor t (bn_train / _BNTrain.backward), if it still matches."""
    s = getattr(t, "_cpd_split", None)
    if s is not None and t.dim() == 2 and s.shape[0] == t.shape[0] and s.shape[2] == t.shape[1] and s.device == t.device:
        return s
    return None


def bn_fusable(bn):
    """A BatchNorm layer whose training-mode forward can take its statistics from the conv epilogue."""
    return isinstance(bn, (nn.BatchNorm1d, nn.BatchNorm2d)) and bn.training and bn.affine and bn.num_features % 4 == 0 \
        and bn.num_features <= 1024 and torch.is_grad_enabled()


class SparseModule(nn.Module):
    """Marker base class: modules that consume/produce SparseConvTensor."""


class SparseConvolution(SparseModule):
    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, subm=False, output_padding=0, transposed=False, inverse=False, indice_key=None,
                 algo=None, **kwargs):
        super().__init__()
        assert ndim == 3 and groups == 1
        assert _triple(dilation) == [1, 1, 1], "dilation is not used by the CPD backbones"
        self.ndim, self.in_channels, self.out_channels = ndim, in_channels, out_channels
        self.kernel_size, self.stride, self.padding = _triple(kernel_size), _triple(stride), _triple(padding)
        self.subm, self.inverse, self.transposed = subm, inverse, transposed
        self.indice_key = indice_key
        self.algo = ops.ALGO_AUTO if algo is None else algo
        self.weight = nn.Parameter(torch.empty(out_channels, *self.kernel_size, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.in_channels * int(np.prod(self.kernel_size))
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"padding={self.padding}, subm={self.subm}, indice_key={self.indice_key}")

    # -- rulebook ---------------------------------------------------------------------
    def get_rulebook(self, inp):
        rb = inp.find_indice_pair(self.indice_key)
        if rb is not None and rb.kind == ("subm" if self.subm else "strided") and rb.in_coords is inp.indices \
                and rb.ksize == self.kernel_size:
            return rb, rb.out_hash
        # a rulebook only depends on the geometry: convs with different indice_keys over the same sites share one
        # (conv_input 'subm1' and the first residual blocks 'res1' are the same 3x3x3 SubM table)
        geo = ("_geometry", self.subm, tuple(self.kernel_size), tuple(self.stride) if not self.subm else None,
               tuple(self.padding) if not self.subm else None)
        rb = inp.indice_dict.get(geo)
        if rb is not None and rb.in_coords is inp.indices:
            if self.indice_key is not None:
                inp.indice_dict[self.indice_key] = rb
            return rb, rb.out_hash
        out_hash = None
        if self.subm:
            nbr = ops.subm_table(inp.indices, inp.spatial_shape, inp.batch_size, self.kernel_size, inp.coord_hash())
            rb = Rulebook("subm", nbr, None, inp.indices, inp.indices, inp.spatial_shape, inp.spatial_shape,
                          self.kernel_size, [1, 1, 1], [0, 0, 0])
        else:
            ocoords, oshape = ops.strided_outputs(inp.indices, inp.spatial_shape, inp.batch_size, self.kernel_size,
                                                  self.stride, self.padding)
            ocoords = ocoords.clone() if ocoords.storage_offset() else ocoords  # keep 16 B alignment
            out_hash = ops.build_hash(ocoords, oshape, inp.batch_size)
            fwd, bwd = ops.strided_tables(inp.indices, inp.spatial_shape, inp.coord_hash(), ocoords, oshape, out_hash,
                                          inp.batch_size, self.kernel_size, self.stride, self.padding, want_bwd=True)
            rb = Rulebook("strided", fwd, bwd, inp.indices, ocoords, inp.spatial_shape, oshape, self.kernel_size,
                          self.stride, self.padding)
            rb.out_hash = out_hash
        if self.indice_key is not None:
            inp.indice_dict[self.indice_key] = rb
        inp.indice_dict[geo] = rb
        return rb, out_hash

    def _wrap_output(self, inp, rb, out_hash, feats):
        if self.subm:
            return inp.replace_feature(feats)
        out = SparseConvTensor(feats, rb.out_coords, rb.out_shape, inp.batch_size, inp.grid, inp.voxel_num,
                               inp.indice_dict, inp.benchmark)
        out._hash = out_hash
        return out

    def forward(self, inp):
        if self.inverse or self.transposed:
            raise NotImplementedError("SparseInverseConv3d forward is not part of the CPD hot path")
        rb, out_hash = self.get_rulebook(inp)
        x, w = inp.features, self.weight
        if self.in_channels % 8:          # e.g. the 5 raw point features: zero-pad to a multiple of 8 so the
            pad = 8 - self.in_channels % 8    # vectorised / tensor-core kernels apply (zeros contribute nothing)
            x, w = torch.nn.functional.pad(x, (0, pad)), torch.nn.functional.pad(w, (0, pad))
        feats = _GatherConv.apply(x, w, self.bias, rb, self.algo)
        return self._wrap_output(inp, rb, out_hash, feats)

    def forward_bn_train(self, inp, bn, relu, residual=None):
        """Training: conv (epilogue emits the batch statistics) -> ONE normalise(+residual)(+ReLU) pass."""
        rb, out_hash = self.get_rulebook(inp)
        x, w = inp.features, self.weight
        if self.in_channels % 8:
            pad = 8 - self.in_channels % 8
            x, w = torch.nn.functional.pad(x, (0, pad)), torch.nn.functional.pad(w, (0, pad))
        y, stats = _GatherConv.apply(x, w, self.bias, rb, self.algo, True)
        if y.shape[0] == 0:
            return self._wrap_output(inp, rb, out_hash, y)
        feats = bn_train(y, stats, bn, relu, residual, dx_split=self.bias is None)
        return self._wrap_output(inp, rb, out_hash, feats)

    def forward_fused(self, inp, scale, shift, relu, residual=None):
        """Inference-only: conv + folded BatchNorm affine (+ residual) (+ ReLU) in one kernel."""
        rb, out_hash = self.get_rulebook(inp)
        x, w = inp.features, self.weight
        if self.in_channels % 8:
            pad = 8 - self.in_channels % 8
            x, w = torch.nn.functional.pad(x, (0, pad)), torch.nn.functional.pad(w, (0, pad))
        cin, K = w.shape[-1], rb.nbr_fwd.shape[1]
        srt = rb.sorted_table("fwd", cin) if (self.algo != ops.ALGO_SIMT and ops.tc_gemm_ok(cin, K, self.out_channels)) else None
        if srt is not None:                                    # tap-pattern order, rows scattered back by the epilogue
            feats = ops.gather_gemm(x, w, srt[0], bias=self.bias, scale=scale, shift=shift, residual=residual, relu=relu,
                                    algo=self.algo, tile_masks=srt[2], out_rows=srt[1])
        else:
            feats = ops.gather_gemm(x, w, rb.nbr_fwd, bias=self.bias, scale=scale, shift=shift,
                                    residual=residual, relu=relu, algo=self.algo, tile_masks=rb.masks_fwd)
        return self._wrap_output(inp, rb, out_hash, feats)


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, **kwargs):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, True,
                         indice_key=indice_key, algo=algo)


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, **kwargs):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, False,
                         indice_key=indice_key, algo=algo)


class SparseInverseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key=None, bias=True, algo=None, **kwargs):
        super().__init__(3, in_channels, out_channels, kernel_size, bias=bias, inverse=True, indice_key=indice_key,
                         algo=algo)


def fold_bn(bn):
    """eval-mode BatchNorm -> per-channel (scale, shift) float32 tensors."""
    inv = torch.rsqrt(bn.running_var.float() + bn.eps)
    g = bn.weight.float() if bn.weight is not None else torch.ones_like(inv)
    b = bn.bias.float() if bn.bias is not None else torch.zeros_like(inv)
    scale = (g * inv).contiguous()
    shift = (b - bn.running_mean.float() * g * inv).contiguous()
    return scale, shift


class SparseSequential(SparseModule):
    """nn.Sequential that threads SparseConvTensor through sparse modules and applies dense
    layers (BatchNorm1d, ReLU, ...) to ``.features`` (spconv_backbone.py:29,153-193).
    In eval mode ``conv -> BatchNorm1d -> ReLU`` runs as ONE kernel (folded affine epilogue)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], dict):
            for k, m in args[0].items():
                self.add_module(k, m)
        else:
            for i, m in enumerate(args):
                self.add_module(str(i), m)
        for k, m in kwargs.items():
            self.add_module(k, m)

    def __getitem__(self, idx):
        return list(self._modules.values())[idx]

    def __len__(self):
        return len(self._modules)

    def add(self, module, name=None):
        self.add_module(name if name is not None else str(len(self._modules)), module)

    def forward(self, x):
        mods = list(self._modules.values())
        i = 0
        while i < len(mods):
            m = mods[i]
            if not self.training and not torch.is_grad_enabled() \
                    and isinstance(m, SparseConvolution) and not m.inverse and isinstance(x, SparseConvTensor) \
                    and i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d):
                relu = i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
                scale, shift = fold_bn(mods[i + 1])
                x = m.forward_fused(x, scale, shift, relu)
                i += 3 if relu else 2
                continue
            if isinstance(m, SparseConvolution) and not m.inverse and isinstance(x, SparseConvTensor) \
                    and i + 1 < len(mods) and bn_fusable(mods[i + 1]):
                relu = i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
                x = m.forward_bn_train(x, mods[i + 1], relu)
                i += 3 if relu else 2
                continue
            if isinstance(m, SparseModule):
                x = m(x)
            elif isinstance(x, SparseConvTensor):
                if x.indices.shape[0] != 0:
                    x = x.replace_feature(m(x.features))
            else:
                x = m(x)
            i += 1
        return x


class ToDense(SparseModule):
    def forward(self, x):
        return x.dense()
