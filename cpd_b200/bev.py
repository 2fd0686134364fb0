"""Dense BEV backbone + CenterHead on the gather-GEMM kernels (NHWC, no cuDNN).

Host-side mirrors of
  * cpd/models/backbones_2d/base_bev_backbone.py:6-122  (``BaseBEVBackbone``)
  * cpd/models/dense_heads/center_head.py:11-354         (``SeparateHead``, ``CenterHead``)
  * cpd/models/model_utils/centernet_utils.py:9-216      (gaussian targets, top-K decode)
  * cpd/utils/loss_utils.py:265-386                      (CenterNet focal / L1 losses)
with the same config keys, module names and parameter shapes (nn.Conv2d-style
(cout, cin, kh, kw) weights => reference checkpoints load), but every convolution runs as
``cpd_gather_gemm`` over a pixel neighbour table (``cpd_conv2d_table``): activations live as
(N*H*W, C) row matrices, i.e. NHWC, so a 3x3 conv is 9 gathered GEMM taps accumulated in TMEM.
Target assignment is vectorised on the device (the reference loops over boxes on the CPU and
syncs, center_head.py:159-219); losses and decode stay plain torch, as in the reference.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import iou3d_nms_utils, ops
from .backbone import _cfg
from .sparse import Rulebook, _GatherConv, _carried_split, bn_fusable, bn_train, fold_bn


class DenseMap:
    """An image batch stored as rows: data (n*h*w, c) contiguous == NHWC."""

    def __init__(self, data, n, h, w):
        self.data, self.n, self.h, self.w = data, n, h, w

    @property
    def c(self):
        return self.data.shape[1]

    @staticmethod
    def from_nchw(x):
        n, c, h, w = x.shape
        return DenseMap(x.permute(0, 2, 3, 1).contiguous().view(n * h * w, c), n, h, w)

    def nchw(self):
        """(n, c, h, w) view with channels_last strides (zero copy)."""
        return self.data.view(self.n, self.h, self.w, self.c).permute(0, 3, 1, 2)


_TABLES = {}


def pixel_tables(n, h, w, kh, kw, stride, pad, device):
    """Cached (nbr_fwd, nbr_bwd, ho, wo) of a conv2d geometry -- the dense analogue of an indice_key."""
    key = (n, h, w, kh, kw, stride, pad, device.index)
    if key not in _TABLES:
        ho, wo = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
        if kh == 1 and kw == 1 and stride == 1 and pad == 0:
            fwd = torch.arange(n * h * w, dtype=torch.int32, device=device).view(-1, 1)
            bwd = fwd
        else:
            fwd = ops.conv2d_table(n, h, w, kh, kw, stride, pad, False, ho, wo, device)
            bwd = ops.conv2d_table(n, ho, wo, kh, kw, stride, pad, True, h, w, device)
        rb = Rulebook("strided", fwd, bwd, None, None, None, None, [1, kh, kw], [1, stride, stride], [0, pad, pad])
        if kh == kw:                                     # geometry for the table-free TMA kernel (cpd_conv2d_fwd / _dgrad)
            rb.geo = dict(n=n, h=h, w=w, k=kh, stride=stride, pad=pad, ho=ho, wo=wo)
        _TABLES[key] = (rb, ho, wo)
    return _TABLES[key]


class DenseConv2d(nn.Module):
    """nn.Conv2d look-alike (square kernels, zero padding) computing with gather-GEMM on DenseMap."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, algo=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.k, self.stride, self.padding = int(kernel_size), int(stride), int(padding)
        self.algo = ops.ALGO_AUTO if algo is None else algo
        # nn.Conv2d's (cout, cin, kh, kw) SHAPE (checkpoints load unchanged) over (cout, kh, kw, cin) MEMORY (channels_last):
        # the kernels' (cout, K, cin) operand is then a view -- no per-step permute copy forward, none for the gradient
        self.weight = nn.Parameter(torch.empty(out_channels, self.k, self.k, in_channels).permute(0, 3, 1, 2))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        if bias:
            bound = 1.0 / (in_channels * self.k * self.k) ** 0.5
            nn.init.uniform_(self.bias, -bound, bound)

    def w_kc(self):
        """(cout, cin, kh, kw) -> (cout, kh*kw, cin): the operand layout of cpd_gather_gemm."""
        return self.weight.permute(0, 2, 3, 1).reshape(self.out_channels, self.k * self.k, self.in_channels)

    def forward(self, x, scale=None, shift=None, relu=False):
        rb, ho, wo = pixel_tables(x.n, x.h, x.w, self.k, self.k, self.stride, self.padding, x.data.device)
        if scale is not None or relu:   # inference: folded BatchNorm (+ReLU) in the epilogue
            if self.stride == 1 and self.algo != ops.ALGO_SIMT and ops.conv2d_ok(self.in_channels, self.k, self.out_channels):
                y = ops.conv2d_fwd(ops.split_rows(x.data), x.n, x.h, x.w, self.w_kc().contiguous(), self.k, self.padding, bias=self.bias,
                                   scale=scale, shift=shift, relu=relu)
            else:
                y = ops.gather_gemm(x.data, self.w_kc().contiguous(), rb.nbr_fwd, bias=self.bias, scale=scale, shift=shift,
                                    relu=relu, algo=self.algo)
        else:
            y = _GatherConv.apply(x.data, self.w_kc().contiguous(), self.bias, rb, self.algo)
        return DenseMap(y, x.n, ho, wo)

    def forward_bn_train(self, x, bn, relu):
        """Training: conv emitting batch statistics, then one fused normalise(+ReLU) pass."""
        rb, ho, wo = pixel_tables(x.n, x.h, x.w, self.k, self.k, self.stride, self.padding, x.data.device)
        y, stats = _GatherConv.apply(x.data, self.w_kc().contiguous(), self.bias, rb, self.algo, True)
        return DenseMap(bn_train(y, stats, bn, relu, dx_split=self.bias is None), x.n, ho, wo)


class _ConvT2d(torch.autograd.Function):
    """nn.ConvTranspose2d with kernel == stride == s > 1 on row matrices.  Forward: s*s 1x1 GEMMs whose epilogues write the
    pixel-shuffled map directly (cpd_convt2d_fwd).  Backward: ONE pixel-unshuffle copy of dy to (m, s*s*cout) rows, then the
    input-gradient is a single 1x1 GEMM over s*s*cout channels and the weight-gradient a single K = 1 reduction."""

    @staticmethod
    def forward(ctx, x, weight, n, h, w, want_stats):
        ctx.set_materialize_grads(False)
        s = weight.shape[2]
        xs = _carried_split(x)
        if xs is None:
            xs = ops.split_rows(x)
        stats = torch.empty((2, weight.shape[1]), dtype=torch.float32, device=x.device) if want_stats else None
        y = ops.convt2d_fwd(xs, n, h, w, weight, s, stats=stats)
        ctx.save_for_backward(x, weight)
        ctx.xs, ctx.geo = xs, (n, h, w, s)
        if not want_stats:
            return y
        ctx.mark_non_differentiable(stats)
        return y, stats

    @staticmethod
    def backward(ctx, dy, *unused):
        if dy is None:
            return None, None, None, None, None, None
        x, weight = ctx.saved_tensors
        n, h, w, s = ctx.geo
        cin, cout = weight.shape[0], weight.shape[1]
        m = n * h * w
        dy = dy.contiguous()
        # (n, h, s, w, s, cout) -> (n, h, w, s, s, cout): row p holds the s*s output pixels input pixel p feeds
        dy4 = dy.view(n, h, s, w, s, cout).permute(0, 1, 3, 2, 4, 5).reshape(m, s * s * cout)
        dys = ops.split_rows(dy4)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            # dx[p, ci] = sum_{k, co} dy4[p, (k, co)] W[ci, co, k]: a 1x1 conv with weight (cin, 1, s*s*cout)
            wd = weight.permute(0, 2, 3, 1).reshape(cin, 1, s * s * cout).contiguous()
            dx = ops.conv2d_fwd(dys, n, h, w, wd, 1, 0)
        if ctx.needs_input_grad[1]:
            ident = pixel_tables(n, h, w, 1, 1, 1, 0, x.device)[0].nbr_fwd
            dwk, _ = ops.gather_wgrad(x, dy4, ident, x_split=ctx.xs, dy_split=dys)          # (s*s*cout, 1, cin)
            dw = dwk.view(s, s, cout, cin).permute(3, 2, 0, 1).contiguous()
        ctx.xs = None
        return dx, dw, None, None, None, None


class DenseConvTranspose2d(nn.Module):
    """nn.ConvTranspose2d with kernel == stride (base_bev_backbone.py:48-59): every output pixel
    has exactly one contributing input pixel, so it is stride^2 1x1 GEMMs + a pixel shuffle."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, bias=False, algo=None):
        super().__init__()
        assert int(kernel_size) == int(stride) and not bias
        self.in_channels, self.out_channels, self.s = in_channels, out_channels, int(stride)
        self.algo = ops.ALGO_AUTO if algo is None else algo
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels, self.s, self.s))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        self.bias = None

    def _tma(self):
        return self.algo != ops.ALGO_SIMT and ops.conv2d_ok(self.in_channels, 1, self.out_channels) and \
            ops.conv2d_ok(self.s * self.s * self.out_channels, 1, self.in_channels)

    def forward(self, x, scale=None, shift=None, relu=False):
        s = self.s
        rb, _, _ = pixel_tables(x.n, x.h, x.w, 1, 1, 1, 0, x.data.device)
        fused = scale is not None or relu
        if s == 1:
            w = self.weight[:, :, 0, 0].t().reshape(self.out_channels, 1, self.in_channels).contiguous()
            if fused and self._tma():
                y = ops.conv2d_fwd(ops.split_rows(x.data), x.n, x.h, x.w, w, 1, 0, scale=scale, shift=shift, relu=relu)
            else:
                y = (ops.gather_gemm(x.data, w, rb.nbr_fwd, scale=scale, shift=shift, relu=relu, algo=self.algo) if fused
                     else _GatherConv.apply(x.data, w, None, rb, self.algo))
            return DenseMap(y, x.n, x.h, x.w)
        if self._tma():
            y = (ops.convt2d_fwd(ops.split_rows(x.data), x.n, x.h, x.w, self.weight, s, scale=scale, shift=shift, relu=relu) if fused
                 else _ConvT2d.apply(x.data, self.weight, x.n, x.h, x.w, False))
            return DenseMap(y, x.n, x.h * s, x.w * s)
        out = x.data.new_empty((x.n, x.h * s, x.w * s, self.out_channels))
        for ky in range(s):
            for kx in range(s):
                w = self.weight[:, :, ky, kx].t().reshape(self.out_channels, 1, self.in_channels).contiguous()
                y = (ops.gather_gemm(x.data, w, rb.nbr_fwd, scale=scale, shift=shift, relu=relu, algo=self.algo) if fused
                     else _GatherConv.apply(x.data, w, None, rb, self.algo))
                out[:, ky::s, kx::s, :] = y.view(x.n, x.h, x.w, self.out_channels)
        return DenseMap(out.view(-1, self.out_channels), x.n, x.h * s, x.w * s)

    def forward_bn_train(self, x, bn, relu):
        """Training: the GEMM epilogues emit the batch statistics of the whole output map, then one fused normalise(+ReLU) pass."""
        s = self.s
        if s == 1:
            rb, _, _ = pixel_tables(x.n, x.h, x.w, 1, 1, 1, 0, x.data.device)
            w = self.weight[:, :, 0, 0].t().reshape(self.out_channels, 1, self.in_channels).contiguous()
            y, stats = _GatherConv.apply(x.data, w, None, rb, self.algo, True)
        else:
            y, stats = _ConvT2d.apply(x.data, self.weight, x.n, x.h, x.w, True)
        return DenseMap(bn_train(y, stats, bn, relu, dx_split=s == 1), x.n, x.h * s, x.w * s)


class DenseSequential(nn.Sequential):
    """nn.Sequential over DenseMap: conv -> BatchNorm2d -> ReLU groups; in eval mode each group is
    ONE kernel (BN folded into the conv epilogue)."""

    def forward(self, x):
        mods = list(self)
        fuse = not self.training and not torch.is_grad_enabled()
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, (DenseConv2d, DenseConvTranspose2d)):
                bn = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm2d) else None
                relu = mods[i + (2 if bn is not None else 1)] if i + (2 if bn is not None else 1) < len(mods) else None
                has_relu = isinstance(relu, nn.ReLU)
                if fuse:
                    sc, sh = fold_bn(bn) if bn is not None else (None, None)
                    x = m(x, scale=sc, shift=sh, relu=has_relu)
                    i += 1 + (bn is not None) + has_relu
                    continue
                if bn is not None and bn_fusable(bn) and (isinstance(m, DenseConv2d) or m.s == 1 or m._tma()):
                    x = m.forward_bn_train(x, bn, has_relu)
                    i += 2 + has_relu
                    continue
                x = m(x)
            elif isinstance(m, nn.BatchNorm2d):
                y = F.batch_norm(x.data, m.running_mean, m.running_var, m.weight, m.bias, m.training,
                                 m.momentum if m.momentum is not None else 0.1, m.eps)
                if m.training and m.num_batches_tracked is not None:
                    m.num_batches_tracked += 1
                x = DenseMap(y, x.n, x.h, x.w)
            elif isinstance(m, nn.ReLU):
                x = DenseMap(F.relu(x.data), x.n, x.h, x.w)
            elif isinstance(m, (nn.Identity, nn.ZeroPad2d)):
                pass                      # the padding is folded into the following conv's table
            elif isinstance(m, DenseSequential):
                x = m(x)
            else:
                raise TypeError(f"DenseSequential cannot run {type(m).__name__}")
            i += 1
        return x


class BaseBEVBackbone(nn.Module):
    """base_bev_backbone.py:6-122 -- same LAYER_NUMS / LAYER_STRIDES / NUM_FILTERS / UPSAMPLE_STRIDES /
    NUM_UPSAMPLE_FILTERS keys, same ``blocks.i.j`` / ``deblocks.i.j`` parameter names."""

    def __init__(self, model_cfg, num_frames=1, input_channels=256, **kwargs):
        super().__init__()
        cfg = _cfg(model_cfg)
        self.model_cfg, self.num_frames = cfg, num_frames
        layer_nums, layer_strides, num_filters = cfg.get("LAYER_NUMS", []), cfg.get("LAYER_STRIDES", []), cfg.get("NUM_FILTERS", [])
        ups, num_up = cfg.get("UPSAMPLE_STRIDES", []), cfg.get("NUM_UPSAMPLE_FILTERS", [])
        assert len(layer_nums) == len(layer_strides) == len(num_filters) and len(ups) == len(num_up)
        c_in = [input_channels, *num_filters[:-1]]
        bn = lambda c: nn.BatchNorm2d(c, eps=1e-3, momentum=0.01)
        self.blocks, self.deblocks = nn.ModuleList(), nn.ModuleList()
        for i in range(len(layer_nums)):
            layers = [nn.Identity(),   # slot of the reference's nn.ZeroPad2d(1): padding lives in the conv table
                      DenseConv2d(c_in[i], num_filters[i], 3, stride=layer_strides[i], padding=1, bias=False),
                      bn(num_filters[i]), nn.ReLU()]
            for _ in range(layer_nums[i]):
                layers += [DenseConv2d(num_filters[i], num_filters[i], 3, padding=1, bias=False), bn(num_filters[i]), nn.ReLU()]
            self.blocks.append(DenseSequential(*layers))
            if len(ups) > 0:
                assert ups[i] >= 1, "fractional UPSAMPLE_STRIDES are not used by the CPD configs"
                self.deblocks.append(DenseSequential(
                    DenseConvTranspose2d(num_filters[i], num_up[i], ups[i], stride=ups[i], bias=False), bn(num_up[i]), nn.ReLU()))
        self.num_bev_features_post = sum(num_up)

    def forward(self, data_dict):
        sf = data_dict["temporal_features"] if "temporal_features" in data_dict else data_dict["spatial_features"]
        x = sf if isinstance(sf, DenseMap) else DenseMap.from_nchw(sf)
        outs = []
        for i, blk in enumerate(self.blocks):
            x = blk(x)
            outs.append(self.deblocks[i](x) if len(self.deblocks) > 0 else x)
        y = outs[0] if len(outs) == 1 else DenseMap(torch.cat([o.data for o in outs], dim=1), outs[0].n, outs[0].h, outs[0].w)
        data_dict["st_features_2d_map"] = y
        data_dict["st_features_2d"] = y.nchw()
        return data_dict


# --------------------------------------------------------------------------------------------
# CenterHead
# --------------------------------------------------------------------------------------------
class SeparateHead(nn.Module):
    """center_head.py:11-45."""

    def __init__(self, input_channels, sep_head_dict, init_bias=-2.19, use_bias=False):
        super().__init__()
        self.sep_head_dict = sep_head_dict
        for name, spec in sep_head_dict.items():
            layers = []
            for _ in range(spec["num_conv"] - 1):
                layers.append(DenseSequential(DenseConv2d(input_channels, input_channels, 3, padding=1, bias=use_bias),
                                              nn.BatchNorm2d(input_channels), nn.ReLU()))
            layers.append(DenseConv2d(input_channels, spec["out_channels"], 3, padding=1, bias=True))
            fc = DenseSequential(*layers)
            if "hm" in name:
                fc[-1].bias.data.fill_(init_bias)
            else:
                for m in fc.modules():
                    if isinstance(m, DenseConv2d):
                        nn.init.kaiming_normal_(m.weight.data)
                        if m.bias is not None:
                            nn.init.constant_(m.bias, 0)
            setattr(self, name, fc)

    def forward(self, x):
        return {name: getattr(self, name)(x).nchw() for name in self.sep_head_dict}


def gaussian_radius(height, width, min_overlap=0.5):
    """centernet_utils.py:9-36."""
    b1 = height + width
    c1 = width * height * (1 - min_overlap) / (1 + min_overlap)
    r1 = (b1 + (b1 ** 2 - 4 * c1).sqrt()) / 2
    b2 = 2 * (height + width)
    c2 = (1 - min_overlap) * width * height
    r2 = (b2 + (b2 ** 2 - 16 * c2).sqrt()) / 2
    a3 = 4 * min_overlap
    b3 = -2 * min_overlap * (height + width)
    c3 = (min_overlap - 1) * width * height
    r3 = (b3 + (b3 ** 2 - 4 * a3 * c3).sqrt()) / 2
    return torch.min(torch.min(r1, r2), r3)


def assign_targets_single(gt_boxes, num_classes, fmap_xy, stride, pc_range, voxel_size, num_max_objs=500,
                          gaussian_overlap=0.1, min_radius=2, max_radius=24):
    """center_head.py:103-157 for one frame/head, vectorised on gt_boxes.device (no host loop, no sync).

    gt_boxes (M, 8) [x,y,z,dx,dy,dz,heading,cls(1-based; 0 = padding)].  Drawing a gaussian is a
    max-reduction, so all boxes are rasterised at once into (2R+1)^2 patches and merged with
    scatter_reduce(amax); per-box radii are honoured by masking the patch.
    Returns heatmap (C, H, W), ret_boxes (num_max_objs, 8), inds (num_max_objs) long, mask (num_max_objs) long."""
    W, H = int(fmap_xy[0]), int(fmap_xy[1])
    dev = gt_boxes.device
    gt = gt_boxes[:num_max_objs]
    m = gt.shape[0]
    heatmap = gt.new_zeros(num_classes, H, W)
    ret_boxes = gt.new_zeros((num_max_objs, 8))
    inds = torch.zeros(num_max_objs, dtype=torch.long, device=dev)
    mask = torch.zeros(num_max_objs, dtype=torch.long, device=dev)
    if m == 0:
        return heatmap, ret_boxes, inds, mask
    cx = ((gt[:, 0] - pc_range[0]) / voxel_size[0] / stride).clamp(min=0, max=W - 0.5)
    cy = ((gt[:, 1] - pc_range[1]) / voxel_size[1] / stride).clamp(min=0, max=H - 0.5)
    ix, iy = cx.int(), cy.int()
    dx, dy = gt[:, 3] / voxel_size[0] / stride, gt[:, 4] / voxel_size[1] / stride
    radius = torch.clamp_min(gaussian_radius(dx, dy, min_overlap=gaussian_overlap).int(), min_radius)
    cls = gt[:, 7].long() - 1
    valid = (dx > 0) & (dy > 0) & (cls >= 0) & (cls < num_classes)
    radius = radius.clamp(max=max_radius)
    R = max_radius
    off = torch.arange(-R, R + 1, device=dev)
    oy, ox = torch.meshgrid(off, off, indexing="ij")                         # (D, D)
    sigma = (2 * radius + 1).float() / 6.0
    g = torch.exp(-(ox * ox + oy * oy).float()[None] / (2 * sigma * sigma)[:, None, None])
    px, py = ix[:, None, None] + ox[None], iy[:, None, None] + oy[None]
    inside = (ox.abs()[None] <= radius[:, None, None]) & (oy.abs()[None] <= radius[:, None, None]) & \
             (px >= 0) & (px < W) & (py >= 0) & (py < H) & valid[:, None, None]
    flat = (cls.clamp(0, num_classes - 1)[:, None, None] * H + py.clamp(0, H - 1)) * W + px.clamp(0, W - 1)
    vals = torch.where(inside, g, torch.zeros_like(g))
    heatmap.view(-1).scatter_reduce_(0, flat.reshape(-1).long(), vals.reshape(-1), reduce="amax", include_self=True)
    k = torch.arange(m, device=dev)
    v = valid.long()
    inds[k] = (iy.long() * W + ix.long()) * v
    mask[k] = v
    rb = torch.stack([cx - ix.float(), cy - iy.float(), gt[:, 2], gt[:, 3].clamp_min(1e-6).log(), gt[:, 4].clamp_min(1e-6).log(),
                      gt[:, 5].clamp_min(1e-6).log(), torch.cos(gt[:, 6]), torch.sin(gt[:, 6])], dim=1)
    ret_boxes[k] = rb * valid[:, None].float()
    return heatmap, ret_boxes, inds, mask


def assign_targets_batched(gt_boxes, num_classes, fmap_xy, stride, pc_range, voxel_size, num_max_objs=500,
                           gaussian_overlap=0.1, min_radius=2, max_radius=24, strict=False):
    """center_head.py:103-157,188-207 over a whole batch at once: gt_boxes (B, M, 8+) with the HEAD-LOCAL class in the last
    column (0 = not in this head / padding) -> heatmap (B, C, H, W), ret_boxes (B, num_max_objs, 8+), inds / mask
    (B, num_max_objs).  Same arithmetic as assign_targets_single, one set of launches instead of one per frame; the frame index
    is folded into the scatter index.
    Like the reference, the boxes of the head are COMPACTED first (stable) and then the first num_max_objs of them are used, so
    slot k of ret_boxes / inds / mask is the k-th box of this head, also with several heads or more than num_max_objs rows;
    extra box columns (velocity, gt[..., 7:-1]) are carried into ret_boxes[..., 8:].
    The one deliberate bound: a gaussian is rasterised into a (2 max_radius + 1)^2 patch, so radii are exact up to max_radius
    (24 feature-map pixels = a 19 m object at stride 8 x 0.1 m; the reference has no cap).  strict=True checks it (one host sync)."""
    W, H = int(fmap_xy[0]), int(fmap_xy[1])
    dev = gt_boxes.device
    B, ncol = gt_boxes.shape[0], gt_boxes.shape[-1]
    heatmap = gt_boxes.new_zeros(B, num_classes, H, W)
    ret_boxes = gt_boxes.new_zeros((B, num_max_objs, ncol))
    inds = torch.zeros((B, num_max_objs), dtype=torch.long, device=dev)
    mask = torch.zeros((B, num_max_objs), dtype=torch.long, device=dev)
    if gt_boxes.shape[1] == 0 or B == 0:
        return heatmap, ret_boxes, inds, mask
    in_head = (gt_boxes[..., -1] >= 1) & (gt_boxes[..., -1] <= num_classes)
    order = torch.sort((~in_head).to(torch.int8), dim=1, stable=True)[1]          # this head's boxes first, original order kept
    gt = torch.gather(gt_boxes, 1, order[..., None].expand(-1, -1, ncol))[:, :num_max_objs]
    in_head = torch.gather(in_head, 1, order)[:, :num_max_objs]
    m = gt.shape[1]
    cx = ((gt[..., 0] - pc_range[0]) / voxel_size[0] / stride).clamp(min=0, max=W - 0.5)
    cy = ((gt[..., 1] - pc_range[1]) / voxel_size[1] / stride).clamp(min=0, max=H - 0.5)
    ix, iy = cx.int(), cy.int()
    dx, dy = gt[..., 3] / voxel_size[0] / stride, gt[..., 4] / voxel_size[1] / stride
    radius = torch.clamp_min(gaussian_radius(dx, dy, min_overlap=gaussian_overlap).int(), min_radius)
    cls = gt[..., -1].long() - 1
    valid = (dx > 0) & (dy > 0) & in_head
    if strict:
        assert int(torch.where(valid, radius, torch.zeros_like(radius)).max()) <= max_radius, "gaussian radius exceeds max_radius"
    radius = radius.clamp(max=max_radius)
    R = max_radius
    off = torch.arange(-R, R + 1, device=dev)
    oy, ox = torch.meshgrid(off, off, indexing="ij")                         # (D, D)
    sigma = (2 * radius + 1).float() / 6.0
    g = torch.exp(-(ox * ox + oy * oy).float()[None, None] / (2 * sigma * sigma)[:, :, None, None])      # (B, m, D, D)
    px, py = ix[:, :, None, None] + ox[None, None], iy[:, :, None, None] + oy[None, None]
    rad = radius[:, :, None, None]
    inside = (ox.abs()[None, None] <= rad) & (oy.abs()[None, None] <= rad) & (px >= 0) & (px < W) & (py >= 0) & (py < H) & \
             valid[:, :, None, None]
    frame = torch.arange(B, device=dev)[:, None, None, None]
    flat = ((frame * num_classes + cls.clamp(0, num_classes - 1)[:, :, None, None]) * H + py.clamp(0, H - 1)) * W + px.clamp(0, W - 1)
    vals = torch.where(inside, g, torch.zeros_like(g))
    heatmap.view(-1).scatter_reduce_(0, flat.reshape(-1).long(), vals.reshape(-1), reduce="amax", include_self=True)
    v = valid.long()
    inds[:, :m] = (iy.long() * W + ix.long()) * v
    mask[:, :m] = v
    rb = torch.stack([cx - ix.float(), cy - iy.float(), gt[..., 2], gt[..., 3].clamp_min(1e-6).log(), gt[..., 4].clamp_min(1e-6).log(),
                      gt[..., 5].clamp_min(1e-6).log(), torch.cos(gt[..., 6]), torch.sin(gt[..., 6])], dim=2)
    if ncol > 8:
        rb = torch.cat([rb, gt[..., 7:-1]], dim=2)
    ret_boxes[:, :m] = rb * valid[..., None].float()
    return heatmap, ret_boxes, inds, mask


def focal_loss_centernet(pred, gt):
    """loss_utils.py:265-302 (neg_loss_cornernet), without the host-side `if num_pos == 0` sync."""
    pos = gt.eq(1).float()
    neg = gt.lt(1).float()
    pos_loss = (torch.log(pred) * torch.pow(1 - pred, 2) * pos).sum()
    neg_loss = (torch.log(1 - pred) * torch.pow(pred, 2) * torch.pow(1 - gt, 4) * neg).sum()
    num_pos = pos.sum()
    return torch.where(num_pos == 0, -neg_loss, -(pos_loss + neg_loss) / num_pos.clamp_min(1.0))


def reg_loss_centernet(output, mask, ind, target):
    """loss_utils.py:318-386: L1 over gathered predictions, per regression channel."""
    b, c = output.shape[0], output.shape[1]
    pred = output.permute(0, 2, 3, 1).reshape(b, -1, c).gather(1, ind.unsqueeze(2).expand(-1, -1, c))
    num = mask.float().sum()
    mk = mask.unsqueeze(2).expand_as(target).float() * (~torch.isnan(target)).float()
    loss = torch.abs(pred * mk - target * mk).sum(dim=(0, 1))
    return loss / torch.clamp_min(num, 1.0)


def topk_decode(heatmap, rot_cos, rot_sin, center, center_z, dim, pc_range, voxel_size, stride, K, score_thresh,
                post_center_limit_range):
    """centernet_utils.py:136-216 (_topk + decode_bbox_from_heatmap), batched; returns per-frame dicts."""
    b, c, h, w = heatmap.shape
    K = min(K, h * w)
    s1, i1 = torch.topk(heatmap.flatten(2, 3), K)
    ys, xs = (i1 // w).float(), (i1 % w).float()
    score, i2 = torch.topk(s1.view(b, -1), K)
    cls = (i2 // K).int()
    pick = lambda t: t.view(b, -1).gather(1, i2)
    inds, ys, xs = pick(i1), pick(ys), pick(xs)

    def at(feat):
        ch = feat.shape[1]
        return feat.permute(0, 2, 3, 1).reshape(b, -1, ch).gather(1, inds.unsqueeze(2).expand(-1, -1, ch))
    ctr, rs, rc, cz, dm = at(center), at(rot_sin), at(rot_cos), at(center_z), at(dim)
    angle = torch.atan2(rs, rc)
    x = (xs.unsqueeze(2) + ctr[:, :, 0:1]) * stride * voxel_size[0] + pc_range[0]
    y = (ys.unsqueeze(2) + ctr[:, :, 1:2]) * stride * voxel_size[1] + pc_range[1]
    boxes = torch.cat([x, y, cz, dm, angle], dim=-1)
    mask = (boxes[..., :3] >= post_center_limit_range[:3]).all(2) & (boxes[..., :3] <= post_center_limit_range[3:]).all(2)
    if score_thresh is not None:
        mask &= score > score_thresh
    return [dict(pred_boxes=boxes[k, mask[k]], pred_scores=score[k, mask[k]], pred_labels=cls[k, mask[k]]) for k in range(b)]


DEFAULT_HEAD_CFG = dict(
    CLASS_NAMES_EACH_HEAD=[["Vehicle", "Pedestrian", "Cyclist"]], SHARED_CONV_CHANNEL=64, USE_BIAS_BEFORE_NORM=True,
    NUM_HM_CONV=2,
    SEPARATE_HEAD_CFG=dict(HEAD_ORDER=["center", "center_z", "dim", "rot"],
                           HEAD_DICT={"center": {"out_channels": 2, "num_conv": 2}, "center_z": {"out_channels": 1, "num_conv": 2},
                                      "dim": {"out_channels": 3, "num_conv": 2}, "rot": {"out_channels": 2, "num_conv": 2}}),
    TARGET_ASSIGNER_CONFIG=dict(FEATURE_MAP_STRIDE=8, NUM_MAX_OBJS=500, GAUSSIAN_OVERLAP=0.1, MIN_RADIUS=2),
    LOSS_CONFIG=dict(LOSS_WEIGHTS={"cls_weight": 1.0, "loc_weight": 2.0, "code_weights": [1.0] * 8}),
    POST_PROCESSING=dict(SCORE_THRESH=0.1, POST_CENTER_LIMIT_RANGE=[-75.2, -75.2, -2, 75.2, 75.2, 4], MAX_OBJ_PER_SAMPLE=500,
                         NMS_CONFIG=dict(NMS_TYPE="nms_gpu", NMS_THRESH=0.8, NMS_PRE_MAXSIZE=4096, NMS_POST_MAXSIZE=500)))


class CenterHead(nn.Module):
    """center_head.py:48-354 with the config of tools/cfgs/models/waymo_unsupervised/voxel_rcnn_cproto_center.yaml:41-80."""

    def __init__(self, model_cfg=None, num_frames=1, input_channels=512, num_class=3, class_names=("Vehicle", "Pedestrian", "Cyclist"),
                 grid_size=None, point_cloud_range=None, voxel_size=None, predict_boxes_when_training=True):
        super().__init__()
        cfg = _cfg(model_cfg if model_cfg is not None else DEFAULT_HEAD_CFG)
        self.model_cfg, self.num_class, self.class_names = cfg, num_class, list(class_names)
        self.grid_size, self.point_cloud_range, self.voxel_size = grid_size, point_cloud_range, voxel_size
        self.feature_map_stride = _cfg(cfg["TARGET_ASSIGNER_CONFIG"]).get("FEATURE_MAP_STRIDE", None)
        self.class_names_each_head, self.class_id_mapping_each_head = [], []
        for names in cfg["CLASS_NAMES_EACH_HEAD"]:
            cur = [x for x in names if x in self.class_names]
            self.class_names_each_head.append(cur)
            self.class_id_mapping_each_head.append(torch.tensor([self.class_names.index(x) for x in cur], dtype=torch.long))
        assert sum(len(x) for x in self.class_names_each_head) == len(self.class_names)
        sc = cfg["SHARED_CONV_CHANNEL"]
        use_bias = cfg.get("USE_BIAS_BEFORE_NORM", False)
        self.shared_conv = DenseSequential(DenseConv2d(input_channels, sc, 3, padding=1, bias=use_bias), nn.BatchNorm2d(sc), nn.ReLU())
        self.heads_list = nn.ModuleList()
        self.separate_head_cfg = _cfg(cfg["SEPARATE_HEAD_CFG"])
        for cur in self.class_names_each_head:
            hd = {k: dict(v) for k, v in self.separate_head_cfg["HEAD_DICT"].items()}
            hd["hm"] = dict(out_channels=len(cur), num_conv=cfg["NUM_HM_CONV"])
            self.heads_list.append(SeparateHead(sc, hd, init_bias=-2.19, use_bias=use_bias))
        self.predict_boxes_when_training = predict_boxes_when_training
        self.forward_ret_dict = {}

    # ---- targets / loss ----------------------------------------------------------------
    def assign_targets(self, gt_boxes, feature_map_size):
        """center_head.py:159-219 on the device.  gt_boxes (B, M, 8), class in the last column (0 = pad)."""
        tcfg = _cfg(self.model_cfg["TARGET_ASSIGNER_CONFIG"])
        fm_xy = [int(feature_map_size[1]), int(feature_map_size[0])]
        ret = dict(heatmaps=[], target_boxes=[], inds=[], masks=[])
        for names in self.class_names_each_head:
            local = torch.zeros(len(self.class_names) + 1, dtype=gt_boxes.dtype, device=gt_boxes.device)
            for j, nme in enumerate(names):
                local[self.class_names.index(nme) + 1] = j + 1
            g = gt_boxes.clone()
            g[..., -1] = local[g[..., -1].long().clamp(0, len(self.class_names))]
            per = assign_targets_batched(g, len(names), fm_xy, tcfg["FEATURE_MAP_STRIDE"], self.point_cloud_range, self.voxel_size,
                                         tcfg["NUM_MAX_OBJS"], tcfg["GAUSSIAN_OVERLAP"], tcfg["MIN_RADIUS"])   # all frames at once
            for key, idx in (("heatmaps", 0), ("target_boxes", 1), ("inds", 2), ("masks", 3)):
                ret[key].append(per[idx])
        return ret

    def get_loss(self):
        """center_head.py:225-250; tb_dict holds tensors (no .item() syncs inside the step)."""
        preds, tgt = self.forward_ret_dict["pred_dicts"], self.forward_ret_dict["target_dicts"]
        lw = self.model_cfg["LOSS_CONFIG"]["LOSS_WEIGHTS"]
        loss, tb = 0, {}
        for i, pd in enumerate(preds):
            hm = torch.clamp(pd["hm"].sigmoid(), min=1e-4, max=1 - 1e-4)
            hm_loss = focal_loss_centernet(hm, tgt["heatmaps"][i]) * lw["cls_weight"]
            pred_boxes = torch.cat([pd[n] for n in self.separate_head_cfg["HEAD_ORDER"]], dim=1)
            reg = reg_loss_centernet(pred_boxes, tgt["masks"][i], tgt["inds"][i], tgt["target_boxes"][i])
            loc_loss = (reg * reg.new_tensor(lw["code_weights"])).sum() * lw["loc_weight"]
            loss = loss + hm_loss + loc_loss
            tb[f"hm_loss_head_{i}"], tb[f"loc_loss_head_{i}"] = hm_loss.detach(), loc_loss.detach()
        tb["rpn_loss"] = loss.detach()
        return loss, tb

    # ---- decode + NMS --------------------------------------------------------------------
    @torch.no_grad()
    def generate_predicted_boxes(self, batch_size, pred_dicts):
        """center_head.py:252-303: top-K decode, score/range mask, class-agnostic rotated NMS per frame."""
        pp = _cfg(self.model_cfg["POST_PROCESSING"])
        nms = _cfg(pp["NMS_CONFIG"])
        dev = pred_dicts[0]["hm"].device
        limit = torch.tensor(pp["POST_CENTER_LIMIT_RANGE"], dtype=torch.float32, device=dev)
        ret = [dict(pred_boxes=[], pred_scores=[], pred_labels=[]) for _ in range(batch_size)]
        for i, pd in enumerate(pred_dicts):
            finals = topk_decode(pd["hm"].sigmoid(), pd["rot"][:, 0:1], pd["rot"][:, 1:2], pd["center"], pd["center_z"], pd["dim"].exp(),
                                 self.point_cloud_range, self.voxel_size, self.feature_map_stride, pp["MAX_OBJ_PER_SAMPLE"],
                                 pp["SCORE_THRESH"], limit)
            cmap = self.class_id_mapping_each_head[i].to(dev)
            for k, fd in enumerate(finals):
                labels = cmap[fd["pred_labels"].long()]
                boxes, scores = fd["pred_boxes"], fd["pred_scores"]
                if nms["NMS_TYPE"] != "circle_nms" and scores.shape[0] > 0:       # model_nms_utils.py:115-134
                    top_s, top_i = torch.topk(scores, k=min(nms["NMS_PRE_MAXSIZE"], scores.shape[0]))
                    keep, _ = getattr(iou3d_nms_utils, nms["NMS_TYPE"])(boxes[top_i][:, 0:7], top_s, nms["NMS_THRESH"])
                    sel = top_i[keep[:nms["NMS_POST_MAXSIZE"]]]
                    boxes, scores, labels = boxes[sel], scores[sel], labels[sel]
                ret[k]["pred_boxes"].append(boxes); ret[k]["pred_scores"].append(scores); ret[k]["pred_labels"].append(labels)
        for k in range(batch_size):
            ret[k]["pred_boxes"] = torch.cat(ret[k]["pred_boxes"], 0)
            ret[k]["pred_scores"] = torch.cat(ret[k]["pred_scores"], 0)
            ret[k]["pred_labels"] = torch.cat(ret[k]["pred_labels"], 0) + 1
        return ret

    @staticmethod
    def reorder_rois_for_refining(batch_size, pred_dicts):
        n = max(1, max(len(d["pred_boxes"]) for d in pred_dicts))
        pb = pred_dicts[0]["pred_boxes"]
        rois = pb.new_zeros((batch_size, n, pb.shape[-1]))
        scores, labels = pb.new_zeros((batch_size, n)), pb.new_zeros((batch_size, n)).long()
        for b in range(batch_size):
            m = len(pred_dicts[b]["pred_boxes"])
            rois[b, :m], scores[b, :m], labels[b, :m] = pred_dicts[b]["pred_boxes"], pred_dicts[b]["pred_scores"], pred_dicts[b]["pred_labels"]
        return rois, scores, labels

    def forward(self, data_dict, pred_dicts=None, fmap=None):
        """pred_dicts / fmap: head outputs computed elsewhere (the CUDA-graphed dense stack) and the feature map size."""
        if pred_dicts is None:
            x = data_dict.get("st_features_2d_map")
            if x is None:
                x = DenseMap.from_nchw(data_dict["st_features_2d"])
            x = self.shared_conv(x)
            pred_dicts = [head(x) for head in self.heads_list]
            fmap = (x.h, x.w)
        if self.training:
            self.forward_ret_dict["target_dicts"] = self.assign_targets(data_dict["gt_boxes"], fmap)
        self.forward_ret_dict["pred_dicts"] = pred_dicts
        if not self.training or self.predict_boxes_when_training:
            boxes = self.generate_predicted_boxes(data_dict["batch_size"], pred_dicts)
            if self.predict_boxes_when_training:
                rois, roi_scores, roi_labels = self.reorder_rois_for_refining(data_dict["batch_size"], boxes)
                data_dict.update(rois=rois, roi_scores=roi_scores, roi_labels=roi_labels, has_class_labels=True)
            if not self.training or not self.predict_boxes_when_training:
                data_dict["final_box_dicts"] = boxes          # no RoI head downstream on this path
        return data_dict
