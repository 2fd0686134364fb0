"""Oracle-side interpreter: runs a cpd_b200 backbone module tree with the CPU oracle
(eval-mode BatchNorm folded to its affine form).  TEST INFRASTRUCTURE ONLY -- used by
tests/ and by bench.py's cpu_baseline / `--impl reference` legs, never by the product."""
import numpy as np
import torch
import torch.nn as nn

from . import oracle as O


def _bn_affine(m):
    s = (m.weight / torch.sqrt(m.running_var + m.eps)).detach().cpu().numpy()
    return s, m.bias.detach().cpu().numpy() - m.running_mean.detach().cpu().numpy() * s


def _conv(m, feats, coords, shape, cache):
    key = (m.indice_key, len(coords), tuple(shape))
    if key not in cache:
        cache[key] = (O.rulebook_subm(coords, shape, m.kernel_size) if m.subm else
                      O.rulebook_strided(coords, shape, m.kernel_size, m.stride, m.padding))
    rb = cache[key]
    y = O.spconv_fwd(feats, m.weight.detach().cpu().numpy(), None if m.bias is None else m.bias.detach().cpu().numpy(), rb)
    return y, rb.out_coords, rb.out_shape


def run_sequential(seq, feats, coords, shape, cache):
    from cpd_b200 import sparse as sp
    from cpd_b200.backbone import SparseBasicBlock
    for m in seq._modules.values():
        if isinstance(m, sp.SparseSequential):
            feats, coords, shape = run_sequential(m, feats, coords, shape, cache)
        elif isinstance(m, SparseBasicBlock):
            ident = feats
            y, _, _ = _conv(m.conv1, feats, coords, shape, cache)
            s, b = _bn_affine(m.bn1)
            y = np.maximum(y * s + b, 0)
            y, _, _ = _conv(m.conv2, y, coords, shape, cache)
            s, b = _bn_affine(m.bn2)
            feats = np.maximum(y * s + b + ident, 0)
        elif isinstance(m, sp.SparseConvolution):
            feats, coords, shape = _conv(m, feats, coords, shape, cache)
        elif isinstance(m, nn.BatchNorm1d):
            s, b = _bn_affine(m)
            feats = feats * s + b
        elif isinstance(m, nn.ReLU):
            feats = np.maximum(feats, 0)
        else:
            raise TypeError(type(m))
    return feats, coords, shape


def decompose(feats, coords, shape, i=0):
    """VoxelBackBone8x.decompose_tensor (spconv_backbone.py:241-260): strict `<` on both slab
    bounds, so x == i*W/4 is dropped -- a reference quirk that is reproduced, not fixed."""
    w4 = shape[2] // 4
    keep = (i * w4 < coords[:, 3]) & (coords[:, 3] < (i + 1) * w4)
    c = coords[keep].copy()
    c[:, 3] -= i * w4
    return feats[keep], c, [shape[0], shape[1], w4]


def backbone_forward(net, frames, pc_range, voxel_size, max_pts=5, max_voxels=1000000, sfx="", eval_wide=False):
    """frames: list of (n_i, C) numpy clouds -> (features, coords (M,4), shape, stage dict).
    eval_wide: VoxelBackBone8x's eval path (spconv_backbone.py:332-393): the tower runs on a
    [D, H, 4*W] grid and the outputs are cut back with decompose()."""
    feats, coords = [], []
    for b, pts in enumerate(frames):
        v, c, n = O.voxelize(pts, pc_range, voxel_size, max_pts, max_voxels)
        feats.append(O.mean_vfe(v, n))
        coords.append(np.concatenate([np.full((len(c), 1), b, np.int32), c], 1))
    feats, coords = np.concatenate(feats, 0), np.concatenate(coords, 0)
    shape, cache, stages = list(net.sparse_shape), {}, {}
    if eval_wide:
        shape[2] *= 4
    names = ["conv_input", "conv1", "conv2", "conv3", "conv4"] + (["conv_out"] if sfx == "" else [])
    for name in names:
        feats, coords, shape = run_sequential(getattr(net, name + sfx if name != "conv_out" else name), feats, coords, shape, cache)
        stages[name] = (feats, coords, list(shape))
    if eval_wide:
        feats, coords, shape = decompose(feats, coords, shape, 0)
    return feats, coords, shape, stages


def backbone_forward_tta(net, stage_frames, pc_range, voxel_size, max_pts=5, max_voxels=1000000):
    """VoxelBackBone8x's multi-stage eval path (spconv_backbone.py:332-393): stage_frames[i] = list of clouds (one per frame) of
    TTA stage i.  Stage i's voxels are shifted by i * W along x, ONE tower pass runs on the [D, H, 4 W] grid (neighbouring slabs
    touch: voxels at a slab border see the next stage's voxels -- reference behaviour, reproduced) and every output is cut back with
    decompose().  -> list over stages of dict(out=(f, c, shape), x_conv3=..., x_conv4=...)."""
    w = net.sparse_shape[2]
    feats, coords = [], []
    for i, frames in enumerate(stage_frames):
        for b, pts in enumerate(frames):
            v, c, n = O.voxelize(pts, pc_range, voxel_size, max_pts, max_voxels)
            feats.append(O.mean_vfe(v, n))
            cc = np.concatenate([np.full((len(c), 1), b, np.int32), c], 1)
            cc[:, 3] += i * w
            coords.append(cc)
    feats, coords = np.concatenate(feats, 0), np.concatenate(coords, 0)
    shape, cache, stages = [net.sparse_shape[0], net.sparse_shape[1], 4 * w], {}, {}
    for name in ["conv_input", "conv1", "conv2", "conv3", "conv4", "conv_out"]:
        feats, coords, shape = run_sequential(getattr(net, name), feats, coords, shape, cache)
        stages[name] = (feats, coords, list(shape))
    return [dict(out=decompose(*stages["conv_out"], i), x_conv3=decompose(*stages["conv3"], i), x_conv4=decompose(*stages["conv4"], i))
            for i in range(len(stage_frames))]


def bev_dense(feats, coords, batch, shape):
    d = O.dense(feats, coords, batch, shape)
    n, c, dd, h, w = d.shape
    return d.reshape(n, c * dd, h, w)


# ----------------------------------------------------------------------------------------------
# CPU training step of the hot-path detector (bench.py cpu_baseline / --impl reference only):
# oracle voxelizer + oracle sparse conv fwd/bwd wrapped for autograd, torch.nn (CPU, MKLDNN)
# for the dense BEV backbone / CenterHead exactly as the reference builds them.
# ----------------------------------------------------------------------------------------------
class _OracleConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, rb):
        ctx.rb = rb
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        y = O.spconv_fwd(x.detach().numpy(), w.detach().numpy(), None if b is None else b.detach().numpy(), rb)
        return torch.from_numpy(y)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx, dw, db = O.spconv_bwd(x.detach().numpy(), w.detach().numpy(), dy.contiguous().numpy(), ctx.rb, need_bias=ctx.has_bias)
        return torch.from_numpy(dx), torch.from_numpy(dw), (torch.from_numpy(db) if ctx.has_bias else None), None


def _rb(m, coords, shape, cache):
    key = (m.indice_key, len(coords), tuple(shape))
    if key not in cache:
        cache[key] = (O.rulebook_subm(coords, shape, m.kernel_size) if m.subm else
                      O.rulebook_strided(coords, shape, m.kernel_size, m.stride, m.padding))
    return cache[key]


def run_sequential_torch(seq, feats, coords, shape, cache):
    """Like run_sequential but on torch CPU tensors with autograd and the modules' own BatchNorm state."""
    from cpd_b200 import sparse as sp
    from cpd_b200.backbone import SparseBasicBlock
    for m in seq._modules.values():
        if isinstance(m, sp.SparseSequential):
            feats, coords, shape = run_sequential_torch(m, feats, coords, shape, cache)
        elif isinstance(m, SparseBasicBlock):
            rb = _rb(m.conv1, coords, shape, cache)
            y = m.relu(m.bn1(_OracleConvFn.apply(feats, m.conv1.weight, m.conv1.bias, rb)))
            y = m.bn2(_OracleConvFn.apply(y, m.conv2.weight, m.conv2.bias, rb))
            feats = m.relu(y + feats)
        elif isinstance(m, sp.SparseConvolution):
            rb = _rb(m, coords, shape, cache)
            feats = _OracleConvFn.apply(feats, m.weight, m.bias, rb)
            coords, shape = rb.out_coords, rb.out_shape
        else:
            feats = m(feats)
    return feats, coords, shape


def torch_bev_reference(bb, head):
    """The reference's dense module structure in plain torch.nn (nn.Conv2d / nn.ConvTranspose2d /
    nn.BatchNorm2d, base_bev_backbone.py:31-59, center_head.py:11-45,73-94) sharing the mirrors' weights."""
    from cpd_b200 import bev
    pairs = []                     # (mirror module, plain torch module) for every parametrised layer

    def seq(ds):
        layers = []
        for m in ds:
            if isinstance(m, bev.DenseSequential):
                layers.append(seq(m))
            elif isinstance(m, bev.DenseConv2d):
                c = nn.Conv2d(m.in_channels, m.out_channels, m.k, m.stride, m.padding, bias=m.bias is not None)
                c.weight.data.copy_(m.weight.data)
                if m.bias is not None:
                    c.bias.data.copy_(m.bias.data)
                layers.append(c)
                pairs.append((m, c))
            elif isinstance(m, bev.DenseConvTranspose2d):
                c = nn.ConvTranspose2d(m.in_channels, m.out_channels, m.s, stride=m.s, bias=False)
                c.weight.data.copy_(m.weight.data)
                layers.append(c)
                pairs.append((m, c))
            elif isinstance(m, nn.BatchNorm2d):
                b = nn.BatchNorm2d(m.num_features, eps=m.eps, momentum=m.momentum)
                b.load_state_dict(m.state_dict())
                layers.append(b)
                pairs.append((m, b))
            elif isinstance(m, nn.ReLU):
                layers.append(nn.ReLU())
            else:
                layers.append(nn.Identity())
        return nn.Sequential(*layers)
    blocks, deblocks = [seq(b) for b in bb.blocks], [seq(d) for d in bb.deblocks]
    shared = seq(head.shared_conv)
    heads = [{n: seq(getattr(h, n)) for n in h.sep_head_dict} for h in head.heads_list]

    def run(x):
        ups = []
        for b, d in zip(blocks, deblocks):
            x = b(x)
            ups.append(d(x))
        f = torch.cat(ups, 1)
        s = shared(f)
        return f, [{n: m(s) for n, m in hd.items()} for hd in heads]
    mods = nn.ModuleList(blocks + deblocks + [shared] + [m for hd in heads for m in hd.values()])
    run.pairs = pairs
    return run, mods


class CpuDetector:
    """CPU restatement of cpd_b200.detector.CPDHotPathDetector's training step."""

    def __init__(self, det):
        import copy
        self.det = copy.deepcopy(det).cpu().train()
        self.run_dense, self.dense_mods = torch_bev_reference(self.det.backbone_2d, self.det.dense_head)
        self.dense_mods.train()
        self.opt = None

    def _tower(self, sfx, frames, names):
        det = self.det
        feats, coords = [], []
        for b, pts in enumerate(frames):
            v, c, n = O.voxelize(pts, det.pc_range, det.voxel_size, det.max_pts, det.max_voxels)
            feats.append(O.mean_vfe(v, n))
            coords.append(np.concatenate([np.full((len(c), 1), b, np.int32), c], 1))
        f, co = torch.from_numpy(np.concatenate(feats, 0)), np.concatenate(coords, 0)
        shape, cache, outs = list(det.backbone_3d.sparse_shape), {}, []
        for name in names:
            mod = getattr(det.backbone_3d, name if name == "conv_out" else name + sfx)
            f, co, shape = run_sequential_torch(mod, f, co, shape, cache)
            outs.append(f)
        return f, co, shape, outs

    def train_step(self, frames, frames1, gt_boxes, optimizer=False):
        """forward + backward (+ grad-clip + Adam update with optimizer=True, as bench.py's CUDA arm does)."""
        from cpd_b200 import bev
        det = self.det
        f, co, shape, _ = self._tower("", frames, ["conv_input", "conv1", "conv2", "conv3", "conv4", "conv_out"])
        bs = len(frames)
        idx = torch.from_numpy(co).long()
        dense = f.new_zeros(bs, shape[0], shape[1], shape[2], f.shape[1])          # (B, D, H, W, C) then permute, like .dense()
        dense = dense.index_put((idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]), f).permute(0, 4, 1, 2, 3)
        bev_in = dense.reshape(bs, f.shape[1] * shape[0], shape[1], shape[2])
        _, heads = self.run_dense(bev_in)
        head = det.dense_head
        head.forward_ret_dict["pred_dicts"] = heads
        head.forward_ret_dict["target_dicts"] = head.assign_targets(torch.from_numpy(gt_boxes), bev_in.shape[2:])
        loss, _ = head.get_loss()
        if frames1 is not None:
            _, _, _, outs = self._tower("_2", frames1, ["conv_input", "conv1", "conv2", "conv3", "conv4"])
            loss = loss + 1e-3 * sum(o.square().mean() for o in outs[1:])
        params = [p for p in list(det.parameters()) + list(self.dense_mods.parameters()) if p.requires_grad]
        for p in params:
            p.grad = None
        loss.backward()
        if optimizer:
            live = [p for p in params if p.grad is not None]
            if self.opt is None:
                self.opt = torch.optim.Adam(live, lr=1e-4)
            torch.nn.utils.clip_grad_norm_(live, 10.0)
            self.opt.step()
        return float(loss)

    def named_grads(self):
        """Gradients of the last train_step keyed by the parameter names of CPDHotPathDetector (the dense layers'
        gradients live on their plain-torch twins)."""
        twin = {}
        for mirror, ref in self.run_dense.pairs:
            for n, p in mirror.named_parameters(recurse=False):
                twin[id(p)] = getattr(ref, n)
        out = {}
        for name, p in self.det.named_parameters():
            q = twin.get(id(p), p)
            if q.grad is not None:
                out[name] = q.grad.detach().clone()
        return out
