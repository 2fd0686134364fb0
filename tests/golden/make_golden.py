"""Regenerates the committed golden fixtures.  Run in the BUILD container (where
/root/reference exists): `python tests/golden/make_golden.py`.

  iou_ref.npz      boxes + IoU matrix computed by the REFERENCE's own iou3d_cpu.cpp
                   (compiled alone into oracle/_ref by `make -C oracle ref`) -- pins the oracle.
  voxel_small.npz  a 3000-point cloud + the voxelizer oracle's output (self-generated:
                   the reference ships no voxelizer source or vectors -- "parity unpinned").
  spconv_small.npz sparse-conv oracle outputs on a small random active set, cross-checked at
                   generation time against torch.nn.functional.conv3d (independent pin).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from cpd_b200.synth import PC_RANGE, VOXEL_SIZE, synth_nms_boxes, synth_scan  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    O.build(ref=True)
    ref = O.ref_cpu_module()
    assert ref is not None, "oracle/_ref/iou3d_ref_cpu.so missing (run make -C oracle ref)"
    a, _ = synth_nms_boxes(96, 1, clusters=12)
    b, _ = synth_nms_boxes(80, 2, clusters=12)
    b[:40] = a[:40] + np.random.default_rng(3).normal(0, 0.05, (40, 7)).astype(np.float32)
    out = torch.zeros(96, 80)
    ref.boxes_iou_bev_cpu(torch.from_numpy(a), torch.from_numpy(b), out)
    np.savez_compressed(os.path.join(HERE, "iou_ref.npz"), a=a, b=b, iou=out.numpy())

    pts = synth_scan(3000, 11)
    pts[5] = [80.0, 0, 0, 0.5, 0]          # out of range in x
    pts[6] = [0.0, 0.0, 4.0, 0.5, 0]       # z == max edge: dropped by the voxelizer
    pts[7] = [75.2, 1.0, 0.0, 0.5, 0]      # x == max edge
    v, c, n = O.voxelize(pts, PC_RANGE, VOXEL_SIZE, 5, 1000000)
    np.savez_compressed(os.path.join(HERE, "voxel_small.npz"), points=pts, voxels=v, coords=c, num=n)

    rng = np.random.default_rng(5)
    shape = [9, 14, 12]
    cells = rng.choice(2 * 9 * 14 * 12, 500, replace=False)
    coords = np.stack([cells // (9 * 14 * 12), (cells // (14 * 12)) % 9, (cells // 12) % 14, cells % 12], 1).astype(np.int32)
    x = rng.normal(0, 1, (500, 8)).astype(np.float32)
    w = rng.normal(0, 0.2, (16, 3, 3, 3, 8)).astype(np.float32)
    bias = rng.normal(0, 0.1, 16).astype(np.float32)
    rb = O.rulebook_subm(coords, shape, 3)
    y_subm = O.spconv_fwd(x, w, bias, rb)
    rs = O.rulebook_strided(coords, shape, 3, 2, 1)
    y_str = O.spconv_fwd(x, w, None, rs)
    # independent pin: dense conv3d
    dense = torch.zeros(2, 8, *shape)
    dense[coords[:, 0], :, coords[:, 1], coords[:, 2], coords[:, 3]] = torch.from_numpy(x)
    wt = torch.from_numpy(w).permute(0, 4, 1, 2, 3).contiguous()
    ref_subm = torch.nn.functional.conv3d(dense, wt, torch.from_numpy(bias), padding=1)
    got = ref_subm[coords[:, 0], :, coords[:, 1], coords[:, 2], coords[:, 3]].numpy()
    assert np.abs(got - y_subm).max() < 1e-4
    ref_str = torch.nn.functional.conv3d(dense, wt, None, stride=2, padding=1)
    oc = rs.out_coords
    got = ref_str[oc[:, 0], :, oc[:, 1], oc[:, 2], oc[:, 3]].numpy()
    assert np.abs(got - y_str).max() < 1e-4
    np.savez_compressed(os.path.join(HERE, "spconv_small.npz"), coords=coords, shape=np.array(shape), x=x, w=w,
                        bias=bias, y_subm=y_subm, y_strided=y_str, out_coords=oc, out_shape=np.array(rs.out_shape))
    print("golden fixtures written")


if __name__ == "__main__":
    main()


# ----------------------------------------------------------------------------------------------
# CenterHead pieces: run the REFERENCE's own Python (functions lifted by name out of its source
# files with ast, because `import cpd.models` cannot resolve here -- SURVEY.md section 8c) and
# store inputs + outputs.
# ----------------------------------------------------------------------------------------------
def _lift(path, names, glb):
    import ast
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            code = ast.get_source_segment(src, node)
            import textwrap
            exec(textwrap.dedent(code), glb)
            out[node.name] = glb[node.name]
    return out


def head_golden():
    import importlib.util
    ref = "/root/reference/cpd"
    spec = importlib.util.spec_from_file_location("ref_centernet_utils", f"{ref}/models/model_utils/centernet_utils.py")
    cu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cu)
    glb = {"torch": torch, "np": np, "centernet_utils": cu, "nn": torch.nn}
    fn = _lift(f"{ref}/models/dense_heads/center_head.py", {"assign_target_of_single_head"}, glb)["assign_target_of_single_head"]
    losses = _lift(f"{ref}/utils/loss_utils.py", {"neg_loss_cornernet", "_reg_loss", "_gather_feat", "_transpose_and_gather_feat"}, glb)

    class Self:
        point_cloud_range = [-75.2, -75.2, -2, 75.2, 75.2, 4]
        voxel_size = [0.1, 0.1, 0.15]

    from cpd_b200.synth import synth_gt_boxes
    gt = torch.from_numpy(synth_gt_boxes(40, 3))
    gt[5, 3] = 0.0                     # degenerate box: skipped by the reference
    gt[7, :2] = torch.tensor([75.0, -75.1])   # near the border: clipped gaussian
    hm, rb, inds, mask = fn(Self(), 3, gt.clone(), [188, 188], 8, num_max_objs=500, gaussian_overlap=0.1, min_radius=2)
    g = torch.Generator().manual_seed(0)
    S = 40                                             # small maps keep the fixture small
    pred_hm = torch.rand(2, 3, S, S, generator=g).clamp(1e-4, 1 - 1e-4)
    tgt_hm = torch.stack([hm[:, :S, :S], hm[:, 60:60 + S, 60:60 + S]], 0).contiguous()
    tgt_hm[0, 1, 3, 4] = 1.0
    focal = losses["neg_loss_cornernet"](pred_hm, tgt_hm)
    out8 = torch.randn(2, 8, S, S, generator=g)
    ind2 = torch.stack([inds[:64] % (S * S), inds[:64].flip(0) % (S * S)], 0)
    mask2 = torch.stack([mask[:64], mask[:64].flip(0)], 0)
    tb2 = torch.stack([rb[:64], rb[:64].flip(0)], 0)
    pred = losses["_transpose_and_gather_feat"](out8, ind2)
    reg = losses["_reg_loss"](pred, tb2, mask2)
    # decode
    heat = torch.rand(2, 3, S, S, generator=g) ** 8
    rot = torch.randn(2, 2, S, S, generator=g)
    ctr, cz = torch.rand(2, 2, S, S, generator=g), torch.randn(2, 1, S, S, generator=g)
    dim = torch.rand(2, 3, S, S, generator=g) * 3 + 0.5
    dec = cu.decode_bbox_from_heatmap(heatmap=heat, rot_cos=rot[:, 0:1], rot_sin=rot[:, 1:2], center=ctr, center_z=cz, dim=dim,
                                      point_cloud_range=Self.point_cloud_range, voxel_size=Self.voxel_size, feature_map_stride=8,
                                      K=100, circle_nms=False, score_thresh=0.1,
                                      post_center_limit_range=torch.tensor(Self.point_cloud_range).float())
    np.savez_compressed(os.path.join(HERE, "center_head_ref.npz"), gt=gt.numpy(), heatmap=hm.numpy(), ret_boxes=rb.numpy(),
                        inds=inds.numpy(), mask=mask.numpy(), pred_hm=pred_hm.numpy(), tgt_hm=tgt_hm.numpy(), focal=focal.numpy(),
                        out8=out8.numpy(), ind2=ind2.numpy(), mask2=mask2.numpy(), tb2=tb2.numpy(), reg=reg.numpy(),
                        heat=heat.numpy(), rot=rot.numpy(), ctr=ctr.numpy(), cz=cz.numpy(), dim=dim.numpy(),
                        dec_boxes0=dec[0]["pred_boxes"].numpy(), dec_scores0=dec[0]["pred_scores"].numpy(),
                        dec_labels0=dec[0]["pred_labels"].numpy(), dec_boxes1=dec[1]["pred_boxes"].numpy())
    print("center_head_ref.npz written:", int(mask.sum()), "valid boxes,", dec[0]["pred_boxes"].shape[0], "decoded")


if __name__ == "__main__":
    head_golden()
