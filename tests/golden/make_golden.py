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
