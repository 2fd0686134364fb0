"""AnchorHeadSingleV2 on the GPU: its DenseConv2d layers against plain torch nn.Conv2d / BatchNorm2d with the same state_dict
(float64), and one training step (targets, losses, backward) end to end.  The torch logic after the convolutions is pinned to the
reference's own class in tests/test_anchor_head_cpu.py."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from cpd_b200.synth import synth_gt_boxes, synth_scan

pytestmark = pytest.mark.gpu


def _torch_layer(dim, out_dim):
    return nn.Sequential(nn.Conv2d(dim, dim, 3, padding=1, bias=True), nn.BatchNorm2d(dim), nn.ReLU(), nn.Conv2d(dim, out_dim, 1, bias=True))


def test_anchor_head_convs_match_torch_and_train_step(cuda):
    from cpd_b200 import anchor_head
    GRID, RANGE, NAMES = [1504, 1504, 40], [-75.2, -75.2, -2.0, 75.2, 75.2, 4.0], ["Vehicle", "Pedestrian", "Cyclist"]
    _cfg = anchor_head.default_cfg
    torch.manual_seed(0)
    head = anchor_head.AnchorHeadSingleV2(_cfg(), 1, 128, 3, NAMES, GRID, RANGE).to(cuda)
    n = head.num_anchors_per_location
    for m in head.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.uniform_(-0.2, 0.2); m.running_var.uniform_(0.5, 1.5); m.weight.data.uniform_(0.7, 1.3); m.bias.data.uniform_(-0.2, 0.2)
        if isinstance(m, (anchor_head.DenseConv2d,)):
            m.weight.data.normal_(0, 0.05)
    ref = nn.ModuleDict(dict(shared_conv=nn.Sequential(nn.Conv2d(128, 64, 3, padding=1, bias=True), nn.BatchNorm2d(64), nn.ReLU()),
                             conv_cls=_torch_layer(64, n * 3), conv_reg=_torch_layer(64, n * 2), conv_height=_torch_layer(64, n),
                             conv_dim=_torch_layer(64, n * 3), conv_ang=_torch_layer(64, n), conv_dir_cls=nn.Conv2d(128, n * 2, 1)))
    ref.load_state_dict(head.state_dict())                       # identical names and (cout, cin, kh, kw) shapes
    ref = ref.to(cuda).double().eval()
    B, H, W = 2, 60, 52
    feat = torch.randn(B, 128, H, W, device=cuda)
    pts = [torch.from_numpy(synth_scan(8000, 5 + i)).to(cuda) for i in range(B)]
    gt = torch.from_numpy(np.stack([synth_gt_boxes(30, 21 + i) for i in range(B)])).float().to(cuda)
    # shrink the anchor grid to the test map: same generator, smaller grid
    head.grid_size = [W * 8, H * 8, 40]
    head.anchors_root, _ = anchor_head.generate_anchors(head.model_cfg["ANCHOR_GENERATOR_CONFIG"], head.grid_size, head.range, 7)
    head.voxel_size = (head.range[3] - head.range[0]) / head.grid_size[0]
    head.eval()
    with torch.no_grad():
        bd = head(dict(st_features_2d=feat, points=pts, batch_size=B))
        x = ref["shared_conv"](feat.double())
        want_cls = ref["conv_cls"](x).permute(0, 2, 3, 1)
        want_box = torch.cat([ref[k](x) for k in ("conv_reg", "conv_height", "conv_dim", "conv_ang")], 1).permute(0, 2, 3, 1)
        want_dir = ref["conv_dir_cls"](feat.double()).permute(0, 2, 3, 1)
        mask = head.get_anchor_mask(torch.cat([p[:, :2] for p in pts]), (H, W))
    assert 0 < int(mask.sum()) <= H * W
    for key, want in (("cls_preds", want_cls), ("box_preds", want_box), ("dir_cls_preds", want_dir)):
        got, w = head.forward_ret_dict[key].double(), want[:, mask, :]
        assert got.shape == w.shape and float((got - w).abs().max()) <= 1e-4 * max(1.0, float(w.abs().max())), key
    assert bd["batch_box_preds"].shape == (B, int(mask.sum()) * n, 7) and torch.isfinite(bd["batch_box_preds"]).all()
    # one training step: targets on the device, three losses, gradients to every parameter
    head.train()
    head(dict(st_features_2d=feat, points=pts, gt_boxes=gt, batch_size=B))
    loss, tb = head.get_loss()
    assert torch.isfinite(loss) and set(tb) == {"rpn_loss_cls", "rpn_loss_loc", "rpn_loss_dir", "rpn_loss"}
    loss.backward()
    for name, p in head.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
    assert float(head.shared_conv[0].weight.grad.abs().sum()) > 0
