"""CenterHead host logic (device-vectorised target assignment, losses, top-K decode) against the
REFERENCE's own Python, captured in tests/golden/center_head_ref.npz by tests/golden/make_golden.py
(center_head.py:103-157, loss_utils.py:265-346, centernet_utils.py:136-216).  Pure torch => runs on CPU."""
import os

import numpy as np
import torch

from cpd_b200 import bev

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "center_head_ref.npz"))
RANGE, VS = [-75.2, -75.2, -2, 75.2, 75.2, 4], [0.1, 0.1, 0.15]


def test_target_assignment_matches_reference_loop():
    gt = torch.from_numpy(G["gt"])
    hm, rb, inds, mask = bev.assign_targets_single(gt, 3, [188, 188], 8, RANGE, VS, 500, 0.1, 2)
    assert np.array_equal(mask.numpy(), G["mask"]) and np.array_equal(inds.numpy(), G["inds"])
    assert np.allclose(rb.numpy(), G["ret_boxes"], atol=1e-6)
    assert np.allclose(hm.numpy(), G["heatmap"], atol=1e-6)
    assert int(mask.sum()) == 39                      # the degenerate box is skipped
    e = bev.assign_targets_single(gt[:0], 3, [188, 188], 8, RANGE, VS)
    assert float(e[0].abs().sum()) == 0 and int(e[3].sum()) == 0


def test_losses_match_reference():
    f = bev.focal_loss_centernet(torch.from_numpy(G["pred_hm"]), torch.from_numpy(G["tgt_hm"]))
    assert np.allclose(f.numpy(), G["focal"], rtol=1e-5)
    r = bev.reg_loss_centernet(torch.from_numpy(G["out8"]), torch.from_numpy(G["mask2"]), torch.from_numpy(G["ind2"]),
                               torch.from_numpy(G["tb2"]))
    assert np.allclose(r.numpy(), G["reg"], rtol=1e-5, atol=1e-6)
    z = bev.focal_loss_centernet(torch.from_numpy(G["pred_hm"]), torch.zeros_like(torch.from_numpy(G["tgt_hm"])))
    assert torch.isfinite(z) and z > 0                # num_pos == 0 branch


def test_topk_decode_matches_reference():
    t = lambda k: torch.from_numpy(G[k])
    rot = t("rot")
    out = bev.topk_decode(t("heat"), rot[:, 0:1], rot[:, 1:2], t("ctr"), t("cz"), t("dim"), RANGE, VS, 8, 100, 0.1,
                          torch.tensor(RANGE).float())
    assert np.allclose(out[0]["pred_boxes"].numpy(), G["dec_boxes0"], atol=1e-5)
    assert np.allclose(out[0]["pred_scores"].numpy(), G["dec_scores0"])
    assert np.array_equal(out[0]["pred_labels"].numpy(), G["dec_labels0"])
    assert np.allclose(out[1]["pred_boxes"].numpy(), G["dec_boxes1"], atol=1e-5)


def test_state_dict_names_match_reference_layout():
    """Parameter names/shapes of the mirrors equal those of the reference modules (checkpoint drop-in)."""
    bb = bev.BaseBEVBackbone(dict(LAYER_NUMS=[5, 5], LAYER_STRIDES=[1, 2], NUM_FILTERS=[128, 256], UPSAMPLE_STRIDES=[1, 2],
                                  NUM_UPSAMPLE_FILTERS=[256, 256]), 1, 256)
    sd = bb.state_dict()
    assert tuple(sd["blocks.0.1.weight"].shape) == (128, 256, 3, 3) and tuple(sd["blocks.1.1.weight"].shape) == (256, 128, 3, 3)
    assert tuple(sd["blocks.0.16.weight"].shape) == (128, 128, 3, 3) and "blocks.0.17.running_mean" in sd
    assert tuple(sd["deblocks.0.0.weight"].shape) == (128, 256, 1, 1) and tuple(sd["deblocks.1.0.weight"].shape) == (256, 256, 2, 2)
    head = bev.CenterHead(None, 1, 512, 3, ["Vehicle", "Pedestrian", "Cyclist"], [1504, 1504, 40], RANGE, VS)
    hs = head.state_dict()
    assert tuple(hs["shared_conv.0.weight"].shape) == (64, 512, 3, 3) and "shared_conv.0.bias" in hs
    assert tuple(hs["heads_list.0.hm.1.weight"].shape) == (3, 64, 3, 3) and tuple(hs["heads_list.0.center_z.1.weight"].shape) == (1, 64, 3, 3)
    assert float(hs["heads_list.0.hm.1.bias"][0]) == np.float32(-2.19)
    assert "heads_list.0.dim.0.1.running_var" in hs


def test_batched_target_assignment_equals_per_frame():
    """assign_targets_batched (one set of launches for the whole batch) == assign_targets_single per frame, bit for bit."""
    torch.manual_seed(3)
    gt0 = torch.from_numpy(G["gt"])
    frames = [gt0, gt0.flip(0).clone(), gt0.clone()]
    frames[1][:, 0:2] += 3.7                                   # moved boxes
    frames[2][10:, :] = 0                                      # padded frame (class 0 rows)
    frames[2][3, 3] = 0.0                                      # a degenerate box
    gt = torch.stack(frames, 0)
    hb, rb, ib, mb = bev.assign_targets_batched(gt, 3, [188, 188], 8, RANGE, VS, 500, 0.1, 2)
    for b in range(gt.shape[0]):
        h1, r1, i1, m1 = bev.assign_targets_single(gt[b], 3, [188, 188], 8, RANGE, VS, 500, 0.1, 2)
        assert torch.equal(hb[b], h1) and torch.equal(rb[b], r1) and torch.equal(ib[b], i1) and torch.equal(mb[b], m1)
    e = bev.assign_targets_batched(gt[:, :0], 3, [188, 188], 8, RANGE, VS)
    assert float(e[0].abs().sum()) == 0 and int(e[3].sum()) == 0 and tuple(e[0].shape) == (3, 3, 188, 188)
    capped = bev.assign_targets_batched(gt, 3, [188, 188], 8, RANGE, VS, num_max_objs=7)
    assert tuple(capped[1].shape) == (3, 7, 8) and int(capped[3].sum()) <= 21
