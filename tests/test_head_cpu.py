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


def test_multi_head_targets_compact_per_head_like_the_reference():
    """center_head.py:188-207: each head first collects ITS boxes (order kept, class made head-local), then the first
    NUM_MAX_OBJS of them fill slots 0, 1, ... -- also when other heads' boxes and padding rows sit in between, when there are
    more boxes than slots, and with extra (velocity) columns.  Compared with the per-frame routine on hand-compacted boxes."""
    torch.manual_seed(5)
    names = ["Vehicle", "Pedestrian", "Cyclist"]
    cfg = dict(bev.DEFAULT_HEAD_CFG)
    cfg["CLASS_NAMES_EACH_HEAD"] = [["Vehicle", "Cyclist"], ["Pedestrian"]]
    cfg["TARGET_ASSIGNER_CONFIG"] = dict(cfg["TARGET_ASSIGNER_CONFIG"], NUM_MAX_OBJS=9)
    head = bev.CenterHead(cfg, 1, 512, 3, names, [1504, 1504, 40], RANGE, VS)
    gt0 = torch.from_numpy(G["gt"])[:24].clone()
    gt0[:, 7] = torch.tensor([1, 2, 3, 0, 2, 1, 1, 3, 0, 2, 3, 1] * 2, dtype=gt0.dtype)      # interleaved classes and padding
    gt = torch.stack([gt0, gt0.flip(0).clone()], 0)
    gt10 = torch.cat([gt[..., :7], torch.randn(2, 24, 2), gt[..., 7:]], -1)                    # + (vx, vy) before the class
    for g in (gt, gt10):
        ret = head.assign_targets(g, (188, 188))
        for h, local_names in enumerate(cfg["CLASS_NAMES_EACH_HEAD"]):
            for b in range(2):
                rows = [r.clone() for r in g[b] if int(r[-1]) >= 1 and names[int(r[-1]) - 1] in local_names]
                for r in rows:
                    r[-1] = local_names.index(names[int(r[-1]) - 1]) + 1
                sel = torch.stack(rows)[:9]
                sel8 = torch.cat([sel[:, :7], sel[:, -1:]], 1)
                hm, rb, inds, mask = bev.assign_targets_single(sel8, len(local_names), [188, 188], 8, RANGE, VS, 9, 0.1, 2)
                assert torch.equal(ret["heatmaps"][h][b], hm) and torch.equal(ret["inds"][h][b], inds) and torch.equal(ret["masks"][h][b], mask)
                assert torch.equal(ret["target_boxes"][h][b][:, :8], rb)
                if g.shape[-1] > 8:
                    assert torch.equal(ret["target_boxes"][h][b][:len(sel), 8:], sel[:, 7:-1] * mask[:len(sel), None])
                assert int(mask.sum()) > 0
