"""Host-side mirror of cpd/ops/iou3d_nms/iou3d_nms_utils.py (same names and argument
meaning), with the on-device keep list used directly: no CPU LongTensor, no D2H of the
mask, one small D2H only where the reference API returns a python-sized result."""
import torch

from . import ops


def boxes_iou_bev(boxes_a, boxes_b):
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    return ops.iou_bev(boxes_a.contiguous(), boxes_b.contiguous())


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """iou3d_nms_utils.py:67-100: BEV overlap (our kernel) x height overlap / union volume."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    a_max = (boxes_a[:, 2] + boxes_a[:, 5] / 2).view(-1, 1)
    a_min = (boxes_a[:, 2] - boxes_a[:, 5] / 2).view(-1, 1)
    b_max = (boxes_b[:, 2] + boxes_b[:, 5] / 2).view(1, -1)
    b_min = (boxes_b[:, 2] - boxes_b[:, 5] / 2).view(1, -1)
    overlaps_bev = ops.iou_bev(boxes_a.contiguous(), boxes_b.contiguous(), overlap=True)
    overlaps_h = torch.clamp(torch.min(a_max, b_max) - torch.max(a_min, b_min), min=0)
    overlaps_3d = overlaps_bev * overlaps_h
    vol_a = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5]).view(-1, 1)
    vol_b = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5]).view(1, -1)
    return overlaps_3d / torch.clamp(vol_a + vol_b - overlaps_3d, min=1e-6)


def _nms(boxes, scores, thresh, pre_maxsize, rotated):
    assert boxes.shape[1] == 7
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    boxes = boxes[order].contiguous()
    keep, n_keep = ops.nms(boxes, thresh, rotated=rotated)
    return order[keep[:int(n_keep.item())]].contiguous(), None


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    return _nms(boxes, scores, thresh, pre_maxsize, True)


def nms_normal_gpu(boxes, scores, thresh, **kwargs):
    return _nms(boxes, scores, thresh, None, False)
