"""Module with the five entry points of the reference's pybind extension
``cpd.ops.iou3d_nms.iou3d_nms_cuda`` (cpd/ops/iou3d_nms/src/iou3d_nms_api.cpp:11-17), same
names, argument order and in-place output convention, backed by libcpd_b200.so.

Differences, all at the error boundary only: bad inputs raise instead of ``exit(-1)``
(iou3d_nms.cpp:14-26).  ``keep`` may be the reference's CPU LongTensor
(iou3d_nms_utils.py:116) -- it is then filled by one D2H copy -- or a CUDA LongTensor,
in which case nothing leaves the device.
"""
import torch

from . import ops


def _check(boxes, name):
    if not boxes.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor")
    if not boxes.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if boxes.dim() != 2 or boxes.shape[1] != 7:
        raise ValueError(f"{name} must be (N, 7)")


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    _check(boxes_a, "boxes_a"); _check(boxes_b, "boxes_b")
    ops.iou_bev(boxes_a, boxes_b, out=ans_overlap, overlap=True)
    return 1


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    _check(boxes_a, "boxes_a"); _check(boxes_b, "boxes_b")
    ops.iou_bev(boxes_a, boxes_b, out=ans_iou, overlap=False)
    return 1


def _nms(boxes, keep, thresh, rotated):
    _check(boxes, "boxes")
    k, n = ops.nms(boxes, thresh, rotated=rotated)
    num = int(n.item())
    keep[:num] = k[:num].to(keep.device)
    return num


def nms_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(boxes, keep, nms_overlap_thresh, True)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(boxes, keep, nms_overlap_thresh, False)


def boxes_iou_bev_cpu(boxes_a, boxes_b, ans_iou):
    """The reference's CPU entry (iou3d_cpu.h:9) takes CPU tensors.  cpd_b200 has no CPU
    arithmetic: the boxes make a round trip through the GPU kernel."""
    dev = torch.device("cuda", torch.cuda.current_device())
    out = ops.iou_bev(boxes_a.to(dev).contiguous(), boxes_b.to(dev).contiguous())
    ans_iou.copy_(out.to(ans_iou.device))
    return 1
