"""Drop-in namespaces so that `cpd/models/**` and `cpd/datasets/processor/data_processor.py`
run unchanged on top of cpd_b200 (SURVEY.md sections 2d, 8b, 8c).

``install()`` registers in ``sys.modules``:
  * ``spconv``, ``spconv.pytorch`` (+ ``.conv``, ``.utils``), ``spconv.utils``  -> cpd_b200.sparse / cpd_b200.voxel
  * ``cumm``, ``cumm.tensorview``                                             -> numpy pass-through
  * ``cpd.ops.iou3d_nms.iou3d_nms_cuda`` is provided by ``cpd_b200.iou3d_nms_cuda``
    (see INTEGRATION.md for the one-line import the reference tree needs).
Nothing here computes anything on the CPU: every call lands in libcpd_b200.so.
"""
import sys
import types

from .. import iou3d_nms_cuda, sparse, voxel


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def install(force=False):
    if "spconv" in sys.modules and not force and not getattr(sys.modules["spconv"], "__cpd_b200__", False):
        raise RuntimeError("a real spconv is already imported; refusing to shadow it (pass force=True)")
    conv = _module("spconv.pytorch.conv", SparseConvolution=sparse.SparseConvolution, SubMConv3d=sparse.SubMConv3d,
                   SparseConv3d=sparse.SparseConv3d, SparseInverseConv3d=sparse.SparseInverseConv3d)
    putils = _module("spconv.pytorch.utils", PointToVoxel=voxel.PointToVoxel,
                     gather_features_by_pc_voxel_id=voxel.gather_features_by_pc_voxel_id)
    pyt = _module("spconv.pytorch", SparseConvTensor=sparse.SparseConvTensor, SparseModule=sparse.SparseModule,
                  SparseSequential=sparse.SparseSequential, SubMConv3d=sparse.SubMConv3d,
                  SparseConv3d=sparse.SparseConv3d, SparseInverseConv3d=sparse.SparseInverseConv3d,
                  ToDense=sparse.ToDense, conv=conv, utils=putils, __cpd_b200__=True)
    sutils = _module("spconv.utils", Point2VoxelCPU3d=voxel.Point2VoxelCPU3d)
    root = _module("spconv", pytorch=pyt, utils=sutils, __version__="2.1.22+cpd_b200", __cpd_b200__=True)
    tv = _module("cumm.tensorview", from_numpy=voxel.tv_from_numpy, Tensor=voxel.TvTensor)
    cumm = _module("cumm", tensorview=tv, __cpd_b200__=True)
    sys.modules.update({"spconv": root, "spconv.pytorch": pyt, "spconv.pytorch.conv": conv,
                        "spconv.pytorch.utils": putils, "spconv.utils": sutils, "cumm": cumm,
                        "cumm.tensorview": tv})
    sys.modules.setdefault("iou3d_nms_cuda", iou3d_nms_cuda)
    return root
