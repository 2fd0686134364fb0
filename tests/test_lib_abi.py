"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/cpd_b200.h
declares with the signature table the ctypes binding uses (no compute calls: no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "cpd_b200.h")).read()
    return sorted(set(re.findall(r"CPD_API[^;(]*?\b(cpd_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported_and_bound():
    from cpd_b200 import _lib, build
    build.build()
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/cpd_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert L.cpd_version() >= 100
    assert L.cpd_launch_count() == 0
    # size queries are host arithmetic and safe without a GPU
    assert L.cpd_nms_workspace_bytes(500) >= 500 * 8 * 8
    assert L.cpd_coord_hash_bytes(1000) >= 2 * 1000 * 8
    assert L.cpd_voxelize_workspace_bytes(1000, 1, 5, 1000) > 0
    # tcgen05 workspaces: weight image (+ split-row images unless the caller brings them)
    assert L.cpd_gather_gemm_workspace_bytes(1000, 1000, 32, 27, 32, 0, 0) >= 1000 * 32 * 4 + 14 * 32 * 256
    assert L.cpd_gather_gemm_workspace_bytes(1000, 1000, 32, 27, 32, 0, 1) < 1000 * 32 * 4
    assert L.cpd_gather_wgrad_workspace_bytes(1000, 900, 32, 27, 64, 0, 0) >= 1000 * 32 * 4 + 900 * 64 * 4


def test_bad_arguments_return_status_not_exit():
    from cpd_b200 import _lib
    L = _lib.lib()
    st = L.cpd_gather_gemm(None, None, 0, 4, None, 27, 4, None, None, None, 10, None, None, None, None, 0, None, None, 0, None, 0, None)
    assert st == -1 and b"null" in L.cpd_last_error_string()
    with pytest.raises(_lib.CpdError):
        _lib.check(st, "cpd_gather_gemm")


def test_ops_refuse_cpu_tensors():
    import torch
    from cpd_b200 import _lib, ops
    with pytest.raises(_lib.CpdError):
        ops.iou_bev(torch.zeros(2, 7), torch.zeros(2, 7))


def test_compat_namespace_resolves():
    from cpd_b200 import compat
    compat.install()
    import spconv.pytorch as spconv
    from spconv.pytorch.utils import PointToVoxel, gather_features_by_pc_voxel_id  # noqa: F401
    from spconv.utils import Point2VoxelCPU3d  # noqa: F401
    import cumm.tensorview as tv
    assert hasattr(spconv.conv, "SparseConvolution") and issubclass(spconv.SubMConv3d, spconv.conv.SparseConvolution)
    conv = spconv.SparseConv3d(4, 8, (3, 1, 1), stride=(2, 1, 1), padding=0, bias=False, indice_key="k")
    assert tuple(conv.weight.shape) == (8, 3, 1, 1, 4)           # spconv 2.x layout (cout, kz, ky, kx, cin)
    spconv.SparseInverseConv3d(4, 4, 3, indice_key="k", bias=False)
    import numpy as np
    assert tv.from_numpy(np.zeros((2, 5), np.float32)).numpy().shape == (2, 5)


def test_python_dispatch_predicates_mirror_the_library():
    """ops.tc_gemm_ok / tc_wgrad_ok decide when the host code builds split-row images: they must agree with the
    library's own dispatch (a workspace size > 0 means the tcgen05 kernel takes the shape)."""
    from cpd_b200 import _lib, ops
    L = _lib.lib()
    for cin in (3, 5, 8, 16, 24, 32, 40, 64, 128, 256, 512):
        for cout in (1, 3, 8, 16, 32, 48, 64, 128, 256, 512):
            for K in (1, 4, 9, 27, 32):
                lib_gemm = L.cpd_gather_gemm_workspace_bytes(1000, 900, cin, K, cout, 0, 1) > 0
                assert ops.tc_gemm_ok(cin, K, cout) == lib_gemm, (cin, K, cout)
                lib_wg = L.cpd_gather_wgrad_workspace_bytes(1000, 900, cin, K, cout, 0, 0) > 0
                assert ops.tc_wgrad_ok(cin, K, cout) == lib_wg, (cin, K, cout)
