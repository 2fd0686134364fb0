"""ctypes binding of libcpd_b200.so (the C ABI in include/cpd_b200.h).

There is no CPU fallback: if the library is missing or an entry fails, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcpd_b200.so")

_vp, _i32, _i64, _sz, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t, C.c_float

# name -> (restype, argtypes): mirrors include/cpd_b200.h one to one
SIGNATURES = {
    "cpd_version": (_i32, []),
    "cpd_last_error_string": (C.c_char_p, []),
    "cpd_launch_count": (_i64, []),
    "cpd_voxelize_workspace_bytes": (_sz, [_i64, _i32, _i32, _i64]),
    "cpd_voxelize": (_i32, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cpd_voxelize_cpu": (_i64, [_vp, _i64, _i32, _vp, _vp, _i32, _i64, _vp, _vp, _vp]),
    "cpd_coord_hash_bytes": (_sz, [_i64]),
    "cpd_coord_hash_build": (_i32, [_vp, _i64, _vp, _i32, _vp, _sz, _vp]),
    "cpd_rulebook_subm": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp, _sz, _vp, _vp]),
    "cpd_rulebook_strided_workspace_bytes": (_sz, [_vp, _i32, _vp, _vp, _vp]),
    "cpd_rulebook_strided_outputs": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "cpd_rulebook_strided_tables": (_i32, [_vp, _i64, _vp, _vp, _sz, _vp, _i64, _vp, _vp, _sz, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cpd_split_rows": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp]),
    "cpd_tile_tap_masks": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "cpd_tap_block_keys": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp]),
    "cpd_table_permute": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "cpd_table_transpose": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "cpd_gather_gemm": (_i32, [_vp, _vp, _i64, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _sz, _vp]),
    "cpd_gather_gemm_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32, _i32, _i32, _i32]),
    "cpd_gather_wgrad": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _i64, _i32, _vp, _i32, _i32, _vp, _vp, _i32, _vp, _sz, _vp]),
    "cpd_gather_wgrad_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32, _i32, _i32, _i32]),
    "cpd_bn_train_fwd": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _i32, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cpd_bn_train_bwd": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "cpd_weight_transpose": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "cpd_conv2d_table": (_i32, [_i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "cpd_conv2d_supported": (_i32, [_i32, _i32, _i32, _i32]),
    "cpd_conv2d_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "cpd_conv2d_fwd": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _sz, _vp]),
    "cpd_conv2d_dgrad_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "cpd_conv2d_dgrad": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "cpd_conv2d_wgrad_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "cpd_conv2d_wgrad": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "cpd_convt2d_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "cpd_convt2d_fwd": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _sz, _vp]),
    "cpd_sparse_to_dense": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _i32, _vp, _vp]),
    "cpd_sparse_to_dense_bwd": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _i32, _vp, _vp]),
    "cpd_voxel_query": (_i32, [_vp, _vp, _i64, _vp, _vp, _sz, _vp, _vp, _i32, _vp, _f, _i32, _vp, _vp, _vp]),
    "cpd_group_points": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "cpd_group_points_bwd": (_i32, [_vp, _vp, _i64, _i32, _i32, _i64, _vp, _vp]),
    "cpd_overlap_bev": (_i32, [_vp, _i32, _vp, _i32, _vp, _vp]),
    "cpd_iou_bev": (_i32, [_vp, _i32, _vp, _i32, _vp, _vp]),
    "cpd_nms_workspace_bytes": (_sz, [_i32]),
    "cpd_nms_rotated": (_i32, [_vp, _i32, _f, _vp, _vp, _vp, _sz, _vp]),
    "cpd_nms_normal": (_i32, [_vp, _i32, _f, _vp, _vp, _vp, _sz, _vp]),
    "cpd_nms_mask": (_i32, [_vp, _i32, _f, _i32, _vp, _vp]),
}

_LIB = None


class CpdError(RuntimeError):
    pass


def lib():
    """Load libcpd_b200.so; raises if it has not been built (no fallback exists)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise CpdError(f"{LIB_PATH} is missing: build it with `python -m cpd_b200.build` "
                           "(or __graft_entry__.build()); cpd_b200 has no CPU or PyTorch fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _LIB = L
    return _LIB


def check(status, what=""):
    if status != 0:
        msg = lib().cpd_last_error_string().decode("utf-8", "replace")
        raise CpdError(f"{what or 'libcpd_b200'} failed with status {status}: {msg}")


def launch_count():
    return int(lib().cpd_launch_count())
