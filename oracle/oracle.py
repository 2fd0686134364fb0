"""numpy/ctypes front-end of the CPU oracle (oracle/cpd_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / `--impl reference` legs of bench.py.  Never imported by cpd_b200/.
Pinning status: rotated IoU/NMS pinned against the reference's own iou3d_cpu.cpp
(oracle/_ref); voxelizer + sparse conv "parity unpinned" (spconv 2.1.22 is not in
/root/reference nor in this image) -- see the header of cpd_oracle.c and DESIGN.md.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(ref=False):
    """(Re)build liboracle.so, and oracle/_ref when the reference tree is present."""
    subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    if ref and os.path.isdir("/root/reference/cpd/ops/iou3d_nms/src"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.cpd_oracle_voxelize.restype = C.c_int64
        L.cpd_oracle_rulebook_strided.restype = C.c_int64
        L.cpd_oracle_nms.restype = C.c_int64
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def set_threads(n):
    lib().cpd_oracle_set_threads(int(n))


def num_threads():
    return int(lib().cpd_oracle_num_threads())


def voxelize(points, pc_range, voxel_size, max_pts=5, max_voxels=1000000):
    """Point2VoxelCPU3d.point_to_voxel: -> voxels (M,max_pts,C), coords (M,3) zyx, num (M,)."""
    pts = _f32(points)
    n, c = pts.shape
    cap = int(min(max_voxels, max(n, 1)))
    voxels = np.empty((cap, max_pts, c), np.float32)
    coords = np.empty((cap, 3), np.int32)
    num = np.empty((cap,), np.int32)
    m = lib().cpd_oracle_voxelize(_p(pts), C.c_int64(n), c, _p(_f32(pc_range)), _p(_f32(voxel_size)),
                                  int(max_pts), C.c_int64(max_voxels), _p(voxels), _p(coords), _p(num))
    assert m >= 0
    return voxels[:m].copy(), coords[:m].copy(), num[:m].copy()


def mean_vfe(voxels, num):
    v = _f32(voxels)
    m, k, c = v.shape
    out = np.empty((m, c), np.float32)
    lib().cpd_oracle_mean_vfe(_p(v), _p(_i32(num)), C.c_int64(m), k, c, _p(out))
    return out


class Rulebook:
    """spconv-style indice pairs: pair_in/pair_out (K, stride) int32, pair_cnt (K,)."""

    def __init__(self, pair_in, pair_out, pair_cnt, m_in, m_out, out_coords, out_shape, ksize):
        self.pair_in, self.pair_out, self.pair_cnt = pair_in, pair_out, pair_cnt
        self.m_in, self.m_out = m_in, m_out
        self.out_coords, self.out_shape, self.ksize = out_coords, out_shape, ksize

    @property
    def K(self):
        return int(np.prod(self.ksize))

    @property
    def n_pairs(self):
        return int(self.pair_cnt.sum())


def _tri(v):
    return [int(v)] * 3 if np.isscalar(v) else [int(t) for t in v]


def rulebook_subm(coords, shape, ksize=3):
    co = _i32(coords)
    m = co.shape[0]
    ks = _tri(ksize)
    K = ks[0] * ks[1] * ks[2]
    pin = np.empty((K, max(m, 1)), np.int32)
    pout = np.empty((K, max(m, 1)), np.int32)
    cnt = np.zeros((K,), np.int32)
    r = lib().cpd_oracle_rulebook_subm(_p(co), C.c_int64(m), _p(_i32(_tri(shape))), _p(_i32(ks)), _p(pin), _p(pout), _p(cnt))
    assert r == 0
    return Rulebook(pin, pout, cnt, m, m, co, _tri(shape), ks)


def rulebook_strided(coords, shape, ksize, stride, padding):
    co = _i32(coords)
    m = co.shape[0]
    ks, st, pd = _tri(ksize), _tri(stride), _tri(padding)
    K = ks[0] * ks[1] * ks[2]
    pin = np.empty((K, max(m, 1)), np.int32)
    pout = np.empty((K, max(m, 1)), np.int32)
    cnt = np.zeros((K,), np.int32)
    oshape = np.zeros((3,), np.int32)
    ocoords = np.empty((max(m * K, 1), 4), np.int32)
    mo = lib().cpd_oracle_rulebook_strided(_p(co), C.c_int64(m), _p(_i32(_tri(shape))), _p(_i32(ks)), _p(_i32(st)),
                                           _p(_i32(pd)), _p(oshape), _p(ocoords), _p(pin), _p(pout), _p(cnt))
    assert mo >= 0
    return Rulebook(pin, pout, cnt, m, int(mo), ocoords[:mo].copy(), [int(t) for t in oshape], ks)


def spconv_fwd(x, w, bias, rb):
    """w: (Cout, kz, ky, kx, Cin) spconv-2.x layout."""
    x = _f32(x)
    w = _f32(w)
    cout, cin = w.shape[0], w.shape[-1]
    y = np.empty((rb.m_out, cout), np.float32)
    b = _f32(bias) if bias is not None else None
    lib().cpd_oracle_spconv_fwd(_p(x), C.c_int64(rb.m_in), cin, _p(w), _p(b), cout, rb.K, _p(rb.pair_in), _p(rb.pair_out),
                                _p(rb.pair_cnt), C.c_int64(rb.pair_in.shape[1]), _p(y), C.c_int64(rb.m_out))
    return y


def spconv_bwd(x, w, dy, rb, need_bias=True):
    x, w, dy = _f32(x), _f32(w), _f32(dy)
    cout, cin = w.shape[0], w.shape[-1]
    dx = np.empty((rb.m_in, cin), np.float32)
    dw = np.empty(w.shape, np.float32)
    db = np.empty((cout,), np.float32) if need_bias else None
    lib().cpd_oracle_spconv_bwd(_p(x), C.c_int64(rb.m_in), cin, _p(w), cout, rb.K, _p(rb.pair_in), _p(rb.pair_out),
                                _p(rb.pair_cnt), C.c_int64(rb.pair_in.shape[1]), _p(dy), C.c_int64(rb.m_out),
                                _p(dx), _p(dw), _p(db))
    return dx, dw, db


def dense(feat, coords, batch, shape):
    f, co = _f32(feat), _i32(coords)
    m, c = f.shape
    sh = _tri(shape)
    out = np.empty((batch, c, sh[0], sh[1], sh[2]), np.float32)
    lib().cpd_oracle_dense(_p(f), _p(co), C.c_int64(m), c, int(batch), _p(_i32(sh)), _p(out))
    return out


def iou_bev(a, b):
    a, b = _f32(a), _f32(b)
    out = np.empty((a.shape[0], b.shape[0]), np.float32)
    lib().cpd_oracle_iou_bev(_p(a), a.shape[0], _p(b), b.shape[0], _p(out))
    return out


def overlap_bev(a, b):
    a, b = _f32(a), _f32(b)
    out = np.empty((a.shape[0], b.shape[0]), np.float32)
    lib().cpd_oracle_overlap_bev(_p(a), a.shape[0], _p(b), b.shape[0], _p(out))
    return out


def nms(boxes_sorted, thresh, rotated=True, return_mask=False):
    """Greedy NMS over boxes already sorted by descending score -> kept row ids (int64)."""
    b = _f32(boxes_sorted)
    n = b.shape[0]
    keep = np.empty((max(n, 1),), np.int64)
    mask = np.zeros((max(n, 1), max((n + 63) // 64, 1)), np.uint64) if return_mask else None
    k = lib().cpd_oracle_nms(_p(b), n, C.c_float(thresh), int(bool(rotated)), _p(keep), _p(mask))
    return (keep[:k].copy(), mask[:n, :(n + 63) // 64]) if return_mask else keep[:k].copy()


# ---- the reference's own binaries (oracle/_ref), when built ----------------------
def ref_cpu_module():
    """The reference's iou3d_cpu.cpp compiled alone (torch extension) or None."""
    path = os.path.join(_HERE, "_ref", "iou3d_ref_cpu.so")
    if not os.path.exists(path):
        return None
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location("iou3d_ref_cpu", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_gpu_lib():
    """The reference's iou3d_nms_kernel.cu compiled unmodified for sm_100a, or None."""
    path = os.path.join(_HERE, "_ref", "libiou3d_ref_gpu.so")
    return C.CDLL(path) if os.path.exists(path) else None


def ref_pointnet2_lib():
    """The reference's voxel_query_gpu.cu + group_points_gpu.cu compiled unmodified for sm_100a, or None."""
    path = os.path.join(_HERE, "_ref", "libpointnet2_ref_gpu.so")
    return C.CDLL(path) if os.path.exists(path) else None
