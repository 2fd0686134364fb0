"""HeightCompression.bev_align + ALIGN methods of the mirror (cpd_b200/backbone.py) against the REFERENCE's own
height_compression.py (:5-36, :48-105, :142-166) imported unmodified through cpd_b200.compat.reference.
Pure torch on both sides, so it runs on the CPU; needs the reference checkout (skipped on the GPU box)."""
import importlib
import os

import numpy as np
import pytest
import torch

REF = os.environ.get("CPD_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "cpd")), reason="reference checkout not present")

PCR = [0.0, -40.0, -3.0, 70.4, 40.0, 1.0]
VS = [0.05, 0.05, 0.1]


@pytest.fixture(scope="module")
def ref_hc():
    from cpd_b200.compat import reference
    reference.install_reference(REF)
    hc = importlib.import_module("cpd.models.backbones_2d.map_to_bev.height_compression")
    yield hc
    reference.uninstall_reference()


class _FakeSparse:
    def __init__(self, dense5):
        self._d = dense5

    def dense(self):
        return self._d

    def dense_bev_nhwc(self):
        n, c, d, h, w = self._d.shape
        return self._d.reshape(n, c * d, h, w).permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("method", ["first", "max", "mean", "weighted_max"])
def test_bev_align_matches_reference(ref_hc, monkeypatch, method):
    from cpd_b200 import backbone
    from cpd_b200.compat.reference import EasyDict
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)         # get_pseudo_points calls .cuda()
    torch.manual_seed(3)
    stride, bs, stages = 8, 2, 3
    h, w = int(round((PCR[4] - PCR[1]) / VS[1] / stride)), int(round((PCR[3] - PCR[0]) / VS[0] / stride))
    cfg = dict(NUM_BEV_FEATURES=12, ALIGN=True, ALIGN_METHOD=method, W1=0.8, W2=0.2)
    ref = ref_hc.HeightCompression(EasyDict(cfg), 1, voxel_size=VS, point_cloud_range=PCR)
    for nhwc in (True, False):
        mine = backbone.HeightCompression(cfg, nhwc=nhwc, voxel_size=VS, point_cloud_range=PCR)
        g = np.random.default_rng(5)
        tp = np.stack([np.stack([np.array([g.uniform(-0.78, 0.78), float(g.integers(0, 2)), g.uniform(0.95, 1.05)]) for _ in range(stages)])
                       for _ in range(bs)])
        tp[0, 1, 1], tp[1, 1, 1] = 1.0, 0.0                                       # both flip states occur
        tp_t = torch.from_numpy(tp).float()
        bd_ref = dict(transform_param=tp_t, encoded_spconv_tensor_stride=stride)
        bd_mine = dict(transform_param=tp_t, encoded_spconv_tensor_stride=stride)
        for i in range(stages):
            sid = "" if i == 0 else str(i)
            dense5 = torch.randn(bs, 6, 2, h, w) * (torch.rand(bs, 1, 1, h, w) < 0.3)
            bd_ref["encoded_spconv_tensor" + sid] = _FakeSparse(dense5)
            bd_mine["encoded_spconv_tensor" + sid] = _FakeSparse(dense5)
        out_ref = ref(bd_ref)
        out_mine = mine(bd_mine)
        for i in range(stages):
            sid = "" if i == 0 else str(i)
            if i > 0 or method == "first":
                assert torch.equal(out_mine["spatial_features" + sid], out_ref["spatial_features" + sid])
        a, b = out_mine["spatial_features"], out_ref["spatial_features"]
        assert a.shape == b.shape and float((a - b).abs().max()) <= 1e-5 * max(1.0, float(b.abs().max()))
        # the alignment itself, stage by stage
        for i in range(1, stages):
            al_ref = ref.bev_align(bd_ref["spatial_features" + str(i)].clone(), tp_t, stride, i)
            al_mine = mine.bev_align(bd_mine["spatial_features" + str(i)], tp_t, stride, i)
            assert float((al_mine - al_ref).abs().max()) <= 1e-5 * max(1.0, float(al_ref.abs().max()))
            assert float(al_ref.abs().max()) > 0
