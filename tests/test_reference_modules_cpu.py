"""The reference's OWN module files, unmodified, on the cpd_b200 shim.

`cpd_b200.compat.reference.install_reference` registers the reference checkout's directories as bare packages (no
`__init__` side effects), the spconv / cumm / iou3d_nms_cuda namespaces of the shim and import stubs for what is out of
scope; then `cpd/models/backbones_3d/spconv_backbone.py`, `vfe/mean_vfe.py`, `map_to_bev/height_compression.py`,
`backbones_2d/base_bev_backbone.py`, `dense_heads/center_head.py` and `cpd/datasets/processor/data_processor.py` are
imported and RUN as they are:

  * VoxelGeneratorWrapper.generate -> spconv.utils.Point2VoxelCPU3d of the shim (cpd_voxelize_cpu, host code: what the
    reference's forked DataLoader workers need) == oracle, bit for bit;
  * VoxelResBackBone8x / SparseBasicBlock / post_act_block / MeanVFE / HeightCompression forward with the mirror's
    state_dict loaded == the mirror == the oracle pipeline.  No GPU here, so the kernels behind cpd_b200.ops are replaced
    by tests/cpu_backend.py (oracle-backed, test-only): what is exercised is the HOST logic the reference code drives --
    SparseConvTensor, indice_key rulebook caching, replace_feature, SparseSequential, dense().  The CUDA arithmetic behind
    the same ops entry points is pinned by the `-m gpu` parity tests;
  * BaseBEVBackbone / CenterHead (plain torch in the reference): the mirrors' state_dicts load into them and the
    reference's outputs equal an independent float64 evaluation.

Needs the reference checkout (/root/reference; absent on the GPU box => skipped there)."""
import importlib
import os

import numpy as np
import pytest
import torch

from cpd_b200.synth import PC_RANGE, VOXEL_SIZE, synth_scan

REF = os.environ.get("CPD_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "cpd")), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    from cpd_b200.compat import reference
    reference.install_reference(REF)
    yield reference
    reference.uninstall_reference()


def test_reference_voxel_generator_runs_unchanged(ref, oracle):
    dp = importlib.import_module("cpd.datasets.processor.data_processor")
    gen = dp.VoxelGeneratorWrapper(vsize_xyz=list(VOXEL_SIZE), coors_range_xyz=list(PC_RANGE), num_point_features=5,
                                   max_num_points_per_voxel=5, max_num_voxels=150000)
    for n, seed in ((16000, 0), (40000, 7)):
        pts = synth_scan(n, seed)
        voxels, coords, num = gen.generate(pts)                     # data_processor.py:43-59, through cumm.tensorview of the shim
        ov, oc, on = oracle.voxelize(pts, PC_RANGE, VOXEL_SIZE, 5, 150000)
        assert isinstance(voxels, np.ndarray) and np.array_equal(voxels, ov) and np.array_equal(coords, oc) and np.array_equal(num, on)


def test_reference_modules_run_unchanged(ref, oracle):
    from cpd_b200 import backbone, voxel
    from oracle import pipeline
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cpu_backend import cpu_ops
    sb = importlib.import_module("cpd.models.backbones_3d.spconv_backbone")
    hc = importlib.import_module("cpd.models.backbones_2d.map_to_bev.height_compression")
    vfe = importlib.import_module("cpd.models.backbones_3d.vfe.mean_vfe")
    import spconv.pytorch as spconv
    assert spconv.SparseConvTensor.__module__.startswith("cpd_b200")           # the reference's `spconv` IS the shim
    cfg = ref.EasyDict(NUM_FILTERS=[16, 32, 64, 128], OUT_FEATURES=128, RETURN_NUM_FEATURES_AS_DICT=True, MM=True)
    torch.manual_seed(0)
    net_ref = sb.VoxelResBackBone8x(cfg, input_channels=5, grid_size=np.array([1504, 1504, 40]), num_frames=1)
    mirror = backbone.VoxelResBackBone8x(dict(cfg), 5, [1504, 1504, 40])
    for m in mirror.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.uniform_(-0.2, 0.2); m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.7, 1.3); m.bias.data.uniform_(-0.2, 0.2)
    assert {k: tuple(v.shape) for k, v in net_ref.state_dict().items()} == {k: tuple(v.shape) for k, v in mirror.state_dict().items()}
    net_ref.load_state_dict(mirror.state_dict())
    net_ref.eval(); mirror.eval()
    assert isinstance(net_ref.conv1[0], sb.SparseBasicBlock) and isinstance(net_ref.conv2[0][0], spconv.SparseConv3d)
    # input exactly as the reference pipeline makes it: VoxelGeneratorWrapper -> collate (batch column) -> float32 -> MeanVFE
    frames = [synth_scan(9000, 21), synth_scan(7000, 22)]
    gen = voxel.Point2VoxelCPU3d(list(VOXEL_SIZE), list(PC_RANGE), 5, 5, 150000)
    vs, cs, ns = [], [], []
    for b, pts in enumerate(frames):
        v, c, n = gen.point_to_voxel(voxel.tv_from_numpy(pts))
        vs.append(v.numpy()); ns.append(n.numpy())
        cs.append(np.pad(c.numpy(), ((0, 0), (1, 0)), mode="constant", constant_values=b))            # dataset.py:262-266
    bd = dict(batch_size=2, voxels=torch.from_numpy(np.concatenate(vs)), voxel_num_points=torch.from_numpy(np.concatenate(ns)).float(),
              voxel_coords=torch.from_numpy(np.concatenate(cs)).float())                               # load_data_to_gpu casts to float
    with torch.no_grad(), cpu_ops():
        bd = vfe.MeanVFE(ref.EasyDict(), 5, 1)(bd)
        out_ref = hc.HeightCompression(ref.EasyDict(NUM_BEV_FEATURES=256), num_frames=1)(net_ref(dict(bd)))
        out_mir = backbone.HeightCompression(None, nhwc=False)(mirror(dict(bd)))
    feats, coords, shape, _ = pipeline.backbone_forward(mirror, frames, PC_RANGE, VOXEL_SIZE)
    t = out_ref["encoded_spconv_tensor"]
    assert t.spatial_shape == shape == [2, 188, 188] and np.array_equal(t.indices.numpy(), coords)
    scale = max(1.0, float(np.abs(feats).max()))
    assert np.abs(t.features.numpy() - feats).max() <= 1e-4 * scale
    assert np.abs(out_mir["encoded_spconv_tensor"].features.numpy() - feats).max() <= 1e-4 * scale
    assert tuple(out_ref["spatial_features"].shape) == (2, 256, 188, 188)
    assert np.abs(out_ref["spatial_features"].numpy() - pipeline.bev_dense(feats, coords, 2, shape)).max() <= 1e-4 * scale
    assert torch.equal(out_ref["spatial_features"], out_mir["spatial_features"].contiguous()) or \
        float((out_ref["spatial_features"] - out_mir["spatial_features"]).abs().max()) <= 1e-4 * scale
    for k in ("x_conv2", "x_conv3", "x_conv4"):                     # multi-scale outputs the RoI head consumes
        a, b = out_ref["multi_scale_3d_features"][k], out_mir["multi_scale_3d_features"][k]
        assert torch.equal(a.indices, b.indices) and float((a.features - b.features).abs().max()) <= 1e-4 * scale


def test_reference_dense_modules_take_the_mirrors_weights(ref):
    """BaseBEVBackbone / CenterHead of the reference are plain torch: the mirrors' state_dicts load into them 1:1 and the
    reference forward on those weights equals the float64 twin the GPU parity tests compare the CUDA path with."""
    from cpd_b200 import bev, detector
    from oracle import pipeline
    bb_ref_mod = importlib.import_module("cpd.models.backbones_2d.base_bev_backbone")
    torch.manual_seed(1)
    cfg = detector.MODEL_CFG["BACKBONE_2D"]
    mirror = bev.BaseBEVBackbone(cfg, 1, 256)
    net_ref = bb_ref_mod.BaseBEVBackbone(ref.EasyDict(cfg), 1, 256)
    assert {k: tuple(v.shape) for k, v in net_ref.state_dict().items()} == {k: tuple(v.shape) for k, v in mirror.state_dict().items()}
    net_ref.load_state_dict(mirror.state_dict())
    net_ref.eval()
    head = bev.CenterHead(None, 1, 512, 3, ["Vehicle", "Pedestrian", "Cyclist"], [1504, 1504, 40], list(PC_RANGE), list(VOXEL_SIZE))
    run, mods = pipeline.torch_bev_reference(mirror, head)
    mods.eval()
    x = torch.randn(1, 256, 40, 36)
    with torch.no_grad():
        y_ref = net_ref({"spatial_features": x})["st_features_2d"]
        y_twin, _ = run(x)
    assert tuple(y_ref.shape) == (1, 512, 40, 36) and float((y_ref - y_twin).abs().max()) <= 1e-5 * max(1.0, float(y_twin.abs().max()))
