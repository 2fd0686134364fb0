"""Run single sparse layers on REAL rulebooks (bs=4 synthetic sweeps) for ncu captures / timing.
usage: python tools/prof_layer.py [stage] [reps]   stage in {1,2,3,4}: channels 16/32/64/128"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cpd_b200 import ops, voxel, sparse as sp
from cpd_b200.synth import PC_RANGE, VOXEL_SIZE, synth_scan

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
frames = [torch.from_numpy(synth_scan(160000, i)).to(dev) for i in range(4)]
bd = voxel.voxelize_batch(frames, PC_RANGE, VOXEL_SIZE)
t = sp.SparseConvTensor(bd["voxel_features"], bd["voxel_coords"], [41, 1504, 1504], 4)
chans = [16, 32, 64, 128]
pads = [1, 1, (0, 1, 1)]
with torch.no_grad():
    for s in range(1, stage):
        conv = sp.SparseConv3d(t.features.shape[1], chans[s], 3, stride=2, padding=pads[s - 1], bias=False, indice_key=f"d{s}").to(dev)
        t = conv(t)
c = chans[stage - 1]
m = t.indices.shape[0]
nbr = ops.subm_table(t.indices, t.spatial_shape, 4, 3, t.coord_hash())
P = int((nbr >= 0).sum())
x, dy = torch.randn(m, c, device=dev), torch.randn(m, c, device=dev)
w = torch.randn(c, 27, c, device=dev) * 0.05
print(f"stage {stage}: M={m} C={c} P={P} ({P / m:.1f} nbrs/row), shape {t.spatial_shape}")
byt = 4.0 * (2 * m * c) + 8.0 * P + 4.0 * 27 * c * c
nbr_t = nbr.t().contiguous()
for name, fn in (("gather_gemm", lambda: ops.gather_gemm(x, w, nbr)), ("gather_wgrad", lambda: ops.gather_wgrad(x, dy, nbr)),
                 ("gather_wgrad tap-major", lambda: ops.gather_wgrad(x, dy, nbr_t, tap_major=True))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"  {name}: {ms * 1e3:8.1f} us   algorithmic {byt / ms / 1e6:7.0f} GB/s   useful {2.0 * P * c * c / ms / 1e9:6.1f} TFLOP/s   gather stream {(P * 2 * c * 4 + 8 * P) / ms / 1e6:7.0f} GB/s")
