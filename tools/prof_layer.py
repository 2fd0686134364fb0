"""Run single sparse / dense layers on REAL rulebooks (bs=4 synthetic sweeps) for ncu captures / timing.
usage: python tools/prof_layer.py [stage] [reps] [only]
   stage in {1,2,3,4}: SubM channels 16/32/64/128 on that stage's sites; stage 5: dense 256->256 3x3 @ 4x94x94, 6: 128->128 @ 4x188x188
   only: run just that variant (for `ncu -k ... -c 1`): sorted | spatial | wgrad | tma | table"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cpd_b200 import ops, voxel, sparse as sp
from cpd_b200.synth import PC_RANGE, VOXEL_SIZE, synth_scan

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
only = sys.argv[3] if len(sys.argv) > 3 else None
dev = torch.device("cuda:0")


def timeit(name, fn, byt, flops, extra=""):
    if only is not None and only != name.split()[0]:
        return
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"  {name:28s}: {ms * 1e3:8.1f} us   algorithmic {byt / ms / 1e6:7.0f} GB/s   useful {flops / ms / 1e9:6.1f} TFLOP/s {extra}")


if stage >= 5:
    from cpd_b200 import bev
    n, h, w, c = (4, 94, 94, 256) if stage == 5 else (4, 188, 188, 128)
    x = torch.randn(n * h * w, c, device=dev) * (torch.rand(n * h * w, 1, device=dev) < 0.3)
    wk = torch.randn(c, 9, c, device=dev) * 0.02
    xs = ops.split_rows(x)
    rb, _, _ = bev.pixel_tables(n, h, w, 3, 3, 1, 1, dev)
    m = n * h * w
    byt, fl = 4.0 * 2 * m * c + 4.0 * 9 * c * c, 2.0 * m * 9 * c * c
    print(f"dense {c}->{c} 3x3 @ {n}x{h}x{w}: {m} pixels")
    timeit("tma", lambda: ops.conv2d_fwd(xs, n, h, w, wk, 3, 1), byt, fl)
    timeit("table", lambda: ops.gather_gemm(x, wk, rb.nbr_fwd, x_split=xs), byt, fl)
    timeit("wgrad", lambda: ops.gather_wgrad(x, x, rb.nbr_fwd_t, tap_major=True, x_split=xs, dy_split=xs), byt, fl)
    sys.exit(0)

frames = [torch.from_numpy(synth_scan(160000, i)).to(dev) for i in range(4)]
bd = voxel.voxelize_batch(frames, PC_RANGE, VOXEL_SIZE)
c1 = bd["voxel_coords"]
key = ((c1[:, 0].long() * 41 + c1[:, 1]) * 1504 + c1[:, 2]) * 1504 + c1[:, 3]       # the backbone visits stage 1 in linear-key order
perm = torch.argsort(key)
t = sp.SparseConvTensor(bd["voxel_features"][perm].contiguous(), c1[perm].contiguous(), [41, 1504, 1504], 4)
chans = [16, 32, 64, 128]
pads = [1, 1, (0, 1, 1)]
with torch.no_grad():
    for s in range(1, stage):
        conv = sp.SparseConv3d(t.features.shape[1], chans[s], 3, stride=2, padding=pads[s - 1], bias=False, indice_key=f"d{s}").to(dev)
        t = conv(t)
c = chans[stage - 1]
m = t.indices.shape[0]
nbr = ops.subm_table(t.indices, t.spatial_shape, 4, 3, t.coord_hash())
rb = sp.Rulebook("subm", nbr, None, t.indices, t.indices, t.spatial_shape, t.spatial_shape, [3, 3, 3], [1, 1, 1], [0, 0, 0])
P = int((nbr >= 0).sum())
x, dy = torch.randn(m, c, device=dev), torch.randn(m, c, device=dev)
w = torch.randn(c, 27, c, device=dev) * 0.05
xs, dys = ops.split_rows(x), ops.split_rows(dy)
print(f"stage {stage}: M={m} C={c} P={P} ({P / m:.1f} nbrs/row), shape {t.spatial_shape}")
byt = 4.0 * (2 * m * c) + 8.0 * P + 4.0 * 27 * c * c
fl = 2.0 * P * c * c
srt = rb.sorted_table("fwd", c)
masks = rb.masks_fwd
tpb = sp.taps_per_block(c)
nkb = (27 + tpb - 1) // tpb


def active(mk):
    bits = torch.zeros_like(mk, dtype=torch.float32)
    for j in range(nkb):
        grp = 0
        for k in range(j * tpb, min(27, (j + 1) * tpb)):
            grp |= 1 << k
        bits += ((mk.long() & grp) != 0).float()
    return float(bits.mean()) / nkb


nbr_t = nbr.t().contiguous()
timeit("spatial", lambda: ops.gather_gemm(x, w, nbr, x_split=xs, tile_masks=masks), byt, fl, f"active k-blocks {active(masks):.3f}")
timeit("sorted", lambda: ops.gather_gemm(x, w, srt[0], x_split=xs, tile_masks=srt[2], out_rows=srt[1]), byt, fl,
       f"active k-blocks {active(srt[2]):.3f}")
timeit("wgrad", lambda: ops.gather_wgrad(x, dy, nbr_t, tap_major=True, x_split=xs, dy_split=dys), byt, fl)
