"""Kernel-time breakdown of the detector train step with torch.profiler (diagnostics, not a bench number)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import bench

a = bench.parse()
dev = torch.device("cuda:0")
net = bench.make_detector(dev)
opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
fr = [torch.from_numpy(f).to(dev) for f in bench.make_frames(0, a.batch, a.points)]
fr1 = [torch.from_numpy(f).to(dev) for f in bench.make_frames(500, a.batch, a.points)]
gt = torch.from_numpy(np.stack(bench.make_gt(0, a.batch))).to(dev)

def step():
    loss, tb = net(dict(points=fr, points1=fr1, gt_boxes=gt))
    opt.zero_grad(set_to_none=True)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(net.parameters(), 10.0)
    opt.step()
    return loss

for _ in range(3):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print(f"wall per step (no profiler): {(time.perf_counter() - t0) / 5 * 1e3:.1f} ms")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70)
print(tab)
