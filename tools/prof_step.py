"""Kernel-time / host-time breakdown of the detector train step with torch.profiler (diagnostics, not a bench number)."""
import os, sys, time
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
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print(f"wall per step (no profiler): {(time.perf_counter() - t0) / 5 * 1e3:.1f} ms")
# host-only pacing: how long the Python side takes to enqueue one step's forward / backward
torch.cuda.synchronize()
t0 = time.perf_counter()
loss, tb = net(dict(points=fr, points1=fr1, gt_boxes=gt))
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
opt.zero_grad(set_to_none=True)
loss.backward()
t3 = time.perf_counter()
torch.cuda.synchronize()
t4 = time.perf_counter()
print(f"forward: host enqueue {1e3 * (t1 - t0):.1f} ms, +drain {1e3 * (t2 - t1):.1f} ms; backward: host enqueue {1e3 * (t3 - t2):.1f} ms, "
      f"+drain {1e3 * (t4 - t3):.1f} ms")
NSTEP = 2
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(NSTEP):
        step()
    torch.cuda.synchronize()
ka = prof.key_averages()
kern = [e for e in ka if e.device_type == torch.autograd.DeviceType.CUDA or getattr(e, "self_device_time_total", 0) > 0]
tot = sum(getattr(e, "self_device_time_total", 0) for e in ka)
print(f"sum of device time over {NSTEP} steps: {tot / 1e3:.1f} ms = {tot / 1e3 / NSTEP:.1f} ms/step")
print(ka.table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70))
print(ka.table(sort_by="self_cpu_time_total", row_limit=40, max_name_column_width=70))
