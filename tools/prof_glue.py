"""Where do the torch (non-libcpd) kernel launches of a train step come from?  torch.profiler with Python stacks: every
aten op that launched device work is attributed to the innermost frame inside this repository (diagnostics only)."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

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
NSTEP = 2
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    for _ in range(NSTEP):
        step()
    torch.cuda.synchronize()

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for e in prof.key_averages(group_by_stack_n=30):
    dt = getattr(e, "self_device_time_total", 0)
    if dt <= 0:
        continue
    frame = "?"
    for f in e.stack:
        if root in f and "site-packages" not in f:
            frame = f.replace(root + "/", "").split(" ")[0] if " " in f else f.replace(root + "/", "")
            frame = f.replace(root + "/", "")
            break
    k = (e.key, frame)
    agg[k][0] += e.count
    agg[k][1] += dt
    agg[k][2] += e.self_cpu_time_total
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
tot = sum(v[1] for v in agg.values())
print(f"device time attributed to aten ops with device work: {tot / 1e3 / NSTEP:.2f} ms/step")
print(f"{'calls/step':>10} {'dev us/step':>12} {'cpu us/step':>12}  op @ frame")
for (op, frame), (n, dt, ct) in rows[:80]:
    print(f"{n / NSTEP:10.1f} {dt / NSTEP:12.1f} {ct / NSTEP:12.1f}  {op[:40]} @ {frame[:110]}")
