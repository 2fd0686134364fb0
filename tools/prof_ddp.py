"""Where does the step time go under DDP?  Run under torchrun (any N, also N=1): every rank runs the bench's train loop
(prefetched input stage, graphed dense stack, DDP with the bench's options); rank 0 reports wall time per step, the GPU-busy
time per step (sum of kernel durations from torch.profiler) and the host-side hot spots.  GPU-busy ~= wall: device-bound;
GPU-busy << wall: the host is pacing the step.   usage: torchrun --nproc-per-node N tools/prof_ddp.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import numpy as np
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

import bench

sys.argv = [sys.argv[0]]
a = bench.parse()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    cores = sorted(os.sched_getaffinity(0))
    per = len(cores) // world
    if per >= 2:
        os.sched_setaffinity(0, cores[local * per:(local + 1) * per])
    dist.init_process_group("nccl", device_id=dev)
net = bench.make_detector(dev)
net.capture_dense_graph(a.batch)
model = net
if world > 1:
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], broadcast_buffers=False, gradient_as_bucket_view=True, static_graph=True)
opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
pool = []
for k in range(2):
    fr = [torch.from_numpy(f).to(dev) for f in bench.make_frames(rank, a.batch * 2, a.points)[k * a.batch:(k + 1) * a.batch]]
    fr1 = [torch.from_numpy(f).to(dev) for f in bench.make_frames(rank + 500, a.batch * 2, a.points)[k * a.batch:(k + 1) * a.batch]]
    gt = torch.from_numpy(np.stack(bench.make_gt(rank, a.batch * 2)[k * a.batch:(k + 1) * a.batch])).to(dev)
    pool.append(dict(points=fr, points1=fr1, gt_boxes=gt))
pending = [None]
count = [0]


def step():
    i = count[0]
    count[0] += 1
    prep = pending[0] if pending[0] is not None else net.prepare(pool[i % 2])
    loss, tb = model(pool[i % 2], prepared=prep)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    pending[0] = net.prepare(pool[(i + 1) % 2])
    torch.nn.utils.clip_grad_norm_(net.parameters(), 10.0)
    opt.step()
    return loss


import gc
for _ in range(6):
    keep = step()
gc.collect(); gc.disable()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
N = 10
for _ in range(N):
    keep = step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / N * 1e3
NS = 3
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(NS):
        keep = step()
    torch.cuda.synchronize()
if rank == 0:
    ka = prof.key_averages()
    busy = sum(getattr(e, "self_device_time_total", 0) for e in ka) / 1e3 / NS
    cpu = sum(e.self_cpu_time_total for e in ka) / 1e3 / NS
    print(f"# N={world} ranks, host cores visible to rank 0: {len(os.sched_getaffinity(0))} (box: {os.cpu_count()})")
    print(f"wall per step (no profiler): {wall:.2f} ms;  GPU-busy per step (sum of kernel time, under the profiler): {busy:.2f} ms;  "
          f"host self time of profiled ops per step (all threads): {cpu:.2f} ms")
    rows = sorted(ka, key=lambda e: -e.self_cpu_time_total)[:14]
    for e in rows:
        print(f"   host {e.self_cpu_time_total / 1e3 / NS:8.2f} ms/step  x{e.count / NS:7.1f}  {e.key[:70]}")
    nccl = [e for e in ka if "nccl" in e.key.lower()]
    for e in nccl:
        print(f"   nccl  {getattr(e, 'self_device_time_total', 0) / 1e3 / NS:8.3f} ms/step  x{e.count / NS:5.1f}  {e.key[:80]}")
if world > 1:
    dist.destroy_process_group()
