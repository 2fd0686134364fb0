"""Diagnostics: per-step times of the train step with the input stage prefetched, host vs resident inputs, with / without a
host sync per step (what the e2e leg of bench.py does)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if "--expandable" in sys.argv:
    os.environ["PYTORCH_CUDA_ALLOC_CONF"] = "expandable_segments:True"
import numpy as np, torch
import bench


dev = torch.device("cuda:0")
B, P = 4, 160000
net = bench.make_detector(dev)
if "--graph" in sys.argv:
    net.capture_dense_graph(B)
opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
pool = 2
host = [torch.from_numpy(f).pin_memory() for f in bench.make_frames(0, B * pool, P)]
host1 = [torch.from_numpy(f).pin_memory() for f in bench.make_frames(500, B * pool, P)]
gts = bench.make_gt(0, B * pool)
host_gt = [torch.from_numpy(np.stack(gts[k:k + B])).pin_memory() for k in range(0, B * pool, B)]
res, res1, res_gt = [h.to(dev) for h in host], [h.to(dev) for h in host1], [g.to(dev) for g in host_gt]


def mk(j, src, src1, sgt):
    k = (j % pool) * B
    return dict(points=src[k:k + B], points1=src1[k:k + B], gt_boxes=sgt[j % pool])


def run(name, src, src1, sgt, prefetch, sync, n=8):
    pend = None
    times = []
    torch.cuda.synchronize()
    for i in range(n):
        t0 = time.perf_counter()
        prep = pend if prefetch else None
        if prefetch and prep is None:
            prep = net.prepare(mk(i, src, src1, sgt))
        t1 = time.perf_counter()
        loss, _ = net(mk(i, src, src1, sgt), prepared=prep)
        t2 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        t3 = time.perf_counter()
        if prefetch:
            pend = net.prepare(mk(i + 1, src, src1, sgt))
        t4 = time.perf_counter()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 10.0)
        opt.step()
        if sync:
            float(loss)
        t5 = time.perf_counter()
        times.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4))
    torch.cuda.synchronize()
    tt = np.array(times[2:]) * 1e3
    print(f"{name:44s} step {tt.sum(1).mean():6.1f} ms = prep-miss {tt[:,0].mean():5.1f} + fwd {tt[:,1].mean():5.1f} + bwd {tt[:,2].mean():5.1f} + prepare {tt[:,3].mean():5.1f} + opt/sync {tt[:,4].mean():5.1f}", flush=True)


for _ in range(2):
    run("warm resident no-prefetch", res, res1, res_gt, False, True, n=4)
run("resident, prefetch, sync each step", res, res1, res_gt, True, True)
run("host,     prefetch, sync each step", host, host1, host_gt, True, True)
run("resident, prefetch, no sync", res, res1, res_gt, True, False)
run("host,     prefetch, no sync", host, host1, host_gt, True, False)
run("host,     inline,   sync each step", host, host1, host_gt, False, True)
run("resident, inline,   sync each step", res, res1, res_gt, False, True)
