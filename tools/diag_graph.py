"""Diagnostics: how far apart are two EAGER runs of the same train step (split-K / statistics atomics reorder), and how far
is the CUDA-graph replay of the dense stack from them?  Prints per-parameter relative gradient differences."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cpd_b200 import detector
from cpd_b200.synth import synth_gt_boxes, synth_scan

dev = torch.device("cuda:0")
torch.manual_seed(0)
bs, npts = 2, 20000
batch = dict(points=[torch.from_numpy(synth_scan(npts, 70 + i)).to(dev) for i in range(bs)],
             points1=[torch.from_numpy(synth_scan(npts, 570 + i)).to(dev) for i in range(bs)],
             gt_boxes=torch.from_numpy(np.stack([synth_gt_boxes(30, 70 + i) for i in range(bs)])).to(dev))
det = detector.CPDHotPathDetector().to(dev).train()
state = {k: v.clone() for k, v in det.state_dict().items()}


def run():
    det.load_state_dict(state)
    det.zero_grad(set_to_none=True)
    loss, _ = det(batch)
    loss.backward()
    g = {n: p.grad.clone() for n, p in det.named_parameters()}
    l = float(loss.detach())
    del loss
    return l, g


def cmp(tag, a, b):
    worst = sorted(((float((a[n] - b[n]).abs().max()) / max(1e-12, float(b[n].abs().max())), n) for n in a), reverse=True)
    print(f"{tag}: worst relative gradient differences (max|d| / max|ref|)")
    for r, n in worst[:8]:
        print(f"   {r:9.2e}  {n}")
    print(f"   median {np.median([w[0] for w in worst]):.2e}")


l0, g0 = run()
l1, g1 = run()
print("eager losses", l0, l1)
cmp("eager vs eager", g1, g0)
det.last_batch_dict = None
det.capture_dense_graph(bs)
l2, g2 = run()
l3, g3 = run()
print("graph losses", l2, l3)
cmp("graph vs eager", g2, g0)
cmp("graph vs graph", g3, g2)
