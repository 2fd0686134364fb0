"""World-size-2 `gloo` coverage of the multi-process path (runs on CPU): frame sharding by rank, the
barrier + max-over-ranks timing reduction bench.py uses, and DDP gradient averaging over the dense
CenterHead loss pieces (pure torch => CPU-runnable; the CUDA kernels are covered by -m gpu)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cpd_b200 import bev


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    frames = bench.make_frames(rank, 1, 2000)                 # seed = 1000 * rank + i  => disjoint frames per rank
    gts = bench.make_gt(rank, 1, boxes=6)
    torch.manual_seed(0)
    head = torch.nn.Linear(8, 8)                              # stand-in parameters: the collective is what is under test
    ddp = torch.nn.parallel.DistributedDataParallel(head)
    gt = torch.from_numpy(gts[0])
    hm, rb, inds, mask = bev.assign_targets_single(gt, 3, [188, 188], 8, [-75.2, -75.2, -2, 75.2, 75.2, 4], [0.1, 0.1, 0.15])
    feat = ddp(rb[: int(mask.sum())])                         # (n_valid, 8)
    loss = (feat - rb[: int(mask.sum())]).abs().mean() * (rank + 1)
    loss.backward()
    g = head.weight.grad.clone()
    t = torch.tensor([10.0 + rank, 5.0 - rank], dtype=torch.float64)     # per-rank (value, e2e) times
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        torch.save(dict(grad=g, t=t, frame_sum=float(frames[0].sum())), out)
    else:
        torch.save(dict(grad=g, frame_sum=float(frames[0].sum())), out + ".1")
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_reduction(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    a, b = torch.load(out), torch.load(out + ".1")
    assert torch.allclose(a["grad"], b["grad"])               # DDP averaged the gradients across ranks
    assert a["t"].tolist() == [11.0, 5.0]                     # max over ranks, as bench.py reports
    assert a["frame_sum"] != b["frame_sum"]                   # ranks work on different frames (weak scaling)
