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


# ---------------------------------------------------------------------------------------------------------------------
# The sparse modules themselves under DistributedDataParallel (gloo, world size 2).  No GPU here, so the kernels behind
# cpd_b200.ops are replaced by tests/cpu_backend.py (oracle-backed, test-only); what is under test is the product's HOST
# path: the custom autograd Functions of cpd_b200.sparse, per-rank rulebooks of different sizes, and DDP's gradient hooks
# averaging the parameter gradients they return.
class _TinyTower(torch.nn.Module):
    def __init__(self):
        super().__init__()
        from cpd_b200 import sparse as sp
        self.net = sp.SparseSequential(sp.SubMConv3d(4, 8, 3, padding=1, bias=True, indice_key="s1"),
                                       sp.SparseConv3d(8, 8, 3, stride=2, padding=1, bias=False, indice_key="d1"),
                                       sp.SubMConv3d(8, 8, 3, padding=1, bias=True, indice_key="s2"))

    def forward(self, feats, coords):
        from cpd_b200 import sparse as sp
        return self.net(sp.SparseConvTensor(feats, coords, [6, 12, 12], 1)).features


def _tower_shard(rank):
    g = np.random.default_rng(100 + rank)
    n = 150 + 60 * rank                                            # ranks hold clouds of different sizes
    cells = np.sort(g.choice(6 * 12 * 12, size=n, replace=False))
    z, r = np.divmod(cells, 144)
    y, x = np.divmod(r, 12)
    coords = torch.from_numpy(np.stack([np.zeros_like(z), z, y, x], 1).astype(np.int32))
    feats = torch.from_numpy(g.normal(0, 1, (n, 4)).astype(np.float32))
    return feats, coords


def _tower_grads(model, rank):
    feats, coords = _tower_shard(rank)
    model.zero_grad(set_to_none=True)
    out = model(feats, coords)
    (out.square().mean() + out.sum() * 0.01).backward()


def _tower_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cpu_backend import cpu_ops
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(7)
    tower = _TinyTower()
    with cpu_ops():
        ddp = torch.nn.parallel.DistributedDataParallel(tower)
        _tower_grads(ddp, rank)
    torch.save({n: p.grad.clone() for n, p in tower.named_parameters()}, f"{out}.{rank}")
    dist.destroy_process_group()


def test_sparse_tower_under_ddp_averages_gradients(tmp_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cpu_backend import cpu_ops
    out = str(tmp_path / "tower")
    mp.spawn(_tower_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    g0, g1 = torch.load(out + ".0"), torch.load(out + ".1")
    torch.manual_seed(7)
    tower = _TinyTower()
    local = []
    with cpu_ops():
        for rank in range(2):
            _tower_grads(tower, rank)
            local.append({n: p.grad.clone() for n, p in tower.named_parameters()})
    for n in g0:
        assert torch.equal(g0[n], g1[n]), n                                   # every rank ends with the same gradient ...
        want = (local[0][n] + local[1][n]) / 2
        assert float(want.abs().max()) > 0
        assert torch.allclose(g0[n], want, rtol=1e-5, atol=1e-6), n           # ... the mean of the per-shard gradients
