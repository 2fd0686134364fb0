#!/usr/bin/env python
"""bench.py -- frames/s of the CPD detection hot path on B200 (contract: see task brief).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, libcpd_b200.so)
  python bench.py --impl reference ...                     # CPU arm: the oracle port of the
                                                           # reference path on the host cores
One "step" = one pass of the hot path over one batch of synthetic Waymo-shaped sweeps.
`value` has the point clouds already resident in HBM; `e2e` goes through the public API with
HOST (pinned) buffers, H2D and D2H inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames_per_sec"
L2_FLUSH_BYTES = 256 << 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="detector_train", choices=["detector_train", "backbone_fwd", "stress"])
    ap.add_argument("--batch", type=int, default=4, help="frames per step per GPU (CPD trains with 4)")
    ap.add_argument("--points", type=int, default=160000, help="points per frame")
    ap.add_argument("--pool", type=int, default=2, help="distinct batches cycled through")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the dense stack (BEV backbone + CenterHead convs) into CUDA graphs")
    ap.add_argument("--prefetch", action="store_true", help="(default) run the input stage (voxelize + rulebooks) one step ahead on a side "
                                                            "stream (detector.prepare), like a prefetching DataLoader")
    ap.add_argument("--no-prefetch", action="store_true", help="run the input stage inline at the start of every forward")
    ap.add_argument("--prefetch-thread", action="store_true", help="run the prefetched input stage on a worker thread (detector.prepare_async), "
                                                                   "submitted at the START of the step before")
    a = ap.parse_args()
    if a.workload == "stress" and a.points == 160000:
        a.points = 300000                      # BASELINE configs[4]: 300k points / 200k active voxels per frame
    a.prefetch = not a.no_prefetch             # measured on one B200: 47.1 ms/step prefetched vs 52.0 inline (profiles/r2_*)
    if a.prefetch:
        # the side stream has its own caching-allocator pool; with the default allocator its growth (cudaMalloc per new
        # segment size) made the first ~10 steps 1.5-2.5x slower -- expandable segments removed that (measured)
        os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
    return a


def workload_name(a):
    if a.workload == "stress":
        return (f"dense-scene stress: {a.points // 1000}k pts/frame at ~200k active voxels/frame, bs={a.batch}/GPU: voxelize + the whole "
                f"rulebook chain of VoxelResBackBone8x + SubM gather-GEMM C->C sweep (C = 16,32,64,128 on the 200k-voxel level and on the "
                f"level each width runs at) + dense() [BASELINE configs[4]]")
    if a.workload == "backbone_fwd":
        return (f"CPD VoxelBackBone8x fwd (voxelize+MeanVFE+12 sparse convs+BEV dense), {a.points // 1000}k pts/frame, "
                f"bs={a.batch}/GPU [BASELINE configs[1]]")
    return (f"CPD full hot-path detector train step fwd+bwd (voxelize x2 + VoxelResBackBone8x with MM tower + BEV backbone + "
            f"CenterHead loss/decode + iou3d_nms + clip + Adam), {a.points // 1000}k pts/frame, bs={a.batch}/GPU "
            f"[BASELINE configs[2]; configs[3] under torchrun]")


def make_frames(rank, count, points, dense=False):
    from cpd_b200.synth import synth_dense_scan, synth_scan
    if dense:
        return [synth_dense_scan(points, 1000 * rank + i)[0] for i in range(count)]
    return [synth_scan(points, 1000 * rank + i) for i in range(count)]


class StressSweep:
    """BASELINE configs[4]: the front end and the gather/scatter stage alone, on dense scenes.  One step = voxelize the batch,
    build every rulebook of the VoxelResBackBone8x chain (SubM tables, strided output sets + tables, tap-pattern order,
    tile masks), run SubM C->C gather-GEMMs for C in {16, 32, 64, 128} on the stage-1 (200 k voxels / frame) table and on
    the level each width runs at in the model, and scatter the last level to the dense BEV map."""
    CH = (16, 32, 64, 128)
    PADS = (1, 1, (0, 1, 1))

    def __init__(self, dev):
        import torch
        self.dev = dev
        g = torch.Generator(device="cpu").manual_seed(7)
        self.w = {c: (torch.randn(c, 27, c, generator=g) * (27 * c) ** -0.5).to(dev) for c in self.CH}
        self.feat = {}
        self.info = {}
        from cpd_b200 import sparse as sp
        self.subm = [sp.SubMConv3d(c, c, 3, bias=False, indice_key=f"subm{i}") for i, c in enumerate(self.CH)]      # geometry only
        self.down = [sp.SparseConv3d(self.CH[i], self.CH[i + 1], 3, stride=2, padding=self.PADS[i], bias=False, indice_key=f"down{i}")
                     for i in range(3)]
        self.out = sp.SparseConv3d(128, 128, (3, 1, 1), stride=(2, 1, 1), padding=0, bias=False, indice_key="out")

    def _x(self, key, m, c):
        import torch
        k = (key, m, c)
        if k not in self.feat:                     # synthetic features of the level's size, made once per distinct batch
            g = torch.Generator(device="cpu").manual_seed(m % 65521 + c)
            x = torch.randn(m, c, generator=g).to(self.dev)
            self.feat[k] = (x, None)
        x, xs = self.feat[k]
        return x, xs

    def step(self, points):
        import torch
        from cpd_b200 import ops, sparse as sp, voxel
        from cpd_b200.synth import PC_RANGE, VOXEL_SIZE
        bd = voxel.voxelize_batch([p if p.is_cuda else p.to(self.dev, non_blocking=True) for p in points], PC_RANGE, VOXEL_SIZE)
        bs = len(points)
        c1 = bd["voxel_coords"]
        key = ((c1[:, 0].long() * 41 + c1[:, 1]) * 1504 + c1[:, 2]) * 1504 + c1[:, 3]
        coords = c1.index_select(0, torch.argsort(key))
        t = sp.SparseConvTensor(None, coords, [41, 1504, 1504], bs)
        acc = bd["voxel_features"].sum()
        levels = []
        for s in range(4):
            rb, _ = self.subm[s].get_rulebook(t)
            m = t.indices.shape[0]
            levels.append((m, rb))
            widths = self.CH if s == 0 else (self.CH[s],)
            for c in widths:
                x, _ = self._x(s, m, c)
                xs = ops.split_rows(x)
                srt = rb.sorted_table("fwd", c)
                if srt is not None:
                    y = ops.gather_gemm(x, self.w[c], srt[0], x_split=xs, tile_masks=srt[2], out_rows=srt[1])
                else:
                    y = ops.gather_gemm(x, self.w[c], rb.nbr_fwd, x_split=xs, tile_masks=rb.masks_fwd)
                acc = acc + y[0, 0]
            if s < 3:
                rbd, oh = self.down[s].get_rulebook(t)
                rbd.sorted_table("fwd", self.CH[s]); rbd.bwd_sorted(self.CH[s + 1])
                t = self.down[s]._wrap_output(t, rbd, oh, None)
        rbo, oh = self.out.get_rulebook(t)
        to = self.out._wrap_output(t, rbo, oh, None)
        x, _ = self._x("out", to.indices.shape[0], 128)
        dense = ops.sparse_to_dense(x, to.indices, bs, to.spatial_shape, channels_last=True)
        self.info = {"voxels_per_level": [m for m, _ in levels] + [int(to.indices.shape[0])]}
        return acc + dense[0, 0, 0, 0], to.replace_feature(x)


def make_net(device):
    import torch
    from cpd_b200 import backbone
    torch.manual_seed(1234)
    net = backbone.VoxelBackBone8x(dict(NUM_FILTERS=[16, 32, 64, 128], OUT_FEATURES=128), 5, [1504, 1504, 40])
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.uniform_(-0.1, 0.1)
            m.running_var.uniform_(0.8, 1.2)
    return net.to(device).eval()


def make_detector(device):
    import torch
    from cpd_b200 import detector
    torch.manual_seed(1234)
    return detector.CPDHotPathDetector().to(device).train()


def make_gt(rank, count, boxes=30):
    from cpd_b200.synth import synth_gt_boxes
    return [synth_gt_boxes(boxes, 1000 * rank + i) for i in range(count)]


# ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.rows, self.gpu, self.proc, self.first = [], gpu_index, None, 0

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([t.strip() for t in line.split(",")])
        except Exception:
            pass

    def mark(self):
        """Statistics are taken over the samples from here on (the timed phases), not over the idle set-up."""
        self.first = len(self.rows)

    def wait_first(self, timeout):
        t0 = time.perf_counter()
        while not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.05)

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows[self.first:]:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


def run_ours(a):
    import torch
    import torch.distributed as dist
    from cpd_b200 import _lib, backbone, ops, voxel
    from cpd_b200.synth import PC_RANGE, VOXEL_SIZE

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # each rank runs a launching thread, an autograd thread and NCCL's proxy: give every rank its own slice of the host's
        # cores instead of letting 8 x 3 busy threads migrate over all of them (the step is host-paced when they collide)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = len(cores) // world
            if per >= 2:
                os.sched_setaffinity(0, cores[local * per:(local + 1) * per])
        except (AttributeError, OSError):
            pass
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    sampler = ClockSampler(local)
    if rank == 0:                                  # one nvidia-smi poller per job (rank 0's GPU), at the recipe's 200 ms period: every
        sampler.start()                            # query takes driver locks that the launching threads of ALL ranks contend for
    train = a.workload == "detector_train"
    if a.workload == "stress":
        a.pool = 1                                 # generating a 200k-voxel scene takes seconds: one batch, L2 flushed between steps
    nfr = a.batch * a.pool
    frames = make_frames(rank, nfr, a.points, dense=a.workload == "stress")
    host = [torch.from_numpy(f).pin_memory() for f in frames]
    resident = [h.to(dev) for h in host]
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def batch_of(i, src):
        k = (i % a.pool) * a.batch
        return src[k:k + a.batch]

    if train:
        frames1 = make_frames(rank + 500, nfr, a.points)          # second cloud per frame (STAGES=2: `points1`, MM tower)
        host1 = [torch.from_numpy(f).pin_memory() for f in frames1]
        resident1 = [h.to(dev) for h in host1]
        gts = make_gt(rank, nfr)
        host_gt = [torch.from_numpy(np.stack(gts[k:k + a.batch])).pin_memory() for k in range(0, nfr, a.batch)]
        resident_gt = [g.to(dev) for g in host_gt]
        net = make_detector(dev)
        if not a.no_graph:
            net.capture_dense_graph(a.batch)                       # static-shape part of the step: forward + backward as CUDA graphs
        model = net
        if world > 1:
            # DDP as tools/train.py:143 builds it, plus three options that change no result (measured at N=2 on B200s: 46.3 ->
            # 44.6 ms/step): gradients live in the all-reduce buckets (no copy in / out), the autograd graph is declared static,
            # and the per-forward broadcast of rank 0's BatchNorm running statistics is dropped -- training-mode BatchNorm never
            # reads them and the checkpoint rank 0 writes holds its own statistics either way.
            ddp_kw = dict(broadcast_buffers=False, gradient_as_bucket_view=True, static_graph=True)
            ddp_kw.update(json.loads(os.environ.get("CPD_DDP_KWARGS", "{}")))
            model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], **ddp_kw)
        opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)

        pending = {}                                               # input stage of the NEXT step, already enqueued on a side stream
        counter = [0]

        def step(_, src, src1, src_gt):
            i = counter[0]
            counter[0] += 1
            mk = lambda j: dict(points=batch_of(j, src), points1=batch_of(j, src1), gt_boxes=src_gt[j % a.pool])
            prep = pending.pop((id(src), i), None)
            if prep is None or a.no_prefetch:
                prep = None if a.no_prefetch else net.prepare(mk(i))
            if a.prefetch_thread and not a.no_prefetch:            # step i+1's input stage: a worker thread, the whole step to finish
                pending.clear()
                pending[(id(src), i + 1)] = net.prepare_async(mk(i + 1))
            loss, tb = model(mk(i), prepared=prep)
            opt.zero_grad(set_to_none=True)
            loss.backward()                                        # DDP: NCCL all-reduce of the gradients overlaps here
            if not a.no_prefetch and not a.prefetch_thread:        # like a prefetching DataLoader: H2D + voxelize + rulebooks of step i+1
                pending.clear()                                    # run on a side stream while this step's backward executes
                pending[(id(src), i + 1)] = net.prepare(mk(i + 1))
            torch.nn.utils.clip_grad_norm_(net.parameters(), 10.0)
            opt.step()
            return loss, net.last_batch_dict["encoded_spconv_tensor"]

        run_resident = lambda i: step(i, resident, resident1, resident_gt)
        run_host = lambda i: step(i, host, host1, host_gt)         # .to(device, non_blocking) happens inside the detector
        ctx = torch.enable_grad()
        h2d = sum(h.numel() * 4 for h in host[:a.batch]) + sum(h.numel() * 4 for h in host1[:a.batch]) + host_gt[0].numel() * 4
    elif a.workload == "stress":
        net = None
        sweep = StressSweep(dev)
        run_resident = lambda i: sweep.step(batch_of(i, resident))
        run_host = lambda i: sweep.step(batch_of(i, host))
        ctx = torch.no_grad()
        h2d = sum(h.numel() * 4 for h in host[:a.batch])
    else:
        net = make_net(dev)
        to_bev = backbone.HeightCompression()

        def fwd(points_list):
            bd = voxel.voxelize_batch([p if p.is_cuda else p.to(dev, non_blocking=True) for p in points_list], PC_RANGE, VOXEL_SIZE)
            bd["batch_size"] = len(points_list)
            out = to_bev(net(bd))
            return out["spatial_features"].sum(), out["encoded_spconv_tensor"]

        run_resident = lambda i: fwd(batch_of(i, resident))
        run_host = lambda i: fwd(batch_of(i, host))
        ctx = torch.no_grad()
        h2d = sum(h.numel() * 4 for h in host[:a.batch])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi needs 0.5-2.5 s to initialise NVML and deliver its first sample and takes driver-wide locks while it does: it is
    # started early (above) and the warm-up only begins once its first sample has arrived, so that start-up cannot land inside a
    # timed step; its statistics are taken over the timed phases only (sampler.mark()).
    if rank == 0:
        sampler.wait_first(15.0)
    # Python's cyclic garbage collector is driven by hand, as training loops at scale do (a generation-2 pass over the
    # autograd / rulebook object graph takes 10-30 ms of host time and showed up as isolated 54-78 ms steps): collected
    # between the phases below, never inside a timed step.
    import gc
    gc.collect()
    gc.disable()
    with ctx:
        for i in range(a.warmup):
            flush.zero_()
            res, enc = run_resident(i)             # (kept across the next step exactly as in the timed loop: the previous step's
        barrier()                                  # outputs -- and the rulebooks they reference -- set the allocator's high-water mark)
        gc.collect()
        if train:
            # Keep the caching allocator from growing inside a timed step.  A device allocation (cuMemCreate / cuMemMap under
            # expandable segments) drains the queued work and stalled single steps by 10-2300 ms.  Two causes were found: the
            # timed loop keeps the previous step's outputs (and the rulebooks they reference) alive across the next step while
            # the warm-up loop did not -- now it does, so the warm-up reaches the same high-water mark -- and the two batches of
            # the pool differ in size, so both streams' pools keep creeping up by a few tens of MB for ~25 steps: allocating and
            # releasing headroom once leaves it cached (mapped) in each stream's pool.
            for st in (torch.cuda.current_stream(dev), getattr(net, "_side", None)):
                if st is not None:
                    with torch.cuda.stream(st):
                        pad = [torch.empty(1 << 30, dtype=torch.uint8, device=dev) for _ in range(3)]
                        del pad
            barrier()
        # ---- timed region: K steps, inputs resident in HBM, L2 flushed between steps ----
        l0 = _lib.launch_count()
        evs = []
        alloc_trace = []
        sampler.mark()
        barrier()
        for i in range(a.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res, enc = run_resident(i)
            e1.record()
            evs.append((e0, e1))
            ms_ = torch.cuda.memory_stats(dev)
            alloc_trace.append((ms_.get("reserved_bytes.all.current", 0), ms_.get("num_device_alloc", 0), ms_.get("num_device_free", 0),
                                ms_.get("num_alloc_retries", 0), time.perf_counter()))
        barrier()
        launches = _lib.launch_count() - l0
        if train and not a.no_graph:            # kernels replayed from the captured graphs do not pass through the library's counter
            launches += net.dense_graph_launches * a.steps
        step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
        total_ms = sum(step_ms)
        if os.environ.get("CPD_BENCH_DIAG"):          # which steps grew the allocator / how long the host took to enqueue them
            for i, (rb_, na, nf, nr, tp) in enumerate(alloc_trace):
                sys.stderr.write(f"[diag] step {i}: {step_ms[i]:7.2f} ms  reserved {rb_ / 1e9:6.2f} GB  device allocs {na} frees {nf} retries {nr}  "
                                 f"host dt {1e3 * (tp - (alloc_trace[i - 1][4] if i else tp)):7.2f} ms\n")
        # ---- roofline pass: the same K steps again with CUDA events around every gather-GEMM / wgrad launch
        #      (per-launch events perturb the step, so `value` above comes from the clean pass) ----
        ops.PROFILE = []
        evs_p = []
        gc.collect()
        barrier()
        for i in range(a.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run_resident(i)
            e1.record()
            evs_p.append((e0, e1))
        barrier()
        prof, ops.PROFILE = ops.PROFILE, None
        prof_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs_p)
        # ---- end-to-end: host (pinned) inputs in, result scalar out, every step ----
        for i in range(a.warmup):                # the host path has its own first-use costs (staging buffers of the H2D copies in the
            res, enc = run_host(i)               # stream's allocator pool): W untimed steps of it, like the resident leg
        evs2 = []
        d2h = 0
        gc.collect()
        barrier()
        for i in range(a.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res, enc = run_host(i)
            out = torch.stack([res.detach().float(), enc.features.detach().abs().max()]).cpu()
            e1.record()
            evs2.append((e0, e1))
            d2h = out.numel() * 4
        barrier()
        e2e_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs2)
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])
    frames_total = a.batch * a.steps * world

    # ---- roofline of the dominant kernel family (gather-GEMM), per layer shape; front-end (voxelizer, rulebooks) entries ----
    hbm, bf16, src = peaks()
    groups, front, pcache, stages = {}, {}, {}, {}

    def pairs(t):                       # device scalar made by ops._prof_end
        if id(t) not in pcache:
            pcache[id(t)] = int(t.item())
        return pcache[id(t)]

    for e0, e1, m in prof:
        kind, ms = m["kind"], e0.elapsed_time(e1)
        if kind in ("gather_gemm", "gather_wgrad", "conv2d", "convt2d"):
            key = (kind, m["cin"], m["cout"], m["K"])
            P = m["P"] if "P" in m else pairs(m["pairs"])        # dense TMA convs: every tap of every output pixel
            g = groups.setdefault(key, dict(ms=0.0, n=0, bytes=0.0, flops=0.0))
            g["ms"] += ms; g["n"] += 1
            g["bytes"] += 4.0 * (m["m_in"] * m["cin"] + m["m_out"] * m["cout"]) + 8.0 * P + 4.0 * m["K"] * m["cin"] * m["cout"]
            g["flops"] += 2.0 * P * m["cin"] * m["cout"]
            continue
        # front end: algorithmic bytes per SURVEY 8d, in the layouts this implementation writes
        if kind == "voxelize":
            M = int(m["counts"][m["batch"]].item())
            byt = 4.0 * m["n"] * m["c"] + M * (16.0 + 4.0 + (4.0 * m["c"] if m["want_mean"] else 0.0) +
                                               (4.0 * m["max_pts"] * m["c"] if m["want_voxels"] else 0.0))
            note = f"N={m['n']} M={M}"
        elif kind == "rulebook_subm":
            P = pairs(m["pairs"])
            byt = 16.0 * m["m_in"] + 4.0 * m["m_out"] * m["K"]
            stages[(m["m_out"], m["K"])] = dict(rows=m["m_out"], taps=m["K"], pairs=P)
            note = f"M={m['m_out']} P={P}"
        elif kind == "rulebook_strided_outputs":
            mo = int(m["n_out"].item())
            byt = 16.0 * m["m_in"] + 16.0 * mo
            note = f"M_in={m['m_in']} M_out={mo}"
        else:                                                    # rulebook_strided_tables
            P = pairs(m["pairs"])
            byt = 16.0 * (m["m_in"] + m["m_out"]) + 4.0 * m["K"] * (m["m_out"] + (m["m_in"] if m["both"] else 0))
            note = f"M_in={m['m_in']} M_out={m['m_out']} P={P}"
        f = front.setdefault(kind, dict(ms=0.0, n=0, bytes=0.0, last=""))
        f["ms"] += ms; f["n"] += 1; f["bytes"] += byt; f["last"] = note
    gg_ms = sum(g["ms"] for g in groups.values())
    front_end = [dict(kernel=k, bound="hbm", launches_per_step=f["n"] / a.steps, ms_per_step=f["ms"] / a.steps, avg_launch_us=1e3 * f["ms"] / f["n"],
                      achieved=f["bytes"] / f["ms"] / 1e6, unit="GB/s", peak=hbm, frac=f["bytes"] / f["ms"] / 1e6 / hbm,
                      share_of_step=f["ms"] / total_ms, last_launch=f["last"]) for k, f in sorted(front.items(), key=lambda kv: -kv[1]["ms"])]
    if os.environ.get("CPD_BENCH_GROUPS"):
        rows = sorted(((k, g) for k, g in groups.items()), key=lambda kv: -kv[1]["ms"])
        with open(os.environ["CPD_BENCH_GROUPS"], "w") as f:
            f.write(f"# per-step totals over {a.steps} steps of the roofline pass; step = {prof_ms / a.steps:.2f} ms "
                    f"(clean pass: {total_ms / a.steps:.2f} ms)\n")
            f.write("kind cin cout K launches/step ms/step us/launch GB/s(algorithmic) TFLOP/s(useful)\n")
            for k, g in rows:
                f.write(f"{k[0]} {k[1]} {k[2]} {k[3]} {g['n'] / a.steps:.1f} {g['ms'] / a.steps:.3f} {1e3 * g['ms'] / g['n']:.1f} "
                        f"{g['bytes'] / g['ms'] / 1e6:.0f} {g['flops'] / g['ms'] / 1e9:.1f}\n")
            for fe in front_end:
                f.write(f"{fe['kernel']} - - - {fe['launches_per_step']:.1f} {fe['ms_per_step']:.3f} {fe['avg_launch_us']:.1f} {fe['achieved']:.0f} - "
                        f"({fe['last_launch']})\n")
    top_key, top = max(groups.items(), key=lambda kv: kv[1]["ms"])
    # The kernels run bf16x3 (3 bf16 tensor products per useful fp32-class product): the tensor roofline for USEFUL flops is
    # the measured bf16 peak / 3; a layer is HBM bound when its arithmetic intensity sits below that ridge.
    tensor_peak = bf16 / 3.0

    def entry(g):
        ai = g["flops"] / max(g["bytes"], 1.0)
        if ai * hbm / 1e3 < tensor_peak:
            return dict(bound="hbm", achieved=g["bytes"] / g["ms"] / 1e6, peak=hbm, unit="GB/s")
        return dict(bound="tensor", achieved=g["flops"] / g["ms"] / 1e9, peak=tensor_peak, unit="TFLOP/s")

    roof = entry(top)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    kname = f"{top_key[0]} {top_key[1]}->{top_key[2]} K={top_key[3]}"
    if os.path.exists(tpath):
        ent = json.load(open(tpath)).get(kname)
        if ent:
            traffic = ent["dram_bytes_per_launch"]
    others = []
    n_next = 3 if a.workload != "stress" else 12                 # stress: the whole C sweep
    for k, g in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])[1:1 + n_next]:      # the next layer shapes, for context
        e = entry(g)
        others.append(dict(kernel=f"{k[0]} {k[1]}->{k[2]} K={k[3]}", bound=e["bound"], achieved=e["achieved"], unit=e["unit"],
                           frac=e["achieved"] / e["peak"], avg_launch_us=1e3 * g["ms"] / g["n"], share_of_step=g["ms"] / total_ms))
    roof.update(frac=roof["achieved"] / roof["peak"], traffic=traffic,
                peak_source=f"{src} ({'HBM copy' if roof['bound'] == 'hbm' else 'bf16 dense / 3 (bf16x3 products), useful flops'})",
                kernel=f"{top_key[0]} {top_key[1]}->{top_key[2]} K={top_key[3]}", launches=top["n"],
                avg_launch_us=1e3 * top["ms"] / top["n"], share_of_step=top["ms"] / total_ms,
                gather_family_share_of_step=gg_ms / total_ms,
                timing="CUDA events around every launch, second pass of K steps on the launching stream; shares are against the clean pass "
                       f"({total_ms / a.steps:.1f} ms/step; the instrumented pass ran {prof_ms / a.steps:.1f} ms/step)",
                algorithmic_bytes_per_launch=top["bytes"] / top["n"], flops_per_launch=top["flops"] / top["n"], next=others)

    line = {
        "metric": METRIC, "value": frames_total / (total_ms / 1e3), "unit": "frames/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "ms_steps_rank0": [round(v, 2) for v in step_ms],
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "arith": "bf16x3", "data": "synthetic",
        "config": {"workload": workload_name(a), "frames_per_step_per_gpu": a.batch, "points_per_frame": a.points,
                   "voxel_size": [0.1, 0.1, 0.15], "grid": [1504, 1504, 40], "parallelism": f"dp{world}" + (" (DDP: gradient_as_bucket_view, static_graph, broadcast_buffers=False)" if world > 1 and train else ""),
                   "arithmetic": "fp32 in / fp32 out; convolution products on tcgen05 as bf16x3 (hi.hi + hi.lo + lo.hi, fp32 accumulate): "
                                 "2^-16 per product, not 2^-24",
                   "l2": f"flushed between steps ({L2_FLUSH_BYTES >> 20} MiB write)",
                   "cuda_graphs": ("dense stack (BEV backbone + CenterHead convolutions) forward and backward replayed as CUDA graphs"
                                   if (train and not a.no_graph) else "none"),
                   "input_stage": "inline" if (a.no_prefetch or not train) else "prefetched one step ahead on a side stream (inside the timed region)",
                   "python_gc": "disabled inside the timed steps, gc.collect() between the phases",
                   "active_voxels_last_batch": int(enc.indices.shape[0])},
        "e2e": {"value": frames_total / (e2e_ms / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "allocator_rank0": {"device_allocs_during_timed_steps": int(alloc_trace[-1][1] - alloc_trace[0][1]) if alloc_trace else None,
                            "reserved_gb": round(alloc_trace[-1][0] / 1e9, 2) if alloc_trace else None},
        "peak_hbm_gb_rank0": round(torch.cuda.max_memory_allocated(dev) / 1e9, 2), "clocks": clocks, "roofline": roof, "front_end": front_end,
        "sparse_levels": sorted(stages.values(), key=lambda d: -d["rows"]),
    }
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(a, frames[:1], net if a.workload == "backbone_fwd" else None)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(a, frames, net=None):
    """CPU port of the same step on the host cores (oracle/ for voxelizer + sparse convs, torch CPU for the
    dense BEV head exactly as the reference builds it), all threads, on a bounded sample."""
    import torch
    from cpd_b200.synth import PC_RANGE, VOXEL_SIZE, synth_gt_boxes, synth_scan
    from oracle import oracle as O
    from oracle import pipeline
    cores = os.cpu_count() or 1
    O.set_threads(cores)
    torch.set_num_threads(cores)
    # the reference's voxelizer is single-threaded per DataLoader worker: its 1-thread figure, on the first frame
    tv = time.perf_counter()
    for _ in range(3):
        O.voxelize(frames[0], PC_RANGE, VOXEL_SIZE)
    vox_1t = 3.0 / (time.perf_counter() - tv)
    if a.workload == "stress":
        rng = np.random.default_rng(3)
        ws = {c: (rng.normal(0, 1, (c, 27, c)) * (27 * c) ** -0.5).astype(np.float32) for c in StressSweep.CH}
        t0 = time.perf_counter()
        n = 0
        while True:
            pts = frames[n % len(frames)]
            v, c3, nn = O.voxelize(pts, PC_RANGE, VOXEL_SIZE)
            O.mean_vfe(v, nn)
            coords, shape = np.concatenate([np.zeros((len(c3), 1), np.int32), c3], 1), [41, 1504, 1504]
            for st in range(4):
                rb = O.rulebook_subm(coords, shape, 3)
                for c in (StressSweep.CH if st == 0 else (StressSweep.CH[st],)):
                    O.spconv_fwd(rng.normal(0, 1, (len(coords), c)).astype(np.float32), ws[c], None, rb)
                if st < 3:
                    rbd = O.rulebook_strided(coords, shape, 3, 2, StressSweep.PADS[st])
                    coords, shape = rbd.out_coords, rbd.out_shape
            rbo = O.rulebook_strided(coords, shape, (3, 1, 1), (2, 1, 1), 0)
            O.dense(rng.normal(0, 1, (rbo.m_out, 128)).astype(np.float32), rbo.out_coords, 1, rbo.out_shape)
            n += 1
            dt = time.perf_counter() - t0
            if dt > 10.0 or n >= 4:
                break
        return {"value": n / dt, "unit": "frames/s", "cores": cores, "kind": "port", "voxelizer_1thread_frames_per_s": vox_1t,
                "sample": f"{n} frame(s) of the same sweep ({a.points} pts/frame): oracle/cpd_oracle.c voxelizer (1 thread, as the reference's) + "
                          f"rulebooks + SubM convolutions + dense() (OpenMP, {cores} threads)"}
    if a.workload == "backbone_fwd":
        if net is None:
            net = make_net(torch.device("cpu"))
        pipeline.backbone_forward(net, [frames[0][:20000]], PC_RANGE, VOXEL_SIZE, eval_wide=True)      # warm caches / page in
        t0 = time.perf_counter()
        n = 0
        while True:
            feats, coords, shape, _ = pipeline.backbone_forward(net, [frames[n % len(frames)]], PC_RANGE, VOXEL_SIZE, eval_wide=True)
            pipeline.bev_dense(feats, coords, 1, shape)
            n += 1
            dt = time.perf_counter() - t0
            if dt > 10.0 or n >= 8:
                break
        what = "fwd"
    else:
        import warnings
        warnings.filterwarnings("ignore")
        cpu = pipeline.CpuDetector(make_detector(torch.device("cpu")))
        f1 = [synth_scan(a.points, 777)]
        gt = np.stack([synth_gt_boxes(30, 0)])
        cpu.train_step([frames[0][:16000]], [f1[0][:16000]], gt)                                        # warm-up on a small cloud
        t0 = time.perf_counter()
        n = 0
        while True:
            cpu.train_step([frames[n % len(frames)]], f1, gt)
            n += 1
            dt = time.perf_counter() - t0
            if dt > 12.0 or n >= 4:
                break
        what = "fwd+bwd train step (bs=1, no optimizer update)"
    return {"value": n / dt, "unit": "frames/s", "cores": cores, "kind": "port", "voxelizer_1thread_frames_per_s": vox_1t,
            "sample": f"{n} frame(s) of the same workload ({a.points} pts/frame), {what}; oracle/cpd_oracle.c (OpenMP) + torch CPU "
                      f"for the dense head, {cores} threads"}


def run_reference(a):
    """The reference arm: the CPU port of the SAME step (bs = --batch frames of --points points, forward + backward +
    grad-clip + Adam) on all host cores.  A full step takes ~15-20 s on the box's cores, so the run does one reduced
    warm-up step and then as many FULL steps as fit a fixed time budget (at least one); `steps` reports how many."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import warnings
    import torch
    warnings.filterwarnings("ignore")
    from cpd_b200.synth import PC_RANGE, VOXEL_SIZE
    from oracle import oracle as O
    from oracle import pipeline
    cores = os.cpu_count() or 1
    O.set_threads(cores)
    torch.set_num_threads(cores)
    budget_s = float(os.environ.get("CPD_REF_BUDGET_S", "150"))
    frames = make_frames(0, a.batch, a.points)
    if a.workload == "backbone_fwd":
        net = make_net(torch.device("cpu"))
        pipeline.backbone_forward(net, [frames[0][:20000]], PC_RANGE, VOXEL_SIZE, eval_wide=True)

        def step():
            feats, coords, shape, _ = pipeline.backbone_forward(net, frames, PC_RANGE, VOXEL_SIZE, eval_wide=True)
            pipeline.bev_dense(feats, coords, a.batch, shape)
        what = "forward"
    else:
        frames1 = make_frames(500, a.batch, a.points)
        gt = np.stack(make_gt(0, a.batch))
        cpu = pipeline.CpuDetector(make_detector(torch.device("cpu")))
        cpu.train_step([frames[0][:16000]], [frames1[0][:16000]], gt[:1], optimizer=False)      # warm-up: page in, spin up the thread pools
        step = lambda: cpu.train_step(frames, frames1, gt, optimizer=True)
        what = "forward + backward + grad-clip + Adam"
    times = []
    t_start = time.perf_counter()
    for _ in range(max(1, a.steps)):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start + times[-1] > budget_s:
            break
    n = len(times)
    v = a.batch * n / sum(times)
    base = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{n} full step(s) of the same workload: bs={a.batch} x {a.points} pts/frame, {what}; oracle/cpd_oracle.c (OpenMP) + torch CPU "
                      f"for the dense head, {cores} threads; one process (rank 0) regardless of --gpus"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": a.gpus, "steps": n,
        "warmup": 1, "ms_per_step": 1e3 * sum(times) / n, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "frames_per_step_per_gpu": a.batch, "points_per_frame": a.points,
                   "steps_requested": a.steps, "steps_timed": n, "time_budget_s": budget_s,
                   "warmup_note": "one reduced warm-up step (bs=1, 16k points) instead of --warmup full steps: a full CPU step takes ~15-20 s",
                   "note": "spconv-cu111 is not installable here; this arm is the CPU oracle port of the same path, ONE host process "
                           "(rank 0) also when --gpus > 1"},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
