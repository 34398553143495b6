#!/usr/bin/env python
"""torchrun --nproc-per-node G tools/mgpu_breakdown.py : where a batch-sharded step spends its device time (CUDA events on
the main stream): head forward (+ gather launch) | loss + backward | wait for the side-stream exchange + enqueue | EM."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                    # noqa: E402
from mgproto_b200 import parallel               # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_%h_%p.log")
    dist.init_process_group("nccl", device_id=dev)
net = bench.build_model(dev)
if world > 1:
    parallel.attach(net)
c = bench.CFG
g = torch.Generator().manual_seed(1 + rank)
xs = [torch.randn(c["B"], c["D"], c["H"], c["W"], generator=g).to(dev) for _ in range(4)]
gts = [torch.randint(0, c["C"], (c["B"],), generator=g).to(dev) for _ in range(4)]
for overlap in (True, False):
    net.overlap_enqueue = overlap
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(30)]
    for it in range(35):
        x, gt = xs[it % 4], gts[it % 4]
        x.grad = None
        x.requires_grad_(True)
        e = ev[it - 5] if it >= 5 else None
        if e: e[0].record()
        out = net.head(x, gt)
        if e: e[1].record()
        bench.loss_fn(out, gt).backward()
        if e: e[2].record()
        net.wait_enqueue()
        if e: e[3].record()
        net.update_GMM()
        if e: e[4].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    seg = [sum(e[i].elapsed_time(e[i + 1]) for e in ev) / len(ev) * 1e3 for i in range(4)]
    tot = sum(ev[i][0].elapsed_time(ev[i + 1][0]) for i in range(len(ev) - 1)) / (len(ev) - 1) * 1e3
    if rank == 0:
        print("world %d overlap %s: head %.0f | loss+bwd %.0f | wait side %.0f | update_GMM %.0f | step (start to start) %.0f us"
              % (world, overlap, seg[0], seg[1], seg[2], seg[3], tot), flush=True)
if world > 1:
    dist.destroy_process_group()
