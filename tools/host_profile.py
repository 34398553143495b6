#!/usr/bin/env python
"""Where the HOST time of one training step goes (Python / ctypes / torch dispatch): cProfile over 200 enqueued steps
of the bench's step (no synchronisation inside), top functions by cumulative and by own time."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                   # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
net = bench.build_model(dev)
c = bench.CFG
g = torch.Generator().manual_seed(1)
feats = [torch.randn(c["B"], c["D"], 14, 14, generator=g).to(dev) for _ in range(4)]
gts = [torch.randint(0, c["C"], (c["B"],), generator=g).to(dev) for _ in range(4)]


def step(x, gt):
    x.grad = None
    x.requires_grad_(True)
    out = net.head(x, gt)
    loss = bench.loss_fn(out, gt)
    loss.backward()
    net.update_GMM()


for i in range(10):
    step(feats[i % 4], gts[i % 4])
torch.cuda.synchronize()
n = 200
t0 = time.perf_counter()
for i in range(n):
    step(feats[i % 4], gts[i % 4])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue %.1f us/step; with the final sync %.1f us/step" % ((t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
pr = cProfile.Profile()
pr.enable()
for i in range(n):
    step(feats[i % 4], gts[i % 4])
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(35)
st.sort_stats("tottime").print_stats(25)
