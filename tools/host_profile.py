#!/usr/bin/env python
"""Host-side cost of enqueueing one hot-path step (cProfile over N steps, GPU left to run behind)."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda:0")
net = bench.build_model(dev)
c = bench.CFG
g = torch.Generator().manual_seed(1)
x = torch.randn(c["B"], c["D"], c["H"], c["W"], generator=g).to(dev)
gt = torch.randint(0, c["C"], (c["B"],), generator=g).to(dev)


def step():
    x.grad = None
    x.requires_grad_(True)
    out = net.head(x, gt)
    bench.loss_fn(out, gt).backward()
    net.update_GMM()


for _ in range(5):
    step()
torch.cuda.synchronize()
n = 50
parts = {"head": 0.0, "loss+bwd": 0.0, "update_GMM": 0.0}
t_all0 = time.perf_counter()
for _ in range(n):
    x.grad = None
    x.requires_grad_(True)
    t0 = time.perf_counter()
    out = net.head(x, gt)
    t1 = time.perf_counter()
    bench.loss_fn(out, gt).backward()
    t2 = time.perf_counter()
    net.update_GMM()
    t3 = time.perf_counter()
    parts["head"] += t1 - t0
    parts["loss+bwd"] += t2 - t1
    parts["update_GMM"] += t3 - t2
t_all1 = time.perf_counter()
torch.cuda.synchronize()
t_all2 = time.perf_counter()
print({k: round(v / n * 1e6, 1) for k, v in parts.items()}, "host us/step", round((t_all1 - t_all0) / n * 1e6, 1),
      "incl. drain", round((t_all2 - t_all0) / n * 1e6, 1))
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
