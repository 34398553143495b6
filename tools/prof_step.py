#!/usr/bin/env python
"""Run a few hot-path steps at the bench workload (driver for ncu captures of the non-GEMM kernels)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
net = bench.build_model(dev)
c = bench.CFG
g = torch.Generator().manual_seed(1)
x = torch.randn(c["B"], c["D"], c["H"], c["W"], generator=g).to(dev)
gt = torch.randint(0, c["C"], (c["B"],), generator=g).to(dev)
for _ in range(steps):
    xr = x.clone().requires_grad_(True)
    out = net.head(xr, gt)
    bench.loss_fn(out, gt).backward()
    net.update_GMM()
torch.cuda.synchronize()
print("ok")
