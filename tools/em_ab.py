#!/usr/bin/env python
"""Time mgp_em_stats at the bench shapes for several n_split values (MGP_EM_LEGACY=1 selects the generic kernel)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mgproto_b200 import ops  # noqa: E402

C, K, D, cap = 200, 10, 128, 800
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
bank = torch.nn.functional.normalize(torch.randn(C, cap, D, generator=g), dim=-1).to(dev)
mu = torch.nn.functional.normalize(torch.randn(C, K, D, generator=g), dim=-1).to(dev)
sg = torch.full((C, K, D), 0.3989, device=dev)
wt = torch.zeros(C, C * K, device=dev)
for c in range(C):
    wt[c, c * K:(c + 1) * K] = 1.0 / K
order = torch.arange(C, dtype=torch.int32, device=dev)
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
for ns in (1, 2, 3, 4, 8):
    stats = torch.empty(C, ns, ops.em_stat_stride(K, D), device=dev)
    ts = []
    for it in range(8):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.em_stats(bank, order, mu, sg, wt, 0.1, stats, ns)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print("n_split", ns, "us min %.1f med %.1f" % (min(ts[2:]), sorted(ts[2:])[len(ts[2:]) // 2]),
          "S0 sum %.3f" % float(stats[:, :, :K].sum()))
