#!/usr/bin/env python
"""Launch the log-likelihood kernel a few times at the bench size (for ncu captures)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mgproto_b200 import ops  # noqa: E402

math = sys.argv[1] if len(sys.argv) > 1 else "tc"
layout = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
B, HW, P, D = 256, 196, 2000, 128
g = torch.Generator().manual_seed(1)
dev = torch.device("cuda:0")
x = F.normalize(torch.randn(B * HW, D, generator=g), dim=1).to(dev)
mu = F.normalize(torch.rand(P, D, generator=g), dim=1).to(dev)
sg = torch.full((P, D), 0.3989422804, device=dev)
for _ in range(reps):
    out = ops.logprob(x, mu, sg, layout, B=B, HW=HW, math=math)
torch.cuda.synchronize()
print("ok", float(out[0, 0] if layout == 0 else out[0, 0, 0]))
