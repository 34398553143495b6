#!/usr/bin/env python
"""Time compute_log_prob's [N,P] op at cfg2 (B=256, P=2000, D=128): TMEM-resident kernel (fused split) vs the
pre-pass kernel, whole op (as called) with rotating inputs/outputs; optional team sizes."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mgproto_b200 import _lib, ops                 # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
lib = _lib.load()
B, HW, P, D = 256, 196, 2000, int(os.environ.get("KA_D", "128"))
N = B * HW
g = torch.Generator().manual_seed(0)
xs = [F.normalize(torch.randn(N, D, generator=g), dim=1).to(dev) for _ in range(6)]
mu = F.normalize(torch.rand(P, D, generator=g), dim=1).to(dev)
sg = torch.full((P, D), 0.3989, device=dev)
outs = [torch.empty(N, P, device=dev) for _ in range(2)]
abytes = 4.0 * (N * D + 2 * P * D + N * P)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def run(tag):
    for i in range(3):
        ops.logprob(xs[i % 6], mu, sg, 0, math="auto", out=outs[i % 2])
    torch.cuda.synchronize()
    e0.record()
    for i in range(20):
        ops.logprob(xs[i % 6], mu, sg, 0, math="auto", out=outs[i % 2])
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20 * 1e-3
    print("%-28s %.1f us/op  %.0f GB/s  frac %.3f" % (tag, t * 1e6, abytes / t / 1e9, abytes / t / 1e9 / 6575.1))


for z in (1, 0):
    lib.mgp_set_option(b"tc_z", z)
    for team in (["0", "4", "1"] if z else ["4"]):
        os.environ["MGP_TC_TEAM"] = team
        for dbg in (os.environ.get("KA_DEBUGS", "0,1,4").split(",") if (z and team == "0") else ["0"]):
            os.environ["MGP_TC_DEBUG"] = dbg
            run("z=%d team=%s debug=%s" % (z, team, dbg))
os.environ["MGP_TC_DEBUG"] = "0"
