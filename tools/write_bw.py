#!/usr/bin/env python
"""Pure-write and copy bandwidth probes (what an output-dominated kernel can hope for)."""
import torch


def t(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


dev = torch.device("cuda:0")
n = 50176 * 2000
a = torch.empty(n, device=dev)
b = torch.empty(n, device=dev)
big = torch.empty(4 * n, device=dev)
print("fill_ 401MB       : %.1f us  %.2f TB/s" % (t(lambda: a.fill_(1.0)) * 1e6, 4 * n / t(lambda: a.fill_(1.0)) / 1e12))
print("zero_ 401MB       : %.1f us  %.2f TB/s" % (t(lambda: a.zero_()) * 1e6, 4 * n / t(lambda: a.zero_()) / 1e12))
print("fill_ 1.6GB       : %.2f TB/s" % (16 * n / t(lambda: big.fill_(1.0)) / 1e12))
print("copy 401MB->401MB : %.2f TB/s (r+w)" % (8 * n / t(lambda: b.copy_(a)) / 1e12))
x = torch.randn(50176, 128, device=dev)
print("mul out=401MB (x[:, :1] * ones): %.2f TB/s" % (4 * n / t(lambda: torch.mul(x[:, :1], 2.0, out=None).expand(50176, 2000).contiguous()) / 1e12))
