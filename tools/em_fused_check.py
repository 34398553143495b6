#!/usr/bin/env python
"""update_GMM: fused cluster kernel vs the multi-launch path (MGP_EM_UNFUSED=1 in a child process) -- same state in, compare out; time both."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(tag):
    import bench
    dev = torch.device("cuda:0")
    net = bench.build_model(dev)
    ts = []
    for it in range(6):
        net.queue.updated.fill_(1)
        if it == 3:
            net.queue.updated[::3] = 0          # some inactive classes
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        net.update_GMM()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    net.sync_optimizer_state()
    st = net.prototype_optimizer.state[net.prototype_means]
    torch.save({"mu": net.prototype_means.data.cpu(), "w": net.last_layer.weight.data.cpu(), "m": st["exp_avg"].cpu(),
                "v": st["exp_avg_sq"].cpu(), "step": float(st["step"])}, "/tmp/em_%s.pt" % tag)
    print(tag, "update_GMM us:", [round(t, 1) for t in ts])


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
        sys.exit(0)
    env = dict(os.environ)
    subprocess.check_call([sys.executable, __file__, "fused"], env=env)
    env["MGP_EM_UNFUSED"] = "1"
    subprocess.check_call([sys.executable, __file__, "unfused"], env=env)
    a, b = torch.load("/tmp/em_fused.pt"), torch.load("/tmp/em_unfused.pt")
    for k in ("mu", "w", "m", "v"):
        d = (a[k] - b[k]).abs().max().item()
        print(k, "max abs diff", d, "max abs", b[k].abs().max().item())
    print("step", a["step"], b["step"])
