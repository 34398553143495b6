#!/usr/bin/env python
"""Per-stage device times of one hot-path step at the bench workload (CUDA events, median of reps)."""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mgproto_b200 import ops  # noqa: E402


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


def main():
    math = sys.argv[1] if len(sys.argv) > 1 else "auto"
    dev = torch.device("cuda:0")
    c = bench.CFG
    net = bench.build_model(dev)
    net.math_mode = math
    B, D, H, W = c["B"], c["D"], c["H"], c["W"]
    HW, C, K, T = H * W, c["C"], c["K"], c["T"]
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, D, H, W, generator=g).to(dev)
    gt = torch.randint(0, C, (B,), generator=g).to(dev)
    mu = net.prototype_means.detach().reshape(C * K, D).contiguous()
    sg = net.prototype_covs.detach().reshape(C * K, D).contiguous()
    wt = net.last_layer.weight.data
    res = {}
    res["normalize_fwd"] = timeit(lambda: ops.normalize_fwd(x))
    xhat, inv, _ = ops.normalize_fwd(x)
    for m in ("fp32", "tc"):
        try:
            res["logprob_NP_" + m] = timeit(lambda: ops.logprob(xhat, mu, sg, 0, math=m))
            res["logprob_BPHW_" + m] = timeit(lambda: ops.logprob(xhat, mu, sg, 1, B=B, HW=HW, math=m))
        except Exception as e:  # noqa: BLE001
            res["logprob_" + m] = "unavailable: %s" % e
    lp = ops.logprob(xhat, mu, sg, 1, B=B, HW=HW, math=math)
    res["head_select"] = timeit(lambda: ops.head_select(lp, wt, gt, T, C, K))
    lpn = ops.logprob(xhat, mu, sg, 0, math=math)
    res["head_select_np"] = timeit(lambda: ops.head_select(lpn, wt, gt, T, C, K, B=B, HW=HW))
    del lpn
    logits, vals, idx = ops.head_select(lp, wt, gt, T, C, K)
    xr = x.clone().requires_grad_(True)

    def fwd_bwd():
        xr.grad = None
        out = net.head(xr, gt)
        bench.loss_fn(out, gt).backward()
    res["head_fwd+loss+bwd+enqueue"] = timeit(fwd_bwd)

    def fwd_only():
        with torch.no_grad():
            ops.head_forward(x, net.prototype_means, net.prototype_covs, wt, gt, T, math)
    res["head_fwd_only"] = timeit(fwd_only)
    out = net.head(xr, gt)
    res["loss_fwd_bwd_torch(20 CE)"] = timeit(lambda: torch.autograd.grad(bench.loss_fn(out, gt), out, retain_graph=True))
    gl = torch.randn_like(out)
    res["head_bwd_only"] = timeit(lambda: torch.autograd.grad(out, xr, gl, retain_graph=True))
    top1, rows = ops.mined_gather(xhat, idx, gt, HW, C, K)

    def enq():
        ops.mined_gather(xhat, idx, gt, HW, C, K)
        q = net.queue
        ops.bank_enqueue(q.bank, q.mem_len, q.head, q.updated, rows, top1, gt)
    res["gather+enqueue"] = timeit(enq)

    def em():
        net.queue.updated.fill_(1)
        net.update_GMM()
    res["update_GMM(200 classes)"] = timeit(em)
    order = torch.arange(C, dtype=torch.int32, device=dev)
    stats = torch.empty(C, net.em_n_split, ops.em_stat_stride(K, D), device=dev)
    res["em_stats"] = timeit(lambda: ops.em_stats(net.queue.bank, order, net.prototype_means.data,
                                                  net.prototype_covs.data, wt, 0.1, stats, net.em_n_split))
    # host cost of enqueueing one whole step (python + ctypes + autograd), GPU kept busy so nothing blocks:
    # if this exceeds the device time per step the loop is launch-bound
    import time

    def step():
        xr.grad = None
        out = net.head(xr, gt)
        bench.loss_fn(out, gt).backward()
        net.update_GMM()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    n = 40
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    t1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    res["step_host_enqueue_us"] = (t1 - t0) / n * 1e6
    res["step_device_us"] = e0.elapsed_time(e1) / n * 1e3
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
