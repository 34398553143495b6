#!/usr/bin/env python
"""torchrun --nproc-per-node G tools/check_multigpu.py
Parity of the batch-sharded path: G ranks, each with B/G images, must end with the same
bank / mu / pi as ONE rank processing the concatenated batch (checked on rank 0 against a local replay)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mgproto_b200 as M  # noqa: E402
from mgproto_b200 import parallel  # noqa: E402


def make(dev, C, K, D, T, cap):
    torch.manual_seed(0)
    net = M.MGProto(features=nn.Sequential(nn.Conv2d(3, 8, 1)), img_size=14, prototype_shape=(C * K, D, 1, 1),
                    proto_layer_rf_info=None, num_classes=C, add_on_layers_type="regular", sz_embedding=8,
                    mem_capacity=cap, mine_K=T).to(dev)
    net.prototype_optimizer = torch.optim.Adam([{"params": net.prototype_means, "lr": 3e-3}])
    net.train()
    return net


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    C, K, D, T, cap, H, W = 12, 5, 128, 8, 16, 14, 14
    Bg = 8 * world
    net = parallel.attach(make(dev, C, K, D, T, cap))
    net.em_shard = len(sys.argv) > 1 and sys.argv[1] == "shard"      # default: every rank runs the whole EM on its replica
    ref = make(dev, C, K, D, T, cap) if rank == 0 else None
    g = torch.Generator().manual_seed(3)
    ok = True
    for it in range(6):
        x = torch.randn(Bg, D, H, W, generator=g)
        gt = torch.randint(0, C, (Bg,), generator=g)
        b0, b1 = parallel.shard_batch(Bg, world, rank)
        with torch.no_grad():
            net.head(x[b0:b1].to(dev), gt[b0:b1].to(dev))
        net.update_GMM()
        if rank == 0:
            with torch.no_grad():
                ref.head(x.to(dev), gt.to(dev))
            ref.update_GMM()
            c1 = torch.equal(net.queue.mem_len, ref.queue.mem_len)
            c2 = torch.equal(net.queue.linear(), ref.queue.linear())
            dmu = (net.prototype_means - ref.prototype_means).abs().max().item()
            dw = (net.last_layer.weight - ref.last_layer.weight).abs().max().item()
            # the shards sum the statistics in a different order (fp32): allow Adam-amplified rounding
            c3 = dmu < 2e-5 and dw < 1e-6
            print("it %d: mem_len %s bank %s |dmu| %.2e |dpi| %.2e" % (it, c1, c2, dmu, dw), flush=True)
            ok &= c1 and c2 and c3
    # every rank holds the same replica
    t = net.prototype_means.detach().clone()
    dist.broadcast(t, 0)
    same = torch.tensor([float(torch.equal(t, net.prototype_means.detach()))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("em_shard", net.em_shard, "multi-gpu parity:", "OK" if ok else "MISMATCH", "| replicas identical:", bool(same.item()),
              "| mem_len", net.queue.mem_len.tolist())
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
