"""world_size-2 gloo tests (CPU) of the N>1 host logic: batch / bank-row sharding, the
rank-ordered all-gather of the mined rows, and the EM sufficient-statistics all-reduce
(sum of the shards' statistics == statistics of the whole bank)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_stats(x, mu, sigma, pi, alpha):
    from oracle import mgproto_oracle as O
    _, log_resp = O.e_step(x, mu, sigma, pi)
    r = O.smooth_resp(log_resp, alpha)
    return np.concatenate([r.sum(0), (r.T @ x).reshape(-1)])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mgproto_b200 import parallel
    ok = True
    # 1. shards partition the range
    cap, B = 11, 8
    r0, r1 = parallel.shard_rows(cap, world, rank)
    b0, b1 = parallel.shard_batch(B, world, rank)
    cover = torch.zeros(cap)
    cover[r0:r1] = 1
    dist.all_reduce(cover)
    ok &= bool((cover == 1).all()) and (b1 - b0) == B // world
    # 2. all-gather keeps global image order
    K, D = 3, 4
    top1 = torch.arange(b0 * K, b1 * K, dtype=torch.int32).view(-1, K)
    rows = torch.arange(b0 * K * D, b1 * K * D, dtype=torch.float32).view(-1, K, D)
    gt = torch.arange(b0, b1, dtype=torch.int64)
    t, r, g = parallel.all_gather_mined(top1, rows, gt)
    ok &= t.flatten().tolist() == list(range(B * K)) and g.tolist() == list(range(B))
    ok &= r.flatten().tolist() == [float(i) for i in range(B * K * D)]
    # the packed record bit-casts int32 / int64 fields through fp32 slots: patterns that are NaN / Inf as floats
    # (labels -1, indices 0x7fc00000) must come back unchanged
    gt2 = torch.full((b1 - b0,), -1, dtype=torch.int64)
    top2 = torch.full((b1 - b0, K), 0x7fc00000, dtype=torch.int32)
    t2, r2, g2 = parallel.all_gather_mined(top2, rows, gt2)
    ok &= bool((g2 == -1).all()) and bool((t2 == 0x7fc00000).all()) and g2.numel() == B and t2.shape == (B, K)
    ok &= r2.shape == (B, K, D) and t2.dtype == torch.int32 and g2.dtype == torch.int64
    # 3. sharded EM statistics, all-reduced == full-bank statistics
    rng = np.random.default_rng(0)
    Kc, Dd = 4, 8
    x = rng.standard_normal((cap, Dd))
    mu = rng.standard_normal((Kc, Dd))
    sg = 0.3 + rng.random((Kc, Dd))
    pi = np.full(Kc, 1.0 / Kc)
    part = _oracle_stats(x[r0:r1], mu, sg, pi, 0.1) if r1 > r0 else np.zeros(Kc + Kc * Dd)
    tt = torch.from_numpy(part)
    dist.all_reduce(tt)
    ok &= bool(np.allclose(tt.numpy(), _oracle_stats(x, mu, sg, pi, 0.1), rtol=1e-10))
    out[rank] = int(ok)
    dist.destroy_process_group()


def test_world2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1] * world
