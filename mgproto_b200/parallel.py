"""Multi-GPU plumbing for the hot path: one process per GPU,
``torch.distributed`` (NCCL on the B200 box, gloo in CPU tests).

The path shards by image: the head (log-likelihood, top-T, logits, backward) touches one
image's patches against replicated prototypes -- no communication.  Two exchanges keep every
rank's memory bank and prototypes identical to a single-GPU run on the concatenated batch:

* ``all_gather_mined``: the per-image mined rows (top1 [B,K], rows [B,K,D], gt [B]; ~1.3 MB at
  B=256, K=10, D=128) are all-gathered in rank (= global image) order before the enqueue, so
  every replica of the bank receives the whole global batch in the reference's order;
* ``update_GMM`` needs no exchange: the bank is replicated, the EM kernels are deterministic, so every rank runs
  the whole (single-launch) update on its replica and all replicas stay bit-identical.  With
  ``model.em_shard = True`` the rows are sharded instead (``shard_rows``) and the packed statistics
  [C, 1, K + K*D + 1] are all-reduced once per EM loop -- the layout for banks too large to re-reduce per rank.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_rows(cap: int, world: int, rank: int):
    """Contiguous bank-slot range [r0, r1) reduced by ``rank`` (the last ranks may be empty)."""
    per = (cap + world - 1) // world
    return min(cap, rank * per), min(cap, (rank + 1) * per)


def shard_batch(global_batch: int, world: int, rank: int):
    """Image range [b0, b1) of ``rank`` for an evenly divisible global batch."""
    if global_batch % world != 0:
        raise ValueError("global batch %d not divisible by world size %d" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


def all_gather_mined(top1, rows, gt, group=None):
    """Concatenate every rank's (top1 [b,K] int32, rows [b,K,D] fp32, gt [b] int64) in rank order.
    One collective: the three tensors travel as one [b, K*D + K + 2] fp32 record per image (bit-casts)."""
    world = dist.get_world_size(group)
    if world == 1:
        return top1, rows, gt
    b, K, D = rows.shape
    if top1.dtype != torch.int32 or gt.dtype != torch.int64 or rows.dtype != torch.float32:
        raise TypeError("all_gather_mined expects int32 top1, fp32 rows, int64 gt")
    rec = torch.cat([rows.reshape(b, K * D), top1.contiguous().view(torch.float32),
                     gt.contiguous().view(torch.float32).reshape(b, 2)], dim=1)
    buf = torch.empty((world * b, K * D + K + 2), dtype=torch.float32, device=rows.device)
    dist.all_gather_into_tensor(buf, rec, group=group)
    rows_all = buf[:, :K * D].reshape(world * b, K, D).contiguous()
    top1_all = buf[:, K * D:K * D + K].contiguous().view(torch.int32)
    gt_all = buf[:, K * D + K:].contiguous().view(torch.int64).reshape(world * b)
    return top1_all, rows_all, gt_all


def attach(model, group=None):
    """Turn on the sharded enqueue / EM for ``model`` (an ``MGProto``) on this process group."""
    model.em_group = group if group is not None else dist.group.WORLD
    return model
