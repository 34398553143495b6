"""Multi-GPU plumbing for the hot path: one process per GPU,
``torch.distributed`` (NCCL on the B200 box, gloo in CPU tests).

The path shards by image: the head (log-likelihood, top-T, logits, backward) touches one
image's patches against replicated prototypes -- no communication.  Two exchanges keep every
rank's memory bank and prototypes identical to a single-GPU run on the concatenated batch:

* ``all_gather_mined``: the per-image mined rows (top1 [B,K], rows [B,K,D], gt [B]; ~1.3 MB at
  B=256, K=10, D=128) are all-gathered in rank (= global image) order before the enqueue, so
  every replica of the bank receives the whole global batch in the reference's order;
* the EM statistics: each rank reduces its row shard ``shard_rows`` of every active class and the
  packed [C, 1, K + K*D + 1] buffer is all-reduced once per EM loop (``MGProto.update_GMM``);
  every rank then applies the identical gradient / Adam / pi step.
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
    """Concatenate every rank's (top1 [b,K], rows [b,K,D], gt [b]) in rank order."""
    world = dist.get_world_size(group)
    if world == 1:
        return top1, rows, gt
    outs = []
    for t in (top1, rows, gt):
        t = t.contiguous()
        buf = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t, group=group)
        outs.append(buf)
    return tuple(outs)


def attach(model, group=None):
    """Turn on the sharded enqueue / EM for ``model`` (an ``MGProto``) on this process group."""
    model.em_group = group if group is not None else dist.group.WORLD
    return model
