"""Per-class FIFO feature bank (ref utils/memory.py:6-151), stored as ONE ring tensor.

Reference layout: C registered buffers ``cls%d [cap, D]`` kept physically oldest->newest by
shifting on every push, ``mem_len [C]`` int64.  Here: ``bank [C, cap, D]`` + ``head [C]`` +
``mem_len [C]`` (+ ``updated [C]``, the reference's ``memory_updated_cls``), written by one
planner + one scatter launch per iteration (``mgp_bank_enqueue``) and read in place by the EM
kernels -- the reference's ``pull_all`` clone of the whole bank disappears.

Wire format: ``state_dict()`` emits exactly the reference's keys (``cls0..cls{C-1}``,
``mem_len``) with rows in oldest->newest order, and ``load_state_dict`` accepts them, so
checkpoints are interchangeable with the reference.
"""
from __future__ import annotations

import re

import torch
import torch.nn as nn

from . import ops

_CLS_RE = re.compile(r"^cls(\d+)$")


class MemoryBank(nn.Module):
    def __init__(self, num_classes, dim_feature, capacity=1024, mode="all", fix_length_mult=4):
        super().__init__()
        assert capacity % num_classes == 0, (capacity, num_classes)        # ref memory.py:16
        if mode != "all":
            raise NotImplementedError("only mode='all' (the one the reference model uses, model.py:165)")
        self.num_classes = num_classes
        self.dim_feature = dim_feature
        self.capacity = capacity
        self.cap_cls = capacity // num_classes
        self.mode = mode
        self.pull = self.pull_all
        self.register_buffer("bank", torch.zeros(num_classes, self.cap_cls, dim_feature), persistent=False)
        self.register_buffer("head", torch.zeros(num_classes, dtype=torch.int32), persistent=False)
        self.register_buffer("updated", torch.zeros(num_classes, dtype=torch.uint8), persistent=False)
        self.register_buffer("mem_len", torch.zeros(num_classes, dtype=torch.int64))
        # tensor-core operand copy of the bank (fp16 hi / lo of 256 * row, |row|^2), built on first use by update_GMM
        # and kept in step by the enqueue scatter; plain attributes (not buffers): derived data, never checkpointed
        self._shadow = None
        self._shadow_key = None
        # a bank enqueue still running on a side stream (multi-GPU overlap): (event, keep-alive tensors)
        self._pending = None

    # -- cross-stream safety: every reader / writer of the bank tensors on another stream waits for a pending enqueue
    def set_pending(self, event, keepalive=None):
        self._pending = (event, keepalive)

    def wait_pending(self):
        pend = self.__dict__.get("_pending")
        if pend is not None:
            self._pending = None             # (cleared first: fetching self.bank below re-enters __getattr__)
            torch.cuda.current_stream(self._buffers["bank"].device).wait_event(pend[0])

    # -- shadow ---------------------------------------------------------------------------------------------
    def _bank_key(self):
        b = self._buffers["bank"]
        return (b.data_ptr(), b._version, str(b.device))

    def shadow_if_valid(self):
        """(shadow_h, shadow_l, shadow_xx) if the shadow exists and the fp32 bank has not been written by anything but
        the enqueue kernel since it was built, else None."""
        if self._shadow is not None and self._shadow_key == self._bank_key():
            return self._shadow
        return None

    def ensure_shadow(self):
        """Build (or rebuild after an external write to ``bank``: torch bumps its version counter) the shadow."""
        sh = self.shadow_if_valid()
        if sh is None:
            b = self._buffers["bank"]
            if self._shadow is None or self._shadow[0].device != b.device:
                self._shadow = (torch.empty(b.shape, dtype=torch.float16, device=b.device),
                                torch.empty(b.shape, dtype=torch.float16, device=b.device),
                                torch.empty(b.shape[:2], dtype=torch.float32, device=b.device))
            ops.bank_shadow_sync(b, *self._shadow)
            self._shadow_key = self._bank_key()
            sh = self._shadow
        return sh

    # -- reference-compatible views ---------------------------------------------------------
    def linear(self) -> torch.Tensor:
        """[C, cap, D] copy in the reference's oldest->newest row order."""
        if self.bank.is_cuda:
            return ops.bank_linearize(self.bank, self.mem_len, self.head)
        # construction / checkpoint handling on CPU (no compute): plain index arithmetic
        idx = (self.head.long()[:, None] + torch.arange(self.cap_cls)[None, :]) % self.cap_cls
        lin = torch.gather(self.bank, 1, idx[:, :, None].expand(-1, -1, self.dim_feature))
        mask = torch.arange(self.cap_cls)[None, :] < self.mem_len[:, None]
        return lin * mask[:, :, None]

    def __getattr__(self, name):
        if name in ("bank", "mem_len", "head", "updated") and self.__dict__.get("_pending") is not None:
            self.wait_pending()              # a side-stream enqueue may still be writing these (model.py head())
        m = _CLS_RE.match(name)
        if m is not None and "_buffers" in self.__dict__ and "bank" in self._buffers:
            return self.linear()[int(m.group(1))]
        return super().__getattr__(name)

    @torch.no_grad()
    def pull_all(self):
        """ref memory.py:135-151: (features [sum len, D], labels [sum len]) class-ascending."""
        if int(self.mem_len.sum()) == 0:
            return None, None
        lin = self.linear()
        mask = torch.arange(self.cap_cls, device=lin.device)[None, :] < self.mem_len[:, None]
        labels = torch.arange(self.num_classes, device=lin.device)[:, None].expand(-1, self.cap_cls)[mask]
        return lin[mask], labels

    @torch.no_grad()
    def push(self, feature, label):
        """ref memory.py:31-73 for explicit rows (API parity; the model's forward uses the fused
        enqueue kernel instead).  Rows of one class keep their order."""
        assert feature.dim() == 2 and label.dim() == 1 and feature.size(0) == label.size(0)
        cap = self.cap_cls
        for c in torch.unique(label).tolist():
            feat = feature[label == c][:cap]
            m = feat.size(0)
            ln, hd = int(self.mem_len[c]), int(self.head[c])
            slots = (hd + ln + torch.arange(m, device=feat.device)) % cap
            self.bank[c, slots] = feat.to(self.bank.dtype)
            if ln + m > cap:
                self.head[c] = (hd + ln + m - cap) % cap
                self.mem_len[c] = cap
            else:
                self.mem_len[c] = ln + m
            self.updated[c] = 1

    # -- wire format --------------------------------------------------------------------------
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)        # mem_len
        lin = self.linear().detach()
        for i in range(self.num_classes):
            destination[prefix + "cls%d" % i] = lin[i].clone()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        found = False
        with torch.no_grad():
            for i in range(self.num_classes):
                key = prefix + "cls%d" % i
                if key in state_dict:
                    found = True
                    t = state_dict.pop(key)
                    if tuple(t.shape) != (self.cap_cls, self.dim_feature):
                        error_msgs.append("size mismatch for %s: %s vs %s" % (key, tuple(t.shape),
                                                                              (self.cap_cls, self.dim_feature)))
                        continue
                    self.bank[i].copy_(t)
                elif strict:
                    missing_keys.append(key)
            if found:
                self.head.zero_()
                self.updated.zero_()
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
