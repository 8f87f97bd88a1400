"""Data-parallel gradient exchange: one process per GPU, NCCL over NVLink via ``torch.distributed``.

The reference trains with Lightning ``strategy="ddp"`` (``src/cultionet/model.py:101``, ``:184``): per-rank BatchNorm statistics
(no SyncBN), gradient mean over ranks.  Here gradients already live in one flat fp32 buffer (``optim.FlatAdamW``), so the exchange
is an all-reduce over contiguous buckets of that buffer, issued on NCCL's stream as soon as backward has produced every gradient
of a bucket (parameters are laid out in forward order, so buckets complete from the back) and overlapping the rest of backward.
"""
from __future__ import annotations

import datetime
import os
from typing import List, Optional

import torch
import torch.distributed as dist


def init_distributed(backend: Optional[str] = None) -> tuple:
    """(rank, local_rank, world_size); initialises the default process group from the torchrun environment if needed."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        # a short timeout: a mismatched collective must fail in minutes, not hold a multi-GPU box for the 10-minute default
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                timeout=datetime.timedelta(seconds=int(os.environ.get("CNB_DIST_TIMEOUT_S", "180"))))
    return rank, local, world


class BucketedGradSync:
    """Overlapped mean all-reduce of ``optimizer.flat_grad`` in ~``bucket_mb`` slices aligned to parameter boundaries."""

    def __init__(self, optimizer, bucket_mb: float = 32.0, group=None):
        self.opt = optimizer
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        cap = int(bucket_mb * 1024 * 1024 / 4)
        # buckets over the flat buffer, built back-to-front so that the last layers (first to finish in backward) go first
        self.buckets: List[tuple] = []
        end = optimizer.numel
        cur_start = end
        self.param_bucket = [0] * len(optimizer.params)
        members: List[int] = []
        for i in range(len(optimizer.params) - 1, -1, -1):
            off, k = optimizer.offsets[i]
            cur_start = off
            members.append(i)
            if end - cur_start >= cap or i == 0:
                b = len(self.buckets)
                self.buckets.append((cur_start, end, len(members)))
                for m in members:
                    self.param_bucket[m] = b
                members = []
                end = cur_start
        self.pending = [0] * len(self.buckets)
        self.handles: list = []
        self.enabled = self.world > 1
        self._hooks: list = []  # RemovableHandles of the per-parameter hooks (close() removes them)
        if self.enabled:
            for i, p in enumerate(optimizer.params):
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))
        self.reset()

    def close(self) -> None:
        """Remove the parameter hooks: a later exchange object on the same parameters must be the only one issuing all-reduces."""
        for h in self._hooks:
            h.remove()
        self._hooks = []
        self.enabled = False
        self.handles = []

    def reset(self) -> None:
        self.pending = [n for (_, _, n) in self.buckets]
        self.handles = []

    def _make_hook(self, index: int):
        def hook(param):
            b = self.param_bucket[index]
            self.pending[b] -= 1
            if self.pending[b] == 0:
                self._launch(b)
        return hook

    def _launch(self, b: int) -> None:
        s, e, _ = self.buckets[b]
        chunk = self.opt.flat_grad[s:e]
        self.handles.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self) -> None:
        """Call after backward: launches any bucket whose hooks did not all fire (unused parameters) and waits."""
        if not self.enabled:
            return
        for b, n in enumerate(self.pending):
            if n > 0:
                self._launch(b)
        for h in self.handles:
            h.wait()
        self.opt.grad_scale = 1.0 / self.world
        self.reset()
