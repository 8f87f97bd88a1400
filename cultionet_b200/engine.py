"""Minimal training / prediction loops for ``CultionetLitModel`` when Lightning is not installed: what Lightning's ``Trainer``
does around ``training_step`` on the hot path (``src/cultionet/model.py:168-186``: DDP, gradient clipping 1.0, optimizer +
per-step OneCycle schedule), nothing else.

``TrainStep(..., cuda_graph=True)`` captures the whole step (zero-grad, forward, loss, backward, gradient all-reduce, AdamW: ~1200
kernel launches at BASELINE config 2) into ONE CUDA graph after a few eager warm-up steps and replays it afterwards.  Measured on
a B200 the eager step needs 64 ms of host time to enqueue 68 ms of device work, so any further kernel speed-up would be hidden behind
Python; a replay costs the host a few hundred microseconds.  Everything the graph needs is already static: parameters, gradients
and optimizer state live in flat buffers, the learning rate and step count are read from a device buffer, the C ABI allocates
nothing and only enqueues work on the stream it is given (``include/cultionet_b200.h``).
"""
from __future__ import annotations

import warnings
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from . import functional as F
from .data import Data
from .nn.modules.convolution import deferred_batch_counters
from .parallel import BucketedGradSync


def batch_to_device(batch: Data, device, non_blocking: bool = True) -> Data:
    out = {}
    for k, v in batch.__dict__.items():
        out[k] = v.to(device, non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v
    return Data(**out)


class DevicePrefetcher:
    """Iterates over HOST batches (pinned ``Data`` objects) and yields device copies, issuing each batch's host-to-device copies on
    a side stream ``depth - 1`` steps ahead so that they overlap the previous step's kernels (what ``DataLoader(pin_memory=True)``
    + non-blocking copies do in the reference's Lightning loop).  The device buffers are reused round-robin; a buffer is refilled
    only after the work the consumer enqueued while holding it has finished (event on the consumer's stream)."""

    def __init__(self, batches, device, depth: int = 2):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.depth = max(1, depth)
        self.stream = torch.cuda.Stream(self.device)
        self.slots = [None] * self.depth
        self.free = [None] * self.depth
        self.ready: list = []

    def _issue(self, slot: int) -> bool:
        try:
            host = next(self.it)
        except StopIteration:
            return False
        with torch.cuda.stream(self.stream):
            if self.free[slot] is not None:
                self.stream.wait_event(self.free[slot])
            buf = self.slots[slot]
            if buf is None or _signature(buf) != _signature(host):
                buf = Data(**{k: (torch.empty(v.shape, dtype=v.dtype, device=self.device) if isinstance(v, torch.Tensor) else v)
                              for k, v in host.__dict__.items()})
                self.slots[slot] = buf
            for k, v in host.__dict__.items():
                if isinstance(v, torch.Tensor):
                    getattr(buf, k).copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.ready.append((slot, ev))
        return True

    def __iter__(self):
        for slot in range(self.depth):
            if not self._issue(slot):
                break
        while self.ready:
            slot, ev = self.ready.pop(0)
            torch.cuda.current_stream(self.device).wait_event(ev)
            yield self.slots[slot]
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))
            self.free[slot] = done
            self._issue(slot)


class LossReader:
    """Reads every step's (scalar) result back to the host WITHOUT stalling the device: ``push(t)`` enqueues a device->host copy of
    ``t`` into a pinned slot on the current stream and returns the value pushed ``depth - 1`` calls earlier (``None`` until then), so the
    host is already enqueuing step i + 1 while step i runs; ``drain()`` returns the values still in flight.  For loops that want every
    step's loss on the host without a synchronisation per step (``model.fit`` itself accumulates the epoch loss on the device).  Measured
    with ``bench.py``'s end-to-end loop it changes nothing at config 2 (532 vs 536 chips/s: the 2-3 % between the end-to-end and the
    resident number is not the read-back), so the bench keeps its plain synchronous read."""

    def __init__(self, depth: int = 2):
        self.depth = max(1, depth)
        self.slots = [torch.zeros(1, dtype=torch.float32).pin_memory() if torch.cuda.is_available() else torch.zeros(1) for _ in range(self.depth)]
        self.pending: list = []  # (slot index, event)
        self._next = 0

    def _pop(self) -> float:
        slot, ev = self.pending.pop(0)
        if ev is not None:
            ev.synchronize()
        return float(self.slots[slot][0])

    def push(self, t: torch.Tensor):
        out = self._pop() if len(self.pending) >= self.depth else None
        slot = self._next
        self._next = (self._next + 1) % self.depth
        src = t.detach().reshape(-1)[:1].float()
        if src.is_cuda:
            self.slots[slot].copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(src.device))
        else:
            self.slots[slot].copy_(src)
            ev = None
        self.pending.append((slot, ev))
        if out is None and len(self.pending) > self.depth - 1 and self.depth == 1:
            out = self._pop()
        return out

    def drain(self) -> list:
        vals = []
        while self.pending:
            vals.append(self._pop())
        return vals


def _signature(batch: Data) -> tuple:
    return tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(batch.__dict__.items()) if isinstance(v, torch.Tensor))


class TrainStep:
    """One data-parallel optimisation step: forward + loss + backward (+ gradient all-reduce) + AdamW.

    Eager mode overlaps the bucketed all-reduce with backward (``parallel.BucketedGradSync``).  Graph mode replays a captured
    step; its gradient exchange is a single all-reduce of the flat gradient buffer after backward on the capture stream (169 MB
    over NVLink: ~0.4 ms, nothing worth overlapping against the host time the graph removes)."""

    def __init__(self, lit_model, total_steps: Optional[int] = None, bucket_mb: float = 32.0, cuda_graph: bool = False,
                 graph_warmup: int = 3, steps_per_epoch: Optional[int] = None):
        self.model = lit_model
        self.optimizer = lit_model.configure_optimizers(total_steps=total_steps, steps_per_epoch=steps_per_epoch)
        self.cuda_graph = bool(cuda_graph) and self.optimizer.flat_param.is_cuda and not _lib.is_emulator()
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        # graph mode exchanges gradients itself (see _device_step); eager mode uses the overlapped bucketed exchange
        self.sync = None if self.cuda_graph else BucketedGradSync(self.optimizer, bucket_mb=bucket_mb)
        self.graph_warmup = graph_warmup
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._static: Optional[Data] = None
        self._static_loss: Optional[torch.Tensor] = None
        self._sig = None
        self._calls = 0
        self.launches_per_step: Optional[int] = None  # C-ABI kernel launches recorded while capturing (replays repeat them)
        self.model.train()

    # ---- the device work of one step (what a graph captures) ----
    def _device_step(self, batch: Data) -> torch.Tensor:
        self.optimizer.zero_grad()
        F.prepack_parameters()  # every convolution weight seen so far: one packing launch instead of one per layer
        with deferred_batch_counters():  # one multi-tensor launch for the 52 BatchNorm step counters
            loss = self.model.training_step(batch, 0)
        # weight / bias gradients go straight into the flat gradient buffer unless per-parameter hooks must see them (overlapped DDP)
        # (zero_grad above has just cleared the flat gradient buffer: every gradient kernel accumulates, none clears its target first)
        with F.direct_param_grads(self.sync is None or not self.sync.enabled, zeroed=True):
            loss.backward()
        if self.sync is not None:
            self.sync.finish()
        elif self.world > 1:
            dist.all_reduce(self.optimizer.flat_grad, op=dist.ReduceOp.SUM)
            self.optimizer.grad_scale = 1.0 / self.world
        self.optimizer.launch_device()
        return loss.detach()

    def eager(self, batch: Data) -> torch.Tensor:
        self.optimizer.advance_host()
        return self._device_step(batch)

    def close(self) -> None:
        """Detach from the model: removes the gradient-exchange hooks (a later ``TrainStep`` / ``fit`` on the same model registers its
        own) and drops the captured graph."""
        if self.sync is not None:
            self.sync.close()
            self.sync = None
        self._graph = None

    def _capture(self, batch: Data) -> None:
        dev = self.optimizer.flat_param.device
        self._static = Data(**{k: (v.to(dev).clone() if isinstance(v, torch.Tensor) else v) for k, v in batch.__dict__.items()})
        self._sig = _signature(batch)
        if self.world > 1:
            self.optimizer.grad_scale = 1.0 / self.world  # a kernel argument: must hold its final value while capturing
        F.invalidate_packed_weights()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count()
        with torch.cuda.graph(graph):
            self._static_loss = self._device_step(self._static)
        self.launches_per_step = _lib.launch_count() - l0
        self._graph = graph
        F.invalidate_packed_weights()

    def __call__(self, batch: Data) -> torch.Tensor:
        if not self.cuda_graph or _lib.TIMER is not None:
            return self.eager(batch)
        if self._graph is None:
            if self._calls < self.graph_warmup:  # lazy initialisation (kernel attributes, NCCL communicators) happens eagerly
                self._calls += 1
                return self.eager(batch)
            try:
                self._capture(batch)
            except Exception as e:  # noqa: BLE001 - a step that cannot be captured still trains, eagerly
                warnings.warn(f"cultionet_b200: CUDA graph capture of the training step failed ({e!r}); running eagerly")
                self.cuda_graph = False
                self._graph = None
                torch.cuda.synchronize()
                if self.sync is not None:
                    self.sync.close()
                self.sync = BucketedGradSync(self.optimizer)
                return self.eager(batch)
        if _signature(batch) != self._sig:
            return self.eager(batch)  # a ragged last batch: same arithmetic, no graph
        for k, v in batch.__dict__.items():
            if isinstance(v, torch.Tensor):
                getattr(self._static, k).copy_(v, non_blocking=True)
        self.optimizer.advance_host()
        self._graph.replay()
        F.invalidate_packed_weights()  # the packed weights inside the graph's pool describe the parameters BEFORE this step
        return self._static_loss


@torch.no_grad()
def predict(lit_model, batch: Data):
    lit_model.eval()
    return lit_model.predict_step(batch, 0)


class PredictStep:
    """``predict_step`` over fixed-shape window batches (the sliding-window inference of ``cultionet predict``): eval-mode BatchNorm,
    no autograd, packed weights cached across calls; with ``cuda_graph=True`` the forward is captured once and replayed, the
    returned dict holds STATIC output tensors that the next call overwrites (copy or consume them first)."""

    def __init__(self, lit_model, cuda_graph: bool = False, graph_warmup: int = 2):
        self.model = lit_model
        self.model.eval()
        p = next(lit_model.parameters())
        self.cuda_graph = bool(cuda_graph) and p.is_cuda and not _lib.is_emulator()
        self.graph_warmup = graph_warmup
        self._graph = None
        self._static = None
        self._out = None
        self._sig = None
        self._calls = 0
        self.launches_per_step: Optional[int] = None

    @torch.no_grad()
    def eager(self, batch: Data):
        return self.model.predict_step(batch, 0)

    @torch.no_grad()
    def __call__(self, batch: Data):
        if not self.cuda_graph or _lib.TIMER is not None:
            return self.eager(batch)
        if self._graph is None:
            if self._calls < self.graph_warmup:
                self._calls += 1
                return self.eager(batch)
            try:
                self._static = Data(**{k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in batch.__dict__.items()})
                self._sig = _signature(batch)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                l0 = _lib.launch_count()
                with torch.cuda.graph(graph):
                    self._out = self.model.predict_step(self._static, 0)
                self.launches_per_step = _lib.launch_count() - l0
                self._graph = graph
            except Exception as e:  # noqa: BLE001
                warnings.warn(f"cultionet_b200: CUDA graph capture of predict_step failed ({e!r}); running eagerly")
                self.cuda_graph = False
                self._graph = None
                torch.cuda.synchronize()
                return self.eager(batch)
        if _signature(batch) != self._sig:
            return self.eager(batch)
        for k, v in batch.__dict__.items():
            if isinstance(v, torch.Tensor):
                getattr(self._static, k).copy_(v, non_blocking=True)
        self._graph.replay()
        return self._out
