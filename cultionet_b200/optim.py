"""Flat-buffer AdamW + global-norm clipping + OneCycle schedule on the sm_100a optimiser kernels.

Mirrors ``LightningModuleMixin.configure_optimizers`` (``src/cultionet/models/lightning.py:611-683``: AdamW, betas (0.9, 0.98),
eps 1e-4, weight decay 1e-3, OneCycleLR stepped per batch) and ``Trainer(gradient_clip_val=1.0)`` (``model.py:168-186``).
All parameters live in ONE fp32 buffer and all gradients in another, so a step is two launches (|g|^2, AdamW) and the
data-parallel gradient exchange is a bucketed all-reduce over contiguous slices of the gradient buffer.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr


def one_cycle_lr(step: int, total_steps: int, max_lr: float, pct_start: float = 0.3, div_factor: float = 25.0,
                 final_div_factor: float = 1e4) -> float:
    """torch.optim.lr_scheduler.OneCycleLR (cosine annealing, two phases) evaluated at ``step`` (0-based)."""
    initial = max_lr / div_factor
    minimum = initial / final_div_factor
    up_end = float(pct_start * total_steps) - 1.0
    down_end = float(total_steps) - 1.0

    def cos(a, b, pct):
        return b + (a - b) / 2.0 * (math.cos(math.pi * pct) + 1.0)

    if step <= up_end or down_end <= up_end:
        return cos(initial, max_lr, step / max(up_end, 1e-12)) if up_end > 0 else max_lr
    return cos(max_lr, minimum, min(1.0, (step - up_end) / (down_end - up_end)))


def make_lr_schedule(name: str, lr: float, total_steps: Optional[int] = None, steps_per_epoch: Optional[int] = None,
                     steplr_step_size: int = 5):
    """The schedulers ``configure_optimizers`` offers (``models/lightning.py:650-672``) as closed forms ``f(step, epoch) -> lr``:
    ``OneCycleLR(max_lr=lr)`` stepped per batch; ``CosineAnnealingLR(T_max=20, eta_min=1e-5)``, ``ExponentialLR(gamma=0.5)`` and
    ``StepLR(step_size, gamma=0.5)`` stepped per epoch (Lightning ``interval='epoch'``).  ``epoch`` = completed epochs; without
    ``steps_per_epoch`` it is taken from ``FlatAdamW.epoch`` (``set_epoch``)."""

    def epoch_of(step: int, epoch: Optional[int]) -> int:
        if epoch is not None:
            return epoch
        return step // steps_per_epoch if steps_per_epoch else 0

    if name == "OneCycleLR":
        if not total_steps:
            return lambda step, epoch=None: lr
        return lambda step, epoch=None: one_cycle_lr(min(step, total_steps - 1), total_steps, lr)
    if name == "CosineAnnealingLR":
        t_max, eta_min = 20, 1e-5
        # the closed form of torch's CosineAnnealingLR (its recursive form follows the same curve, periodic with period 2*T_max)
        return lambda step, epoch=None: eta_min + (lr - eta_min) * (1.0 + math.cos(math.pi * epoch_of(step, epoch) / t_max)) / 2.0
    if name == "ExponentialLR":
        return lambda step, epoch=None: lr * 0.5 ** epoch_of(step, epoch)
    if name == "StepLR":
        return lambda step, epoch=None: lr * 0.5 ** (epoch_of(step, epoch) // max(1, int(steplr_step_size)))
    raise NameError("The learning rate scheduler is not implemented in Cultionet.")


class FlatAdamW:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 0.01, betas=(0.9, 0.98), eps: float = 1e-4,
                 weight_decay: float = 1e-3, clip_norm: float = 1.0, total_steps: Optional[int] = None, lr_schedule=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatAdamW: no trainable parameters")
        dev = self.params[0].device
        _lib.check_device(*self.params)
        n = sum(p.numel() for p in self.params)
        self.numel = n
        self.flat_param = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.offsets = []
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_param[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_param[off:off + k].view(p.shape)
                p.grad = self.flat_grad[off:off + k].view(p.shape)
                self.offsets.append((off, k))
                off += k
        self.lr, self.betas, self.eps, self.weight_decay, self.clip_norm = lr, betas, eps, weight_decay, clip_norm
        self.total_steps = total_steps
        self.lr_schedule = lr_schedule  # f(step, epoch) -> lr; None: OneCycle over total_steps (constant lr without total_steps)
        self.epoch: Optional[int] = None  # set by the training loop for the per-epoch schedulers
        self.step_count = 0
        self.hyper = torch.zeros(2, dtype=torch.float32, device=dev)
        # (lr, step) travel through a small RING of pinned host buffers: the host runs ahead of the device (and of a replaying CUDA
        # graph), so a single staging buffer could be overwritten before its copy has executed
        self._ring_n = 16
        self._ring = [torch.zeros(2, dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.zeros(2) for _ in range(self._ring_n)]
        self._ring_events: list = [None] * self._ring_n
        self._ring_i = 0
        self.norm_ws = torch.zeros(1, dtype=torch.float32, device=dev)
        self.grad_scale = 1.0

    def zero_grad(self) -> None:
        """Gradients are accumulated in place into the flat buffer by autograd; re-attach views torch may have dropped."""
        self.flat_grad.zero_()
        for p, (off, k) in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                p.grad = self.flat_grad[off:off + k].view(p.shape)

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def current_lr(self) -> float:
        if self.lr_schedule is not None:
            return float(self.lr_schedule(self.step_count, self.epoch))
        if self.total_steps:
            return one_cycle_lr(min(self.step_count, self.total_steps - 1), self.total_steps, self.lr)
        return self.lr

    def advance_host(self) -> None:
        """Host half of a step: next (lr, step count) staged in pinned memory and copied to the device on the current stream."""
        lr = self.current_lr()
        self.step_count += 1
        slot = self._ring_i % self._ring_n
        self._ring_i += 1
        ev = self._ring_events[slot]
        if ev is not None:
            ev.synchronize()  # the copy that last used this slot has executed (normally long ago)
        buf = self._ring[slot]
        buf[0] = lr
        buf[1] = float(self.step_count)
        self.hyper.copy_(buf, non_blocking=True)
        if self.hyper.is_cuda:
            ev = torch.cuda.Event()
            ev.record()
            self._ring_events[slot] = ev

    def launch_device(self) -> None:
        """Device half of a step (capturable in a CUDA graph): |g|^2 for the clip, then AdamW reading (lr, step) from ``hyper``."""
        st = stream_ptr(self.flat_param)
        if self.clip_norm and self.clip_norm > 0:
            call("cnb_grad_sqnorm", ptr(self.flat_grad), self.numel, ptr(self.norm_ws), st)
        call("cnb_adamw_step", ptr(self.flat_param), ptr(self.flat_grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.numel,
             ptr(self.hyper), self.betas[0], self.betas[1], self.eps, self.weight_decay, self.grad_scale,
             float(self.clip_norm or 0.0), ptr(self.norm_ws), st)
        # the kernel wrote the parameters through raw pointers (no torch version bump): drop their packed bf16 copies
        from . import functional as F

        F.invalidate_packed_weights()

    def step(self) -> None:
        self.advance_host()
        self.launch_device()

    def state_dict(self) -> dict:
        """``torch.optim.AdamW.state_dict()`` layout -- ``{"state": {i: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [...]}``
        with the parameters in ``cultionet_model.parameters()`` order, the order of the reference's ``params_list``
        (``models/lightning.py:614``) -- so the moments of a checkpoint written here load into the reference's optimizer and back."""
        state = {}
        for i, (p, (off, k)) in enumerate(zip(self.params, self.offsets)):
            state[i] = {"step": torch.tensor(float(self.step_count)),
                        "exp_avg": self.exp_avg[off:off + k].view(p.shape).detach().clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + k].view(p.shape).detach().clone()}
        group = {"lr": self.current_lr(), "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": True, "initial_lr": self.lr / 25.0, "max_lr": self.lr, "min_lr": self.lr / 25.0 / 1e4,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: dict) -> None:
        """Accepts the torch layout above (a reference ``last.ckpt``'s ``optimizer_states[0]`` included) and the flat
        ``{"step", "exp_avg", "exp_avg_sq"}`` layout this class wrote in round 1."""
        if "state" in sd and "param_groups" in sd:
            order = [i for g in sd["param_groups"] for i in g["params"]]
            if len(order) != len(self.params):
                raise ValueError(f"optimizer state holds {len(order)} parameters, the model has {len(self.params)}")
            step = 0
            with torch.no_grad():
                for idx, p, (off, k) in zip(order, self.params, self.offsets):
                    st = sd["state"].get(idx)
                    if st is None:  # a parameter that never received a gradient has no state in torch
                        self.exp_avg[off:off + k].zero_()
                        self.exp_avg_sq[off:off + k].zero_()
                        continue
                    if tuple(st["exp_avg"].shape) != tuple(p.shape):
                        raise ValueError(f"optimizer state {idx}: shape {tuple(st['exp_avg'].shape)} vs parameter {tuple(p.shape)}")
                    self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1))
                    self.exp_avg_sq[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                    step = max(step, int(float(st["step"])))
            self.step_count = step
            return
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
