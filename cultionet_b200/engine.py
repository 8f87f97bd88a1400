"""Minimal training / prediction loops for ``CultionetLitModel`` when Lightning is not installed: what Lightning's ``Trainer``
does around ``training_step`` on the hot path (``src/cultionet/model.py:168-186``: DDP, gradient clipping 1.0, optimizer +
per-step OneCycle schedule), nothing else."""
from __future__ import annotations

from typing import Optional

import torch

from .data import Data
from .parallel import BucketedGradSync


def batch_to_device(batch: Data, device, non_blocking: bool = True) -> Data:
    out = {}
    for k, v in batch.__dict__.items():
        out[k] = v.to(device, non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v
    return Data(**out)


class TrainStep:
    """One data-parallel optimisation step: forward + loss + backward (+ overlapped gradient all-reduce) + AdamW."""

    def __init__(self, lit_model, total_steps: Optional[int] = None, bucket_mb: float = 32.0):
        self.model = lit_model
        self.optimizer = lit_model.configure_optimizers(total_steps=total_steps)
        self.sync = BucketedGradSync(self.optimizer, bucket_mb=bucket_mb)
        self.model.train()

    def __call__(self, batch: Data) -> torch.Tensor:
        self.optimizer.zero_grad()
        loss = self.model.training_step(batch, 0)
        loss.backward()
        self.sync.finish()
        self.optimizer.step()
        return loss.detach()


@torch.no_grad()
def predict(lit_model, batch: Data):
    lit_model.eval()
    return lit_model.predict_step(batch, 0)
