"""Tanimoto-with-complement loss on the sm_100a reduction kernels.

API of ``src/cultionet/losses/losses.py:103-218`` (``TanimotoComplementLoss(smooth, depth, transform_logits,
one_hot_targets)(inputs, targets, mask=None, dim=None)``) with ``LossPreprocessing`` (``:9-59``) folded into the kernel:
one-hot targets, the ``[B,H,W] -> [B,1,H,W]`` unsqueeze and the mask product are evaluated on the fly, never materialised.
"""
from __future__ import annotations

import typing as T

import torch
import torch.nn as nn

from . import functional as F


class TanimotoComplementLoss(nn.Module):
    variant = F.TANIMOTO_COMPLEMENT

    def __init__(self, smooth: float = 1e-5, depth: int = 5, transform_logits: bool = False, one_hot_targets: bool = True):
        super().__init__()
        if transform_logits:
            raise NotImplementedError(
                "cultionet_b200: transform_logits=True is not built (the reference's LOSS_DICT only uses False, lightning.py:48-54)"
            )
        self.smooth = smooth
        self.depth = depth
        self.one_hot_targets = one_hot_targets

    def _spec(self, inputs: torch.Tensor, targets: torch.Tensor, mask: T.Optional[torch.Tensor], weight: float = 1.0):
        if inputs.dim() != 4:
            raise ValueError("inputs must be shaped (B, C, H, W)")
        if self.one_hot_targets and inputs.shape[1] > 1:
            if targets.dim() != 3:
                raise ValueError("one-hot targets must be integer labels shaped (B, H, W)")
            tmode = F.TARGET_ONEHOT
        else:
            if targets.dim() == 4 and targets.shape[1] not in (1, inputs.shape[1]):
                raise ValueError("targets must have 1 or C channels")
            tmode = F.TARGET_FLOAT
        mmode = F.MASK_NONE
        if mask is not None:
            if mask.dim() == 4:
                if mask.shape[1] != 1:
                    raise ValueError("mask must be shaped (B, H, W) or (B, 1, H, W)")
                mask = mask[:, 0]
            mmode = F.MASK_FLOAT
        return F.TanimotoTermSpec(targets, tmode, mask, mmode, weight=weight)

    def forward(self, inputs: torch.Tensor, targets: torch.Tensor, mask: T.Optional[torch.Tensor] = None,
                dim: T.Optional[T.Tuple[int, ...]] = None) -> torch.Tensor:
        if dim is not None and tuple(dim) != (1, 2, 3):
            raise NotImplementedError("cultionet_b200: only the default reduction dim=(1, 2, 3) is built")
        total, _ = F.tanimoto_complement([inputs], [self._spec(inputs, targets, mask)], smooth=self.smooth, depth=self.depth,
                                         variant=self.variant)
        return total


class TanimotoDistLoss(TanimotoComplementLoss):
    """``src/cultionet/losses/losses.py:221-340``: ``T = (P + eps) / (S - P + eps)`` on the pair and on its complement -- the depth-1
    case of the complement form, so the same reduction kernel serves it (``variant`` 1)."""

    variant = F.TANIMOTO_DIST

    def __init__(self, smooth: float = 1e-5, transform_logits: bool = False, one_hot_targets: bool = True):
        super().__init__(smooth=smooth, depth=1, transform_logits=transform_logits, one_hot_targets=one_hot_targets)

    def forward(self, inputs: torch.Tensor, targets: torch.Tensor, mask: T.Optional[torch.Tensor] = None) -> torch.Tensor:
        return super().forward(inputs, targets, mask)


class CombinedLoss(nn.Module):
    """``losses.py:62-100``: the mean of its member losses over the same (inputs, targets, mask).  The pair the reference builds
    (``LOSS_DICT[TANIMOTO_COMBINED]``, ``models/lightning.py:64-82``: TanimotoDistLoss + TanimotoComplementLoss) runs as ONE reduction
    (``variant`` 2); any other combination evaluates its members one after the other."""

    def __init__(self, losses: T.List[T.Callable]):
        super().__init__()
        self.losses = list(losses)

    def _fused(self):
        kinds = sorted(type(m).__name__ for m in self.losses)
        if kinds == ["TanimotoComplementLoss", "TanimotoDistLoss"]:
            a, b = self.losses
            if a.smooth == b.smooth and a.one_hot_targets == b.one_hot_targets:
                return a if type(a) is TanimotoComplementLoss else b
        return None

    def forward(self, inputs: torch.Tensor, targets: torch.Tensor, mask: T.Optional[torch.Tensor] = None) -> torch.Tensor:
        comp = self._fused()
        if comp is not None:
            total, _ = F.tanimoto_complement([inputs], [comp._spec(inputs, targets, mask)], smooth=comp.smooth, depth=comp.depth,
                                             variant=F.TANIMOTO_COMBINED)
            return total
        loss = 0.0
        for fn in self.losses:
            loss = loss + fn(inputs=inputs, targets=targets, mask=mask)
        return loss / len(self.losses)


def tower_unet_loss(predictions: T.Dict[str, torch.Tensor], y: torch.Tensor, bdist: torch.Tensor, edge_class: int = 2,
                    smooth: float = 1e-5, depth: int = 5, variant: int = F.TANIMOTO_COMPLEMENT):
    """The three-term training loss of ``LightningModuleMixin.calc_loss`` (``models/lightning.py:209-354``) in one launch:
    (distance vs bdist, edge vs (y == edge_class), crop vs (0 < y < edge_class)) / 3, with the weak-supervision mask
    ``y != -1`` applied on the fly (the reference only builds it when ``y.min() == -1``, a host sync; all-ones otherwise).

    Returns (loss, tensor[4] = total, dloss, eloss, closs).
    """
    from .enums import InferenceNames

    w = 1.0 / 3.0
    specs = [
        F.TanimotoTermSpec(bdist, F.TARGET_FLOAT, y, F.MASK_FROM_LABELS, edge_class, w),
        F.TanimotoTermSpec(y, F.TARGET_EDGE, y, F.MASK_FROM_LABELS, edge_class, w),
        F.TanimotoTermSpec(y, F.TARGET_CROP, y, F.MASK_FROM_LABELS, edge_class, w),
    ]
    preds = [predictions[InferenceNames.DISTANCE], predictions[InferenceNames.EDGE], predictions[InferenceNames.CROP]]
    return F.tanimoto_complement(preds, specs, smooth=smooth, depth=depth, variant=variant)
