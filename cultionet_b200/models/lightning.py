"""``CultionetLitModel`` -- the reference's LightningModule surface (``src/cultionet/models/lightning.py:91-372``, ``:821-898``)
for the TowerUNet hot path: ``forward`` / ``predict_step`` / ``training_step`` / ``validation_step`` / ``calc_loss`` /
``get_true_labels`` / ``configure_optimizers`` with the same constructor keywords and the same attribute layout
(``cultionet_model`` stored under ``f"{model_name}_{model_type}"`` so checkpoints keep their key prefix).

The class is a plain ``nn.Module`` carrying the LightningModule METHOD surface; ``cultionet_b200.model.fit`` /
``cultionet_b200.engine.TrainStep`` are the loops that drive it (what ``lightning.Trainer`` does for this path in the reference:
DDP, clipping, optimizer + scheduler stepping, best-``val_score`` checkpoint).  It is deliberately NOT a ``lightning.LightningModule``
subclass: the optimiser is a flat-buffer kernel pair that owns parameter and gradient storage and clips inside its step, which a
Lightning ``Trainer`` (torch ``Optimizer`` protocol, its own ``gradient_clip_val``) cannot drive.
"""
from __future__ import annotations

import typing as T

import torch
import torch.nn as nn

from ..data import Data
from ..enums import AttentionTypes, InferenceNames, LearningRateSchedulers, LossTypes, ModelTypes, ResBlockTypes, ValidationNames
from .. import functional as F
from ..losses import tower_unet_loss
from ..optim import FlatAdamW, make_lr_schedule
from .cultionet import CultioNet

HAVE_LIGHTNING = False  # see the module docstring

# LOSS_DICT of the reference (models/lightning.py:38-88), the Tanimoto family: one reduction kernel, three closed forms
LOSS_VARIANTS = {
    str(LossTypes.TANIMOTO_COMPLEMENT): F.TANIMOTO_COMPLEMENT,
    str(LossTypes.TANIMOTO): F.TANIMOTO_DIST,
    str(LossTypes.TANIMOTO_COMBINED): F.TANIMOTO_COMBINED,
}


def _plain(v):
    """Checkpoint-safe hyper-parameter value: enum members become their string value, so that neither side needs the other's enum
    classes importable to unpickle a checkpoint."""
    import enum

    if isinstance(v, enum.Enum):
        return str(v.value)
    if isinstance(v, (list, tuple)):
        return [_plain(i) for i in v]
    return v


def scores_from_counts(c: torch.Tensor) -> T.Dict[str, torch.Tensor]:
    """The scorers of ``configure_scorer`` (models/lightning.py:562-580) from the fp64[12] counts of ``cnb_val_counts``:
    MeanAbsoluteError / MeanSquaredError; ``FBetaScore(task="multiclass", num_classes=2, beta=2)`` -- whose default ``average="micro"``
    makes it (tp + tn) / n, i.e. accuracy, for any beta; ``MatthewsCorrCoef(task="multiclass", num_classes=2)`` from the 2x2 confusion
    matrix (0 when a marginal is empty, 1 / -1 when every pixel is right / wrong, as torchmetrics special-cases it)."""
    n = c[0].clamp_min(1.0)
    out = {"dist_mae": c[1] / n, "dist_mse": c[2] / n}
    for name, o in (("edge", 3), ("crop", 7)):
        tp, fp, fn, tn = c[o], c[o + 1], c[o + 2], c[o + 3]
        out[f"{name}_f1"] = (tp + tn) / n
        den = (tp + fp) * (tp + fn) * (tn + fp) * (tn + fn)
        mcc = (tp * tn - fp * fn) / den.clamp_min(1e-300).sqrt()
        perfect = (fp + fn == 0).to(mcc.dtype) - (tp + tn == 0).to(mcc.dtype)
        out[f"{name}_mcc"] = torch.where(den > 0, mcc, perfect)
    return {k: v.float() for k, v in out.items()}


class LightningModuleMixin(nn.Module):
    def __call__(self, *args, **kwargs):  # the reference overrides __call__ the same way (lightning.py:95-96)
        return self.forward(*args, **kwargs)

    def forward(self, batch: Data, batch_idx: int = None) -> T.Dict[str, torch.Tensor]:
        return self.cultionet_model(batch)

    @property
    def cultionet_model(self) -> CultioNet:
        return getattr(self, self.model_attr)

    def probas_to_labels(self, x: torch.Tensor, thresh: float = 0.5) -> torch.Tensor:
        if x.shape[1] == 1:
            return x.gt(thresh).squeeze(dim=1).long()
        return x.argmax(dim=1).long()

    def predict_step(self, batch: Data, batch_idx: int = None) -> T.Dict[str, torch.Tensor]:
        return self.forward(batch, batch_idx=batch_idx)

    @torch.no_grad()
    def get_true_labels(self, batch: Data, crop_type: torch.Tensor = None) -> T.Dict[str, T.Optional[torch.Tensor]]:
        """Label recoding of the reference (lightning.py:161-207).  Not on the training hot path here: ``calc_loss`` derives the
        same targets inside the loss kernel; this method serves validation / user code."""
        y, ec = batch.y, self.edge_class
        mask = None
        if y.min() == -1:
            mask = (y != -1).long().unsqueeze(1)
        return {
            ValidationNames.TRUE_EDGE: (y == ec).long(),
            ValidationNames.TRUE_CROP: ((y > 0) & (y < ec)).long(),
            ValidationNames.TRUE_CROP_AND_EDGE: (y > 0).long(),
            ValidationNames.TRUE_CROP_OR_EDGE: torch.where((y > 0) & (y < ec), 1, torch.where(y == ec, 2, 0)).long(),
            ValidationNames.TRUE_CROP_TYPE: None,
            ValidationNames.MASK: mask,
        }

    def calc_loss(self, batch: Data, predictions: T.Dict[str, torch.Tensor]):
        """(distance + edge + crop) / 3 with Tanimoto-complement terms (lightning.py:318-354), one fused reduction."""
        loss, parts = tower_unet_loss(predictions, batch.y, batch.bdist, edge_class=self.edge_class,
                                      variant=LOSS_VARIANTS[str(self.loss_name)])
        return loss, {"dloss": parts[1], "eloss": parts[2], "closs": parts[3]}

    def training_step(self, batch: Data, batch_idx: int = None):
        predictions = self(batch)
        loss, _ = self.calc_loss(batch, predictions)
        return loss

    @torch.no_grad()
    def _shared_eval_step(self, batch: Data, batch_idx: int = None) -> dict:
        """``_shared_eval_step`` (models/lightning.py:374-481): loss + one counting pass (``cnb_val_counts``) over the labelled pixels
        -> MAE / MSE of the distance, micro F-beta and MCC of the edge and crop masks, and
        ``score = loss + (1 - edge_f) + (1 - crop_f) + mae + (1 - max(edge_mcc, 0)) + (1 - max(crop_mcc, 0))``."""
        predictions = self(batch)
        loss, report = self.calc_loss(batch, predictions)
        counts = F.validation_counts(predictions[InferenceNames.DISTANCE], predictions[InferenceNames.EDGE],
                                     predictions[InferenceNames.CROP], batch.y, batch.bdist, edge_class=self.edge_class)
        m = scores_from_counts(counts)
        score = (loss + (1.0 - m["edge_f1"]) + (1.0 - m["crop_f1"]) + m["dist_mae"] + (1.0 - m["edge_mcc"].clamp_min(0))
                 + (1.0 - m["crop_mcc"].clamp_min(0)))
        metrics = {"loss": loss, "score": score, "counts": counts, **m}
        metrics.update(report)
        return metrics

    @torch.no_grad()
    def validation_step(self, batch: Data, batch_idx: int = None) -> dict:
        """Keys of the reference's ``validation_step`` (:483-510) plus the MCC / MSE values it computes but does not log."""
        e = self._shared_eval_step(batch, batch_idx)
        return {"vef1": e["edge_f1"], "vcf1": e["crop_f1"], "vmae": e["dist_mae"], "val_score": e["score"], "val_loss": e["loss"],
                "val_dloss": e["dloss"], "val_eloss": e["eloss"], "val_closs": e["closs"], "vmse": e["dist_mse"],
                "vemcc": e["edge_mcc"], "vcmcc": e["crop_mcc"]}

    @torch.no_grad()
    def test_step(self, batch: Data, batch_idx: int = None) -> dict:
        """``test_step`` (:544-560) for the scorers ``_shared_eval_step`` provides (the reference also reads dice / jaccard keys that its
        own ``_shared_eval_step`` never sets)."""
        e = self._shared_eval_step(batch, batch_idx)
        return {"test_loss": e["loss"], "tmae": e["dist_mae"], "tmse": e["dist_mse"], "tef1": e["edge_f1"], "tcf1": e["crop_f1"],
                "temcc": e["edge_mcc"], "tcmcc": e["crop_mcc"], "test_score": e["score"]}

    def configure_optimizers(self, total_steps: T.Optional[int] = None, steps_per_epoch: T.Optional[int] = None):
        """``configure_optimizers`` (:611-683).  Optimisers: AdamW (betas (0.9, 0.98)) and Adam (no decay) on the flat AdamW kernel; the
        reference's RAdam / SGD are not built (``NameError``, as for any unknown name there).  Schedulers: OneCycleLR stepped per batch,
        CosineAnnealingLR(T_max=20, eta_min=1e-5) / ExponentialLR(0.5) / StepLR(step_size, 0.5) stepped per epoch."""
        if self.optimizer == "AdamW":
            betas, wd = (0.9, 0.98), self.weight_decay
        elif self.optimizer == "Adam":
            betas, wd = (0.9, 0.999), 0.0
        else:
            raise NameError("cultionet_b200 builds the reference's 'AdamW' and 'Adam' optimizers only.")
        schedule = make_lr_schedule(str(self.lr_scheduler), self.learning_rate, total_steps=total_steps,
                                    steps_per_epoch=steps_per_epoch, steplr_step_size=self.steplr_step_size)
        return FlatAdamW(self.cultionet_model.parameters(), lr=self.learning_rate, betas=betas, eps=self.eps, weight_decay=wd,
                         clip_norm=1.0, total_steps=total_steps, lr_schedule=schedule)


class CultionetLitModel(LightningModuleMixin):
    def __init__(
        self,
        in_channels: int,
        in_time: int,
        hidden_channels: int = 64,
        model_type: str = ModelTypes.TOWERUNET,
        dropout: float = 0.2,
        activation_type: str = "SiLU",
        dilations: T.Union[int, T.Sequence[int]] = None,
        res_block_type: str = ResBlockTypes.RESA,
        attention_weights: str = AttentionTypes.NATTEN,
        optimizer: str = "AdamW",
        loss_name: str = LossTypes.TANIMOTO_COMPLEMENT,
        learning_rate: float = 0.01,
        lr_scheduler: str = LearningRateSchedulers.ONE_CYCLE_LR,
        steplr_step_size: int = 5,
        weight_decay: float = 1e-3,
        eps: float = 1e-4,
        ckpt_name: str = "last",
        model_name: str = "cultionet",
        pool_by_max: bool = False,
        batchnorm_first: bool = False,
        class_counts: T.Optional[torch.Tensor] = None,
        edge_class: T.Optional[int] = None,
        scale_pos_weight: bool = False,
        save_batch_val_metrics: bool = False,
        compute_dtype: T.Union[torch.dtype, str] = torch.bfloat16,
    ):
        super().__init__()
        if isinstance(compute_dtype, str):  # a checkpoint stores the name
            compute_dtype = getattr(torch, compute_dtype.replace("torch.", ""))
        # what Lightning's save_hyperparameters() records (lightning.py:850) and writes into checkpoints as 'hyper_parameters' -- as
        # plain values (enum members -> their strings, dtype -> its name): a checkpoint must unpickle without this package
        self.hyper_parameters = {k: _plain(v) for k, v in locals().items() if k not in ("self", "__class__", "compute_dtype")}
        self.hyper_parameters["compute_dtype"] = str(compute_dtype).replace("torch.", "")
        if str(loss_name) not in LOSS_VARIANTS:
            raise NotImplementedError(f"cultionet_b200 builds the Tanimoto family of the reference's LOSS_DICT ({sorted(LOSS_VARIANTS)}); "
                                      f"got {loss_name!r}")
        self.optimizer, self.loss_name, self.learning_rate = optimizer, loss_name, learning_rate
        self.lr_scheduler, self.steplr_step_size = lr_scheduler, steplr_step_size
        self.weight_decay, self.eps, self.ckpt_name, self.model_name = weight_decay, eps, ckpt_name, model_name
        self.in_time, self.class_counts = in_time, class_counts
        self.scale_pos_weight, self.save_batch_val_metrics = scale_pos_weight, save_batch_val_metrics
        self.edge_class = edge_class if edge_class is not None else 2
        self.model_attr = f"{model_name}_{model_type}"
        setattr(self, self.model_attr, CultioNet(
            in_channels=in_channels, in_time=in_time, hidden_channels=hidden_channels, model_type=model_type, dropout=dropout,
            activation_type=activation_type, dilations=dilations, res_block_type=res_block_type, attention_weights=attention_weights,
            pool_by_max=pool_by_max, batchnorm_first=batchnorm_first,
        ))
        # reference default is Trainer(precision="16-mixed") (model.py:86); bf16 storage with fp32 accumulation is the B200 analogue
        self.cultionet_model.mask_model.set_compute_dtype(compute_dtype)

    @property
    def is_transfer_model(self) -> bool:
        return False

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location="cpu", strict: bool = True, **kwargs) -> "CultionetLitModel":
        """Lightning's classmethod as the reference calls it (``model.py:398-400``, ``:458-460``)."""
        from ..model import load_from_checkpoint

        return load_from_checkpoint(checkpoint_path, map_location=map_location, strict=strict, **kwargs)

    def freeze(self) -> None:
        """``LightningModule.freeze`` (``model.py:402``): no gradients, eval mode."""
        for p in self.parameters():
            p.requires_grad_(False)
        self.eval()
