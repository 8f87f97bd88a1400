"""``CultionetLitModel`` -- the reference's LightningModule surface (``src/cultionet/models/lightning.py:91-372``, ``:821-898``)
for the TowerUNet hot path: ``forward`` / ``predict_step`` / ``training_step`` / ``validation_step`` / ``calc_loss`` /
``get_true_labels`` / ``configure_optimizers`` with the same constructor keywords and the same attribute layout
(``cultionet_model`` stored under ``f"{model_name}_{model_type}"`` so checkpoints keep their key prefix).

``lightning`` is optional: when it is importable the class derives from ``lightning.LightningModule`` and drops into a Lightning
``Trainer``; otherwise it is a plain ``nn.Module`` and ``cultionet_b200.engine`` provides the training / prediction loops.
"""
from __future__ import annotations

import typing as T

import torch
import torch.nn as nn

from ..data import Data
from ..enums import AttentionTypes, InferenceNames, LearningRateSchedulers, LossTypes, ModelTypes, ResBlockTypes, ValidationNames
from ..losses import tower_unet_loss
from ..optim import FlatAdamW
from .cultionet import CultioNet

try:  # pragma: no cover - lightning is not installed in the build image
    from lightning import LightningModule as _Base

    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    _Base = nn.Module
    HAVE_LIGHTNING = False


class LightningModuleMixin(_Base):
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
        loss, parts = tower_unet_loss(predictions, batch.y, batch.bdist, edge_class=self.edge_class)
        return loss, {"dloss": parts[1], "eloss": parts[2], "closs": parts[3]}

    def training_step(self, batch: Data, batch_idx: int = None):
        predictions = self(batch)
        loss, _ = self.calc_loss(batch, predictions)
        if HAVE_LIGHTNING:  # pragma: no cover
            self.log("loss", loss, on_step=False, on_epoch=True, prog_bar=True, batch_size=batch.num_samples)
        return loss

    @torch.no_grad()
    def validation_step(self, batch: Data, batch_idx: int = None) -> dict:
        predictions = self(batch)
        loss, report = self.calc_loss(batch, predictions)
        labels = self.get_true_labels(batch)
        valid = labels[ValidationNames.MASK]
        valid = torch.ones_like(batch.y, dtype=torch.bool) if valid is None else valid.squeeze(1).bool()
        dist_mae = (predictions[InferenceNames.DISTANCE].squeeze(1) - batch.bdist).abs()[valid].mean()
        metrics = {"val_loss": loss, "vmae": dist_mae, "val_dloss": report["dloss"], "val_eloss": report["eloss"],
                   "val_closs": report["closs"]}
        for name, key, truth in (("vef1", InferenceNames.EDGE, ValidationNames.TRUE_EDGE), ("vcf1", InferenceNames.CROP, ValidationNames.TRUE_CROP)):
            pred = self.probas_to_labels(predictions[key])[valid]
            true = labels[truth][valid]
            tp = ((pred == 1) & (true == 1)).sum().float()
            fp = ((pred == 1) & (true == 0)).sum().float()
            fn = ((pred == 0) & (true == 1)).sum().float()
            metrics[name] = 5.0 * tp / (5.0 * tp + 4.0 * fn + fp).clamp_min(1.0)  # F-beta, beta = 2 (lightning.py:574-576)
        metrics["val_score"] = loss + (1.0 - metrics["vef1"]) + (1.0 - metrics["vcf1"]) + dist_mae
        return metrics

    def configure_optimizers(self, total_steps: T.Optional[int] = None):
        if self.optimizer != "AdamW":
            raise NameError("cultionet_b200 builds the reference's default optimizer only: choose 'AdamW'.")
        if self.lr_scheduler != LearningRateSchedulers.ONE_CYCLE_LR:
            raise NameError("The learning rate scheduler is not implemented in cultionet_b200 (OneCycleLR only).")
        return FlatAdamW(self.cultionet_model.parameters(), lr=self.learning_rate, betas=(0.9, 0.98), eps=self.eps,
                         weight_decay=self.weight_decay, clip_norm=1.0, total_steps=total_steps)


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
        compute_dtype: torch.dtype = torch.bfloat16,
    ):
        super().__init__()
        # what Lightning's save_hyperparameters() records (lightning.py:850) and writes into checkpoints as 'hyper_parameters'
        self.hyper_parameters = {k: v for k, v in locals().items() if k not in ("self", "__class__")}
        if loss_name != LossTypes.TANIMOTO_COMPLEMENT:
            raise NotImplementedError("cultionet_b200 builds the reference's default loss only (TanimotoComplementLoss)")
        if HAVE_LIGHTNING:  # pragma: no cover
            self.save_hyperparameters()
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

    if not HAVE_LIGHTNING:

        def freeze(self) -> None:
            """``LightningModule.freeze`` (``model.py:402``): no gradients, eval mode."""
            for p in self.parameters():
                p.requires_grad_(False)
            self.eval()
