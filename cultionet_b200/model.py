"""``fit`` / ``predict`` entry points around the hot path when Lightning is not installed: what
``src/cultionet/model.py:273-328`` (``fit``) and ``:405-467`` (``predict_lightning``) get from ``lightning.Trainer`` for this path
-- epochs over batches, per-epoch validation, the best-``val_score`` checkpoint (``callbacks.py:238-249``:
``ModelCheckpoint(monitor="val_score", mode="min", save_top_k=1)``), resume from ``ckpt_file`` when it exists (``model.py:308-314``)
-- on ``engine.TrainStep`` and ``tile.TilePredictor``.

Checkpoints use Lightning's file layout so that they interchange with the reference's ``last.ckpt``: a ``torch.save``d dict with
``state_dict`` (keys prefixed by the module attribute ``f"{model_name}_{model_type}"`` = ``cultionet_TowerUNet.``, ``lightning.py:874``),
``hyper_parameters`` (the ``CultionetLitModel`` keyword arguments, ``save_hyperparameters()`` at ``lightning.py:850``), ``epoch``,
``global_step``, ``optimizer_states`` and ``pytorch-lightning_version``.  ``load_from_checkpoint`` accepts files written by either
side (the ``pre_unet._orig_mod.`` infix ``torch.compile`` adds in the reference, ``nunet.py:141``, is handled by the model's
``load_state_dict``)."""
from __future__ import annotations

import io
import pickle
from pathlib import Path
from typing import Iterable, Optional, Sequence, Union

import attr
import torch
import torch.distributed as dist

from . import enums as _enums
from .data import Data
from .engine import TrainStep, batch_to_device
from .enums import LearningRateSchedulers, LossTypes, ModelTypes, ResBlockTypes
from .models.lightning import CultionetLitModel

CKPT_FORMAT_VERSION = "2.1.0"  # the reference pins lightning>=2.1 (setup.cfg)


def _opt_path(v):
    return None if v is None else Path(v)


def _opt_list(v):
    return None if v is None else list(v)


@attr.s
class CultionetParams:
    """``src/cultionet/model.py:46-186``: the one bag of settings ``cultionet train`` builds and hands to ``fit``; same field names,
    defaults and converters.  ``get_lightning_params()`` feeds ``CultionetLitModel(**...)`` exactly as ``model.py:285`` does;
    ``get_trainer_params()`` is what ``fit_params`` below reads instead of a ``lightning.Trainer``.  Fields that configure parts of the
    reference outside this path (SWA, pruning, LR finder, profiler) are carried so that calling code keeps working; ``fit_params``
    rejects the ones that would change training if silently ignored."""

    ckpt_file: Union[str, Path] = attr.ib(converter=_opt_path, default=None)
    spatial_partitions: str = attr.ib(default=None)
    dataset = attr.ib(default=None)
    test_dataset = attr.ib(default=None)
    val_frac: float = attr.ib(converter=float, default=0.2)
    batch_size: int = attr.ib(converter=int, default=4)
    load_batch_workers: int = attr.ib(converter=int, default=0)
    edge_class: int = attr.ib(converter=attr.converters.optional(int), default=None)
    class_counts: torch.Tensor = attr.ib(default=None)
    hidden_channels: int = attr.ib(converter=int, default=64)
    model_type: str = attr.ib(converter=str, default=ModelTypes.TOWERUNET)
    activation_type: str = attr.ib(converter=str, default="SiLU")
    dropout: float = attr.ib(converter=float, default=0.1)
    dilations: Union[int, Sequence[int]] = attr.ib(converter=_opt_list, default=None)
    res_block_type: str = attr.ib(converter=str, default=ResBlockTypes.RESA)
    attention_weights: str = attr.ib(default=None)
    optimizer: str = attr.ib(converter=str, default="AdamW")
    loss_name: str = attr.ib(converter=str, default=LossTypes.TANIMOTO_COMPLEMENT)
    learning_rate: float = attr.ib(converter=float, default=0.01)
    lr_scheduler: str = attr.ib(converter=str, default=LearningRateSchedulers.ONE_CYCLE_LR)
    steplr_step_size: int = attr.ib(converter=int, default=5)
    weight_decay: float = attr.ib(converter=float, default=1e-3)
    eps: float = attr.ib(converter=float, default=1e-4)
    ckpt_name: str = attr.ib(converter=str, default="last")
    model_name: str = attr.ib(converter=str, default="cultionet")
    pool_by_max: bool = attr.ib(default=False)
    batchnorm_first: bool = attr.ib(default=False)
    scale_pos_weight: bool = attr.ib(default=False)
    save_batch_val_metrics: bool = attr.ib(default=False)
    epochs: int = attr.ib(converter=int, default=100)
    accumulate_grad_batches: int = attr.ib(converter=int, default=1)
    gradient_clip_val: float = attr.ib(converter=float, default=1.0)
    gradient_clip_algorithm: str = attr.ib(converter=str, default="norm")
    precision: Union[int, str] = attr.ib(default="16-mixed")
    device: str = attr.ib(converter=str, default="gpu")
    devices: int = attr.ib(converter=int, default=1)
    reset_model: bool = attr.ib(default=False)
    auto_lr_find: bool = attr.ib(default=False)
    stochastic_weight_averaging: bool = attr.ib(default=False)
    stochastic_weight_averaging_lr: float = attr.ib(converter=float, default=0.05)
    stochastic_weight_averaging_start: float = attr.ib(converter=float, default=0.8)
    model_pruning: bool = attr.ib(default=False)
    skip_train: bool = attr.ib(default=False)
    finetune: str = attr.ib(default=None)
    strategy: str = attr.ib(converter=str, default="ddp")
    profiler: str = attr.ib(default=None)

    def check_checkpoint(self) -> None:
        if self.reset_model:
            if self.ckpt_file.is_file():
                self.ckpt_file.unlink()
            model_file = self.ckpt_file.parent / f"{self.model_name}.pt"
            if model_file.is_file():
                model_file.unlink()

    def update_channels(self, data_module=None, in_channels: Optional[int] = None, in_time: Optional[int] = None) -> "CultionetParams":
        """``in_channels`` / ``in_time`` from the data module's training set (``model.py:109-115``) or given directly."""
        if data_module is not None:
            in_channels, in_time = data_module.train_ds.num_channels, data_module.train_ds.num_time
        self.in_channels, self.in_time = int(in_channels), int(in_time)
        return self

    def get_callback_params(self) -> dict:
        return dict(ckpt_file=self.ckpt_file, stochastic_weight_averaging=self.stochastic_weight_averaging,
                    stochastic_weight_averaging_lr=self.stochastic_weight_averaging_lr,
                    stochastic_weight_averaging_start=self.stochastic_weight_averaging_start, model_pruning=self.model_pruning)

    def get_datamodule_params(self) -> dict:
        return dict(dataset=self.dataset, test_dataset=self.test_dataset, val_frac=self.val_frac,
                    spatial_partitions=self.spatial_partitions, batch_size=self.batch_size, load_batch_workers=self.load_batch_workers)

    def get_lightning_params(self) -> dict:
        return dict(
            in_channels=self.in_channels, in_time=self.in_time, hidden_channels=self.hidden_channels, model_type=self.model_type,
            dropout=self.dropout, activation_type=self.activation_type, dilations=self.dilations, res_block_type=self.res_block_type,
            attention_weights=self.attention_weights, optimizer=self.optimizer, loss_name=self.loss_name, learning_rate=self.learning_rate,
            lr_scheduler=self.lr_scheduler, steplr_step_size=self.steplr_step_size, weight_decay=self.weight_decay, eps=self.eps,
            ckpt_name=self.ckpt_name, model_name=self.model_name, pool_by_max=self.pool_by_max, batchnorm_first=self.batchnorm_first,
            class_counts=self.class_counts, edge_class=self.edge_class, scale_pos_weight=self.scale_pos_weight,
            save_batch_val_metrics=self.save_batch_val_metrics,
        )

    def get_trainer_params(self) -> dict:
        return dict(
            default_root_dir=str(self.ckpt_file.parent) if self.ckpt_file is not None else None, enable_checkpointing=True,
            accumulate_grad_batches=self.accumulate_grad_batches, gradient_clip_val=self.gradient_clip_val,
            gradient_clip_algorithm=self.gradient_clip_algorithm, check_val_every_n_epoch=1, min_epochs=5 if self.epochs >= 5 else self.epochs,
            max_epochs=self.epochs, precision=self.precision, devices=self.devices, accelerator=self.device, log_every_n_steps=50,
            deterministic=False, benchmark=False, strategy=self.strategy, profiler=self.profiler,
        )


def compute_dtype_of(precision) -> torch.dtype:
    """Lightning ``precision`` -> compute dtype of this build: any 16-bit setting runs bf16 storage with fp32 accumulation (the B200
    analogue of the reference's default ``"16-mixed"``, ``model.py:86``), 32 runs the fp32 parity kernels."""
    return torch.float32 if str(precision).startswith("32") else torch.bfloat16


def fit_params(params: CultionetParams, train_batches, val_batches=None, **kwargs) -> dict:
    """``cultionet.model.fit(cultionet_params)`` (``model.py:273-328``) for batches already in hand: ``check_checkpoint``, build
    ``CultionetLitModel(**get_lightning_params())`` (``:285``), train ``epochs`` with per-epoch validation and the best-``val_score``
    checkpoint at ``ckpt_file``.  Settings this path does not build raise instead of being ignored."""
    if params.accumulate_grad_batches != 1:
        raise NotImplementedError("cultionet_b200.fit: accumulate_grad_batches != 1 is not built")
    if abs(params.gradient_clip_val - 1.0) > 0 or params.gradient_clip_algorithm != "norm":
        raise NotImplementedError("cultionet_b200.fit: the optimiser kernel clips the global gradient norm at 1.0 (the reference default)")
    for name in ("stochastic_weight_averaging", "model_pruning", "auto_lr_find"):
        if getattr(params, name):
            raise NotImplementedError(f"cultionet_b200.fit: {name} is outside the TowerUNet hot path and not built")
    if not hasattr(params, "in_channels"):
        first = next(iter(train_batches() if callable(train_batches) else train_batches))
        params.update_channels(in_channels=first.x.shape[1], in_time=first.x.shape[2])
    if params.ckpt_file is not None:
        params.check_checkpoint()
    model = CultionetLitModel(**params.get_lightning_params(), compute_dtype=compute_dtype_of(params.precision))
    if params.skip_train:
        return {"model": model, "loss": [], "val_score": [], "checkpoint": None}
    hist = fit(model, train_batches, val_batches, epochs=params.epochs, ckpt_file=params.ckpt_file, **kwargs)
    hist["model"] = model
    return hist


def checkpoint_dict(lit_model: CultionetLitModel, optimizer=None, epoch: int = 0, global_step: int = 0, **extra) -> dict:
    ckpt = {
        "epoch": int(epoch),
        "global_step": int(global_step),
        "pytorch-lightning_version": CKPT_FORMAT_VERSION,
        "state_dict": {k: v.detach().cpu().clone() for k, v in lit_model.state_dict().items()},
        "hyper_parameters": dict(lit_model.hyper_parameters),
        # torch.optim.AdamW layout ({state, param_groups}, parameters in cultionet_model.parameters() order)
        "optimizer_states": [] if optimizer is None else [_to_cpu(optimizer.state_dict())],
        "lr_schedulers": [] if optimizer is None else [{"last_epoch": int(optimizer.step_count), "_step_count": int(optimizer.step_count) + 1,
                                                        "total_steps": optimizer.total_steps, "_last_lr": [optimizer.current_lr()]}],
    }
    ckpt.update(extra)
    return ckpt


def _to_cpu(obj):
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu().clone()
    if isinstance(obj, dict):
        return {k: _to_cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_cpu(v) for v in obj)
    return obj


class _ForeignEnum(str):
    """Stand-in for a ``cultionet.enums`` member pickled into a reference checkpoint's ``hyper_parameters``: keeps the string value."""

    def __new__(cls, value=""):
        return str.__new__(cls, value)


class _CkptUnpickler(pickle.Unpickler):
    """``cultionet.enums.X`` (the reference package, normally not importable next to this one) resolves to this package's enum of the same
    name, or to a plain string type; everything else unpickles as usual."""

    def find_class(self, module, name):
        if module == "cultionet.enums" or module.startswith("cultionet.enums."):
            return getattr(_enums, name, _ForeignEnum)
        return super().find_class(module, name)


class _ckpt_pickle:  # the `pickle_module` protocol torch.load expects
    __name__ = "cultionet_b200_ckpt_pickle"
    Unpickler = _CkptUnpickler

    @staticmethod
    def load(f, **kw):
        return _CkptUnpickler(f, **kw).load()


def read_checkpoint(path: Union[str, Path]) -> dict:
    """``torch.load`` of a Lightning checkpoint written by either side (reference checkpoints pickle ``cultionet.enums`` members)."""
    return torch.load(str(path), map_location="cpu", weights_only=False, pickle_module=_ckpt_pickle)


def save_checkpoint(lit_model: CultionetLitModel, path: Union[str, Path], optimizer=None, epoch: int = 0, global_step: int = 0,
                    **extra) -> Path:
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    tmp = path.with_suffix(path.suffix + ".tmp")
    torch.save(checkpoint_dict(lit_model, optimizer, epoch, global_step, **extra), tmp)
    tmp.replace(path)  # a reader never sees a half-written checkpoint
    return path


def load_from_checkpoint(checkpoint_path: Union[str, Path], map_location="cpu", strict: bool = True, **overrides) -> CultionetLitModel:
    """``CultionetLitModel.load_from_checkpoint`` (``model.py:458-460``): rebuild the module from ``hyper_parameters`` (keyword
    ``overrides`` win, e.g. ``compute_dtype``) and load ``state_dict``."""
    ckpt = read_checkpoint(checkpoint_path)
    if "state_dict" not in ckpt:
        raise KeyError(f"{checkpoint_path} is not a Lightning checkpoint (no 'state_dict')")
    hp = dict(ckpt.get("hyper_parameters", {}))
    hp.update(overrides)
    if "in_channels" not in hp or "in_time" not in hp:
        raise KeyError("the checkpoint holds no in_channels / in_time hyper-parameters; pass them as keyword arguments")
    known = set(CultionetLitModel.__init__.__code__.co_varnames)
    hp = {k: (str(v) if isinstance(v, (_ForeignEnum, _enums.StrEnum)) else v) for k, v in hp.items()}
    model = CultionetLitModel(**{k: v for k, v in hp.items() if k in known})
    model.load_state_dict(ckpt["state_dict"], strict=strict)
    if map_location not in (None, "cpu"):
        model = model.to(map_location)
    model.loaded_checkpoint = {k: ckpt[k] for k in ("epoch", "global_step", "optimizer_states") if k in ckpt}
    return model


def fit(lit_model: CultionetLitModel, train_batches, val_batches=None, epochs: int = 1, ckpt_file: Optional[Union[str, Path]] = None,
        device=None, cuda_graph: bool = True, steps_per_epoch: Optional[int] = None, log=None,
        onecycle_reference_span: bool = True) -> dict:
    """Train for ``epochs`` passes over ``train_batches`` (a re-iterable of ``Data``; a callable returning an iterator also works).
    After every epoch the batch-size-weighted mean of ``validation_step`` over ``val_batches`` gives ``val_score`` (what Lightning's
    ``log_dict(on_epoch=True, batch_size=...)`` reduces to); the best one is written to ``ckpt_file``.  An existing ``ckpt_file`` is
    resumed (weights, AdamW moments, step count, epoch) whether this package or the reference wrote it.  Returns the history.

    ``onecycle_reference_span``: the reference builds ``OneCycleLR(epochs=max_epochs, steps_per_epoch=trainer.estimated_stepping_batches)``
    (``models/lightning.py:658-664``) and ``estimated_stepping_batches`` already counts ALL epochs, so its cycle spans
    ``epochs * (epochs * steps_per_epoch)`` steps and a run only walks the first ``1 / epochs`` of it.  ``True`` reproduces that learning
    rate curve; ``False`` spans the cycle over the run."""
    device = torch.device(device) if device is not None else next(lit_model.parameters()).device
    lit_model.to(device)
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def iterate(src) -> Iterable[Data]:
        return src() if callable(src) else iter(src)

    if steps_per_epoch is None:
        try:
            steps_per_epoch = len(train_batches)  # a list / DataLoader: no pass over the data just to count it
        except TypeError:
            steps_per_epoch = sum(1 for _ in iterate(train_batches))
    start_epoch = 0
    resume = None
    if ckpt_file is not None and Path(ckpt_file).is_file():  # model.py:308-314
        resume = read_checkpoint(ckpt_file)
        lit_model.load_state_dict(resume["state_dict"])
        start_epoch = int(resume.get("epoch", -1)) + 1
    run_steps = max(1, epochs * steps_per_epoch)
    step = TrainStep(lit_model, total_steps=run_steps * (epochs if onecycle_reference_span else 1), cuda_graph=cuda_graph,
                     steps_per_epoch=steps_per_epoch)
    if resume is not None and resume.get("optimizer_states"):
        step.optimizer.load_state_dict(resume["optimizer_states"][0])
    history = {"loss": [], "val_score": [], "best_val_score": float(resume["best_val_score"]) if resume and "best_val_score" in resume
               else float("inf"), "checkpoint": None}
    global_step = step.optimizer.step_count
    for epoch in range(start_epoch, epochs):
        lit_model.train()
        step.optimizer.set_epoch(epoch)
        running, n = None, 0
        for batch in iterate(train_batches):
            loss = step(batch_to_device(batch, device))
            running = loss.clone() if running is None else running + loss  # stays on the device: no per-step host sync
            n += 1
            global_step += 1
        epoch_loss = float(running / max(n, 1)) if running is not None else float("nan")
        history["loss"].append(epoch_loss)
        score = epoch_loss
        if val_batches is not None:
            lit_model.eval()
            sums, m, weight = {}, 0, 0.0
            with torch.no_grad():
                for batch in iterate(val_batches):
                    bs = float(batch.x.shape[0])
                    for k, v in lit_model.validation_step(batch_to_device(batch, device), m).items():
                        sums[k] = sums.get(k, 0.0) + bs * float(v)
                    m += 1
                    weight += bs
            metrics = {k: v / max(weight, 1.0) for k, v in sums.items()}
            score = metrics.get("val_score", epoch_loss)
            history.setdefault("val_metrics", []).append(metrics)
        history["val_score"].append(score)
        if log is not None:
            log(f"epoch {epoch}: loss {epoch_loss:.5f} val_score {score:.5f} lr {step.optimizer.current_lr():.5g}")
        if distributed:  # every rank must take the same save decision: use the mean score of the ranks' validation shards
            t = torch.tensor([score], dtype=torch.float64, device=device if device.type == "cuda" else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            score = float(t) / dist.get_world_size()
            history["val_score"][-1] = score
        if ckpt_file is not None and score <= history["best_val_score"]:
            history["best_val_score"] = score
            if not distributed or dist.get_rank() == 0:  # replicas are identical: one writer (Lightning's rank-zero-only checkpointing)
                history["checkpoint"] = str(save_checkpoint(lit_model, ckpt_file, step.optimizer, epoch, global_step,
                                                            best_val_score=score))
            if distributed:
                dist.barrier()
    step.close()
    return history


def predict_tile(ckpt_file: Union[str, Path], tile: torch.Tensor, norm_values=None, window_size: int = 100, padding: int = 20,
                 batch_size: int = 32, device="cuda", compute_dtype: torch.dtype = torch.bfloat16, gather_to: Optional[int] = 0):
    """``predict_lightning`` for a tile already in memory (``model.py:405-467``): load the checkpoint, run every prediction window of
    this rank through ``tile.TilePredictor`` and return the 3-band uint16 mosaic (on rank ``gather_to`` when a process group is
    initialised; ``None`` on the other ranks).  A host (pinned) tile is streamed in, a device tile is used in place."""
    from .tile import TilePredictor

    model = load_from_checkpoint(ckpt_file, map_location=device, compute_dtype=compute_dtype)
    model.eval()
    device = torch.device(device)
    if tile.device.type == device.type:
        tp = TilePredictor(model, tile, norm_values, window_size, padding, batch_size)
        tp.run()
    else:
        resident = torch.zeros(tile.shape, dtype=torch.int16, device=device)
        tp = TilePredictor(model, resident, norm_values, window_size, padding, batch_size)
        tp.run_streaming(tile.contiguous())
    if device.type == "cuda":
        torch.cuda.synchronize(device)
    return tp.writer.gather(dst=gather_to) if gather_to is not None else tp.writer.mosaic
