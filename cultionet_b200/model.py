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

from pathlib import Path
from typing import Iterable, Optional, Union

import torch
import torch.distributed as dist

from .data import Data
from .engine import TrainStep, batch_to_device
from .models.lightning import CultionetLitModel

CKPT_FORMAT_VERSION = "2.1.0"  # the reference pins lightning>=2.1 (setup.cfg)


def checkpoint_dict(lit_model: CultionetLitModel, optimizer=None, epoch: int = 0, global_step: int = 0, **extra) -> dict:
    ckpt = {
        "epoch": int(epoch),
        "global_step": int(global_step),
        "pytorch-lightning_version": CKPT_FORMAT_VERSION,
        "state_dict": {k: v.detach().cpu().clone() for k, v in lit_model.state_dict().items()},
        "hyper_parameters": dict(lit_model.hyper_parameters),
        "optimizer_states": [] if optimizer is None else [{k: (v.detach().cpu().clone() if isinstance(v, torch.Tensor) else v)
                                                            for k, v in optimizer.state_dict().items()}],
        "lr_schedulers": [],
    }
    ckpt.update(extra)
    return ckpt


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
    ckpt = torch.load(str(checkpoint_path), map_location="cpu", weights_only=False)
    if "state_dict" not in ckpt:
        raise KeyError(f"{checkpoint_path} is not a Lightning checkpoint (no 'state_dict')")
    hp = dict(ckpt.get("hyper_parameters", {}))
    hp.update(overrides)
    if "in_channels" not in hp or "in_time" not in hp:
        raise KeyError("the checkpoint holds no in_channels / in_time hyper-parameters; pass them as keyword arguments")
    known = set(CultionetLitModel.__init__.__code__.co_varnames)
    model = CultionetLitModel(**{k: v for k, v in hp.items() if k in known})
    model.load_state_dict(ckpt["state_dict"], strict=strict)
    if map_location not in (None, "cpu"):
        model = model.to(map_location)
    model.loaded_checkpoint = {k: ckpt[k] for k in ("epoch", "global_step", "optimizer_states") if k in ckpt}
    return model


def fit(lit_model: CultionetLitModel, train_batches, val_batches=None, epochs: int = 1, ckpt_file: Optional[Union[str, Path]] = None,
        device=None, cuda_graph: bool = True, steps_per_epoch: Optional[int] = None, log=None) -> dict:
    """Train for ``epochs`` passes over ``train_batches`` (a re-iterable of ``Data``; a callable returning an iterator also works).
    After every epoch the mean of ``validation_step`` over ``val_batches`` gives ``val_score``; the best one is written to
    ``ckpt_file``.  An existing ``ckpt_file`` is resumed (weights, AdamW moments, step count, epoch).  Returns the history."""
    device = torch.device(device) if device is not None else next(lit_model.parameters()).device
    lit_model.to(device)
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def iterate(src) -> Iterable[Data]:
        return src() if callable(src) else iter(src)

    if steps_per_epoch is None:
        steps_per_epoch = sum(1 for _ in iterate(train_batches))
    start_epoch = 0
    resume = None
    if ckpt_file is not None and Path(ckpt_file).is_file():  # model.py:308-314
        resume = torch.load(str(ckpt_file), map_location="cpu", weights_only=False)
        lit_model.load_state_dict(resume["state_dict"])
        start_epoch = int(resume.get("epoch", -1)) + 1
    step = TrainStep(lit_model, total_steps=max(1, epochs * steps_per_epoch), cuda_graph=cuda_graph)
    if resume is not None and resume.get("optimizer_states"):
        st = resume["optimizer_states"][0]
        step.optimizer.load_state_dict({k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in st.items()})
    history = {"loss": [], "val_score": [], "best_val_score": float(resume["best_val_score"]) if resume and "best_val_score" in resume
               else float("inf"), "checkpoint": None}
    global_step = step.optimizer.step_count
    for epoch in range(start_epoch, epochs):
        lit_model.train()
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
            sums, m = {}, 0
            with torch.no_grad():
                for batch in iterate(val_batches):
                    for k, v in lit_model.validation_step(batch_to_device(batch, device), m).items():
                        sums[k] = sums.get(k, 0.0) + float(v)
                    m += 1
            metrics = {k: v / max(m, 1) for k, v in sums.items()}
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
