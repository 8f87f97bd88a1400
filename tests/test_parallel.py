"""Data-parallel step on CPU: world_size 2 over gloo, kernels on the CPU interpreter.

Two ranks take different batches; after ``TrainStep`` both must hold identical parameters, equal to a single-process step
whose gradient is the mean of the two per-batch gradients (per-rank BatchNorm statistics, mean all-reduce of the flat gradient
-- what Lightning ``strategy="ddp"`` does for the reference, ``src/cultionet/model.py:168-186``)."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
CFG = dict(in_channels=2, in_time=6, hidden_channels=8, dropout=0.0, compute_dtype=torch.float32)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batch(rank: int):
    import cultionet_b200 as cb

    g = torch.Generator().manual_seed(100 + rank)
    return cb.Data(x=torch.rand(1, 2, 6, 16, 16, generator=g), y=torch.randint(-1, 3, (1, 16, 16), generator=g),
                   bdist=torch.rand(1, 16, 16, generator=g))


def _model():
    from cultionet_b200.models.lightning import CultionetLitModel

    torch.manual_seed(7)
    return CultionetLitModel(**CFG)


def _bind_emulator():
    from cultionet_b200 import _lib
    from tests.emu.build_emu import build

    _lib.use_library(build())


def _worker(rank: int, world: int, port: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    _bind_emulator()
    from cultionet_b200.engine import TrainStep
    from cultionet_b200.parallel import init_distributed

    r, _, w = init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    model = _model()
    step = TrainStep(model, total_steps=10, bucket_mb=0.01)  # tiny buckets: several all-reduces, launched from the grad hooks
    assert step.sync.enabled and len(step.sync.buckets) > 3
    loss = step(_batch(rank))
    torch.save({"loss": float(loss), "param": step.optimizer.flat_param.clone(), "grad": step.optimizer.flat_grad.clone(),
                "grad_scale": step.optimizer.grad_scale}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_gloo_step_matches_mean_gradient(tmp_path, dev):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(tmp_path / f"rank{r}.pt") for r in range(world))
    # replicas stay identical
    assert torch.equal(r0["param"], r1["param"])
    assert torch.equal(r0["grad"], r1["grad"])
    assert r0["grad_scale"] == 0.5
    assert r0["loss"] != r1["loss"]  # different shards

    # single-process check: mean of the two per-batch gradients, then the same optimizer step
    from cultionet_b200.engine import TrainStep

    grads = []
    for r in range(world):
        m = _model()
        s = TrainStep(m, total_steps=10)
        s.optimizer.zero_grad()
        m.training_step(_batch(r), 0).backward()
        grads.append(s.optimizer.flat_grad.clone())
    from tests.util import rel_err

    assert rel_err(grads[0] + grads[1], r0["grad"]) < 1e-4  # the buckets hold the SUM; AdamW applies 1/world (fp32 atomics reorder)
    m = _model()
    s = TrainStep(m, total_steps=10)
    s.optimizer.zero_grad()
    s.optimizer.flat_grad.copy_(0.5 * (grads[0] + grads[1]))
    s.optimizer.step()
    assert rel_err(s.optimizer.flat_param, r0["param"]) < 1e-4


def test_bucket_layout_covers_the_flat_gradient(dev):
    from cultionet_b200.engine import TrainStep

    step = TrainStep(_model(), total_steps=10, bucket_mb=0.02)
    spans = sorted((s, e) for s, e, _ in step.sync.buckets)
    assert spans[0][0] == 0 and spans[-1][1] == step.optimizer.numel
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert sum(n for _, _, n in step.sync.buckets) == len(step.optimizer.params)


def _tile_worker(rank: int, world: int, port: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    _bind_emulator()
    from cultionet_b200.parallel import init_distributed
    from tests import cases

    init_distributed(backend="gloo")
    import cultionet_b200.tile as tile_mod

    kept = {}
    orig_run = tile_mod.TilePredictor.run

    def run(self, batches=None):  # keep the predictor of the case to call its writer's gather afterwards
        kept["tp"] = self
        return orig_run(self, batches)

    tile_mod.TilePredictor.run = run
    mine = cases.tile_predictor_case("cpu", rank=rank, world_size=world)  # checks this rank's windows against the reference pipeline
    full = kept["tp"].writer.gather(dst=0)
    assert (full is None) == (rank != 0)
    torch.save({"mine": torch.from_numpy(mine.astype("int32")), "full": None if full is None else full.to(torch.int32).clone()},
               os.path.join(out_dir, f"tile{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_tile_prediction_shards_windows_and_gathers_mosaic(tmp_path, dev):
    """Windows i mod 2 go to rank i (no data-path collective); the writer rank's gathered mosaic equals the single-process one."""
    from tests import cases

    world = 2
    mp.spawn(_tile_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(tmp_path / f"tile{r}.pt") for r in range(world))
    single = torch.from_numpy(cases.tile_predictor_case(dev).astype("int32"))
    assert r1["full"] is None
    assert not torch.equal(r0["mine"], r1["mine"])
    assert int(((r0["mine"] != 0) & (r1["mine"] != 0)).sum()) == 0  # disjoint windows
    assert torch.equal(r0["full"], single)


def _fit_worker(rank: int, world: int, port: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    _bind_emulator()
    from cultionet_b200 import model as M
    from cultionet_b200.parallel import init_distributed

    init_distributed(backend="gloo")
    lit = _model()
    ckpt = Path(out_dir) / "ckpt" / "last.ckpt"
    hist = M.fit(lit, [_batch(rank), _batch(rank + 10)], val_batches=[_batch(rank + 20)], epochs=1, ckpt_file=ckpt, device="cpu",
                 cuda_graph=False)
    flat = torch.cat([p.detach().reshape(-1) for p in lit.parameters()])
    torch.save({"val_score": hist["val_score"], "checkpoint": hist["checkpoint"], "param": flat}, os.path.join(out_dir, f"fit{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_fit_writes_one_checkpoint_with_a_common_score(tmp_path, dev):
    """model.fit under a process group: replicas stay identical, the validation score is the mean over the ranks' shards (every rank
    takes the same save decision) and only rank 0 writes the checkpoint (Lightning's rank-zero-only ModelCheckpoint)."""
    world = 2
    mp.spawn(_fit_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(tmp_path / f"fit{r}.pt") for r in range(world))
    assert torch.equal(r0["param"], r1["param"])
    assert r0["val_score"] == r1["val_score"]
    assert r0["checkpoint"] is not None and r1["checkpoint"] is None
    ck = torch.load(r0["checkpoint"], weights_only=False)
    assert ck["epoch"] == 0 and int(ck["optimizer_states"][0]["state"][0]["step"]) == 2
