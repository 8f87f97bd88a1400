"""BASELINE config 5 for real: sliding-window prediction over ONE synthetic 10980 x 10980 Sentinel-2 tile (5 bands, T = 12, int16,
14.5 GB), 12 100 windows of 100 px + 20 px halo sharded `window i -> rank i mod world` (the reference's DistributedSampler order,
model.py:437-467) over all ranks, every rank's uint16 mosaic gathered onto rank 0 (MosaicWriter.gather).

  torchrun --nproc-per-node 8 tools/predict_full_tile.py [--side 10980] [--batch 32] [--stream]

Prints one JSON line on rank 0: seconds per tile and useful Mpx/s for (a) the tile resident in every rank's HBM (generated on the
device from one seed: identical on all ranks), (b) with --stream: the tile in pinned HOST memory of every rank, rows copied in ahead of the window
batches (TilePredictor.run_streaming) -- both INCLUDING the mosaic gather -- plus a spot check of three windows against the model run
on those windows alone."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
import numpy as np
import torch
import torch.distributed as dist

from cultionet_b200.data import Data
from cultionet_b200.models.lightning import CultionetLitModel
from cultionet_b200.parallel import init_distributed
from cultionet_b200.tile import TilePredictor, predict_windows

ap = argparse.ArgumentParser()
ap.add_argument("--side", type=int, default=10980)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--stream", action="store_true")
args = ap.parse_args()

rank, local, world = init_distributed()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
T, C, ws, pad = 12, 5, 100, 20
H = W = args.side
torch.manual_seed(1234)  # identical replicas
model = CultionetLitModel(in_channels=C, in_time=T, hidden_channels=64, dilations=[1, 2], dropout=0.0, compute_dtype=torch.bfloat16).to(dev).eval()
g = torch.Generator(device=dev).manual_seed(77)
tile = torch.randint(0, 10000, (T, C, H, W), generator=g, device=dev, dtype=torch.int16)  # the same tile on every rank
mean, std = torch.full((C,), 0.5), torch.full((C,), 0.29)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms: float) -> float:
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


tp = TilePredictor(model, tile, (mean, std), ws, pad, args.batch)
for i in range(3):  # warm-up: lazy initialisation + graph capture
    tp.step(i % tp.num_batches)
tp.writer.mosaic.zero_()
barrier()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
tp.run()
e1.record()
full = tp.writer.gather(dst=0) if world > 1 else tp.writer.mosaic
e2.record()
barrier()
if rank == 0:
    full = full.clone()  # the gather reduces in place: keep the result apart from the writer's storage (re-used by the streamed run)
ms_run, ms_total = max_over_ranks(e0.elapsed_time(e1)), max_over_ranks(e0.elapsed_time(e2))

# spot check on rank 0: three windows of the gathered mosaic against the model run on those windows alone
check = None
if rank == 0:
    from cultionet_b200.tile import MosaicWriter, WindowLoader

    wins = predict_windows(H, W, ws, pad)
    pick = [0, len(wins) // 2 + 3, len(wins) - 1]
    loader = WindowLoader(tile, ws, pad, (mean, std))
    sel = torch.from_numpy(np.ascontiguousarray(wins[pick])).to(dev)
    with torch.no_grad():
        batch = loader.load(sel)
        out = model.predict_step(Data(x=batch.x), 0)
        ref = MosaicWriter(H, W, dev, ws)
        ref.write_windows(out, sel, pad)
    worst = 0
    for r, c, h, w in wins[pick].tolist():
        a = full[:, r:r + h, c:c + w].int()
        b = ref.mosaic[:, r:r + h, c:c + w].int()
        worst = max(worst, int((a - b).abs().max()))
    covered = int((full[2].int() > 0).sum())
    check = {"windows_checked": len(pick), "max_abs_diff_uint16": worst, "nonzero_crop_pixels": covered, "pixels": H * W}
    assert worst <= 1, worst

line = None
if rank == 0:
    line = {"workload": f"cfg5 full tile: int16 [{T},{C},{H},{W}] ({tile.numel() * 2 / 1e9:.1f} GB), {len(predict_windows(H, W, ws, pad))} windows of "
                        f"{ws}+2x{pad} px, batches of {args.batch}, window i -> rank i mod {world}, mosaic gathered on rank 0",
            "n_gpus": world, "resident": {"s_per_tile": ms_total / 1e3, "Mpx_per_s": H * W / 1e6 / (ms_total / 1e3),
                                          "predict_s": ms_run / 1e3, "gather_s": (ms_total - ms_run) / 1e3},
            "windows_per_rank": tp.num_windows, "batches_per_rank": tp.num_batches, "launches_per_batch": tp.launches_per_batch,
            "spot_check": check}

if args.stream:
    import psutil

    need = tile.numel() * 2 * world
    ok = psutil.virtual_memory().available > 2.5 * need
    flag = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag):
        host_tile = torch.empty(tile.shape, dtype=torch.int16).pin_memory()
        host_tile.copy_(tile)
        torch.cuda.synchronize()
        tile.zero_()  # nothing can be read before it has arrived from the host
        tp.writer.mosaic.zero_()
        host_mosaic = torch.empty(tuple(tp.writer._store.shape), dtype=torch.uint16).pin_memory()
        barrier()
        e0.record()
        tp.run_streaming(host_tile, host_mosaic)
        e1.record()
        full2 = tp.writer.gather(dst=0) if world > 1 else tp.writer.mosaic
        e2.record()
        barrier()
        ms_total2 = max_over_ranks(e0.elapsed_time(e2))
        if rank == 0:
            same = bool(torch.equal(full2, full)) if full is not full2 else True
            line["streamed_from_host"] = {"s_per_tile": ms_total2 / 1e3, "Mpx_per_s": H * W / 1e6 / (ms_total2 / 1e3),
                                          "h2d_bytes_per_rank": int(tile.numel() * 2), "d2h_bytes_per_rank": int(host_mosaic.numel() * 2),
                                          "mosaic_equals_resident_run": same}
    elif rank == 0:
        line["streamed_from_host"] = {"skipped": f"host memory: {psutil.virtual_memory().available / 1e9:.0f} GB available, {need / 1e9:.0f} GB of pinned tiles needed"}

if rank == 0:
    print(json.dumps(line))
sys.stdout.flush()
torch.cuda.synchronize()
tp._graph = None
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
