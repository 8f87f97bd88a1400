#!/usr/bin/env python
"""bench.py -- train chips/s of the TowerUNet hot path (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|tiny] [--batch B]

A step = one pass of the hot path over one batch of synthetic chips: forward + Tanimoto-complement loss + backward +
data-parallel gradient all-reduce + AdamW, through the reference-facing API (``CultionetLitModel.training_step`` driven by
``cultionet_b200.engine.TrainStep``).  One process per GPU (torchrun); rank 0 prints ONE JSON line.

  value     chips/s with the batch already resident in HBM (CUDA events, barrier + synchronize both sides, max over ranks)
  e2e       the same steps with the batch in pinned HOST memory: H2D copy of x/y/bdist and a D2H read of the loss inside the region
  roofline  the dominant kernel class (implicit-GEMM convolution: fwd/dgrad/wgrad launches) timed per launch with CUDA events
            on the launching stream in a separate instrumented pass: algorithmic FLOPs / summed launch time vs the measured
            bf16 peak in MEASURED_PEAKS.json
  cpu_baseline  the UNMODIFIED reference modules (``baseline/_ref/cultionet``: TowerUNet + TanimotoComplementLoss + torch AdamW,
            fp32; its one absent third-party dependency, natten, is the torch restatement oracle/natten_ref.py) on the host
            cores, rank 0, N=1 only, on a bounded sample (batch 2 of the same chip shape); the oracle port
            (kind "port") when ``baseline/_ref`` is absent
  secondary     N=1 only: short runs of the other single-GPU BASELINE configs (cfg1 fp32, cfg4 long series / NA k7 d2, cfg5
            sliding-window inference Mpx/s) as sub-processes of this script, their own JSON lines nested under this key

``--impl reference`` times that CPU reference alone with all host threads, same metric / unit / config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")

import torch  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]/[2]: TowerUNet bf16 training step, batch 32 per GPU, x=[32,5,24,128,128], hidden 64
    "cfg2": dict(B=32, C=5, T=24, H=128, W=128, hidden=64, dilations=[1, 2], dtype="bf16", fwd_gflop_per_chip=423.0),
    # BASELINE.json configs[0]: fp32, x=[4,3,12,100,100], hidden 32 (the reference's CPU-runnable case)
    "cfg1": dict(B=4, C=3, T=12, H=100, W=100, hidden=32, dilations=[1, 2], dtype="f32", fwd_gflop_per_chip=64.9),
    # BASELINE.json configs[3]: long time series T=36, neighbourhood attention kernel 7 dilation 2, batch 16 of 256x256 chips
    "cfg4": dict(B=16, C=5, T=36, H=256, W=256, hidden=64, dilations=[1, 2], dtype="bf16", fwd_gflop_per_chip=1697.0,
                 natten=dict(natten_kernel_size=7, natten_dilation=2)),
    # BASELINE.json configs[4] on ONE GPU: eval-mode sliding-window prediction, windows of 100 px + 20 px halo = 140x140 chips
    # (C=5, T=12); a step = one batch of 32 windows through predict_step; metric = useful (un-padded) Mpx/s
    "cfg5": dict(B=32, C=5, T=12, H=140, W=140, hidden=64, dilations=[1, 2], dtype="bf16", fwd_gflop_per_chip=505.0, predict=True,
                 useful_px_per_chip=100 * 100),
    "tiny": dict(B=2, C=3, T=8, H=32, W=32, hidden=16, dilations=[1, 2], dtype="bf16", fwd_gflop_per_chip=0.0),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch override")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying the captured CUDA graph")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short cfg1 / cfg4 / cfg5 runs nested under 'secondary'")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.is_file():
        d = json.loads(p.read_text())
        return {"bf16_tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "bf16_tflops_burst": d.get("bf16_tflops"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured"}
    return {"bf16_tflops": 1400.0, "bf16_tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


# ---------------------------------------------------------------------------------------------------------------------
def make_host_batch(w: dict, batch: int, rank: int, pin: bool):
    """Synthetic chips as SURVEY.md 8(d): x = rand, y in {0,1,2} (edge class 2), bdist = rand; seeded per rank."""
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.rand(batch, w["C"], w["T"], w["H"], w["W"], generator=g)
    y = torch.randint(0, 3, (batch, w["H"], w["W"]), generator=g)
    bdist = torch.rand(batch, w["H"], w["W"], generator=g)
    if pin:
        x, y, bdist = x.pin_memory(), y.pin_memory(), bdist.pin_memory()
    return x, y, bdist


def _reference_modules():
    """The unmodified reference (pip-installed from /root/reference into baseline/_ref) through the package-shell
    loader; None when it did not travel."""
    ref_src = ROOT / "baseline" / "_ref" / "cultionet"
    if not (ref_src / "models" / "nunet.py").is_file():
        return None
    os.environ["CULTIONET_REFERENCE_SRC"] = str(ref_src)
    from oracle import ref_loader

    ref_loader.REFERENCE_SRC = ref_src
    return ref_loader.load_reference()


def cpu_reference_chips_per_s(w: dict, steps: int, warmup: int, batch: int = 2) -> dict:
    """fwd + loss + bwd + clip + AdamW on the host cores (fp32), a bounded sample of the workload: the reference's own TowerUNet /
    TanimotoComplementLoss modules when baseline/_ref travelled (kind "reference"), else the oracle port (kind "port")."""
    torch.set_num_threads(os.cpu_count() or 1)
    x, y, bdist = make_host_batch(w, batch, 0, pin=False)
    ref = None
    try:
        ref = _reference_modules()
    except Exception as e:  # noqa: BLE001
        print(f"[bench] reference modules unavailable ({e!r}); timing the oracle port", file=sys.stderr)
    if w.get("natten") and ref is not None:
        for lvl in ("a", "b", "c"):
            ref.unet_parts.NATTEN_PARAMS[lvl].update(w["natten"])
    predict = bool(w.get("predict"))
    if ref is not None:
        kind = "reference"
        torch.manual_seed(0)
        model = ref.TowerUNet(in_channels=w["C"], in_time=w["T"], hidden_channels=w["hidden"], dilations=w["dilations"], dropout=0.0)
        loss_fn_c = ref.TanimotoComplementLoss()
        opt = torch.optim.AdamW(model.parameters(), lr=0.01, betas=(0.9, 0.98), eps=1e-4, weight_decay=1e-3)

        def one_step():
            if predict:
                with torch.no_grad():
                    model(x)
                return
            out = model(x)
            # label recoding + (distance + edge + crop) / 3 as reference models/lightning.py:161-207, :318-354
            true_edge = (y == 2).long()
            true_crop = ((y > 0) & (y < 2)).long()
            loss = (loss_fn_c(out["distance"], bdist) + loss_fn_c(out["edge"], true_edge) + loss_fn_c(out["crop"], true_crop)) / 3.0
            opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()

        model.eval() if predict else model.train()
    else:
        kind = "port"
        from oracle import towerunet_port as port

        spec = port.param_spec(w["C"], w["T"], w["hidden"], w["dilations"])
        sd = port.synth_state_dict(spec, seed=0)
        sd = {k: (v.requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}

        def one_step():
            if predict:
                with torch.no_grad():
                    port.towerunet_forward(sd, x, w["dilations"], training=False)
                return
            out = port.towerunet_forward(sd, x, w["dilations"], training=True)
            loss, _ = port.training_loss(out, y, bdist)
            loss.backward()
            for v in sd.values():
                if v.grad is not None:
                    v.grad = None

    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one_step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = statistics.median(times)
    what = ("the reference's own modules (baseline/_ref: TowerUNet + TanimotoComplementLoss + torch AdamW; natten = torch restatement)"
            if kind == "reference" else "oracle port")
    if predict:
        px = batch * w.get("useful_px_per_chip", w["H"] * w["W"])
        return {"value": px / 1e6 / sec, "unit": "Mpx/s", "cores": torch.get_num_threads(), "kind": kind,
                "sample": f"{what}, fp32 eval forward, batch {batch} of x=[{w['C']},{w['T']},{w['H']},{w['W']}] hidden {w['hidden']}, "
                          f"median of {steps} steps after {warmup} warm-up", "s_per_step": sec}
    return {"value": batch / sec, "unit": "chips/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{what}, fp32 fwd+loss+bwd+clip+AdamW, batch {batch} of x=[{w['C']},{w['T']},{w['H']},{w['W']}] hidden {w['hidden']}, "
                      f"median of {steps} steps after {warmup} warm-up", "s_per_step": sec}


def secondary_runs(args) -> dict:
    """Short runs of the other single-GPU BASELINE configs through this same script (one sub-process each: a workload owns module
    level state such as NATTEN_PARAMS and its own CUDA graphs).  Returns {workload: parsed JSON line or {"error": ...}}."""
    out = {}
    for name, steps in (("cfg1", 10), ("cfg4", 5), ("cfg5", 10)):
        cmd = [sys.executable, str(ROOT / "bench.py"), "--gpus", "1", "--workload", name, "--steps", str(steps), "--warmup", "3",
               "--no-cpu-baseline", "--no-secondary", "--no-roofline"]
        t0 = time.perf_counter()
        try:
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=420, env={**os.environ, "WORLD_SIZE": "1", "RANK": "0",
                                                                                     "LOCAL_RANK": os.environ.get("LOCAL_RANK", "0")})
            lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
            rec = json.loads(lines[-1]) if lines else {"error": f"rc {res.returncode}: {res.stderr[-400:]}"}
        except Exception as e:  # noqa: BLE001
            rec = {"error": repr(e)}
        keep = ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "dtype", "e2e", "gpu_launches", "clocks", "model_tflops",
                "config", "error", "final_loss")
        out[name] = {k: rec[k] for k in keep if k in rec}
        out[name]["wall_s"] = round(time.perf_counter() - t0, 1)
    return out


def workload_label(name: str, w: dict, B: int) -> str:
    if w.get("predict"):
        return (f"{name}: sliding-window prediction, batches of {B} windows (100 px + 20 px halo = x[{B},{w['C']},{w['T']},140,140]) per GPU, "
                f"eval-mode BatchNorm, hidden {w['hidden']}; value counts the un-padded 100x100 pixels of every window")
    return (f"{name}: TowerUNet train step (fwd + Tanimoto-complement loss + bwd + all-reduce + AdamW), "
            f"x=[{B},{w['C']},{w['T']},{w['H']},{w['W']}] per GPU, hidden {w['hidden']}, dilations {w['dilations']}, dropout 0")


def run_reference_arm(args, w: dict) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_b = 1 if (w["H"] >= 256 or w.get("predict")) else 2  # cfg 2 at B=2, cfg 4 at B=1, scaled per chip
    sample_b = min(sample_b, w["B"])
    base = cpu_reference_chips_per_s(w, steps=args.steps, warmup=args.warmup, batch=sample_b)
    predict = bool(w.get("predict"))
    line = {
        "impl": "reference", "metric": "inference Mpx/s" if predict else "train chips/s", "value": base["value"], "unit": base["unit"],
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["s_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_label(args.workload, w, w["B"]), "global_batch": w["B"] * args.gpus,
                   "parallelism": f"dp{args.gpus}", "reference_sample_batch": sample_b},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": base["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_predict(args, w: dict) -> None:
    """Secondary metric of BASELINE.json (inference Mpx/s): sliding-window prediction over a synthetic Sentinel-2 tile strip held as
    int16 [T,C,H,W].  A step = one batch of 32 windows through cnb_window_load -> predict_step -> cnb_predict_pack
    (cultionet_b200.tile.TilePredictor), weak-scaled over ranks (every rank owns a strip; windows are independent: no collective on the
    data path).  `value`: strip resident in HBM; `e2e`: strip in pinned host memory, rows copied in ahead of the batches, finished
    uint16 mosaic rows copied back.  Not the driver's default line (that is the training metric)."""
    import torch.distributed as dist

    from cultionet_b200 import _lib
    from cultionet_b200.models.lightning import CultionetLitModel
    from cultionet_b200.parallel import init_distributed
    from cultionet_b200.tile import TilePredictor

    rank, local, world = init_distributed()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, ws, pad = w["B"], 100, 20
    steps, warm = args.steps, max(args.warmup, 3)
    # one window row of the strip = one batch; enough rows for warm-up + timed steps without revisiting a row inside a region
    n_rows = steps + warm
    H, W = n_rows * ws, B * ws
    torch.manual_seed(1234)
    model = CultionetLitModel(in_channels=w["C"], in_time=w["T"], hidden_channels=w["hidden"], dilations=w["dilations"], dropout=0.0,
                              compute_dtype=torch.bfloat16).to(dev)
    g = torch.Generator().manual_seed(100 + rank)
    host_tile = torch.randint(0, 10000, (w["T"], w["C"], H, W), generator=g, dtype=torch.int16).pin_memory()
    mean, std = torch.full((w["C"],), 0.5), torch.full((w["C"],), 0.29)
    tp = TilePredictor(model, host_tile.to(dev), (mean, std), ws, pad, B, cuda_graph=not args.no_graph, rank=0, world_size=1)
    assert tp.num_batches == n_rows
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(ms: float) -> float:
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    for i in range(warm + 3):
        tp.step(i % n_rows)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        tp.step(warm + i)
        flush.zero_()
    e1.record()
    barrier()
    clocks = sampler.stop()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(steps):
        flush.zero_()
    f1.record()
    torch.cuda.synchronize()
    fl = f0.elapsed_time(f1)
    ms_res = reduce_max(e0.elapsed_time(e1)) - fl

    # the two tile kernels alone, CUDA-event timed (HBM roofline; they are ~1 % of a batch)
    def kernel_ms(fn, n=20):
        fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    tp._win.copy_(tp.windows[:B])
    ms_load = kernel_ms(lambda: tp.loader.load_into(tp._win, tp._x))
    pred = {k: torch.rand(B, 1, ws + 2 * pad, ws + 2 * pad, device=dev) for k in ("distance", "edge", "crop")}
    ms_pack = kernel_ms(lambda: tp.writer.write_windows(pred, tp._win, pad))
    peak = float(measured_peaks().get("hbm_gbs") or 6555.0)
    by_load = 6.0 * B * w["C"] * w["T"] * (ws + 2 * pad) ** 2
    by_pack = 18.0 * B * ws * ws

    # end to end: the strip starts in pinned host memory; zero the device copy so that nothing can be read before it arrives
    tp.loader.tile.zero_()
    host_mosaic = torch.empty(tuple(tp.writer._store.shape), dtype=torch.uint16).pin_memory()
    tp.run_streaming(host_tile, host_mosaic, batches=range(0, warm))
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tp.loader.tile.zero_()
    torch.cuda.synchronize()
    a.record()
    tp.run_streaming(host_tile, host_mosaic, batches=range(warm, warm + steps))
    b.record()
    barrier()
    ms_e2e = reduce_max(a.elapsed_time(b))
    h2d = host_tile.numel() * 2 * (steps * ws + pad) / H / steps  # rows [0, need) of the timed batches, averaged per step
    if rank == 0:
        px = B * world * steps * ws * ws
        line = {
            "metric": "inference Mpx/s", "value": px / 1e6 / (ms_res / 1e3), "unit": "Mpx/s", "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": ms_res / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.workload}: sliding-window prediction over a resident int16 tile strip [{w['T']},{w['C']},{H},{W}] per GPU: "
                                   f"batches of {B} windows (100 px + 20 px halo = x[{B},{w['C']},{w['T']},140,140]) through cnb_window_load -> "
                                   f"predict_step (eval-mode BatchNorm, hidden {w['hidden']}) -> cnb_predict_pack into the uint16 mosaic; value "
                                   "counts the un-padded 100x100 pixels of every window",
                       "windows_per_s": B * world * steps / (ms_res / 1e3), "parallelism": f"dp{world} (a strip per rank, no collective)",
                       "l2": "256 MB flush buffer written between timed steps (its time subtracted)",
                       "execution": "one CUDA graph per window batch (load + forward + pack), replayed" if tp.cuda_graph
                                    else "eager launches through the C ABI"},
            "e2e": {"value": px / 1e6 / (ms_e2e / 1e3), "unit": "Mpx/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(3 * ws * tp.writer._pitch * 2), "ms_per_step": ms_e2e / steps,
                    "path": "TilePredictor.run_streaming: int16 tile rows host->device on a side stream one batch ahead, finished "
                            "mosaic rows device->host on a third stream"},
            "gpu_launches": int(tp.launches_per_batch or 0), "clocks": clocks,
            "model_tflops": w["fwd_gflop_per_chip"] * 1e9 * B * world * steps / (ms_res / 1e3) / 1e12,
            "roofline": {"bound": "hbm", "kernel": "cnb_window_load (int16 tile -> fp32 window batch; 6 B per element)",
                         "achieved": by_load / 1e9 / (ms_load / 1e3), "peak": peak, "unit": "GB/s",
                         "frac": by_load / 1e9 / (ms_load / 1e3) / peak, "traffic": None, "avg_launch_ms": ms_load,
                         "note": "the tile kernels are <1 % of a batch; the batch itself is the tensor-bound TowerUNet forward "
                                 "(model_tflops)",
                         "cnb_predict_pack": {"avg_launch_ms": ms_pack, "GBps": by_pack / 1e9 / (ms_pack / 1e3),
                                              "algorithmic_bytes": by_pack}},
        }
        emit(line)
    clean_exit(tp, getattr(tp, "predict", None))


# ---------------------------------------------------------------------------------------------------------------------
def clean_exit(*graph_owners) -> None:
    """Leave through the interpreter's normal exit path (exit hooks run, the loaded libraries stay visible to whoever inspects the
    process at exit).  Captured CUDA graphs hold references to the NCCL communicator's streams: they are released FIRST, then the
    process group is destroyed (destroying it under live graphs blocked for minutes in round 1).  A watchdog bounds a teardown
    that hangs anyway: it runs the registered exit hooks itself and then leaves."""
    import atexit
    import gc

    import torch.distributed as dist

    sys.stdout.flush()
    sys.stderr.flush()

    def bail():  # pragma: no cover - only when a teardown hangs
        try:
            atexit._run_exitfuncs()
        finally:
            os._exit(0)

    t = threading.Timer(float(os.environ.get("CNB_EXIT_WATCHDOG_S", "90")), bail)
    t.daemon = True
    t.start()
    torch.cuda.synchronize()
    for o in graph_owners:
        for attr in ("_graph", "graph", "_graphs"):
            if hasattr(o, attr):
                try:
                    setattr(o, attr, None)
                except Exception:  # noqa: BLE001
                    pass
    gc.collect()
    torch.cuda.synchronize()
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    t.cancel()


_REAL_STDOUT = None


def claim_stdout() -> None:
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to stdout when NCCL_DEBUG is
    set): keep a private duplicate of fd 1 for the result line and point fd 1 at stderr for everything else."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main() -> None:
    claim_stdout()
    args = parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.batch:
        w["B"] = args.batch
    if args.impl == "reference":
        run_reference_arm(args, w)
        return

    if w.get("predict"):
        run_predict(args, w)
        return

    import torch.distributed as dist

    import cultionet_b200 as cb
    from cultionet_b200 import _lib
    from cultionet_b200.engine import TrainStep
    from cultionet_b200.models.lightning import CultionetLitModel
    from cultionet_b200.parallel import init_distributed

    rank, local, world = init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the TowerUNet hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dtype = torch.bfloat16 if w["dtype"] == "bf16" else torch.float32
    B = w["B"]

    if w.get("natten"):  # the reference configures natten through this module-level dict too (unet_parts.py:19-40)
        from cultionet_b200.nn.modules import unet_parts

        for lvl in ("a", "b", "c"):
            unet_parts.NATTEN_PARAMS[lvl].update(w["natten"])
    torch.manual_seed(1234)  # identical replicas
    model = CultionetLitModel(in_channels=w["C"], in_time=w["T"], hidden_channels=w["hidden"], dilations=w["dilations"], dropout=0.0,
                              compute_dtype=dtype).to(dev)
    step = TrainStep(model, total_steps=10_000, cuda_graph=not args.no_graph)
    hx, hy, hb = make_host_batch(w, B, rank, pin=True)
    dbatch = cb.Data(x=hx.to(dev), y=hy.to(dev), bdist=hb.to(dev))
    h2d_bytes = hx.numel() * hx.element_size() + hy.numel() * hy.element_size() + hb.numel() * hb.element_size()
    # L2 hygiene: every step streams > 126 MB of activations (the input batch alone is larger at cfg2); a 256 MB flush buffer is
    # also written between timed steps
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_enqueue_ms = []  # host time to ENQUEUE one step (no synchronisation): below the device time = the GPU is the bottleneck

    def timed(fn, n) -> float:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host = time.perf_counter()
        for _ in range(n):
            fn()
            flush.zero_()
        host_enqueue_ms.append((time.perf_counter() - t_host) * 1e3 / n)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def flush_ms(n) -> float:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            flush.zero_()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    losses = []
    e2e_step_ms: list = []

    def resident_step():
        losses.append(step(dbatch))

    def timed_e2e(n) -> float:
        """n steps fed from pinned HOST memory through the public loop (engine.DevicePrefetcher + TrainStep): every step's x/y/bdist
        are copied host->device inside the region (on a side stream, overlapping the previous step) and every step's loss is read
        back to the host."""
        from cultionet_b200.engine import DevicePrefetcher

        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_prev = time.perf_counter()
        e2e_step_ms.clear()
        for b in DevicePrefetcher((cb.Data(x=hx, y=hy, bdist=hb) for _ in range(n)), dev):
            loss = step(b)
            losses.append(float(loss.cpu()))  # D2H read of the step's result
            flush.zero_()
            t_now = time.perf_counter()
            e2e_step_ms.append(round((t_now - t_prev) * 1e3, 2))  # host clock per step (diagnostic: shows a stalled copy)
            t_prev = t_now
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(args.warmup, 3) + (3 if step.cuda_graph else 0)):
        resident_step()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.launch_count()
    ms_res = timed(resident_step, args.steps)
    launches = (_lib.launch_count() - l0) // max(args.steps, 1)
    if step.cuda_graph and step.launches_per_step:
        launches = step.launches_per_step  # a replay repeats the launches recorded at capture; the host-side counter does not see them
    clocks = sampler.stop()
    timed_e2e(2)
    # the end-to-end region shares the host (PCIe, memory) with whatever else runs on the box: three regions of K steps each are
    # timed back to back, the MEDIAN is reported and all three are listed.
    e2e_regions, e2e_steps_all = [], []
    for _ in range(3):
        e2e_regions.append(timed_e2e(args.steps))
        e2e_steps_all.append(list(e2e_step_ms))
    ms_e2e = statistics.median(e2e_regions)
    e2e_steps_best = e2e_steps_all[e2e_regions.index(ms_e2e)]
    fl = flush_ms(args.steps)
    ms_res -= fl
    ms_e2e -= fl

    chips = B * world * args.steps
    value = chips / (ms_res / 1e3)
    e2e_value = chips / (ms_e2e / 1e3)
    peaks = measured_peaks()

    roofline = None
    if not args.no_roofline:
        # instrumented pass: every rank runs the same steps (the gradient all-reduce is a collective); only rank 0 records
        if rank == 0:
            _lib.TIMER = _lib.KernelTimer()
        nprof = min(args.steps, 3)
        for _ in range(nprof):
            resident_step()
        barrier()
    if not args.no_roofline and rank == 0:
        summ = _lib.TIMER.summary()
        shapes = _lib.TIMER.by_detail()
        _lib.TIMER = None
        conv = {k: v for k, v in summ.items() if k.startswith("conv") and v["flops"] > 0}
        if conv:
            flops = sum(v["flops"] for v in conv.values())
            ms = sum(v["ms"] for v in conv.values())
            calls = sum(v["calls"] for v in conv.values())
            total_ms = sum(v["ms"] for v in summ.values())
            # the dominant kernel = the convolution shape with the largest summed launch time; its roofline is the headline one
            conv_shapes = {k: v for k, v in shapes.items() if k.startswith("conv") and v["flops"] > 0}
            top_key, top = max(conv_shapes.items(), key=lambda kv: kv[1]["ms"])
            achieved = top["flops"] / (top["ms"] / 1e3) / 1e12
            # DRAM traffic cannot be measured outside a profiler: it is read from the committed ncu --set full capture of the SAME
            # kernel and shape (profiles/ncu_traffic.json, regenerated by tools/ncu_summary.py), not from this run
            traffic, traffic_src = None, None
            tfile = ROOT / "profiles" / "ncu_traffic.json"
            if tfile.is_file():
                ent = json.loads(tfile.read_text()).get(top_key, {})
                traffic, traffic_src = ent.get("traffic_bytes"), ent.get("capture")
            roofline = {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                        "frac": achieved / peaks["bf16_tflops"], "traffic": traffic,
                        "traffic_source": (f"static, not this run: {traffic_src}" if traffic is not None else None),
                        "kernel": f"tcgen05 implicit-GEMM convolution, dominant shape: {top_key}",
                        "launches_per_step": top["calls"] // nprof, "avg_launch_ms": top["ms"] / top["calls"],
                        "algorithmic_flops_per_launch": top["flops"] / top["calls"],
                        "share_of_step_kernel_time": top["ms"] / total_ms, "peak_source": peaks["source"] + " (sustained)",
                        "all_conv_launches": {"achieved": flops / (ms / 1e3) / 1e12, "frac": flops / (ms / 1e3) / 1e12 / peaks["bf16_tflops"],
                                              "launches_per_step": calls // nprof, "share_of_step_kernel_time": ms / total_ms},
                        "by_class": {k: {"calls_per_step": v["calls"] // nprof, "ms_per_step": v["ms"] / nprof,
                                         "tflops": v["flops"] / (v["ms"] / 1e3) / 1e12} for k, v in conv.items()},
                        "top_conv_shapes": [
                            {"shape": k, "calls_per_step": v["calls"] // nprof, "ms_per_step": round(v["ms"] / nprof, 4),
                             "tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1) if v["ms"] > 0 else None}
                            for k, v in sorted(shapes.items(), key=lambda kv: -kv[1]["ms"]) if k.startswith("conv")][:40],
                        "other_ms_per_step": {k: v["ms"] / nprof for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])
                                              if not k.startswith("conv")},
                        # the bandwidth-bound kernels against the measured HBM copy rate: algorithmic bytes of all
                        # launches of an entry point / their summed launch time (eager instrumented pass: small launches include
                        # host gaps, so these are lower bounds; tools/bench_bw.py times the level-a shapes alone)
                        "hbm_kernels": {"peak_GBps": peaks["hbm_gbs"],
                                        "by_entry_point": {k: {"ms_per_step": round(v["ms"] / nprof, 4),
                                                               "GB_per_step": round(v["bytes"] / nprof / 1e9, 3),
                                                               "GBps": round(v["bytes"] / 1e9 / (v["ms"] / 1e3), 0),
                                                               "frac": round(v["bytes"] / 1e9 / (v["ms"] / 1e3) / peaks["hbm_gbs"], 3)}
                                                           for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])
                                                           if v["bytes"] > 0 and v["ms"] > 0}}}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_reference_chips_per_s(w, steps=3, warmup=1, batch=1 if w["H"] >= 256 else min(2, B))
        cpu_base = {k: cpu_base[k] for k in ("value", "unit", "cores", "kind", "sample")}

    secondary = None
    if rank == 0 and world == 1 and args.workload == "cfg2" and not args.no_secondary and not args.batch:
        secondary = secondary_runs(args)

    if rank == 0:
        line = {
            "metric": "train chips/s", "value": value, "unit": "chips/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": w["dtype"],
            "data": "synthetic",
            "config": {"workload": workload_label(args.workload, w, B),
                       "global_batch": B * world, "parallelism": f"dp{world}", "l2": "256 MB flush buffer written between timed steps "
                       "(its time subtracted); per-step activations exceed L2",
                       "train_gflop_per_chip": 3 * w["fwd_gflop_per_chip"],
                       "execution": "one CUDA graph per step (captured after 3 eager steps), replayed" if step.cuda_graph
                       else "eager launches through the C ABI"},
            "e2e": {"value": e2e_value, "unit": "chips/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps, "host_ms_each_step": e2e_steps_best,
                    "regions_ms": [round(r, 2) for r in e2e_regions], "pick": "median of 3 regions of K steps (regions_ms before the flush time is subtracted)"},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_enqueue_ms[0] if host_enqueue_ms else None, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_base,
            "model_tflops": 3 * w["fwd_gflop_per_chip"] * 1e9 * value / 1e12 if w["fwd_gflop_per_chip"] else None,
            "final_loss": float(losses[-1]), "secondary": secondary,
        }
        emit(line)
    clean_exit(step)


if __name__ == "__main__":
    main()
