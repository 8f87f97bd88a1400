"""Tanimoto-complement loss kernels alone, CUDA-event timed, at the BASELINE size (B = 32 of 128x128: 12.6 MB, launch-latency-bound) and at
scaled sizes ("report achieved GB/s at cfg sizes and at a scaled size").  Algorithmic bytes: forward 3 x (4 B prediction) +
8 B labels + 4 B distance target = 24 B per pixel; backward the same reads + 3 x 4 B gradient writes = 36 B per pixel."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from cultionet_b200.losses import tower_unet_loss

dev = "cuda"
out = {}
for B in (32, 512, 2048):
    H = W = 128
    preds = {k: torch.rand(B, 1, H, W, device=dev, requires_grad=True) for k in ("distance", "edge", "crop")}
    y = torch.randint(-1, 3, (B, H, W), device=dev)
    bd = torch.rand(B, H, W, device=dev)

    def fwd():
        return tower_unet_loss(preds, y, bd)[0]

    def t(fn, n=20):
        fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    ms_f = t(fwd)
    loss = fwd()
    ms_b = t(lambda: torch.autograd.grad(loss, list(preds.values()), retain_graph=True))
    px = B * H * W
    out[f"B{B}"] = {"fwd_ms": round(ms_f, 4), "bwd_ms": round(ms_b, 4), "fwd_GBps": round(24 * px / ms_f / 1e6), "bwd_GBps": round(36 * px / ms_b / 1e6),
                    "note": "fwd = memset + sums + finalize launches, bwd = one launch (+ torch bookkeeping)"}
print(json.dumps(out))
