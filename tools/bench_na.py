"""Neighbourhood-attention kernels alone at the BASELINE shapes, CUDA-event timed: python tools/bench_na.py [cfg2|cfg4]
(cfg 2: levels c / b / a = k3 d1 32^2 h8 hd32, k3 d1 64^2, k3 d2 128^2 at B = 32; cfg 4: k7 d2 at 64^2 / 128^2 / 256^2, B = 16)."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from cultionet_b200 import functional as F

which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
shapes = {"cfg2": [(32, 32, 32, 8, 32, 3, 1), (32, 64, 64, 4, 64, 3, 1), (32, 128, 128, 4, 64, 3, 2)],
          "cfg4": [(16, 64, 64, 8, 32, 7, 2), (16, 128, 128, 4, 64, 7, 2), (16, 256, 256, 4, 64, 7, 2)]}[which]
out = {}
for B, H, W, heads, hd, k, d in shapes:
    torch.manual_seed(0)
    qkv = torch.randn(B, H, W, 3 * heads * hd, device="cuda").bfloat16().requires_grad_(True)
    g = torch.randn(B, H, W, heads * hd, device="cuda").bfloat16()
    def fwd():
        return F.na2d(qkv, heads, k, d, hd ** -0.5)
    y = fwd()
    torch.autograd.grad(y, qkv, g)
    def t(fn, n=5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    ms_f = t(fwd)
    y = fwd()
    ms_b = t(lambda: torch.autograd.grad(y, qkv, g, retain_graph=True))
    alg = 4 * B * H * W * heads * hd * 2
    out[f"B{B} {H}x{W} h{heads} hd{hd} k{k} d{d}"] = {"fwd_ms": round(ms_f, 4), "bwd_ms": round(ms_b, 4), "fwd_GBps": round(alg / ms_f / 1e6),
                                                      "bwd_GBps": round(2 * alg / ms_b / 1e6)}
print(json.dumps(out))
