"""Bandwidth kernels at the three pixel levels of config 2 (B 32: 128^2, 64^2, 32^2 pixels x 256 channels bf16): microseconds and algorithmic
GB/s per launch with an L2 flush between launches -- how much of the HBM rate the SMALL tensors of a step reach."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from cultionet_b200 import _lib
from cultionet_b200 import functional as F

dev, dt = "cuda", torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
reps = 5
out = {}
for name, H in (("a", 128), ("b", 64), ("c", 32)):
    B, C = 32, 256
    x = torch.randn(B, H, H, C, device=dev).to(dt).requires_grad_(True)
    res = torch.randn(B, H, H, C, device=dev).to(dt)
    gam = torch.rand(C, device=dev, requires_grad=True)
    bet = torch.rand(C, device=dev, requires_grad=True)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    g = torch.randn(B, H, H, C, device=dev).to(dt)

    def run():
        y = F.batchnorm_act(x, gam, bet, rm, rv, True, 0.1, 1e-5, True, 1, None)
        flush.zero_()
        torch.autograd.grad(y, [x, gam, bet], g)
        flush.zero_()
        F.add_n(x, res, res)
        flush.zero_()
        y = F.layernorm(x, gam, bet, 1e-5)
        flush.zero_()
        torch.autograd.grad(y, [x, gam, bet], g)
        flush.zero_()

    run()
    torch.cuda.synchronize()
    _lib.TIMER = _lib.KernelTimer()
    for _ in range(reps):
        run()
    s = _lib.TIMER.summary()
    _lib.TIMER = None
    out[name] = {k: {"us": round(1e3 * v["ms"] / v["calls"], 1), "GBps": round(v["bytes"] / v["ms"] / 1e6) if v["bytes"] else None} for k, v in s.items()}
    print(name, json.dumps(out[name]), flush=True)
