"""CUDA-event timing of the bandwidth-bound kernels at the cfg2 level-a shape ([32,128,128,256] bf16): achieved GB/s against the
algorithmic bytes of DESIGN.md 4.2.  `python tools/bench_bw.py [reps]`"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from cultionet_b200 import _lib
from cultionet_b200 import functional as F

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev, dt = "cuda", torch.bfloat16
B, H, W, C = 32, 128, 128, 256
unit = B * H * W * C * 2  # bytes of one level-a activation
torch.manual_seed(0)
x = torch.randn(B, H, W, C, device=dev).to(dt).requires_grad_(True)
res = torch.randn(B, H, W, C, device=dev).to(dt)
gam = torch.rand(C, device=dev, requires_grad=True)
bet = torch.rand(C, device=dev, requires_grad=True)
rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
xs = torch.randn(B, H - 1, W - 1, C, device=dev).to(dt).requires_grad_(True)
qkv = torch.randn(B, H, W, 3 * C, device=dev).to(dt).requires_grad_(True)
xin = torch.rand(B, 5, 24, H, W, device=dev)
w1 = torch.randn(5, 5, 3, 1, 1, device=dev, requires_grad=True)
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)


def run():
    y = F.batchnorm_act(x, gam, bet, rm, rv, True, 0.1, 1e-5, True, 1, None)
    torch.autograd.grad(y, [x, gam, bet], torch.ones_like(y))
    y = F.batchnorm_act(x, gam, bet, rm, rv, True, 0.1, 1e-5, True, 1, res)
    y = F.layernorm(x, gam, bet, 1e-5)
    torch.autograd.grad(y, [x, gam, bet], torch.ones_like(y))
    y = F.resize_bilinear(xs, (H, W))
    torch.autograd.grad(y, xs, torch.ones_like(y))
    y = F.na2d(qkv, 4, 3, 2, 0.125)
    torch.autograd.grad(y, qkv, torch.ones_like(y))
    y = F.add_n(x, res, res)
    u = F.pretime_conv(xin, w1, dt)
    torch.autograd.grad(u, w1, torch.ones_like(u))
    flush.zero_()


run()
torch.cuda.synchronize()
_lib.TIMER = _lib.KernelTimer()
for _ in range(reps):
    run()
torch.cuda.synchronize()
summ = _lib.TIMER.summary()
_lib.TIMER = None
# algorithmic bytes per call in units of one level-a activation (268 MB); None = see DESIGN.md
ALG = {"cnb_bn_stats": 1, "cnb_bn_act_fwd": 2.5, "cnb_bn_act_bwd_reduce": 2, "cnb_bn_act_bwd_apply": 3, "cnb_layernorm_fwd": 2,
       "cnb_layernorm_bwd": 3, "cnb_resize_bilinear_fwd": 2, "cnb_resize_bilinear_bwd": 2, "cnb_na2d_fwd": 4, "cnb_na2d_bwd": 8,
       "cnb_add_n": 4}
out = {}
for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
    ms = v["ms"] / v["calls"]
    e = {"calls": v["calls"] // reps, "ms_per_call": round(ms, 4)}
    if k in ALG:
        e["GBps"] = round(ALG[k] * unit / 1e9 / (ms / 1e3), 0)
    out[k] = e
    print(k, e)
# pure-read reference points (library reductions, not ours): is ~4.2 TB/s a property of read-only streams on this part?
def ev_time(fn, n=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        flush.zero_()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


xd = x.detach()
big = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev).normal_()
for name, fn, nbytes in (("torch.sum bf16 268MB", lambda: xd.sum(dtype=torch.float32), xd.numel() * 2),
                         ("torch.sum bf16 2GB", lambda: big.sum(dtype=torch.float32), big.numel() * 2),
                         ("torch.amax bf16 2GB", lambda: big.amax(), big.numel() * 2),
                         ("torch copy bf16 2GB->2GB", lambda: big[: 1 << 29].copy_(big[1 << 29:]), big.numel() * 2)):
    ms = ev_time(fn)
    out[name] = {"ms": round(ms, 4), "GBps": round(nbytes / 1e9 / (ms / 1e3), 0)}
    print(name, out[name])
print(json.dumps(out))
