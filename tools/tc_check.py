"""GPU bring-up check of the tcgen05 convolution against the CUDA-core kernel (same bf16 inputs, both through the C ABI).
Prints one line per case and flushes, so a trap or a hang is attributable.  Run under `timeout`."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from cultionet_b200 import functional as F

dev = "cuda"
torch.manual_seed(0)
CASES = [
    # (kind, B, H, W, cins, cout, k, stride, pad, dil, bias)
    ("conv", 1, 16, 16, [64], 64, 1, 1, 0, 1, False),
    ("conv", 1, 16, 16, [64], 64, 3, 1, 1, 1, False),
    ("conv", 2, 32, 32, [256], 256, 3, 1, 1, 1, True),
    ("conv", 2, 32, 32, [256], 768, 1, 1, 0, 1, True),
    ("conv", 2, 25, 25, [64, 128], 192, 3, 1, 1, 1, False),
    ("conv", 2, 13, 13, [256, 512, 256, 256], 256, 3, 1, 1, 1, False),
    ("conv", 2, 50, 50, [128], 128, 3, 1, 2, 2, False),
    ("conv", 4, 128, 128, [64, 128, 256, 256, 256], 256, 3, 1, 1, 1, False),
    # ragged channel counts: partial 64-channel chunks, partial N tiles
    ("conv", 2, 20, 20, [72], 64, 3, 1, 1, 1, True),
    ("conv", 2, 20, 20, [72, 24, 8], 136, 3, 1, 1, 1, True),
    ("conv", 2, 20, 20, [64], 40, 1, 1, 0, 1, True),
    # skinny Psi-Net stream convolution 256 -> 3
    ("conv", 2, 40, 40, [256], 3, 3, 1, 1, 1, False),
    ("conv", 2, 40, 40, [128], 9, 3, 1, 1, 1, True),
    # strided direct (pool convolutions) incl. odd sizes
    ("conv", 2, 32, 32, [64], 128, 3, 2, 1, 1, False),
    ("conv", 2, 25, 25, [128], 256, 3, 2, 1, 1, True),
    ("conv", 1, 13, 13, [64], 64, 3, 2, 1, 1, False),
    # transposed stride 2 / 4 (ConvTranspose2d k3 p1)
    ("convT", 2, 16, 16, [64], 64, 3, 2, 1, 1, True),
    ("convT", 2, 25, 25, [128], 128, 3, 2, 1, 1, True),
    ("convT", 2, 13, 13, [256], 256, 3, 2, 1, 1, False),
    ("convT", 2, 8, 8, [64], 64, 3, 4, 1, 1, True),
    ("convT", 1, 25, 25, [128], 128, 3, 4, 1, 1, True),
    ("convT", 2, 64, 64, [256], 256, 3, 2, 1, 1, True),
]


def run(kind, backend, xs, w, b, k, stride, pad, dil):
    F.CONV_BACKEND = backend
    try:
        if kind == "conv":
            return F.conv2d(xs, w, b, k, stride, pad, dil)
        return F.conv_transpose2d(xs[0], w, b, k, stride, pad, dil)
    finally:
        F.CONV_BACKEND = "auto"


def rel(a, c):
    return float((a.float() - c.float()).norm() / (c.float().norm() + 1e-30))


ok = True
for case in CASES:
    kind, B, H, W, cins, cout, k, stride, pad, dil, bias = case
    xs = [torch.randn(B, H, W, c, device=dev).bfloat16().requires_grad_(True) for c in cins]
    if kind == "conv":
        w = (torch.randn(cout, sum(cins), k, k, device=dev) / (sum(cins) * k * k) ** 0.5).requires_grad_(True)
    else:
        w = (torch.randn(sum(cins), cout, k, k, device=dev) / (sum(cins) * k * k) ** 0.5).requires_grad_(True)
    b = torch.randn(cout, device=dev, requires_grad=True) if bias else None
    print("case", case, end=" ... ", flush=True)
    yg = run(kind, "generic", xs, w, b, k, stride, pad, dil)
    torch.cuda.synchronize()
    yt = run(kind, "tc", xs, w, b, k, stride, pad, dil)
    torch.cuda.synchronize()
    err = rel(yt, yg)
    mx = float((yt.float() - yg.float()).abs().max())
    g = torch.randn_like(yg)
    F.CONV_BACKEND = "generic"
    gg = torch.autograd.grad(yg, xs, g, retain_graph=True)
    F.CONV_BACKEND = "tc"
    gt = torch.autograd.grad(yt, xs, g, retain_graph=True)
    F.CONV_BACKEND = "auto"
    torch.cuda.synchronize()
    gerr = max(rel(a, c) for a, c in zip(gt, gg))
    print(f"fwd rel {err:.2e} max {mx:.2e} | dgrad rel {gerr:.2e}", end=" ", flush=True)
    F.CONV_BACKEND = "generic"
    wg = torch.autograd.grad(yg, w, g)[0]
    torch.cuda.synchronize()
    F.CONV_BACKEND = "tc"
    wt = torch.autograd.grad(yt, w, g)[0]
    F.CONV_BACKEND = "auto"
    torch.cuda.synchronize()
    werr = rel(wt, wg)
    good = err < 2e-3 and gerr < 2e-3 and werr < 2e-3
    ok &= good
    print(f"| wgrad rel {werr:.2e} {'OK' if good else 'MISMATCH'}", flush=True)

# timing of the hot shapes (cfg2): tower_a 960->256 3x3 at 128^2, B=32 and a 256->256 3x3
for (B, H, W, cins, cout) in [(32, 128, 128, [64, 128, 256, 256, 256], 256), (32, 128, 128, [256], 256), (32, 64, 64, [128, 256, 256, 256, 256], 256)]:
    xs = [torch.randn(B, H, W, c, device=dev).bfloat16() for c in cins]
    w = torch.randn(cout, sum(cins), 3, 3, device=dev) / (sum(cins) * 9) ** 0.5
    for backend in ("tc",):
        for _ in range(2):
            run("conv", backend, xs, w, None, 3, 1, 1, 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 5
        for _ in range(n):
            run("conv", backend, xs, w, None, 3, 1, 1, 1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 2.0 * B * H * W * cout * sum(cins) * 9
        print(f"time {backend} B{B} {H}x{W} {sum(cins)}->{cout}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (incl. weight pack)", flush=True)
# wgrad timing on the hot shapes
from cultionet_b200 import _lib
for (B, H, W, cins, cout) in [(32, 128, 128, [64, 128, 256, 256, 256], 256), (32, 128, 128, [256], 256), (32, 32, 32, [256], 256)]:
    xs = [torch.randn(B, H, W, c, device=dev).bfloat16() for c in cins]
    w = (torch.randn(cout, sum(cins), 3, 3, device=dev) / (sum(cins) * 9) ** 0.5).requires_grad_(True)
    y = run("conv", "tc", xs, w, None, 3, 1, 1, 1)
    g = torch.randn_like(y)
    F.CONV_BACKEND = "tc"
    _lib.TIMER = _lib.KernelTimer()
    for _ in range(4):
        torch.autograd.grad(y, w, g, retain_graph=True)
    summ = _lib.TIMER.summary()
    _lib.TIMER = None
    F.CONV_BACKEND = "auto"
    v = summ["conv_wgrad"]
    print(f"time wgrad tc B{B} {H}x{W} {sum(cins)}->{cout}: {v['ms'] / 4:.3f} ms per conv ({v['calls'] // 4} launches) {v['flops'] / v['ms'] / 1e9:.1f} TFLOP/s", flush=True)
# transposed stride-2 and skinny-head timings (fwd + dgrad + wgrad through autograd, per-launch timed)
for (kind, B, H, W, cin, cout, stride) in [("convT", 32, 64, 64, 256, 256, 2), ("convT", 32, 32, 32, 256, 256, 4), ("conv", 32, 128, 128, 256, 3, 1),
                                           ("conv", 32, 128, 128, 64, 128, 2)]:
    x = torch.randn(B, H, W, cin, device=dev).bfloat16().requires_grad_(True)
    shape = (cout, cin, 3, 3) if kind == "conv" else (cin, cout, 3, 3)
    w = (torch.randn(*shape, device=dev) / (cin * 9) ** 0.5).requires_grad_(True)
    for _ in range(2):
        y = run(kind, "auto", [x], w, None, 3, stride, 1, 1)
        torch.autograd.grad(y, [x, w], torch.ones_like(y))
    _lib.TIMER = _lib.KernelTimer()
    for _ in range(3):
        y = run(kind, "auto", [x], w, None, 3, stride, 1, 1)
        torch.autograd.grad(y, [x, w], torch.ones_like(y))
    summ = _lib.TIMER.summary()
    _lib.TIMER = None
    desc = ", ".join(f"{k2}: {v['ms'] / 3:.3f} ms" + (f" {v['flops'] / v['ms'] / 1e9:.0f} TF/s" if v['flops'] else "") for k2, v in summ.items())
    print(f"time {kind} s{stride} B{B} {H}x{W} {cin}->{cout}: {desc}", flush=True)
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
