"""Per-launch CUDA-event timing of the convolution shapes of config 2 that run below the tensor roofline (narrow N, short K, 1x1,
small levels, strided phases, tiny Psi-Net streams), forward / data gradient / weight gradient separately, with a 256 MB L2 flush
between repetitions.  `python tools/bench_conv_tail.py [reps] [filter]` -> one JSON line per shape."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from cultionet_b200 import _lib
from cultionet_b200 import functional as F

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
flt = sys.argv[2] if len(sys.argv) > 2 else ""
fwd_only = len(sys.argv) > 3 and sys.argv[3] == "fwd"
dev, dt = "cuda", torch.bfloat16
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

# (name, kind, B, H, W, cins, cout, k, stride, pad, bias, want_stats)
SHAPES = [
    ("a64", "conv", 32, 128, 128, [64], 64, 3, 1, 1, False, True),
    ("b128", "conv", 32, 64, 64, [128], 128, 3, 1, 1, False, True),
    ("c256", "conv", 32, 32, 32, [256], 256, 3, 1, 1, False, True),
    ("b256", "conv", 32, 64, 64, [256], 256, 3, 1, 1, False, True),
    ("a256", "conv", 32, 128, 128, [256], 256, 3, 1, 1, False, True),
    ("qkv", "conv", 32, 128, 128, [256], 768, 1, 1, 0, True, False),
    ("proj", "conv", 32, 128, 128, [256], 256, 1, 1, 0, True, False),
    ("a96to256", "conv", 32, 128, 128, [96], 256, 1, 1, 0, False, True),
    ("a256to96", "conv", 32, 128, 128, [256], 96, 1, 1, 0, False, True),
    ("tower1x1", "conv", 32, 128, 128, [64, 128, 256, 256, 256], 256, 1, 1, 0, False, True),
    ("b1x1", "conv", 32, 64, 64, [128, 256, 256, 256, 256], 256, 1, 1, 0, False, True),
    ("pool_a", "conv", 32, 127, 127, [256], 256, 3, 2, 1, False, True),
    ("up_b", "convT", 32, 64, 64, [256], 256, 3, 2, 1, True, False),
    ("up_c", "convT", 32, 32, 32, [256], 256, 3, 2, 1, True, False),
    ("head9to3", "conv", 32, 128, 128, [9], 3, 3, 1, 1, False, False),
    ("head256to3", "conv", 32, 128, 128, [256], 3, 3, 1, 1, False, False),
]


def timed(fn):
    fn()
    torch.cuda.synchronize()
    _lib.TIMER = _lib.KernelTimer()
    for _ in range(reps):
        flush.zero_()
        fn()
    s = _lib.TIMER.by_detail()
    _lib.TIMER = None
    return {k: {"us": round(1e3 * v["ms"] / reps, 1), "calls": v["calls"] // reps,
                **({"TF": round(v["flops"] / v["ms"] / 1e9, 1)} if v["flops"] else {})} for k, v in s.items()}


for name, kind, B, H, W, cins, cout, k, stride, pad, bias, ws in SHAPES:
    if flt and flt != "all" and name not in flt.split(","):
        continue
    torch.manual_seed(0)
    xs = [torch.randn(B, H, W, c, device=dev).to(dt).requires_grad_(True) for c in cins]
    cin = sum(cins)
    shape = (cout, cin, k, k) if kind == "conv" else (cin, cout, k, k)
    w = torch.nn.Parameter(torch.randn(*shape, device=dev) / (cin * k * k) ** 0.5)
    b = torch.nn.Parameter(torch.randn(cout, device=dev)) if bias else None
    w.grad = torch.zeros_like(w)
    if b is not None:
        b.grad = torch.zeros_like(b)

    def fwd():
        if kind == "conv":
            r = F.conv2d(xs, w, b, k, stride, pad, 1, want_stats=ws)
            return r[0] if ws else r
        return F.conv_transpose2d(xs[0], w, b, k, stride, pad, 1)

    y = fwd()
    g = torch.randn_like(y)

    def bwd():
        with F.direct_param_grads():
            torch.autograd.grad(y, xs, g, retain_graph=True)

    def bwd_w():
        with F.direct_param_grads():
            torch.autograd.grad(y, [w] + ([b] if b is not None else []), g, retain_graph=True, allow_unused=True)

    out = {"shape": name, "fwd": timed(fwd)}
    if fwd_only:
        print(json.dumps(out), flush=True)
        continue
    try:
        out["dgrad"] = timed(bwd)
        out["wgrad"] = timed(bwd_w)
    except Exception as e:  # noqa: BLE001
        import traceback
        out["bwd_error"] = traceback.format_exc()[-1500:]
    print(json.dumps(out), flush=True)
