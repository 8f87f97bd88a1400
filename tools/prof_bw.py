"""Driver for ncu captures of the bandwidth-bound kernels at the cfg2 level-a shape ([32,128,128,256] bf16) and of the loss kernels at a
scaled size (B = 512: at the BASELINE size they move 13 MB and are launch-latency-bound).  The second pass runs between
cudaProfilerStart/Stop: `ncu --profile-from-start off --set full --clock-control none -o X python tools/prof_bw.py`."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from cultionet_b200 import functional as F

dev, dt = "cuda", torch.bfloat16
B, H, W, C = 32, 128, 128, 256
torch.manual_seed(0)
x = torch.randn(B, H, W, C, device=dev).to(dt).requires_grad_(True)
res = torch.randn(B, H, W, C, device=dev).to(dt)
gam = torch.rand(C, device=dev, requires_grad=True)
bet = torch.rand(C, device=dev, requires_grad=True)
rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
from cultionet_b200.losses import tower_unet_loss

for it in range(2):
    if it == 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    y = F.batchnorm_act(x, gam, bet, rm, rv, True, 0.1, 1e-5, True, 1, None)
    torch.autograd.grad(y, [x, gam, bet], torch.ones_like(y))
    y = F.layernorm(x, gam, bet, 1e-5)
    torch.autograd.grad(y, [x, gam, bet], torch.ones_like(y))
    xs = torch.randn(B, H - 1, W - 1, C, device=dev).to(dt).requires_grad_(True)
    y = F.resize_bilinear(xs, (H, W))
    torch.autograd.grad(y, xs, torch.ones_like(y))
    qkv = torch.randn(B, H, W, 3 * C, device=dev).to(dt).requires_grad_(True)
    y = F.na2d(qkv, 4, 3, 2, 0.125)
    torch.autograd.grad(y, qkv, torch.ones_like(y))
    y = F.add_n(x, res, res)
    xin = torch.rand(B, 5, 24, H, W, device=dev)
    w1 = torch.randn(5, 5, 3, 1, 1, device=dev, requires_grad=True)
    u = F.pretime_conv(xin, w1, dt)
    torch.autograd.grad(u, w1, torch.ones_like(u))
    preds = {k: torch.rand(512, 1, H, W, device=dev, requires_grad=True) for k in ("distance", "edge", "crop")}
    loss, _ = tower_unet_loss(preds, torch.randint(-1, 3, (512, H, W), device=dev), torch.rand(512, H, W, device=dev))
    loss.backward()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
