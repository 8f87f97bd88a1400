"""Small driver for ncu captures: a level-a 256->256 3x3 convolution (cfg2 shape) forward + backward, a few iterations."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from cultionet_b200 import functional as F

B, H, W, Cin, Cout = 32, 128, 128, 256, 256
if len(sys.argv) > 1:
    Cin = int(sys.argv[1])
torch.manual_seed(0)
x = torch.randn(B, H, W, Cin, device="cuda").bfloat16().requires_grad_(True)
w = (torch.randn(Cout, Cin, 3, 3, device="cuda") / (Cin * 9) ** 0.5).requires_grad_(True)
for _ in range(3):
    y, sums = F.conv2d([x], w, None, 3, 1, 1, 1, want_stats=True)
    gx, gw = torch.autograd.grad(y, [x, w], torch.ones_like(y))
torch.cuda.synchronize()
print("ok", float(gw.float().abs().mean()))
