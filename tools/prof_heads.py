"""Driver for an ncu capture of the Psi-Net stream-head convolution of cfg 2: 256 -> 9 (x 9 taps = 81, pitched to 88) 1x1 GEMM +
shift-and-add, forward + backward, B = 32, 128 x 128."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from cultionet_b200 import functional as F

torch.manual_seed(0)
x = torch.randn(32, 128, 128, 256, device="cuda").bfloat16().requires_grad_(True)
w = (torch.randn(9, 256, 3, 3, device="cuda") / 48).requires_grad_(True)
for _ in range(2):
    y = F.conv2d_skinny(x, w, ksize=3, pad=1, dil=1)
    gx, gw = torch.autograd.grad(y, [x, w], torch.ones_like(y))
torch.cuda.synchronize()
print("ok", float(gw.abs().mean()))
