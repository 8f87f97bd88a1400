"""Driver for ncu captures of the dominant convolution of BASELINE config 2: tower_a's 960 -> 256 3x3 convolution over the five
sources of the UNet3+ full-scale skip (64 + 128 + 256 + 256 + 256 channels, B = 32, 128 x 128), forward + merged data gradient +
weight gradients, plus the BatchNorm statistics / backward-reduce streams of the same tensor."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from cultionet_b200 import functional as F

B, H, W, Cout = 32, 128, 128, 256
chans = [64, 128, 256, 256, 256]
torch.manual_seed(0)
xs = [torch.randn(B, H, W, c, device="cuda").bfloat16().requires_grad_(True) for c in chans]
w = (torch.randn(Cout, sum(chans), 3, 3, device="cuda") / (sum(chans) * 9) ** 0.5).requires_grad_(True)
for _ in range(2):
    y = F.conv2d(xs, w, None, 3, 1, 1, 1)
    grads = torch.autograd.grad(y, [*xs, w], torch.ones_like(y))
torch.cuda.synchronize()
print("ok", float(grads[-1].float().abs().mean()))
