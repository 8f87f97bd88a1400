"""Driver for ncu captures of the tensor-core neighbourhood attention kernels (k_na_tc.cuh) at BASELINE config 4's level-a shape:
B 16, 256 x 256, 4 heads x 64, kernel 7, dilation 2 -- forward, query-side backward, key-side backward."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from cultionet_b200 import functional as F

B, H, W, heads, hd, k, d = 16, 256, 256, 4, 64, 7, 2
if len(sys.argv) > 1 and sys.argv[1] == "b":
    B, H, W = 16, 128, 128
torch.manual_seed(0)
qkv = torch.randn(B, H, W, 3 * heads * hd, device="cuda").bfloat16().requires_grad_(True)
g = torch.randn(B, H, W, heads * hd, device="cuda").bfloat16()
for _ in range(2):
    y = F.na2d(qkv, heads, k, d, hd ** -0.5)
    (dq,) = torch.autograd.grad(y, qkv, g)
torch.cuda.synchronize()
print("ok", float(dq.float().abs().mean()))
