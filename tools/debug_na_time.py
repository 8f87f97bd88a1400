import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cultionet_b200 import functional as F
for (B, H, W, heads, hd, k, d) in [(16, 64, 64, 8, 32, 7, 2), (16, 128, 128, 4, 64, 7, 2), (16, 64, 64, 8, 32, 7, 2)]:
    torch.manual_seed(0)
    qkv = torch.randn(B, H, W, 3 * heads * hd, device="cuda").bfloat16().requires_grad_(True)
    g = torch.randn(B, H, W, heads * hd, device="cuda").bfloat16()
    y = F.na2d(qkv, heads, k, d, hd ** -0.5)
    times = []
    for i in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        torch.autograd.grad(y, qkv, g, retain_graph=True)
        b.record()
        torch.cuda.synchronize()
        times.append(round(a.elapsed_time(b), 3))
    print((B, H, W, heads, hd, k, d), times, flush=True)
