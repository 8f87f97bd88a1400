"""Debug: tensor-core NA forward at BASELINE config 4 level shapes against the oracle, sample by sample."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cultionet_b200 import functional as F
from oracle import natten_ref
from tests.util import rel_err
torch.manual_seed(0)
for (B, H, W, heads, hd, k, d) in [(16, 256, 256, 4, 64, 7, 2), (16, 128, 128, 4, 64, 7, 2), (16, 64, 64, 8, 32, 7, 2)]:
    Cn = heads * hd
    qkv = torch.randn(B, H, W, 3 * Cn, device="cuda").bfloat16()
    y = F.na2d(qkv, heads, k, d, hd ** -0.5)
    torch.cuda.synchronize()
    errs = []
    for b in range(B):
        t = qkv[b:b + 1].float().reshape(1, H, W, 3, heads, hd).permute(3, 0, 4, 1, 2, 5)
        a = natten_ref.na2d_qk(t[0] * hd ** -0.5, t[1], k, d).softmax(-1)
        yr = natten_ref.na2d_av(a, t[2], k, d).permute(0, 2, 3, 1, 4).reshape(1, H, W, Cn)
        errs.append(round(rel_err(y[b:b + 1].float(), yr), 5))
        del a, yr, t
    print((B, H, W, heads, hd, k, d), errs, flush=True)
