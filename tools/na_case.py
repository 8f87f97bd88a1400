"""One neighbourhood-attention parity case (for compute-sanitizer / bring-up): python tools/na_case.py heads hd k d H W [bf16]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from tests import cases

heads, hd, k, d, H, W = (int(a) for a in sys.argv[1:7])
dtype = torch.bfloat16 if len(sys.argv) > 7 else torch.float32
cases.na_case("cuda", dtype, 2, H, W, heads, hd, k, d)
torch.cuda.synchronize()
print("ok")
