"""Per-entry-point CUDA-event timing of ONE eval-mode forward of config 5's window batch (x = [32, 5, 12, 140, 140], hidden 64, bf16):
where the 21.8 ms of a prediction batch go.  `python tools/prof_predict.py [batch]`"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import cultionet_b200 as cb
from cultionet_b200 import _lib
from cultionet_b200.models.lightning import CultionetLitModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda")
torch.manual_seed(0)
model = CultionetLitModel(in_channels=5, in_time=12, hidden_channels=64, dilations=[1, 2], dropout=0.0, compute_dtype=torch.bfloat16).to(dev).eval()
batch = cb.Data(x=torch.rand(B, 5, 12, 140, 140, device=dev))
with torch.no_grad():
    for _ in range(2):
        model.predict_step(batch, 0)
    torch.cuda.synchronize()
    _lib.TIMER = _lib.KernelTimer()
    reps = 3
    for _ in range(reps):
        model.predict_step(batch, 0)
    by = _lib.TIMER.summary()
    det = _lib.TIMER.by_detail()
    _lib.TIMER = None
tot = sum(v["ms"] for v in by.values()) / reps
print(f"batch {B}: {tot:.2f} ms summed over entry points")
for k, v in sorted(by.items(), key=lambda kv: -kv[1]["ms"]):
    tf = f" {v['flops'] / v['ms'] / 1e9:7.0f} TFLOP/s" if v["flops"] else ""
    gb = f" {v['bytes'] / v['ms'] / 1e6:7.0f} GB/s" if v["bytes"] else ""
    print(f"{v['ms'] / reps:8.3f} ms x{v['calls'] // reps:3d}  {k}{tf}{gb}")
print("-- convolution shapes")
for k, v in sorted(det.items(), key=lambda kv: -kv[1]["ms"])[:24]:
    if "conv" in k:
        print(f"{v['ms'] / reps:8.3f} ms x{v['calls'] // reps:3d} {v['flops'] / v['ms'] / 1e9:7.0f} TFLOP/s  {k}")
