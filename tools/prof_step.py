"""One eager cfg2 training step between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv`:
the launch list of exactly one steady-state step (two un-profiled warm-up steps first)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import cultionet_b200 as cb
from cultionet_b200.engine import TrainStep
from cultionet_b200.models.lightning import CultionetLitModel

B, C, T, H, W = 32, 5, 24, 128, 128
if len(sys.argv) > 1:
    B = int(sys.argv[1])
torch.manual_seed(1234)
dev = torch.device("cuda")
model = CultionetLitModel(in_channels=C, in_time=T, hidden_channels=64, dilations=[1, 2], dropout=0.0, compute_dtype=torch.bfloat16).to(dev)
step = TrainStep(model, total_steps=100, cuda_graph=False)
batch = cb.Data(x=torch.rand(B, C, T, H, W, device=dev), y=torch.randint(0, 3, (B, H, W), device=dev), bdist=torch.rand(B, H, W, device=dev))
for _ in range(2):
    loss = step(batch)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
loss = step(batch)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", float(loss))
