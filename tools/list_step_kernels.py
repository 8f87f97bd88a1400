"""Kernel names of ONE eager training step at BASELINE config 2 (torch.profiler / CUPTI): which launches are ours (namespace cnb / natc /
st / tc) and which are torch's own (at::, elementwise, cat, fill ...).  python tools/list_step_kernels.py [B]"""
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import cultionet_b200 as cb
from cultionet_b200.engine import TrainStep
from cultionet_b200.models.lightning import CultionetLitModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
model = CultionetLitModel(in_channels=5, in_time=24, hidden_channels=64, dilations=[1, 2], dropout=0.0, compute_dtype=torch.bfloat16).cuda()
step = TrainStep(model, total_steps=1000, cuda_graph=False)
batch = cb.Data(x=torch.rand(B, 5, 24, 128, 128).cuda(), y=torch.randint(0, 3, (B, 128, 128)).cuda(), bdist=torch.rand(B, 128, 128).cuda())
for _ in range(3):
    step(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(batch)
    torch.cuda.synchronize()
ours, theirs = collections.Counter(), collections.Counter()
t_ours = t_theirs = 0.0
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA:
        continue
    name = ev.name
    mine = any(s in name for s in ("cnb::", "natc::", "st::", "tc::", "naf::", "cnb_"))
    short = name.split("(")[0][:90]
    if mine:
        ours[short] += 1
        t_ours += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    else:
        theirs[short] += 1
        t_theirs += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
print(json.dumps({"batch": B, "our_launches": sum(ours.values()), "other_launches": sum(theirs.values()), "our_us": round(t_ours, 1),
                  "other_us": round(t_theirs, 1), "other": theirs.most_common(40), "ours_top": ours.most_common(12)}, indent=1))
