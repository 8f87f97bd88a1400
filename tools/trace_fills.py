"""Which Python lines still create zero-filled tensors (torch fill kernels) inside one training step.  python tools/trace_fills.py"""
import collections
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cultionet_b200 as cb
from cultionet_b200.engine import TrainStep
from cultionet_b200.models.lightning import CultionetLitModel

dev = "cuda"
torch.manual_seed(0)
model = CultionetLitModel(in_channels=5, in_time=24, hidden_channels=64, dropout=0.0, compute_dtype=torch.bfloat16).to(dev)
step = TrainStep(model, total_steps=100)
batch = cb.Data(x=torch.rand(2, 5, 24, 128, 128).to(dev), y=torch.randint(0, 3, (2, 128, 128)).to(dev), bdist=torch.rand(2, 128, 128).to(dev))
step(batch)
step(batch)
cnt = collections.Counter()


def wrap(name, mod):
    f = getattr(mod, name)

    def g(*a, **k):
        fr = traceback.extract_stack(limit=3)[-2]
        cnt[(name, os.path.basename(fr.filename), fr.lineno)] += 1
        return f(*a, **k)

    setattr(mod, name, g)


for n in ("zeros", "zeros_like", "ones", "full", "cat"):
    wrap(n, torch)
oz = torch.Tensor.zero_


def z(self):
    fr = traceback.extract_stack(limit=3)[-2]
    cnt[("zero_", os.path.basename(fr.filename), fr.lineno)] += 1
    return oz(self)


torch.Tensor.zero_ = z
step(batch)
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1]):
    print(v, k)
