"""``Data`` batch container: the field contract of ``src/cultionet/data/data.py:51-139`` (x, y, bdist, lon, lat, ...)
without the geospatial helpers, which are outside the hot path."""
from __future__ import annotations

from copy import deepcopy
from typing import Optional

import numpy as np
import torch


class Data:
    def __init__(self, x: torch.Tensor, y: Optional[torch.Tensor] = None, **kwargs):
        self.x = x
        self.y = y
        for k, v in kwargs.items():
            if v is not None:
                assert isinstance(v, (torch.Tensor, np.ndarray, list)), "Only tensors, arrays, and lists are supported."
            setattr(self, k, v)

    def _keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def to_dict(self, device=None, dtype=None) -> dict:
        out = {}
        for k in self._keys():
            v = getattr(self, k)
            if isinstance(v, torch.Tensor):
                v = v.clone()
                if device is not None:
                    v = v.to(device=device, dtype=dtype)
            elif isinstance(v, np.ndarray):
                v = v.copy()
            elif v is not None:
                v = deepcopy(v)
            out[k] = v
        return out

    def to(self, device=None, dtype=None) -> "Data":
        return Data(**self.to_dict(device=device, dtype=dtype))

    def copy(self) -> "Data":
        return Data(**self.to_dict())

    @property
    def num_samples(self) -> int:
        return self.x.shape[0]

    @property
    def num_channels(self) -> int:
        return self.x.shape[1]

    @property
    def num_time(self) -> int:
        return self.x.shape[2]

    @property
    def height(self) -> int:
        return self.x.shape[3]

    @property
    def width(self) -> int:
        return self.x.shape[4]

    def __str__(self) -> str:
        parts = [f"x={tuple(self.x.shape)}"]
        if self.y is not None:
            parts.append(f"y={tuple(self.y.shape)}")
        for k in self._keys():
            v = getattr(self, k)
            if k not in ("x", "y") and isinstance(v, (torch.Tensor, np.ndarray)):
                parts.append(f"{k}={tuple(v.shape)}")
        return "Data(" + ", ".join(parts) + ")"

    __repr__ = __str__
