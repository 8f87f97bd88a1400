"""Neighbourhood attention module with natten 0.17.1's parameter names (``qkv``, ``proj``), as configured by the
reference at ``src/cultionet/nn/modules/convolution.py:341-350`` (``rel_pos_bias=False``, ``qkv_bias=True``)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import functional as F


class NeighborhoodAttention2D(nn.Module):
    def __init__(self, dim: int, num_heads: int, kernel_size: int, dilation: int = 1, rel_pos_bias: bool = False,
                 qkv_bias: bool = True, qk_scale=None, attn_drop: float = 0.0, proj_drop: float = 0.0):
        super().__init__()
        if rel_pos_bias:
            raise NotImplementedError("cultionet_b200: rel_pos_bias=True is not used by the reference and is not built")
        assert dim % num_heads == 0
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = qk_scale or self.head_dim ** -0.5
        self.kernel_size = kernel_size
        self.dilation = dilation
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self._rng_sites = (F.new_rng_site(), F.new_rng_site())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: [B, H, W, C] (already pixel-major, as natten expects)."""
        drop = self.attn_drop.p if self.training else 0.0
        qkv = F.linear(x, self.qkv.weight, self.qkv.bias)
        o = F.na2d(qkv, self.num_heads, self.kernel_size, self.dilation, float(self.scale), attn_drop=drop, site=self._rng_sites[0])
        o = F.linear(o, self.proj.weight, self.proj.bias)
        if self.training and self.proj_drop.p > 0:
            o = F.dropout(o, self.proj_drop.p, self._rng_sites[1])
        return o
