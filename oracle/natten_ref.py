"""TEST INFRASTRUCTURE ONLY -- CPU restatement of natten==0.17.1 NeighborhoodAttention2D.

The reference (jgrss/cultionet) calls ``natten.NeighborhoodAttention2D`` at
``src/cultionet/nn/modules/convolution.py:341-350`` with ``rel_pos_bias=False``,
``qkv_bias=True``.  natten is a third-party dependency pinned only in
``README.md:247`` / ``.github/workflows/ci.yml:53`` (``natten==0.17.1``); it is
not vendored under /root/reference and not installed in this image.  This file
restates its published algorithm (Hassani et al., "Neighborhood Attention
Transformer" / "Dilated Neighborhood Attention Transformer"):

* ``qkv = Linear(C, 3C)``; reshape ``[B,H,W,3,heads,hd]``; ``q *= hd**-0.5``
* pixel (i, j) attends over a k x k window of its own dilation group; the window
  is *clamped* inside the group (no zero padding), see ``window_start``
* softmax over the k*k logits, weighted sum of V, merge heads, ``proj = Linear(C, C)``

PARITY UNPINNED at this boundary: no reference test pins NA values
(``tests/test_tower_unet.py`` asserts shapes only) and the real natten binary is
unavailable, so CUDA-vs-oracle parity for NA means "CUDA kernel vs this restatement".

Nothing in the product package imports this file.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def window_start(index: int, length: int, kernel_size: int, dilation: int) -> int:
    """First neighbour index along one axis (NATTEN 0.17 ``get_window_start`` semantics).

    With g = index mod d, p = index div d and L_g = ceil((length - g) / d) the
    dilation group {g, g+d, ...} has L_g members; the window of k members is centred
    on p and clamped to [0, L_g - k].
    """
    nb = kernel_size // 2
    g = index % dilation
    p = index // dilation
    group_len = (length - g + dilation - 1) // dilation
    s = min(max(p - nb, 0), group_len - kernel_size)
    return g + dilation * s


def natten_get_window_start(index: int, length: int, kernel_size: int, dilation: int) -> int:
    """The branchy form natten 0.17 ships in its naive kernels (restated from the paper's
    reference implementation); used only to cross-check ``window_start`` in tests."""
    nb = kernel_size // 2
    if dilation <= 1:
        start = max(index - nb, 0)
        if index + nb >= length:
            start += length - index - nb - 1
        return start
    ni = index - nb * dilation
    if ni < 0:
        return index % dilation
    if index + nb * dilation >= length:
        imodd = index % dilation
        a = (length // dilation) * dilation
        b = length - a
        if imodd < b:
            return length - b + imodd - 2 * nb * dilation
        return a + imodd - kernel_size * dilation
    return ni


_GATHER_BYTES = 6 << 30  # the gathered neighbour tensor [B, heads, H, k, W, k, hd] is materialised: process sample by sample above this


def neighbor_index(length: int, kernel_size: int, dilation: int) -> torch.Tensor:
    """[length, k] long tensor: neighbour coordinates along one axis."""
    assert kernel_size * dilation <= length, "natten requires kernel_size * dilation <= min(H, W)"
    idx = torch.empty(length, kernel_size, dtype=torch.long)
    for i in range(length):
        s = window_start(i, length, kernel_size, dilation)
        for a in range(kernel_size):
            idx[i, a] = s + a * dilation
    return idx


def na2d_qk(q: torch.Tensor, k: torch.Tensor, kernel_size: int, dilation: int) -> torch.Tensor:
    """q, k: [B, heads, H, W, hd] -> logits [B, heads, H, W, k*k]."""
    B, nh, H, W, hd = q.shape
    if B > 1 and B * nh * H * W * kernel_size * kernel_size * hd * 4 > _GATHER_BYTES:  # samples are independent: bound the gather
        return torch.cat([na2d_qk(q[b:b + 1], k[b:b + 1], kernel_size, dilation) for b in range(B)], dim=0)
    iy = neighbor_index(H, kernel_size, dilation).to(q.device)  # [H, k]
    ix = neighbor_index(W, kernel_size, dilation).to(q.device)  # [W, k]
    # gather neighbours: [B, nh, H, k, W, k, hd]
    kk = k[:, :, iy][:, :, :, :, ix]  # [B, nh, H, k, W, k, hd]
    kk = kk.permute(0, 1, 2, 4, 3, 5, 6)  # [B, nh, H, W, k, k, hd]
    attn = torch.einsum("bnhwd,bnhwxyd->bnhwxy", q, kk)
    return attn.reshape(B, nh, H, W, kernel_size * kernel_size)


def na2d_av(attn: torch.Tensor, v: torch.Tensor, kernel_size: int, dilation: int) -> torch.Tensor:
    """attn: [B, heads, H, W, k*k], v: [B, heads, H, W, hd] -> [B, heads, H, W, hd]."""
    B, nh, H, W, hd = v.shape
    if B > 1 and B * nh * H * W * kernel_size * kernel_size * hd * 4 > _GATHER_BYTES:
        return torch.cat([na2d_av(attn[b:b + 1], v[b:b + 1], kernel_size, dilation) for b in range(B)], dim=0)
    iy = neighbor_index(H, kernel_size, dilation).to(v.device)
    ix = neighbor_index(W, kernel_size, dilation).to(v.device)
    vv = v[:, :, iy][:, :, :, :, ix].permute(0, 1, 2, 4, 3, 5, 6)
    a = attn.reshape(B, nh, H, W, kernel_size, kernel_size)
    return torch.einsum("bnhwxy,bnhwxyd->bnhwd", a, vv)


def na2d(q, k, v, kernel_size: int, dilation: int = 1, scale=None):
    """Fused-signature variant over [B, H, W, heads, hd] tensors (natten.functional.na2d)."""
    hd = q.shape[-1]
    scale = scale if scale is not None else hd**-0.5
    qh, kh, vh = (t.permute(0, 3, 1, 2, 4) for t in (q, k, v))
    attn = na2d_qk(qh * scale, kh, kernel_size, dilation).softmax(dim=-1)
    out = na2d_av(attn, vh, kernel_size, dilation)
    return out.permute(0, 2, 3, 1, 4)


class NeighborhoodAttention2D(nn.Module):
    """Module-level restatement; parameter names ``qkv`` / ``proj`` as in natten 0.17.1."""

    def __init__(
        self,
        dim: int,
        num_heads: int,
        kernel_size: int,
        dilation: int = 1,
        is_causal: bool = False,
        rel_pos_bias: bool = False,
        qkv_bias: bool = True,
        qk_scale=None,
        attn_drop: float = 0.0,
        proj_drop: float = 0.0,
    ):
        super().__init__()
        assert not rel_pos_bias and not is_causal, "only the reference's configuration is restated"
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = qk_scale or self.head_dim**-0.5
        self.kernel_size = kernel_size
        self.dilation = dilation
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, H, W, C = x.shape
        qkv = self.qkv(x).reshape(B, H, W, 3, self.num_heads, self.head_dim).permute(3, 0, 4, 1, 2, 5)
        q, k, v = qkv[0], qkv[1], qkv[2]
        q = q * self.scale
        attn = na2d_qk(q, k, self.kernel_size, self.dilation)
        attn = self.attn_drop(attn.softmax(dim=-1))
        x = na2d_av(attn, v, self.kernel_size, self.dilation)
        x = x.permute(0, 2, 3, 1, 4).reshape(B, H, W, C)
        return self.proj_drop(self.proj(x))
