"""TowerUNet -- the reference's ``nn.Module`` surface (``src/cultionet/models/nunet.py:108-265``) over sm_100a kernels.

Same constructor arguments, same ``forward(x[B,C,T,H,W], latlon_coords=None) -> {"distance","edge","crop"}`` (each
``[B,1,H,W]`` float32), same parameter names and shapes.  Inside, activations are pixel-major ``[B,H,W,C]`` in
``compute_dtype`` (float32 = parity mode, bfloat16 = throughput mode) and every operator is a hand-written kernel reached
through the C ABI; there is no PyTorch/cuDNN fallback.
"""
from __future__ import annotations

import typing as T

import torch
import torch.nn as nn

from .. import functional as F
from .. import nn as cunn
from ..enums import AttentionTypes, InferenceNames, ResBlockTypes
from ..layers.weights import init_conv_weights
from ..nn.modules.convolution import batchnorm_act


class Conv3d(nn.Module):
    """Time-reducing stack of ``PreTimeReduction`` (reference ``nunet.py:18-57``):
    Conv3d(C->C,(k,1,1)) -> BN3d -> SiLU -> Conv3d(C->hid,(T-k+1,1,1)) -> BN2d -> SiLU."""

    def __init__(self, in_channels: int, in_time: int, out_channels: int, kernel_size: int, activation_type: str):
        super().__init__()
        from ..nn.modules.convolution import _act_module

        self.act = F.act_code(activation_type)
        remaining_time = in_time - kernel_size + 1
        self.remaining_time = remaining_time
        self.seq = nn.Sequential(
            nn.Conv3d(in_channels, in_channels, kernel_size=(kernel_size, 1, 1), padding=0, bias=False),
            nn.BatchNorm3d(in_channels),
            _act_module(activation_type),
            nn.Conv3d(in_channels, out_channels, kernel_size=(remaining_time, 1, 1), padding=0, bias=False),
            nn.Identity(),  # einops Rearrange('b c 1 h w -> b c h w') in the reference
            nn.BatchNorm2d(out_channels),
            _act_module(activation_type),
        )

    def forward(self, x: torch.Tensor, dtype: torch.dtype, xp: T.Optional[torch.Tensor] = None) -> torch.Tensor:
        conv1, bn1, conv2, bn2 = self.seq[0], self.seq[1], self.seq[3], self.seq[5]
        # u: [B,H,W,pitch >= C*T'], column = c*T' + t', zero row padding
        if xp is not None:  # throughput mode: banded GEMM over the pixel-major copy of x (tensor cores)
            u = F.pretime_conv_gemm(xp, conv1.weight, x.shape[2])
        else:
            u = F.pretime_conv(x, conv1.weight, dtype)
        a = batchnorm_act(bn1, u, act=self.act, ch_div=self.remaining_time)
        w2 = F.tag_derived(conv2.weight.view(conv2.weight.shape[0], -1), conv2.weight, "flat")
        v = F.linear(a, w2, None, in_features=w2.shape[1])
        return batchnorm_act(bn2, v, act=self.act)


class PreTimeReduction(nn.Module):
    """Two temporal stacks (k=3, k=5) -> sum -> LayerNorm over channels (reference ``nunet.py:60-105``)."""

    def __init__(self, in_channels: int, in_time: int, out_channels: int, activation_type: str):
        super().__init__()
        self.conv3 = Conv3d(in_channels, in_time, out_channels, kernel_size=3, activation_type=activation_type)
        self.conv5 = Conv3d(in_channels, in_time, out_channels, kernel_size=5, activation_type=activation_type)
        self.layer_norm = nn.Sequential(nn.Identity(), nn.LayerNorm(out_channels), nn.Identity())

    def forward(self, x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
        ln = self.layer_norm[1]
        # bf16: ONE pixel-major copy of x feeds both temporal branches, each a 1x1 GEMM; fp32 parity mode keeps the direct kernels
        xp = F.time_to_pixel_major(x, dtype) if dtype == torch.bfloat16 else None
        s = F.add_n(self.conv3(x, dtype, xp), self.conv5(x, dtype, xp))
        return F.layernorm(s, ln.weight, ln.bias, ln.eps)


class TowerUNet(nn.Module):
    """Tower U-Net."""

    def __init__(
        self,
        in_channels: int,
        in_time: int,
        hidden_channels: int = 64,
        num_classes: int = 1,
        dilations: T.Optional[T.Sequence[int]] = None,
        activation_type: str = "SiLU",
        dropout: float = 0.0,
        res_block_type: str = ResBlockTypes.RESA,
        attention_weights: str = AttentionTypes.NATTEN,
        pool_by_max: bool = False,
        batchnorm_first: bool = False,
        edge_activation: bool = True,
        mask_activation: bool = True,
        use_latlon: bool = False,
        compute_dtype: torch.dtype = torch.float32,
    ):
        super().__init__()
        if dilations is None:
            dilations = [1, 2]
        channels = [hidden_channels, hidden_channels * 2, hidden_channels * 4, hidden_channels * 8]
        up_channels = int(hidden_channels * len(channels))
        self.in_channels, self.in_time = in_channels, in_time
        self.compute_dtype = compute_dtype
        self.dropout = float(dropout)

        self.pre_unet = PreTimeReduction(in_channels, in_time, channels[0], activation_type)
        self.encoder = cunn.TowerUNetEncoder(channels=channels, dilations=dilations, activation_type=activation_type, dropout=dropout,
                                             res_block_type=res_block_type, attention_weights=None, pool_by_max=pool_by_max,
                                             batchnorm_first=batchnorm_first)
        self.decoder = cunn.TowerUNetDecoder(channels=channels, up_channels=up_channels, dilations=dilations,
                                             activation_type=activation_type, dropout=dropout, res_block_type=res_block_type,
                                             attention_weights=attention_weights, batchnorm_first=batchnorm_first)
        self.tower_fusion = cunn.TowerUNetFusion(channels=channels, up_channels=up_channels, dilations=dilations,
                                                 activation_type=activation_type, dropout=dropout, res_block_type=res_block_type,
                                                 attention_weights=None, batchnorm_first=batchnorm_first, use_latlon=use_latlon)
        self.final_a = cunn.TowerUNetFinal(up_channels, num_classes, activation_type=activation_type)
        self.final_b = cunn.TowerUNetFinal(up_channels, num_classes, activation_type=activation_type, resample_factor=2)
        self.final_c = cunn.TowerUNetFinal(up_channels, num_classes, activation_type=activation_type, resample_factor=4)
        self.final_combine = cunn.TowerUNetFinalCombine(num_classes=num_classes, edge_activation=edge_activation,
                                                        mask_activation=mask_activation)
        self.apply(init_conv_weights)
        # the reference wraps pre_unet in torch.compile, so its checkpoints may carry `pre_unet._orig_mod.` keys
        self._register_load_state_dict_pre_hook(self._strip_compile_prefix)

    @staticmethod
    def _strip_compile_prefix(state_dict, prefix, *args):
        for k in list(state_dict.keys()):
            if k.startswith(prefix) and "._orig_mod." in k:
                state_dict[k.replace("._orig_mod.", ".")] = state_dict.pop(k)

    def set_compute_dtype(self, dtype: torch.dtype) -> "TowerUNet":
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("compute_dtype must be torch.float32 or torch.bfloat16")
        self.compute_dtype = dtype
        return self

    def forward(self, x: torch.Tensor, latlon_coords: T.Optional[torch.Tensor] = None) -> T.Dict[str, torch.Tensor]:
        """x: image time series ``[B, C, T, H, W]`` (float32)."""
        if x.dim() != 5 or x.shape[1] != self.in_channels or x.shape[2] != self.in_time:
            raise ValueError(f"TowerUNet expects x[B,{self.in_channels},{self.in_time},H,W], got {tuple(x.shape)}")
        dtype = self.compute_dtype
        if self.training:
            F.reset_stats_arena(x.device)  # clean [2, N] slices for the BatchNorm sums the convolution epilogues produce
        if self.training and self.dropout > 0:
            F.rng_advance(x.device)  # one new set of dropout masks per forward (a kernel, so CUDA-graph replays advance too)
        embeddings = self.pre_unet(x.float(), dtype)
        encoded = self.encoder(embeddings)
        decoded = self.decoder(encoded)
        towers = self.tower_fusion(encoded=encoded, decoded=decoded, latlon_coords=latlon_coords)
        t_a, t_b, t_c = towers["x_tower_a"], towers["x_tower_b"], towers["x_tower_c"]
        size = tuple(t_a.shape[1:3])
        h_a = self.final_a(t_a)
        h_b = self.final_b(t_b, size=size)
        h_c = self.final_c(t_c, size=size)
        distance, edge, crop = self.final_combine(h_a, h_b, h_c)
        return {InferenceNames.DISTANCE: distance, InferenceNames.EDGE: edge, InferenceNames.CROP: crop}
