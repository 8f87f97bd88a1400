"""TowerUNet encoder / decoder / UNet3+ full-scale towers / Psi-Net heads over pixel-major activations.

Module and parameter names follow ``src/cultionet/nn/modules/unet_parts.py`` (state_dict compatible).  Differences in
mechanics, not in arithmetic:
  * the full-scale skip ``torch.cat`` (reference ``unet_parts.py:733-758``) is never materialised -- the towers hand a
    *source list* to the implicit-GEMM convolution, which walks the sources inside its K loop;
  * ``TowerUNetFinal`` returns its three streams packed as one ``[B, H, W, 3]`` tensor (distance, edge, crop) and
    ``TowerUNetFinalCombine`` fuses the 1/gamma weighting, the 1x1 convs, the sigmoids and SigmoidCrisp in one kernel.
"""
from __future__ import annotations

import typing as T

import torch
import torch.nn as nn

from ... import functional as F
from ...enums import AttentionTypes, ResBlockTypes
from .convolution import ConvBlock2d, ConvTranspose2d, PoolResidualConv, ResidualAConv, ResidualConv, batchnorm_act

# natten settings per resolution level (reference ``unet_parts.py:19-40``); a mutable module-level dict there too
NATTEN_PARAMS = {
    "a": {"natten_num_heads": 4, "natten_kernel_size": 3, "natten_dilation": 2},
    "b": {"natten_num_heads": 4, "natten_kernel_size": 3, "natten_dilation": 1},
    "c": {"natten_num_heads": 8, "natten_kernel_size": 3, "natten_dilation": 1},
    "d": {"natten_num_heads": 8, "natten_kernel_size": 1, "natten_dilation": 1},
}


def _res_block(res_block_type: str, in_channels: int, out_channels: int, kernel_size: int, num_blocks: T.Optional[int], dilations,
               attention_weights, activation_type: str, batchnorm_first: bool, natten_kw: dict) -> nn.Module:
    """ResidualConv for ``res``, ResidualAConv for ``resa`` (reference ``unet_parts.py:340-368``, ``:683-710``)."""
    assert res_block_type in (ResBlockTypes.RES, ResBlockTypes.RESA)
    if res_block_type == ResBlockTypes.RES:
        return ResidualConv(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size,
                            num_blocks=2 if num_blocks is None else num_blocks, attention_weights=attention_weights,
                            activation_type=activation_type, batchnorm_first=batchnorm_first)
    kw = {} if num_blocks is None else {"num_blocks": num_blocks}
    return ResidualAConv(in_channels, out_channels, kernel_size=kernel_size, dilations=dilations, attention_weights=attention_weights,
                         activation_type=activation_type, batchnorm_first=batchnorm_first, **kw, **natten_kw)


class SigmoidCrisp(nn.Module):
    """``sigmoid(x / (smooth + sigmoid(gamma)))`` -- holds ``gamma``; evaluated inside the final-combine kernel."""

    def __init__(self, smooth: float = 1e-2):
        super().__init__()
        self.smooth = smooth
        self.gamma = nn.Parameter(torch.ones(1))


class TowerUNetFinalCombine(nn.Module):
    """Learned 1/gamma-weighted sum of the three towers per task -> 1x1 conv -> sigmoid / SigmoidCrisp
    (reference ``unet_parts.py:101-193``)."""

    def __init__(self, num_classes: int, edge_activation: bool = True, mask_activation: bool = True):
        super().__init__()
        if num_classes != 1:
            raise NotImplementedError("cultionet_b200: the fused head is built for num_classes=1 (what CultioNet uses)")
        self.edge_activation = edge_activation
        self.mask_activation = mask_activation
        self.final_dist = nn.Sequential(nn.Conv2d(1, 1, kernel_size=1, padding=0), nn.Sigmoid())
        self.dist_gamma1 = nn.Parameter(torch.ones(1))
        self.dist_gamma2 = nn.Parameter(torch.ones(1))
        self.dist_gamma3 = nn.Parameter(torch.ones(1))
        self.final_edge = nn.Sequential(nn.Conv2d(1, 1, kernel_size=1, padding=0), SigmoidCrisp() if edge_activation else nn.Identity())
        self.edge_gamma1 = nn.Parameter(torch.ones(1))
        self.edge_gamma2 = nn.Parameter(torch.ones(1))
        self.edge_gamma3 = nn.Parameter(torch.ones(1))
        self.final_crop = nn.Sequential(nn.Conv2d(num_classes, num_classes, kernel_size=1, padding=0),
                                        nn.Sigmoid() if mask_activation else nn.Identity())
        self.crop_gamma1 = nn.Parameter(torch.ones(1))
        self.crop_gamma2 = nn.Parameter(torch.ones(1))
        self.crop_gamma3 = nn.Parameter(torch.ones(1))
        self.register_buffer("_unit", torch.ones(1), persistent=False)

    def forward(self, h_a: torch.Tensor, h_b: torch.Tensor, h_c: torch.Tensor):
        crisp = self.final_edge[1]
        crisp_gamma = crisp.gamma if self.edge_activation else self._unit
        smooth = crisp.smooth if self.edge_activation else 1e-2
        params = [
            self.dist_gamma1, self.dist_gamma2, self.dist_gamma3,
            self.edge_gamma1, self.edge_gamma2, self.edge_gamma3,
            self.crop_gamma1, self.crop_gamma2, self.crop_gamma3,
            self.final_dist[0].weight, self.final_edge[0].weight, self.final_crop[0].weight,
            self.final_dist[0].bias, self.final_edge[0].bias, self.final_crop[0].bias,
            crisp_gamma,
        ]
        return F.final_combine(h_a, h_b, h_c, params, smooth=smooth, edge_activation=self.edge_activation,
                               mask_activation=self.mask_activation)


class StreamConv2d(nn.Module):
    """3x3 C->hidden (+BN+SiLU) -> 3x3 hidden->out with bias (reference ``unet_parts.py:196-224``)."""

    def __init__(self, in_channels: int, hidden_channels: int, out_channels: int, activation_type: str):
        super().__init__()
        self.conv = nn.Sequential(
            ConvBlock2d(in_channels, hidden_channels, kernel_size=3, padding=1, add_activation=True, activation_type=activation_type),
            nn.Conv2d(hidden_channels, out_channels, kernel_size=3, padding=1),
        )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = self.conv[0](x)
        c = self.conv[1]
        return F.conv2d([h], c.weight, c.bias, ksize=3, stride=1, pad=1)


class TowerUNetFinal(nn.Module):
    """Psi-Net head of one tower (reference ``unet_parts.py:227-309``); returns ``[B, H, W, 3]`` = (distance, edge, crop)."""

    def __init__(self, in_channels: int, num_classes: int, activation_type: str = "SiLU", resample_factor: int = 0):
        super().__init__()
        self.in_channels = in_channels
        self.num_classes = num_classes
        if resample_factor > 1:
            self.up_conv = ConvTranspose2d(in_channels, in_channels, kernel_size=3, stride=resample_factor, padding=1)
        self.dist_conv = StreamConv2d(in_channels, 3, 1, activation_type)
        self.edge_conv = StreamConv2d(in_channels, 3, 1, activation_type)
        self.crop_conv = StreamConv2d(in_channels, 3, 1, activation_type)
        self.fuse_conv = ConvBlock2d(3, 3, kernel_size=3, padding=1, add_activation=True, activation_type=activation_type)

    def forward(self, x: torch.Tensor, size=None, suffix: str = "") -> torch.Tensor:
        """The three streams read the same ``[B,H,W,C]`` tensor, so they run as ONE convolution with the three filter banks
        stacked along the output axis (C -> 9), one BatchNorm+SiLU over the 9 channels (per-channel statistics are independent,
        so this equals the three separate BatchNorm2d(3) of the reference) and one block-diagonal 9 -> 3 convolution for the
        three 3 -> 1 stream outputs.  Stacking the (tiny) parameters is plumbing; the arithmetic per channel is unchanged."""
        if size is not None:
            x = self.up_conv(x, size=size)
        streams = (self.dist_conv, self.edge_conv, self.crop_conv)
        blocks = [s.conv[0] for s in streams]
        bns = [b.seq[1] for b in blocks]
        training = bns[0].training
        C = blocks[0].seq[0].weight.shape[1]
        # the stacked filters / BatchNorm parameters come from ONE multi-tensor copy each (functional.stack_params: no torch.cat, no
        # per-parameter gradient adds); the running statistics of the three BatchNorm2d(3) are views of one 9-element buffer
        w1 = F.stack_params([b.seq[0].weight for b in blocks], [0, 27 * C, 54 * C], 81 * C)
        w1 = F.tag_derived(w1.view(9, C, 3, 3), w1)
        gamma = F.stack_params([bn.weight for bn in bns], [0, 3, 6], 9)
        beta = F.stack_params([bn.bias for bn in bns], [0, 3, 6], 9)
        rm, rv = self._stacked_running_stats(bns)
        sums = None
        if x.dtype == torch.bfloat16:
            # throughput mode: C -> 9 as a 1x1 GEMM (C -> 81) + shift-and-add, so the wide tower tensor is read once, not once per tap
            h = F.conv2d_skinny(x, w1, ksize=3, pad=1, dil=1)
        else:
            h = F.conv2d([x], w1, None, ksize=3, stride=1, pad=1, want_stats=training)
            if training:
                h, sums = h
        if training:
            from .convolution import bump_batch_counter

            for bn in bns:
                bump_batch_counter(bn)
        h = F.batchnorm_act(h, gamma, beta, rm, rv, training, momentum=bns[0].momentum if bns[0].momentum is not None else 0.1,
                            eps=bns[0].eps, act=blocks[0].act, sums=sums)
        # 3 x (3 -> 1) as one block-diagonal 9 -> 3 convolution: stream i's [1,3,3,3] filter sits at input channels 3i..3i+2 of output i
        w2 = F.stack_params([s.conv[1].weight for s in streams], [0, 108, 216], 243)
        w2 = F.tag_derived(w2.view(3, 9, 3, 3), w2)
        b2 = F.stack_params([s.conv[1].bias for s in streams], [0, 1, 2], 3)
        z = F.conv2d([h], w2, b2, ksize=3, stride=1, pad=1)
        return self.fuse_conv(z)

    def _stacked_running_stats(self, bns):
        """(running_mean[9], running_var[9]) whose thirds ARE the three modules' buffers: each ``bn.running_mean`` / ``running_var`` is
        re-pointed at a view of one shared tensor (state_dict / load_state_dict see the same names, shapes and values), so the stacked
        BatchNorm kernel updates them in place and nothing is copied per step.  Re-established when ``.to()`` re-homes the buffers."""
        out = []
        for name in ("running_mean", "running_var"):
            bufs = [getattr(bn, name) for bn in bns]
            base = getattr(self, "_stk_" + name, None)
            ok = base is not None and base.device == bufs[0].device and all(
                b.data_ptr() == base.data_ptr() + 12 * i and b.numel() == 3 for i, b in enumerate(bufs))
            if not ok:
                base = torch.cat([b.detach().reshape(-1).float() for b in bufs])
                for i, bn in enumerate(bns):
                    setattr(bn, name, base[3 * i:3 * i + 3])
                object.__setattr__(self, "_stk_" + name, base)
            out.append(base)
        return out


class UNetUpBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, num_blocks: int = 2,
                 attention_weights: T.Optional[str] = None, activation_type: str = "SiLU", res_block_type: str = ResBlockTypes.RESA,
                 dilations: T.Sequence[int] = None, batchnorm_first: bool = False, resample_up: bool = True, natten_num_heads: int = 8,
                 natten_kernel_size: int = 3, natten_dilation: int = 1, natten_attn_drop: float = 0.0, natten_proj_drop: float = 0.0):
        super().__init__()
        if resample_up:
            self.up_conv = ConvTranspose2d(in_channels, in_channels)
        natten_kw = dict(natten_num_heads=natten_num_heads, natten_kernel_size=natten_kernel_size, natten_dilation=natten_dilation,
                         natten_attn_drop=natten_attn_drop, natten_proj_drop=natten_proj_drop)
        # like the reference (unet_parts.py:355-368) the RESA branch does not forward ``num_blocks``; the RES branch does (:342-350)
        self.res_conv = _res_block(res_block_type, in_channels, out_channels, kernel_size,
                                   num_blocks if res_block_type == ResBlockTypes.RES else None, dilations, attention_weights,
                                   activation_type, batchnorm_first, natten_kw)

    def forward(self, x: torch.Tensor, size) -> torch.Tensor:
        if tuple(x.shape[1:3]) != tuple(size):
            x = self.up_conv(x, size=size)
        return self.res_conv(x)


class TowerUNetEncoder(nn.Module):
    def __init__(self, channels: T.Sequence[int], dilations: T.Sequence[int] = None, activation_type: str = "SiLU", dropout: float = 0.0,
                 res_block_type: str = ResBlockTypes.RESA, attention_weights: str = AttentionTypes.NATTEN, pool_by_max: bool = False,
                 batchnorm_first: bool = False):
        super().__init__()
        kw = dict(dropout=dropout, activation_type=activation_type, res_block_type=res_block_type, batchnorm_first=batchnorm_first,
                  pool_by_max=pool_by_max, natten_attn_drop=dropout, natten_proj_drop=dropout)
        self.down_a = PoolResidualConv(channels[0], channels[0], dilations=dilations, pool_first=False, attention_weights=attention_weights,
                                       **{**kw, **NATTEN_PARAMS["a"]})
        self.down_b = PoolResidualConv(channels[0], channels[1], dilations=dilations[:3], attention_weights=attention_weights,
                                       **{**kw, **NATTEN_PARAMS["b"]})
        self.down_c = PoolResidualConv(channels[1], channels[2], dilations=dilations[:2], attention_weights=attention_weights,
                                       **{**kw, **NATTEN_PARAMS["c"]})
        self.down_d = PoolResidualConv(channels[2], channels[3], kernel_size=1, num_blocks=1, dilations=[1], attention_weights=None, **kw)

    def forward(self, x: torch.Tensor) -> T.Dict[str, torch.Tensor]:
        # every level feeds the next level AND one or two towers: F.fanout hands each consumer its own alias so that the backward sums
        # their gradients with one n-ary add kernel (keys: "x_*" = the same-level tower / decoder input, "x_*_down" = the input of the
        # tower one level up)
        a_next, x_a = F.fanout(self.down_a(x), 2)
        b_next, x_b, x_b_down = F.fanout(self.down_b(a_next), 3)
        c_next, x_c, x_c_down = F.fanout(self.down_c(b_next), 3)
        x_d, x_d_down = F.fanout(self.down_d(c_next), 2)
        return {"x_a": x_a, "x_b": x_b, "x_c": x_c, "x_d": x_d, "x_b_down": x_b_down, "x_c_down": x_c_down, "x_d_down": x_d_down}


class TowerUNetDecoder(nn.Module):
    def __init__(self, channels: T.Sequence[int], up_channels: int, dilations: T.Sequence[int] = None, activation_type: str = "SiLU",
                 dropout: float = 0.0, res_block_type: str = ResBlockTypes.RESA, attention_weights: str = AttentionTypes.NATTEN,
                 batchnorm_first: bool = False):
        super().__init__()
        kw = dict(activation_type=activation_type, res_block_type=res_block_type, batchnorm_first=batchnorm_first,
                  natten_attn_drop=dropout, natten_proj_drop=dropout)
        self.over_d = UNetUpBlock(channels[3], up_channels, kernel_size=1, num_blocks=1, dilations=[1], resample_up=False,
                                  attention_weights=None, **kw)
        self.up_cu = UNetUpBlock(up_channels, up_channels, dilations=dilations[:2], attention_weights=attention_weights,
                                 **{**kw, **NATTEN_PARAMS["c"]})
        self.up_bu = UNetUpBlock(up_channels, up_channels, dilations=dilations[:3], attention_weights=attention_weights,
                                 **{**kw, **NATTEN_PARAMS["b"]})
        self.up_au = UNetUpBlock(up_channels, up_channels, dilations=dilations, attention_weights=attention_weights,
                                 **{**kw, **NATTEN_PARAMS["a"]})

    def forward(self, x: T.Dict[str, torch.Tensor]) -> T.Dict[str, torch.Tensor]:
        hw = lambda t: tuple(t.shape[1:3])  # noqa: E731
        du_next, x_du = F.fanout(self.over_d(x["x_d"], size=hw(x["x_d"])), 2)
        cu_next, x_cu, x_cu_down = F.fanout(self.up_cu(du_next, size=hw(x["x_c"])), 3)
        bu_next, x_bu, x_bu_down = F.fanout(self.up_bu(cu_next, size=hw(x["x_b"])), 3)
        x_au = self.up_au(bu_next, size=hw(x["x_a"]))
        return {"x_au": x_au, "x_bu": x_bu, "x_cu": x_cu, "x_du": x_du, "x_cu_down": x_cu_down, "x_bu_down": x_bu_down}


class GeoEmbeddings(nn.Module):
    """``nn/modules/geo_encoding.py:5-26``: (lon, lat) in decimal degrees -> unit-sphere cartesian (x, y, z) -> ``Linear(3, channels)``.
    The degrees -> cartesian step is three trigonometric values per SAMPLE (a ``[B, 2]`` tensor, no gradient in the reference either);
    the Linear runs on the convolution kernel like every other Linear of the model."""

    def __init__(self, channels: int):
        super().__init__()
        self.coord_embedding = nn.Linear(3, channels)

    @torch.no_grad()
    def decimal_degrees_to_cartesian(self, degrees: torch.Tensor) -> torch.Tensor:
        radians = torch.deg2rad(degrees)
        cosine, sine = torch.cos(radians), torch.sin(radians)
        return torch.stack([cosine[:, 1] * cosine[:, 0], cosine[:, 1] * sine[:, 0], sine[:, 1]], dim=-1)

    def forward(self, x: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """``x``: ``[B, 2]`` (lon, lat) -> ``[B, 1, 1, channels]`` pixel-major embedding in the compute dtype."""
        cart = self.decimal_degrees_to_cartesian(x.float())
        cart = cart.to(dtype).view(cart.shape[0], 1, 1, 3)
        return F.linear(cart, self.coord_embedding.weight, self.coord_embedding.bias)


class TowerUNetBlock(nn.Module):
    """UNet3+ full-scale skip: same-level {backbone, decoder} + ConvT-upsampled level below of {backbone, decoder[, tower]}
    -> ResidualAConv over the virtual concatenation (reference ``unet_parts.py:615-760``)."""

    def __init__(self, backbone_side_channels: int, backbone_down_channels: int, up_channels: int, out_channels: int, tower: bool = False,
                 kernel_size: int = 3, num_blocks: int = 2, attention_weights: T.Optional[str] = None,
                 res_block_type: str = ResBlockTypes.RESA, dilations: T.Sequence[int] = None, activation_type: str = "SiLU",
                 batchnorm_first: bool = False, natten_num_heads: int = 8, natten_kernel_size: int = 3, natten_dilation: int = 1,
                 natten_attn_drop: float = 0.0, natten_proj_drop: float = 0.0, use_latlon: bool = False):
        super().__init__()
        self.use_latlon = use_latlon
        in_channels = backbone_side_channels + backbone_down_channels + up_channels * 2
        self.backbone_down_conv = ConvTranspose2d(backbone_down_channels, backbone_down_channels, kernel_size=3, stride=2, padding=1)
        self.decode_down_conv = ConvTranspose2d(up_channels, up_channels, kernel_size=3, stride=2, padding=1)
        if tower:
            self.tower_conv = ConvTranspose2d(up_channels, up_channels, kernel_size=3, stride=2, padding=1)
            in_channels += up_channels
        if use_latlon:  # reference unet_parts.py:676-681 (its torch.compile wrapper changes nothing in the arithmetic)
            self.geo_embeddings = GeoEmbeddings(up_channels)
            in_channels += up_channels
        natten_kw = dict(natten_num_heads=natten_num_heads, natten_kernel_size=natten_kernel_size, natten_dilation=natten_dilation,
                         natten_attn_drop=natten_attn_drop, natten_proj_drop=natten_proj_drop)
        self.res_conv = _res_block(res_block_type, in_channels, out_channels, kernel_size, num_blocks, dilations, attention_weights,
                                   activation_type, batchnorm_first, natten_kw)

    def forward(self, backbone_side, backbone_down, decode_side, decode_down, tower_down=None, latlon_coords=None) -> torch.Tensor:
        size = tuple(decode_side.shape[1:3])
        sources = [
            backbone_side,
            self.backbone_down_conv(backbone_down, size=size),
            decode_side,
            self.decode_down_conv(decode_down, size=size),
        ]
        if self.use_latlon:  # reference :739-750: the per-sample embedding broadcast over the level's pixels, concatenated
            assert latlon_coords is not None, "No lat/lon coordinates given."
            emb = self.geo_embeddings(latlon_coords, decode_side.dtype)
            sources.append(F.broadcast_pixels(emb, size))
        if tower_down is not None:
            sources.append(self.tower_conv(tower_down, size=size))
        return self.res_conv(sources)


class TowerUNetFusion(nn.Module):
    def __init__(self, channels: T.Sequence[int], up_channels: int, dilations: T.Sequence[int] = None, activation_type: str = "SiLU",
                 dropout: float = 0.0, res_block_type: str = ResBlockTypes.RESA, attention_weights: str = AttentionTypes.NATTEN,
                 batchnorm_first: bool = False, use_latlon: bool = False):
        super().__init__()
        kw = dict(up_channels=up_channels, out_channels=up_channels, activation_type=activation_type, res_block_type=res_block_type,
                  batchnorm_first=batchnorm_first, attention_weights=attention_weights, natten_attn_drop=dropout, natten_proj_drop=dropout,
                  use_latlon=use_latlon)
        self.tower_c = TowerUNetBlock(backbone_side_channels=channels[2], backbone_down_channels=channels[3], dilations=dilations[:2],
                                      **{**kw, **NATTEN_PARAMS["c"]})
        self.tower_b = TowerUNetBlock(backbone_side_channels=channels[1], backbone_down_channels=channels[2], tower=True,
                                      dilations=dilations, **{**kw, **NATTEN_PARAMS["b"]})
        self.tower_a = TowerUNetBlock(backbone_side_channels=channels[0], backbone_down_channels=channels[1], tower=True,
                                      dilations=dilations, **{**kw, **NATTEN_PARAMS["a"]})

    def forward(self, encoded, decoded, latlon_coords=None) -> T.Dict[str, torch.Tensor]:
        enc = lambda k: encoded.get(k + "_down", encoded[k])  # noqa: E731 - plain dicts of tensors (no aliases) work too
        dec = lambda k: decoded.get(k + "_down", decoded[k])  # noqa: E731
        t_c_next, t_c = F.fanout(self.tower_c(backbone_side=encoded["x_c"], backbone_down=enc("x_d"), decode_side=decoded["x_cu"],
                                              decode_down=decoded["x_du"], latlon_coords=latlon_coords), 2)
        t_b_next, t_b = F.fanout(self.tower_b(backbone_side=encoded["x_b"], backbone_down=enc("x_c"), decode_side=decoded["x_bu"],
                                              decode_down=dec("x_cu"), tower_down=t_c_next, latlon_coords=latlon_coords), 2)
        t_a = self.tower_a(backbone_side=encoded["x_a"], backbone_down=enc("x_b"), decode_side=decoded["x_au"],
                           decode_down=dec("x_bu"), tower_down=t_b_next, latlon_coords=latlon_coords)
        return {"x_tower_a": t_a, "x_tower_b": t_b, "x_tower_c": t_c}
