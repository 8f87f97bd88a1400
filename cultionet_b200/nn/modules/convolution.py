"""ResUNet-a building blocks over pixel-major activations.

Module / parameter names follow ``src/cultionet/nn/modules/convolution.py`` so reference ``state_dict``s load unchanged:
the ``torch.nn`` layers below only *hold* parameters and buffers in the reference's layouts; every forward runs the
sm_100a kernels of ``cultionet_b200.functional``.  Activations are ``[B, H, W, C]`` (the reference is ``[B, C, H, W]``);
a "source list" stands for the channel concatenation the reference materialises with ``torch.cat``.
"""
from __future__ import annotations

import typing as T

import torch
import torch.nn as nn

from ... import functional as F
from ...enums import AttentionTypes, ResBlockTypes
from .attention import NeighborhoodAttention2D

Sources = T.Union[torch.Tensor, T.Sequence[torch.Tensor]]


def _as_sources(x: Sources) -> T.List[torch.Tensor]:
    return [x] if isinstance(x, torch.Tensor) else list(x)


def _act_module(activation_type: str) -> nn.Module:
    """The parameter-free activation module the reference's ``SetActivation`` builds (``activations.py:18-21``); it only keeps the
    ``nn.Sequential`` indices (and with them the state_dict keys) of the reference -- the arithmetic runs inside the fused kernels."""
    F.act_code(activation_type)  # raises for a name the kernels do not implement
    cls = getattr(nn, str(getattr(activation_type, "value", activation_type)))
    try:
        return cls(inplace=False)
    except TypeError:
        return cls()


# `num_batches_tracked += 1` is one tiny kernel per BatchNorm layer (52 per TowerUNet step).  A training loop may collect the counters
# of a step and bump them with ONE multi-tensor launch instead: `with deferred_batch_counters(): forward(...)`.
_PENDING_COUNTERS: T.Optional[T.List[torch.Tensor]] = None


class deferred_batch_counters:
    def __enter__(self):
        global _PENDING_COUNTERS
        self.prev = _PENDING_COUNTERS
        _PENDING_COUNTERS = []
        return self

    def __exit__(self, *exc):
        global _PENDING_COUNTERS
        pending, _PENDING_COUNTERS = _PENDING_COUNTERS, self.prev
        if pending:
            torch._foreach_add_(pending, 1)
        return False


def bump_batch_counter(bn: nn.modules.batchnorm._BatchNorm) -> None:
    if bn.track_running_stats and bn.num_batches_tracked is not None:
        if _PENDING_COUNTERS is not None:
            _PENDING_COUNTERS.append(bn.num_batches_tracked)
        else:
            bn.num_batches_tracked.add_(1)


def batchnorm_act(bn: nn.modules.batchnorm._BatchNorm, x: torch.Tensor, act: T.Union[bool, int], ch_div: int = 1,
                  sums: T.Optional[torch.Tensor] = None, residual: T.Optional[torch.Tensor] = None) -> torch.Tensor:
    """BatchNorm2d/3d (+ activation: True = SiLU, or a ``F.act_code``) (+ residual) with the module's parameters; batch statistics iff
    the module is in training mode."""
    training = bn.training
    if training:
        bump_batch_counter(bn)
    return F.batchnorm_act(
        x, bn.weight, bn.bias, bn.running_mean, bn.running_var, training,
        momentum=bn.momentum if bn.momentum is not None else 0.1, eps=bn.eps, act=act, ch_div=ch_div, sums=sums, residual=residual,
    )


class ConvTranspose2d(nn.Module):
    """``nn.ConvTranspose2d(k, stride, pad)`` then the bilinear ``align_corners=True`` fix-up to ``size``
    (reference ``convolution.py:45-68`` + ``nn/functional.py:72-81``)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, stride: int = 2, padding: int = 1):
        super().__init__()
        self.up_conv = nn.ConvTranspose2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding)

    def forward(self, x: torch.Tensor, size) -> torch.Tensor:
        m = self.up_conv
        k, s, p = m.kernel_size[0], m.stride[0], m.padding[0]
        out_hw = ((x.shape[1] - 1) * s - 2 * p + (k - 1) + 1, (x.shape[2] - 1) * s - 2 * p + (k - 1) + 1)
        if m.bias is not None and out_hw != tuple(size) and m.bias.requires_grad and torch.is_grad_enabled():
            # the resize backward writes the gradient the bias needs the column sums of: it accumulates them in the same pass
            y = F.conv_transpose2d(x, m.weight, m.bias.detach(), ksize=k, stride=s, pad=p)
            return F.resize_bilinear(y, size, producer_bias=m.bias)
        y = F.conv_transpose2d(x, m.weight, m.bias, ksize=k, stride=s, pad=p)
        return F.resize_bilinear(y, size)


class ConvBlock2d(nn.Module):
    """Conv2d(bias=False) -> BatchNorm2d -> [SiLU], or with ``batchnorm_first``: BatchNorm2d -> SiLU -> Conv2d(bias=True)
    (reference ``convolution.py:71-120``)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, padding: int = 0, dilation: int = 1,
                 stride: int = 1, add_activation: bool = True, activation_type: str = "SiLU", batchnorm_first: bool = False):
        super().__init__()
        self.act = F.act_code(activation_type)
        self.batchnorm_first = batchnorm_first
        if batchnorm_first:
            layers = [
                nn.BatchNorm2d(in_channels),
                _act_module(activation_type),
                nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, padding=padding, dilation=dilation, stride=stride),
            ]
        else:
            layers = [
                nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, padding=padding, dilation=dilation, stride=stride, bias=False),
                nn.BatchNorm2d(out_channels),
            ]
            if add_activation:
                layers.append(_act_module(activation_type))
        self.add_activation = add_activation
        self.seq = nn.Sequential(*layers)

    def forward(self, x: Sources, residual: T.Optional[torch.Tensor] = None) -> torch.Tensor:
        """``residual`` (training-mode, conv-first blocks only): added to the block's output inside the BatchNorm kernel."""
        if self.batchnorm_first:
            assert residual is None
            return self._forward_bn_first(_as_sources(x))
        conv, bn = self.seq[0], self.seq[1]
        if not bn.training and not torch.is_grad_enabled() and residual is None:
            # inference: BatchNorm (running statistics) and SiLU ride in the convolution epilogue -- one launch, one write
            y = F.conv2d_bn_act_eval(_as_sources(x), conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps,
                                     self.act if self.add_activation else 0, ksize=conv.kernel_size[0], stride=conv.stride[0],
                                     pad=conv.padding[0],
                                     dil=conv.dilation[0])
            if y is not None:
                return y
        # in training mode the tcgen05 convolution epilogue also produces BatchNorm's per-channel sums (no separate statistics pass)
        y = F.conv2d(_as_sources(x), conv.weight, None, ksize=conv.kernel_size[0], stride=conv.stride[0], pad=conv.padding[0],
                     dil=conv.dilation[0], want_stats=bn.training)
        sums = None
        if bn.training:
            y, sums = y
        return batchnorm_act(bn, y, act=self.act if self.add_activation else 0, sums=sums, residual=residual)

    def _forward_bn_first(self, sources: T.List[torch.Tensor]) -> torch.Tensor:
        """BatchNorm over the (virtual) channel concatenation = BatchNorm of every source with its slice of the parameters."""
        bn, conv = self.seq[0], self.seq[2]
        if len(sources) == 1:
            normed = [batchnorm_act(bn, sources[0], act=self.act)]
        else:
            if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
                if _PENDING_COUNTERS is not None:
                    _PENDING_COUNTERS.append(bn.num_batches_tracked)
                else:
                    bn.num_batches_tracked.add_(1)
            normed, c0 = [], 0
            for s in sources:
                c1 = c0 + s.shape[-1]
                normed.append(F.batchnorm_act(s, bn.weight[c0:c1], bn.bias[c0:c1], bn.running_mean[c0:c1], bn.running_var[c0:c1],
                                              bn.training, momentum=bn.momentum if bn.momentum is not None else 0.1, eps=bn.eps, act=self.act))
                c0 = c1
        return F.conv2d(normed, conv.weight, conv.bias, ksize=conv.kernel_size[0], stride=conv.stride[0], pad=conv.padding[0],
                        dil=conv.dilation[0])


class ResConvBlock2d(nn.Module):
    """``num_blocks`` ConvBlock2d in sequence; block 0 always runs at dilation 1, later blocks at ``max(1, dilation-1)``
    (reference ``convolution.py:123-176``)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, dilation: int = 1, activation_type: str = "SiLU",
                 num_blocks: int = 2, batchnorm_first: bool = False):
        super().__init__()
        assert num_blocks > 0, "There must be at least one block."
        later = 1 if kernel_size == 1 else max(1, dilation - 1)
        blocks = [
            ConvBlock2d(in_channels, out_channels, kernel_size, padding=0 if kernel_size == 1 else kernel_size // 2, dilation=1,
                        activation_type=activation_type, batchnorm_first=batchnorm_first)
        ]
        for _ in range(num_blocks - 1):
            blocks.append(
                ConvBlock2d(out_channels, out_channels, kernel_size, padding=0 if kernel_size == 1 else later, dilation=later,
                            activation_type=activation_type, batchnorm_first=batchnorm_first)
            )
        self.block = nn.ModuleList(blocks)

    def forward(self, x: Sources, residual: T.Optional[torch.Tensor] = None) -> torch.Tensor:
        last = len(self.block) - 1
        for i, layer in enumerate(self.block):
            x = layer(x, residual=residual) if (i == last and residual is not None) else layer(x)
        return x

    @property
    def takes_residual(self) -> bool:
        """The last block can add a residual in its BatchNorm kernel: conv-first layout, training-mode statistics."""
        blk = self.block[-1]
        return (not blk.batchnorm_first) and blk.seq[1].training


class ChannelAttention(nn.Module):
    """Holds the two channel MLPs (1x1 convs without bias) of the reference ``attention.py:12-63``."""

    def __init__(self, in_channels: int, activation_type: str):
        super().__init__()
        self.act = F.act_code(activation_type)

        def mlp():
            return nn.Sequential(nn.Conv2d(in_channels, in_channels // 2, kernel_size=1, padding=0, bias=False), _act_module(activation_type),
                                 nn.Conv2d(in_channels // 2, in_channels, kernel_size=1, padding=0, bias=False))

        self.fc1 = mlp()
        self.fc2 = mlp()

    def _mlp(self, seq: nn.Sequential, v: torch.Tensor) -> torch.Tensor:
        h = F.activation(F.conv2d([v], seq[0].weight, None, ksize=1, stride=1, pad=0), self.act)
        return F.conv2d([h], seq[2].weight, None, ksize=1, stride=1, pad=0)

    def logits(self, ch_avg: torch.Tensor, ch_max: torch.Tensor) -> torch.Tensor:
        """fc1(avg-pooled) + fc2(max-pooled): ``[B,1,1,C]`` fp32 (the sigmoid is taken by the apply kernel)."""
        return F.add_n(self._mlp(self.fc1, ch_avg), self._mlp(self.fc2, ch_max))


class SpatialAttention(nn.Module):
    """Holds the 3x3 (2 -> 1) convolution of the reference ``attention.py:66-88``."""

    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(2, 1, kernel_size=3, padding=1, bias=False)

    def logits(self, sp: torch.Tensor) -> torch.Tensor:
        return F.conv2d([sp], self.conv.weight, None, ksize=3, stride=1, pad=1)


class SpatialChannelAttention(nn.Module):
    """``1 + gamma * 0.5 * (channel_attention + spatial_attention)`` (reference ``attention.py:91-125``), applied to a second
    tensor by ``scale``: the pooled statistics and both logit maps are small fp32 tensors, the two passes over the activations
    (pooling, apply) are single kernels."""

    def __init__(self, in_channels: int, activation_type: str):
        super().__init__()
        self.channel_attention = ChannelAttention(in_channels=in_channels, activation_type=activation_type)
        self.spatial_attention = SpatialAttention()
        self.gamma = nn.Parameter(torch.zeros(1))

    def scale(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """``y * attention(x)``."""
        sp, ch_avg, ch_max = F.sca_pool(x)
        cl = self.channel_attention.logits(ch_avg, ch_max)
        sl = self.spatial_attention.logits(sp)
        return F.sca_apply(y, cl, sl, self.gamma)


class ResidualConv(nn.Module):
    """``skip(x) + ResConvBlock2d(x)`` (reference ``convolution.py:179-247``)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, num_blocks: int = 2,
                 attention_weights: T.Optional[str] = None, activation_type: str = "SiLU", batchnorm_first: bool = False):
        super().__init__()
        self.attention_weights = attention_weights
        if attention_weights is not None:
            assert attention_weights in [AttentionTypes.SPATIAL_CHANNEL], "The attention method is not supported."
            # the reference constructs SpatialChannelAttention(out_channels=...) here (convolution.py:203-205), a keyword that class
            # does not take: ResidualConv with attention cannot be built there either
            raise TypeError("SpatialChannelAttention.__init__() got an unexpected keyword argument 'out_channels'")
        self.seq = ResConvBlock2d(in_channels, out_channels, kernel_size=kernel_size, num_blocks=num_blocks,
                                  activation_type=activation_type, batchnorm_first=batchnorm_first)
        self.skip = None
        if in_channels != out_channels:
            self.skip = nn.Conv2d(in_channels, out_channels, kernel_size=1, padding=0)

    def forward(self, x: Sources) -> torch.Tensor:
        sources = _as_sources(x)
        if self.skip is None:
            assert len(sources) == 1
            a, b = F.fanout(sources[0], 2)
            return F.add_n(a, self.seq(b))
        refs = [F.fanout(s, 2) for s in sources]
        skip = F.conv2d([r[0] for r in refs], self.skip.weight, self.skip.bias, ksize=1, stride=1, pad=0)
        return F.add_n(skip, self.seq([r[1] for r in refs]))


class ResidualAConv(nn.Module):
    """``skip(x) + sum_d ResConvBlock2d_d(x) [+ LN(NA(LN(skip(x))))]`` (reference ``convolution.py:250-395``)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, num_blocks: int = 2,
                 dilations: T.Optional[T.List[int]] = None, attention_weights: T.Optional[str] = None, activation_type: str = "SiLU",
                 batchnorm_first: bool = False, natten_num_heads: int = 8, natten_kernel_size: int = 3, natten_dilation: int = 1,
                 natten_attn_drop: float = 0.0, natten_proj_drop: float = 0.0):
        super().__init__()
        if dilations is None:
            dilations = [1, 2]
        self.attention_weights = attention_weights
        if in_channels != out_channels:
            self.skip = nn.Conv2d(in_channels, out_channels, kernel_size=1, padding=0)
        else:
            self.skip = nn.Identity()
        if attention_weights is not None:
            assert attention_weights in [AttentionTypes.NATTEN, AttentionTypes.SPATIAL_CHANNEL], "The attention method is not supported."
            if attention_weights == AttentionTypes.NATTEN:
                # indices 1..3 carry the parameters (the reference has einops Rearrange layers at 0 and 4)
                self.attention_conv = nn.Sequential(
                    nn.Identity(),
                    nn.LayerNorm(out_channels),
                    NeighborhoodAttention2D(out_channels, num_heads=natten_num_heads, kernel_size=natten_kernel_size,
                                            dilation=natten_dilation, attn_drop=natten_attn_drop, proj_drop=natten_proj_drop),
                    nn.LayerNorm(out_channels),
                    nn.Identity(),
                )
            else:
                self.attention_conv = SpatialChannelAttention(in_channels=out_channels, activation_type=activation_type)
        self.res_modules = nn.ModuleList([
            ResConvBlock2d(in_channels, out_channels, kernel_size=kernel_size, dilation=d, activation_type=activation_type,
                           num_blocks=num_blocks, batchnorm_first=batchnorm_first)
            for d in dilations
        ])

    def forward(self, x: Sources) -> torch.Tensor:
        sources = _as_sources(x)
        nres = len(self.res_modules)
        attention = self.attention_weights is not None
        # every source feeds the skip path and each dilation branch; F.fanout hands each consumer its own reference so that the
        # backward sums their gradients in one n-ary add instead of autograd's pairwise accumulation
        if isinstance(self.skip, nn.Identity):
            assert len(sources) == 1
            refs = F.fanout(sources[0], nres + 1 + (1 if attention else 0))
            skip, att_in = refs[0], refs[-1]
            branch_in = [[refs[1 + i]] for i in range(nres)]
        else:
            refs = [F.fanout(s, nres + 1) for s in sources]
            skip = F.conv2d([r[0] for r in refs], self.skip.weight, self.skip.bias, ksize=1, stride=1, pad=0)
            branch_in = [[r[1 + i] for r in refs] for i in range(nres)]
            if attention:
                skip, att_in = F.fanout(skip, 2)
        if all(layer.takes_residual for layer in self.res_modules):
            # training: skip + branch_1 + branch_2 ... accumulate inside the branches' last BatchNorm kernels (y = SiLU(BN(x)) + residual)
            # instead of a separate n-ary add pass; the gradient of a residual is the incoming gradient itself (no kernel)
            acc = skip
            for i, layer in enumerate(self.res_modules):
                acc = layer(branch_in[i], residual=acc)
            terms = [acc]
        else:
            terms = [skip] + [layer(branch_in[i]) for i, layer in enumerate(self.res_modules)]
        natten = attention and self.attention_weights == AttentionTypes.NATTEN
        if natten:
            ln1, na, ln2 = self.attention_conv[1], self.attention_conv[2], self.attention_conv[3]
            a = F.layernorm(att_in, ln1.weight, ln1.bias, ln1.eps)
            a = na(a)
            terms.append(F.layernorm(a, ln2.weight, ln2.bias, ln2.eps))
        out = terms[0] if len(terms) == 1 else F.add_n(*terms[:4])
        for i in range(4, len(terms), 3):
            out = F.add_n(out, *terms[i:i + 3])
        if attention and not natten:
            out = self.attention_conv.scale(att_in, out)  # out * (1 + gamma * attention(skip)), convolution.py:392-393
        return out


class PoolResidualConv(nn.Module):
    """[3x3 stride-2 conv (+ BN), or adaptive max pooling to (H//2, W//2)] -> ResidualAConv / ResidualConv -> Dropout2d
    (reference ``convolution.py:398-513``)."""

    def __init__(self, in_channels: int, out_channels: int, dropout: float = 0.0, kernel_size: int = 3, num_blocks: int = 2,
                 attention_weights: T.Optional[str] = None, activation_type: str = "SiLU", res_block_type: str = ResBlockTypes.RESA,
                 dilations: T.Sequence[int] = None, pool_first: bool = True, pool_by_max: bool = False, batchnorm_first: bool = False,
                 natten_num_heads: int = 8, natten_kernel_size: int = 3, natten_dilation: int = 1, natten_attn_drop: float = 0.0,
                 natten_proj_drop: float = 0.0):
        super().__init__()
        assert res_block_type in (ResBlockTypes.RES, ResBlockTypes.RESA)
        self.pool_first = pool_first
        self.pool_by_max = pool_by_max
        if pool_first and not pool_by_max:
            if batchnorm_first:
                self.pool_conv = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1, stride=2)
            else:
                self.pool_conv = ConvBlock2d(in_channels, out_channels, kernel_size=3, padding=1, stride=2, add_activation=False,
                                             batchnorm_first=False)
            in_channels = out_channels
        if res_block_type == ResBlockTypes.RES:
            self.res_conv = ResidualConv(in_channels, out_channels, kernel_size=kernel_size, attention_weights=attention_weights,
                                         num_blocks=num_blocks, activation_type=activation_type, batchnorm_first=batchnorm_first)
        else:
            self.res_conv = ResidualAConv(in_channels, out_channels, kernel_size=kernel_size, dilations=dilations, num_blocks=num_blocks,
                                          attention_weights=attention_weights, activation_type=activation_type,
                                          batchnorm_first=batchnorm_first, natten_num_heads=natten_num_heads,
                                          natten_kernel_size=natten_kernel_size, natten_dilation=natten_dilation,
                                          natten_attn_drop=natten_attn_drop, natten_proj_drop=natten_proj_drop)
        self.dropout_layer = nn.Dropout2d(p=dropout)
        self._rng_site = F.new_rng_site()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.pool_first:
            if self.pool_by_max:
                x = F.adaptive_max_pool2d(x, (x.shape[1] // 2, x.shape[2] // 2))
            elif isinstance(self.pool_conv, nn.Conv2d):
                c = self.pool_conv
                x = F.conv2d([x], c.weight, c.bias, ksize=3, stride=2, pad=1)
            else:
                x = self.pool_conv(x)
        x = self.res_conv(x)
        if self.training and self.dropout_layer.p > 0:
            x = F.dropout2d(x, self.dropout_layer.p, self._rng_site)
        return x
