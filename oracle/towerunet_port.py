"""TEST INFRASTRUCTURE ONLY -- plain-PyTorch fp32 restatement ("port") of the reference TowerUNet hot path.

The reference is pure Python over torch; it cannot travel to the GPU box (``/root/reference`` is absent there), so this
file restates its arithmetic functionally over a ``state_dict`` with the reference's key names.  It is the checker for the
CUDA path in ``tests/`` and the CPU baseline of ``bench.py`` (``cpu_baseline.kind = "port"``); nothing in the product
package imports it.

Pinned (not "parity unpinned") except at the natten boundary:
  * ``tests/test_oracle.py::test_port_matches_reference`` compares it with the real reference modules imported through
    ``oracle/ref_loader.py`` (runs wherever ``/root/reference`` exists);
  * ``tests/golden/*.npz`` hold outputs/loss/gradient digests produced by the real reference (``oracle/make_golden.py``),
    which this port must reproduce everywhere, and the loss reproduces the reference's own known answers
    (``tests/test_loss.py:109-145``: 0.824 / 0.692 / 0.704).
  * neighbourhood attention follows ``oracle/natten_ref.py`` (natten==0.17.1 is unavailable: unpinned there).

Every function cites the reference lines it follows (paths relative to ``src/cultionet`` in jgrss/cultionet).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import natten_ref

# nn/modules/unet_parts.py:19-40
NATTEN_PARAMS = {
    "a": dict(heads=4, k=3, d=2),
    "b": dict(heads=4, k=3, d=1),
    "c": dict(heads=8, k=3, d=1),
}


# ---------------------------------------------------------------------------------------------------------------------
# parameter inventory (key -> shape) of TowerUNet(in_channels, in_time, hidden, dilations), models/nunet.py:111-211
# ---------------------------------------------------------------------------------------------------------------------
def _bn(spec, p, c):
    spec += [(p + ".weight", (c,)), (p + ".bias", (c,)), (p + ".running_mean", (c,)), (p + ".running_var", (c,)),
             (p + ".num_batches_tracked", ())]


def _conv_block(spec, p, cin, cout, k):  # ConvBlock2d, nn/modules/convolution.py:71-120
    spec.append((p + ".seq.0.weight", (cout, cin, k, k)))
    _bn(spec, p + ".seq.1", cout)


def _resa(spec, p, cin, cout, k, num_blocks, dilations, natten):  # ResidualAConv, convolution.py:250-375
    if cin != cout:
        spec += [(p + ".skip.weight", (cout, cin, 1, 1)), (p + ".skip.bias", (cout,))]
    if natten:
        spec += [(p + ".attention_conv.1.weight", (cout,)), (p + ".attention_conv.1.bias", (cout,)),
                 (p + ".attention_conv.2.qkv.weight", (3 * cout, cout)), (p + ".attention_conv.2.qkv.bias", (3 * cout,)),
                 (p + ".attention_conv.2.proj.weight", (cout, cout)), (p + ".attention_conv.2.proj.bias", (cout,)),
                 (p + ".attention_conv.3.weight", (cout,)), (p + ".attention_conv.3.bias", (cout,))]
    for i, _ in enumerate(dilations):
        for b in range(num_blocks):
            _conv_block(spec, f"{p}.res_modules.{i}.block.{b}", cin if b == 0 else cout, cout, k)


def _convT(spec, p, c):  # ConvTranspose2d, convolution.py:45-68
    spec += [(p + ".up_conv.weight", (c, c, 3, 3)), (p + ".up_conv.bias", (c,))]


def param_spec(in_channels: int, in_time: int, hidden: int, dilations: Sequence[int] = (1, 2)) -> List[Tuple[str, tuple]]:
    h = hidden
    ch = [h, 2 * h, 4 * h, 8 * h]
    up = 4 * h
    dil = list(dilations)
    spec: List[Tuple[str, tuple]] = []
    for name, k in (("conv3", 3), ("conv5", 5)):  # models/nunet.py:18-57
        p = f"pre_unet.{name}.seq"
        spec.append((p + ".0.weight", (in_channels, in_channels, k, 1, 1)))
        _bn(spec, p + ".1", in_channels)
        spec.append((p + ".3.weight", (h, in_channels, in_time - k + 1, 1, 1)))
        _bn(spec, p + ".5", h)
    spec += [("pre_unet.layer_norm.1.weight", (h,)), ("pre_unet.layer_norm.1.bias", (h,))]
    # encoder, unet_parts.py:377-449 (attention_weights=None: nunet.py:156)
    _resa(spec, "encoder.down_a.res_conv", ch[0], ch[0], 3, 2, dil, False)
    for lvl, cin, cout, d in (("b", ch[0], ch[1], dil[:3]), ("c", ch[1], ch[2], dil[:2])):
        _conv_block(spec, f"encoder.down_{lvl}.pool_conv", cin, cout, 3)
        _resa(spec, f"encoder.down_{lvl}.res_conv", cout, cout, 3, 2, d, False)
    _conv_block(spec, "encoder.down_d.pool_conv", ch[2], ch[3], 3)
    _resa(spec, "encoder.down_d.res_conv", ch[3], ch[3], 1, 1, [1], False)
    # decoder, unet_parts.py:452-525 (UNetUpBlock ignores num_blocks on the RESA branch, :355-368)
    _resa(spec, "decoder.over_d.res_conv", ch[3], up, 1, 2, [1], False)
    for lvl, d in (("cu", dil[:2]), ("bu", dil[:3]), ("au", dil)):
        _convT(spec, f"decoder.up_{lvl}.up_conv", up)
        _resa(spec, f"decoder.up_{lvl}.res_conv", up, up, 3, 2, d, True)
    # towers, unet_parts.py:528-760
    for lvl, side, down, tower, d in (("c", ch[2], ch[3], False, dil[:2]), ("b", ch[1], ch[2], True, dil), ("a", ch[0], ch[1], True, dil)):
        p = f"tower_fusion.tower_{lvl}"
        _convT(spec, p + ".backbone_down_conv", down)
        _convT(spec, p + ".decode_down_conv", up)
        cin = side + down + 2 * up
        if tower:
            _convT(spec, p + ".tower_conv", up)
            cin += up
        _resa(spec, p + ".res_conv", cin, up, 3, 2, d, False)
    # heads, unet_parts.py:227-309
    for lvl in ("a", "b", "c"):
        p = f"final_{lvl}"
        if lvl != "a":
            _convT(spec, p + ".up_conv", up)
        for s in ("dist", "edge", "crop"):
            _conv_block(spec, f"{p}.{s}_conv.conv.0", up, 3, 3)
            spec += [(f"{p}.{s}_conv.conv.1.weight", (1, 3, 3, 3)), (f"{p}.{s}_conv.conv.1.bias", (1,))]
        _conv_block(spec, p + ".fuse_conv", 3, 3, 3)
    # final combine, unet_parts.py:101-147
    for s in ("dist", "edge", "crop"):
        spec += [(f"final_combine.final_{s}.0.weight", (1, 1, 1, 1)), (f"final_combine.final_{s}.0.bias", (1,))]
        if s == "edge":
            spec.append(("final_combine.final_edge.1.gamma", (1,)))
        spec += [(f"final_combine.{s}_gamma{i}", (1,)) for i in (1, 2, 3)]
    return spec


def synth_state_dict(spec: Sequence[Tuple[str, tuple]], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic weights from numpy's PCG64 (stable across numpy versions), drawn in sorted-key order, with the scale of
    ``layers/weights.py:24-39`` (Kaiming fan_in weights, N(0,1) biases, N(1,0.02) BN scales) and non-trivial running stats."""
    rng = np.random.default_rng(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in sorted(spec):
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros((), dtype=torch.long)
            continue
        if name.endswith("running_var"):
            v = rng.uniform(0.5, 1.5, size=shape)
        elif name.endswith("running_mean"):
            v = rng.normal(0.0, 0.1, size=shape)
        elif "gamma" in name:
            v = rng.uniform(0.8, 1.25, size=shape)
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            v = rng.normal(0.0, np.sqrt(2.0 / fan_in), size=shape)
        elif name.endswith(".bias") and (".seq.1." in name or ".seq.5." in name or "layer_norm" in name or "attention_conv.1." in name
                                         or "attention_conv.3." in name):
            v = rng.normal(0.0, 0.05, size=shape)  # norm-layer shifts
        elif name.endswith(".bias"):
            v = rng.normal(0.0, 1.0, size=shape)
        else:
            v = rng.normal(1.0, 0.02, size=shape)  # norm-layer scales
        out[name] = torch.from_numpy(np.asarray(v, dtype=np.float32).reshape(shape))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# forward
# ---------------------------------------------------------------------------------------------------------------------
class _Ctx:
    def __init__(self, sd: Dict[str, torch.Tensor], training: bool, bn_out: Optional[dict], activation_type: str = "SiLU"):
        self.sd, self.training, self.bn_out = sd, training, bn_out
        # SetActivation (nn/modules/activations.py:15-24): getattr(torch.nn, activation_type)() with default arguments
        self.act = getattr(torch.nn, activation_type)()

    def bn(self, x, p):  # nn.BatchNorm2d/3d: batch statistics in training, running statistics in eval
        sd = self.sd
        if self.training and self.bn_out is not None:
            dims = [0] + list(range(2, x.dim()))
            n = x.numel() // x.shape[1]
            with torch.no_grad():
                mean = x.mean(dims)
                var = x.var(dims, unbiased=False)
                self.bn_out[p + ".running_mean"] = 0.9 * sd[p + ".running_mean"] + 0.1 * mean
                self.bn_out[p + ".running_var"] = 0.9 * sd[p + ".running_var"] + 0.1 * var * (n / max(n - 1, 1))
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                            training=self.training, momentum=0.0, eps=1e-5)


_SAFE_CONV_ELEMS = 1 << 26


def _conv2d(x, w, b, **kw):
    """``F.conv2d``.  On a GPU, inputs above 2**26 elements are convolved in batch chunks and concatenated: torch 2.11 / cuDNN 9 on
    the B200 was caught returning a wrong fp32 result for one BATCHED call of this path (``[8, 960, 140, 140] * [256, 960, 3, 3]``,
    93 % off its own per-sample and fp64 results, which agree with each other and with both CUDA kernels of the product to 2e-6;
    ``tools/debug_eval64b.py``).  Convolution is independent per sample, so the chunked form is the same function."""
    n = x.shape[0]
    if x.is_cuda and n > 1 and x.numel() > _SAFE_CONV_ELEMS:
        step = max(1, int(_SAFE_CONV_ELEMS // (x.numel() // n)))
        if step < n:
            return torch.cat([F.conv2d(x[i:i + step], w, b, **kw) for i in range(0, n, step)], dim=0)
    return F.conv2d(x, w, b, **kw)


def _conv_block_fwd(c: _Ctx, x, p, k, stride=1, pad=None, dil=1, act=True):  # convolution.py:71-120
    pad = (0 if k == 1 else k // 2) if pad is None else pad
    y = _conv2d(x, c.sd[p + ".seq.0.weight"], None, stride=stride, padding=pad, dilation=dil)
    y = c.bn(y, p + ".seq.1")
    return c.act(y) if act else y


def _natten_block(c: _Ctx, skip, p, heads, k, d):  # convolution.py:338-353 + natten 0.17.1 module
    sd = c.sd
    C = skip.shape[1]
    t = skip.permute(0, 2, 3, 1)
    t = F.layer_norm(t, (C,), sd[p + ".1.weight"], sd[p + ".1.bias"], 1e-5)
    B, H, W, _ = t.shape
    hd = C // heads
    qkv = F.linear(t, sd[p + ".2.qkv.weight"], sd[p + ".2.qkv.bias"]).reshape(B, H, W, 3, heads, hd).permute(3, 0, 4, 1, 2, 5)
    q, kk, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
    attn = natten_ref.na2d_qk(q, kk, k, d).softmax(dim=-1)
    o = natten_ref.na2d_av(attn, v, k, d).permute(0, 2, 3, 1, 4).reshape(B, H, W, C)
    o = F.linear(o, sd[p + ".2.proj.weight"], sd[p + ".2.proj.bias"])
    o = F.layer_norm(o, (C,), sd[p + ".3.weight"], sd[p + ".3.bias"], 1e-5)
    return o.permute(0, 3, 1, 2)


def _resa_fwd(c: _Ctx, x, p, k, num_blocks, dilations, natten=None):  # convolution.py:377-395, :142-167
    sd = c.sd
    out = _conv2d(x, sd[p + ".skip.weight"], sd[p + ".skip.bias"]) if (p + ".skip.weight") in sd else x
    skip = out
    for i, d in enumerate(dilations):
        h = x
        for b in range(num_blocks):
            later = 1 if k == 1 else max(1, d - 1)
            h = _conv_block_fwd(c, h, f"{p}.res_modules.{i}.block.{b}", k, pad=(0 if k == 1 else (k // 2 if b == 0 else later)),
                                dil=1 if b == 0 else later)
        out = out + h
    if natten is not None:
        out = out + _natten_block(c, skip, p + ".attention_conv", **natten)
    return out


def _convT_fwd(c: _Ctx, x, p, size, stride=2):  # convolution.py:56-68 + nn/functional.py:72-81
    y = F.conv_transpose2d(x, c.sd[p + ".up_conv.weight"], c.sd[p + ".up_conv.bias"], stride=stride, padding=1)
    if tuple(y.shape[-2:]) != tuple(size):
        y = F.interpolate(y, size=tuple(size), mode="bilinear", align_corners=True)
    return y


def _pre_unet(c: _Ctx, x):  # models/nunet.py:18-105
    sd = c.sd
    outs = []
    for name in ("conv3", "conv5"):
        p = f"pre_unet.{name}.seq"
        h = F.conv3d(x, sd[p + ".0.weight"])
        h = c.act(c.bn(h, p + ".1"))
        h = F.conv3d(h, sd[p + ".3.weight"]).squeeze(2)
        outs.append(c.act(c.bn(h, p + ".5")))
    s = (outs[0] + outs[1]).permute(0, 2, 3, 1)
    s = F.layer_norm(s, (s.shape[-1],), sd["pre_unet.layer_norm.1.weight"], sd["pre_unet.layer_norm.1.bias"], 1e-5)
    return s.permute(0, 3, 1, 2)


def _final(c: _Ctx, x, p, size=None, stride=2):  # unet_parts.py:281-309
    sd = c.sd
    if size is not None:
        x = _convT_fwd(c, x, p + ".up_conv", size, stride=stride)
    hs = []
    for s in ("dist", "edge", "crop"):
        h = _conv_block_fwd(c, x, f"{p}.{s}_conv.conv.0", 3)
        hs.append(F.conv2d(h, sd[f"{p}.{s}_conv.conv.1.weight"], sd[f"{p}.{s}_conv.conv.1.bias"], padding=1))
    return _conv_block_fwd(c, torch.cat(hs, dim=1), p + ".fuse_conv", 3)


def towerunet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, dilations: Sequence[int] = (1, 2), training: bool = True,
                      bn_out: Optional[dict] = None, natten_params: Optional[dict] = None, taps: Optional[dict] = None,
                      activation_type: str = "SiLU") -> Dict[str, torch.Tensor]:
    """TowerUNet.forward (models/nunet.py:213-265).  ``bn_out`` (optional dict) receives the updated running statistics;
    ``taps`` (optional dict) receives named intermediate activations in NCHW."""
    c = _Ctx(sd, training, bn_out, activation_type)
    nat = natten_params or NATTEN_PARAMS
    dil = list(dilations)
    e0 = _pre_unet(c, x)
    # encoder (unet_parts.py:437-449; PoolResidualConv convolution.py:484-513)
    x_a = _resa_fwd(c, e0, "encoder.down_a.res_conv", 3, 2, dil)
    x_b = _resa_fwd(c, _conv_block_fwd(c, x_a, "encoder.down_b.pool_conv", 3, stride=2, pad=1, act=False), "encoder.down_b.res_conv", 3, 2, dil[:3])
    x_c = _resa_fwd(c, _conv_block_fwd(c, x_b, "encoder.down_c.pool_conv", 3, stride=2, pad=1, act=False), "encoder.down_c.res_conv", 3, 2, dil[:2])
    x_d = _resa_fwd(c, _conv_block_fwd(c, x_c, "encoder.down_d.pool_conv", 3, stride=2, pad=1, act=False), "encoder.down_d.res_conv", 1, 1, [1])
    # decoder (unet_parts.py:510-525)
    hw = lambda t: tuple(t.shape[-2:])  # noqa: E731
    x_du = _resa_fwd(c, x_d, "decoder.over_d.res_conv", 1, 2, [1])
    x_cu = _resa_fwd(c, _convT_fwd(c, x_du, "decoder.up_cu.up_conv", hw(x_c)), "decoder.up_cu.res_conv", 3, 2, dil[:2], nat["c"])
    x_bu = _resa_fwd(c, _convT_fwd(c, x_cu, "decoder.up_bu.up_conv", hw(x_b)), "decoder.up_bu.res_conv", 3, 2, dil[:3], nat["b"])
    x_au = _resa_fwd(c, _convT_fwd(c, x_bu, "decoder.up_au.up_conv", hw(x_a)), "decoder.up_au.res_conv", 3, 2, dil, nat["a"])

    # towers (unet_parts.py:576-612, :714-760)
    def tower(p, side, down, dside, ddown, tdown, d):
        size = hw(dside)
        parts = [side, _convT_fwd(c, down, p + ".backbone_down_conv", size), dside, _convT_fwd(c, ddown, p + ".decode_down_conv", size)]
        if tdown is not None:
            parts.append(_convT_fwd(c, tdown, p + ".tower_conv", size))
        return _resa_fwd(c, torch.cat(parts, dim=1), p + ".res_conv", 3, 2, d)

    t_c = tower("tower_fusion.tower_c", x_c, x_d, x_cu, x_du, None, dil[:2])
    t_b = tower("tower_fusion.tower_b", x_b, x_c, x_bu, x_cu, t_c, dil)
    t_a = tower("tower_fusion.tower_a", x_a, x_b, x_au, x_bu, t_b, dil)
    h_a = _final(c, t_a, "final_a")
    h_b = _final(c, t_b, "final_b", size=hw(t_a), stride=2)
    h_c = _final(c, t_c, "final_c", size=hw(t_a), stride=4)
    # final combine (unet_parts.py:148-193) and SigmoidCrisp (:86-98)
    out = {}
    for i, (s, key) in enumerate((("dist", "distance"), ("edge", "edge"), ("crop", "crop"))):
        g = [sd[f"final_combine.{s}_gamma{j}"] for j in (1, 2, 3)]
        z = h_a[:, i:i + 1] / g[0] + h_b[:, i:i + 1] / g[1] + h_c[:, i:i + 1] / g[2]
        z = F.conv2d(z, sd[f"final_combine.final_{s}.0.weight"], sd[f"final_combine.final_{s}.0.bias"])
        if s == "edge":
            z = z * torch.reciprocal(1e-2 + torch.sigmoid(sd["final_combine.final_edge.1.gamma"]))
        out[key] = torch.sigmoid(z)
    if taps is not None:
        taps.update(e0=e0, x_a=x_a, x_b=x_b, x_c=x_c, x_d=x_d, x_du=x_du, x_cu=x_cu, x_bu=x_bu, x_au=x_au, t_c=t_c, t_b=t_b, t_a=t_a,
                    h_a=h_a, h_b=h_b, h_c=h_c)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# loss
# ---------------------------------------------------------------------------------------------------------------------
def tanimoto_distance(y: torch.Tensor, yhat: torch.Tensor, smooth: float = 1e-5, depth: int = 5) -> torch.Tensor:
    """losses/losses.py:152-184 with dim=(1,2,3)."""
    tpl = (y * yhat).sum(dim=(1, 2, 3))
    sq = (y ** 2 + yhat ** 2).sum(dim=(1, 2, 3))
    den = 0.0
    for d in range(depth):
        a = 2.0 ** d
        b = -(2.0 * a - 1.0)
        den = den + torch.reciprocal(a * sq + b * tpl + smooth)
    return 1.0 - (tpl + smooth) * den * (1.0 / depth)


def tanimoto_complement_loss(inputs, targets, mask=None, one_hot_targets=True, smooth=1e-5, depth=5) -> torch.Tensor:
    """TanimotoComplementLoss.forward (losses.py:186-218) after LossPreprocessing (:19-59), transform_logits=False."""
    if one_hot_targets and inputs.shape[1] > 1:
        targets = F.one_hot(targets, num_classes=inputs.shape[1]).permute(0, 3, 1, 2)
    elif targets.dim() == 3:
        targets = targets.unsqueeze(1)
    if mask is not None:
        if mask.dim() == 3:
            mask = mask.unsqueeze(1)
        inputs = inputs * mask
        targets = targets * mask
    l1 = tanimoto_distance(targets, inputs, smooth, depth)
    l2 = tanimoto_distance(1.0 - targets, 1.0 - inputs, smooth, depth)
    return ((l1 + l2) * 0.5).mean()


def training_loss(pred: Dict[str, torch.Tensor], y: torch.Tensor, bdist: torch.Tensor, edge_class: int = 2):
    """get_true_labels + calc_loss (models/lightning.py:161-207, :318-354) for LossTypes.TANIMOTO_COMPLEMENT."""
    true_edge = (y == edge_class).long()
    true_crop = ((y > 0) & (y < edge_class)).long()
    mask = None
    if int(y.min()) == -1:
        mask = (y != -1).long().unsqueeze(1)
    d = tanimoto_complement_loss(pred["distance"], bdist, mask, one_hot_targets=False)
    e = tanimoto_complement_loss(pred["edge"], true_edge, mask)
    c = tanimoto_complement_loss(pred["crop"], true_crop, mask)
    return (d + e + c) / 3.0, {"dloss": d, "eloss": e, "closs": c}


# ---------------------------------------------------------------------------------------------------------------------
# optional block variants (constructor arguments the reference's own tests use, tests/test_cultionet.py:67-78)
# ---------------------------------------------------------------------------------------------------------------------
def spatial_channel_attention(x: torch.Tensor, fc1_0, fc1_2, fc2_0, fc2_2, sp_conv, gamma) -> torch.Tensor:
    """SpatialChannelAttention.forward over NCHW ``x`` (nn/modules/attention.py:54-63, :78-88, :118-125): returns the
    ``1 + gamma * attention`` map that ResidualAConv multiplies its output with (nn/modules/convolution.py:392-393)."""
    avg = F.conv2d(F.silu(F.conv2d(F.adaptive_avg_pool2d(x, 1), fc1_0)), fc1_2)  # attention.py:56
    mx = F.conv2d(F.silu(F.conv2d(F.adaptive_max_pool2d(x, 1), fc2_0)), fc2_2)  # attention.py:57
    channel = torch.sigmoid(avg + mx)  # attention.py:58-61
    sp = torch.cat([x.mean(dim=1, keepdim=True), x.amax(dim=1, keepdim=True)], dim=1)  # attention.py:81-83 (einops reduce -> amax)
    spatial = torch.sigmoid(F.conv2d(sp, sp_conv, padding=1))  # attention.py:84-85
    return 1.0 + gamma * ((channel + spatial) * 0.5)  # attention.py:121-123


def adaptive_max_pool_half(x: torch.Tensor) -> torch.Tensor:
    """PoolResidualConv's ``pool_by_max`` down-sampling over NCHW ``x`` (nn/modules/convolution.py:499-503)."""
    return F.adaptive_max_pool2d(x, output_size=(x.shape[-2] // 2, x.shape[-1] // 2))
