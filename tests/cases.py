"""Operator and model parity cases shared by the CPU-interpreter tests and the `-m gpu` tests.

Operator checks compare a kernel with the plain PyTorch fp32 op of the same name on the same device; in bf16 mode the inputs
are rounded to bf16 first and the comparison tolerance is the bf16 one.  Model checks compare with the oracle port."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as TF

from cultionet_b200 import functional as F
from oracle import natten_ref
from tests.util import rel_err


def _tol(dtype):
    return 2e-5 if dtype == torch.float32 else 1.5e-2


def _mk(shape, dev, dtype, grad=True, scale=1.0, shift=0.0):
    t = (torch.randn(*shape, device=dev) * scale + shift).to(dtype)
    return t.requires_grad_(grad)


def _f(t):
    return t.detach().float().requires_grad_(t.requires_grad)


def _check(tag, got, want, tol):
    e = rel_err(got.float(), want)
    assert e < tol, f"{tag}: rel err {e:.3e} >= {tol:.1e}"


def conv_case(dev, dtype, B, H, W, cins, cout, k, stride, pad, dil, seed=0):
    torch.manual_seed(seed)
    xs = [_mk((B, H, W, c), dev, dtype) for c in cins]
    w = torch.randn(cout, sum(cins), k, k, device=dev) / (sum(cins) * k * k) ** 0.5
    w.requires_grad_(True)
    b = torch.randn(cout, device=dev, requires_grad=True)
    y = F.conv2d(xs, w, b, k, stride, pad, dil)
    xr = [_f(x) for x in xs]
    yr = TF.conv2d(torch.cat(xr, -1).permute(0, 3, 1, 2), w, b, stride=stride, padding=pad, dilation=dil).permute(0, 2, 3, 1)
    tol = _tol(dtype)
    _check("conv fwd", y, yr, tol)
    g = torch.randn_like(yr)
    gm = torch.autograd.grad(y, [*xs, w, b], g.to(dtype))
    gr = torch.autograd.grad(yr, [*xr, w, b], g.to(dtype).float())
    for i, (a, c) in enumerate(zip(gm, gr)):
        _check(f"conv grad {i}", a, c, tol)


def convT_case(dev, dtype, B, H, W, cin, cout, stride, seed=0):
    torch.manual_seed(seed)
    x = _mk((B, H, W, cin), dev, dtype)
    w = torch.randn(cin, cout, 3, 3, device=dev) / (cin * 9) ** 0.5
    w.requires_grad_(True)
    b = torch.randn(cout, device=dev, requires_grad=True)
    y = F.conv_transpose2d(x, w, b, 3, stride, 1)
    xr = _f(x)
    yr = TF.conv_transpose2d(xr.permute(0, 3, 1, 2), w, b, stride=stride, padding=1).permute(0, 2, 3, 1)
    tol = _tol(dtype)
    _check("convT fwd", y, yr, tol)
    g = torch.randn_like(yr)
    for i, (a, c) in enumerate(zip(torch.autograd.grad(y, [x, w, b], g.to(dtype)), torch.autograd.grad(yr, [xr, w, b], g.to(dtype).float()))):
        _check(f"convT grad {i}", a, c, tol)


def upconv_case(dev, dtype, B, H, W, C, direct: bool, seed=0):
    """``nn.modules.convolution.ConvTranspose2d``: ConvTranspose2d(k3, s2, p1) then the bilinear fix-up to 2H x 2W, whose backward also
    produces the bias gradient (cnb_resize_bilinear_bwd_colsum); ``direct``: gradients written straight into ``param.grad``."""
    from cultionet_b200.nn.modules.convolution import ConvTranspose2d

    torch.manual_seed(seed)
    m = ConvTranspose2d(C, C).to(dev)
    x = _mk((B, H, W, C), dev, dtype)
    g = torch.randn(B, 2 * H, 2 * W, C, device=dev).to(dtype)
    xr = _f(x)
    w, b = m.up_conv.weight, m.up_conv.bias
    yr = TF.conv_transpose2d(xr.permute(0, 3, 1, 2), w, b, stride=2, padding=1)
    yr = TF.interpolate(yr, size=(2 * H, 2 * W), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    gr = torch.autograd.grad(yr, [xr, w, b], g.float())
    y = m(x, size=(2 * H, 2 * W))
    tol = _tol(dtype)
    _check("upconv fwd", y, yr, tol)
    if direct:
        for p in (w, b):
            p.grad = torch.full_like(p, 7.0)  # stale contents: the first contribution overwrites
        with F.direct_param_grads():
            y.backward(g, inputs=[x, w, b])
        got = [x.grad, w.grad, b.grad]
    else:
        got = torch.autograd.grad(y, [x, w, b], g)
    for i, (a, c) in enumerate(zip(got, gr)):
        _check(f"upconv grad {i}", a, c, tol * (2 if dtype == torch.bfloat16 else 1))


def linear_case(dev, dtype, B, H, W, cin, cout, seed=0):
    torch.manual_seed(seed)
    x = _mk((B, H, W, cin), dev, dtype)
    w = (torch.randn(cout, cin, device=dev) / cin ** 0.5).requires_grad_(True)
    b = torch.randn(cout, device=dev, requires_grad=True)
    y = F.linear(x, w, b)
    xr = _f(x)
    yr = TF.linear(xr, w, b)
    tol = _tol(dtype)
    _check("linear fwd", y, yr, tol)
    g = torch.randn_like(yr)
    for i, (a, c) in enumerate(zip(torch.autograd.grad(y, [x, w, b], g.to(dtype)), torch.autograd.grad(yr, [xr, w, b], g.to(dtype).float()))):
        _check(f"linear grad {i}", a, c, tol)


def batchnorm_case(dev, dtype, B, H, W, C, training=True, act=True, with_res=True, seed=0):
    torch.manual_seed(seed)
    x = _mk((B, H, W, C), dev, dtype, scale=2.0, shift=0.7)
    gam = (torch.rand(C, device=dev) + 0.5).requires_grad_(True)
    bet = torch.randn(C, device=dev, requires_grad=True)
    rm, rv = torch.randn(C, device=dev) * 0.1, torch.rand(C, device=dev) + 0.5
    rm2, rv2 = rm.clone(), rv.clone()
    res = _mk((B, H, W, C), dev, dtype) if with_res else None
    y = F.batchnorm_act(x, gam, bet, rm, rv, training, 0.1, 1e-5, act, 1, res)
    xr = _f(x)
    rr = _f(res) if with_res else None
    z = TF.batch_norm(xr.permute(0, 3, 1, 2), rm2, rv2, gam, bet, training, 0.1, 1e-5).permute(0, 2, 3, 1)
    yr = TF.silu(z) if act else z
    if with_res:
        yr = yr + rr
    tol = _tol(dtype)
    _check("bn fwd", y, yr, tol)
    if training:
        _check("bn running_mean", rm, rm2, 1e-4)
        _check("bn running_var", rv, rv2, 1e-4)
    g = torch.randn_like(yr)
    ins_m = [x, gam, bet] + ([res] if with_res else [])
    ins_r = [xr, gam, bet] + ([rr] if with_res else [])
    for i, (a, c) in enumerate(zip(torch.autograd.grad(y, ins_m, g.to(dtype)), torch.autograd.grad(yr, ins_r, g.to(dtype).float()))):
        _check(f"bn grad {i}", a, c, tol * (3 if dtype == torch.bfloat16 else 5))


def activation_case(dev, dtype, name, seed=0):
    """Every activation code: stand-alone kernel and the fused BatchNorm + activation (+ residual) kernels, forward and backward, against
    ``getattr(torch.nn, name)()`` -- what the reference's SetActivation builds (nn/modules/activations.py:15-24)."""
    torch.manual_seed(seed)
    code = F.act_code(name)
    ref = getattr(torch.nn, name)()
    tol = _tol(dtype)
    x = _mk((2, 7, 9, 16), dev, dtype, scale=2.5)
    y = F.activation(x, code)
    xr = _f(x)
    yr = ref(xr)
    _check(f"{name} fwd", y, yr, tol)
    g = torch.randn_like(yr)
    _check(f"{name} grad", torch.autograd.grad(y, x, g.to(dtype))[0], torch.autograd.grad(yr, xr, g.to(dtype).float())[0], tol * 2)
    # fused with BatchNorm (training statistics) and a residual
    x = _mk((2, 9, 11, 16), dev, dtype, scale=2.0, shift=0.4)
    gam = (torch.rand(16, device=dev) + 0.5).requires_grad_(True)
    bet = torch.randn(16, device=dev, requires_grad=True)
    res = _mk((2, 9, 11, 16), dev, dtype)
    rm, rv = torch.zeros(16, device=dev), torch.ones(16, device=dev)
    y = F.batchnorm_act(x, gam, bet, rm, rv, True, 0.1, 1e-5, code, 1, res)
    xr, rr = _f(x), _f(res)
    yr = ref(TF.batch_norm(xr.permute(0, 3, 1, 2), None, None, gam, bet, True, 0.1, 1e-5).permute(0, 2, 3, 1)) + rr
    _check(f"bn+{name} fwd", y, yr, tol)
    g = torch.randn_like(yr)
    for i, (a, c) in enumerate(zip(torch.autograd.grad(y, [x, gam, bet, res], g.to(dtype)),
                                   torch.autograd.grad(yr, [xr, gam, bet, rr], g.to(dtype).float()))):
        _check(f"bn+{name} grad {i}", a, c, tol * (4 if dtype == torch.bfloat16 else 5))


def batchnorm3d_case(dev, dtype, B, H, W, C, Tp, seed=0):
    torch.manual_seed(seed)
    u = _mk((B, H, W, C * Tp), dev, dtype)
    gam = (torch.rand(C, device=dev) + 0.5).requires_grad_(True)
    bet = torch.randn(C, device=dev, requires_grad=True)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    y = F.batchnorm_act(u, gam, bet, rm, rv, True, 0.1, 1e-5, True, Tp, None)
    ur = _f(u)
    yr = TF.silu(TF.batch_norm(ur.reshape(B, H, W, C, Tp).permute(0, 3, 4, 1, 2), None, None, gam, bet, True, 0.1, 1e-5))
    yr = yr.permute(0, 3, 4, 1, 2).reshape(B, H, W, C * Tp)
    tol = _tol(dtype)
    _check("bn3d fwd", y, yr, tol)
    g = torch.randn_like(yr)
    for i, (a, c) in enumerate(zip(torch.autograd.grad(y, [u, gam, bet], g.to(dtype)), torch.autograd.grad(yr, [ur, gam, bet], g.to(dtype).float()))):
        _check(f"bn3d grad {i}", a, c, tol * 5)


def layernorm_case(dev, dtype, B, H, W, C, seed=0):
    torch.manual_seed(seed)
    x = _mk((B, H, W, C), dev, dtype, scale=1.5, shift=0.3)
    gam = torch.randn(C, device=dev, requires_grad=True)
    bet = torch.randn(C, device=dev, requires_grad=True)
    y = F.layernorm(x, gam, bet, 1e-5)
    xr = _f(x)
    yr = TF.layer_norm(xr, (C,), gam, bet, 1e-5)
    tol = _tol(dtype)
    _check("ln fwd", y, yr, tol)
    g = torch.randn_like(yr)
    for i, (a, c) in enumerate(zip(torch.autograd.grad(y, [x, gam, bet], g.to(dtype)), torch.autograd.grad(yr, [xr, gam, bet], g.to(dtype).float()))):
        _check(f"ln grad {i}", a, c, tol * 3)


def na_case(dev, dtype, B, H, W, heads, hd, k, d, seed=0):
    torch.manual_seed(seed)
    Cn = heads * hd
    qkv = _mk((B, H, W, 3 * Cn), dev, dtype)
    y = F.na2d(qkv, heads, k, d, hd ** -0.5)
    qr = _f(qkv)
    t = qr.reshape(B, H, W, 3, heads, hd).permute(3, 0, 4, 1, 2, 5)
    a = natten_ref.na2d_qk(t[0] * hd ** -0.5, t[1], k, d).softmax(-1)
    yr = natten_ref.na2d_av(a, t[2], k, d).permute(0, 2, 3, 1, 4).reshape(B, H, W, Cn)
    tol = _tol(dtype)
    _check("na fwd", y, yr, tol)
    g = torch.randn_like(yr)
    _check("na grad", torch.autograd.grad(y, qkv, g.to(dtype))[0], torch.autograd.grad(yr, qr, g.to(dtype).float())[0], tol * 2)


def na_module_case(dev, dtype, B, H, W, dim, heads, k, d, direct: bool, seed=0):
    """``nn.modules.attention.NeighborhoodAttention2D`` (qkv Linear -> neighbourhood attention -> proj Linear) against the same arithmetic
    in torch fp32 (oracle/natten_ref.py), every parameter gradient included.  ``direct``: parameter gradients written straight into
    ``param.grad``."""
    from cultionet_b200.nn.modules.attention import NeighborhoodAttention2D

    torch.manual_seed(seed)
    m = NeighborhoodAttention2D(dim, heads, k, d).to(dev)
    hd = dim // heads
    x = _mk((B, H, W, dim), dev, dtype)
    g = torch.randn(B, H, W, dim, device=dev).to(dtype)
    xr = _f(x)
    ps = [m.qkv.weight, m.qkv.bias, m.proj.weight, m.proj.bias]
    qkv = TF.linear(xr, m.qkv.weight, m.qkv.bias)
    if dtype == torch.bfloat16:
        qkv = qkv.to(dtype).float()  # the product stores qkv in bf16 between the projection and the attention
    t = qkv.reshape(B, H, W, 3, heads, hd).permute(3, 0, 4, 1, 2, 5)
    a = natten_ref.na2d_qk(t[0] * hd ** -0.5, t[1], k, d).softmax(-1)
    o = natten_ref.na2d_av(a, t[2], k, d).permute(0, 2, 3, 1, 4).reshape(B, H, W, dim)
    yr = TF.linear(o, m.proj.weight, m.proj.bias)
    gr = torch.autograd.grad(yr, [xr, *ps], g.float())
    m.train()
    y = m(x)
    tol = _tol(dtype)
    _check("na module fwd", y, yr, tol * 2)
    if direct:
        for p in ps:
            p.grad = torch.full_like(p, 3.0)  # stale contents: the first contribution overwrites
        with F.direct_param_grads():
            y.backward(g, inputs=[x, *ps])
        got = [x.grad] + [p.grad for p in ps]
    else:
        got = torch.autograd.grad(y, [x, *ps], g)
    for i, (a_, c) in enumerate(zip(got, gr)):
        _check(f"na module grad {i}", a_, c, tol * 3)


def resize_case(dev, dtype, B, hi, wi, ho, wo, C, seed=0):
    torch.manual_seed(seed)
    x = _mk((B, hi, wi, C), dev, dtype)
    y = F.resize_bilinear(x, (ho, wo))
    xr = _f(x)
    yr = TF.interpolate(xr.permute(0, 3, 1, 2), size=(ho, wo), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    tol = _tol(dtype)
    _check("resize fwd", y, yr, tol)
    g = torch.randn_like(yr)
    _check("resize grad", torch.autograd.grad(y, x, g.to(dtype))[0], torch.autograd.grad(yr, xr, g.to(dtype).float())[0], tol)


def pretime_case(dev, dtype, B, C, T, H, W, k, seed=0):
    torch.manual_seed(seed)
    x = torch.randn(B, C, T, H, W, device=dev)
    w1 = torch.randn(C, C, k, 1, 1, device=dev, requires_grad=True)
    u = F.pretime_conv(x, w1, dtype)
    Tp = T - k + 1
    ur = TF.conv3d(x, w1).permute(0, 3, 4, 1, 2).reshape(B, H, W, C * Tp)
    tol = _tol(dtype)
    assert u.shape[-1] % 8 == 0 and u.shape[-1] >= C * Tp
    assert float(u[..., C * Tp:].float().abs().sum()) == 0.0  # row padding is zero
    _check("pretime fwd", u[..., :C * Tp], ur, tol)
    g = torch.randn_like(ur)
    gp = torch.zeros_like(u, dtype=torch.float32)
    gp[..., :C * Tp] = g
    _check("pretime wgrad", torch.autograd.grad(u, w1, gp.to(dtype))[0], torch.autograd.grad(ur, w1, g.to(dtype).float())[0], tol)


def conv_skinny_case(dev, dtype, B, H, W, cin, cout, k=3, pad=1, dil=1, seed=0):
    """1x1 GEMM + shift-and-add form of a skinny convolution against torch's conv2d (forward, data and weight gradients)."""
    torch.manual_seed(seed)
    x = _mk((B, H, W, cin), dev, dtype)
    w = (torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5).requires_grad_(True)
    y = F.conv2d_skinny(x, w, ksize=k, pad=pad, dil=dil)
    xr = _f(x)
    yr = TF.conv2d(xr.permute(0, 3, 1, 2), w, None, stride=1, padding=pad, dilation=dil).permute(0, 2, 3, 1)
    tol = _tol(dtype)
    assert y.shape == yr.shape
    _check("skinny conv fwd", y, yr, tol)
    g = torch.randn_like(yr)
    for i, (a, c) in enumerate(zip(torch.autograd.grad(y, [x, w], g.to(dtype)), torch.autograd.grad(yr, [xr, w], g.to(dtype).float()))):
        _check(f"skinny conv grad {i}", a, c, tol * 2)


def pretime_gemm_case(dev, dtype, B, C, T, H, W, k, seed=0):
    """The banded-GEMM form of the temporal convolution (throughput mode) against torch's conv3d."""
    torch.manual_seed(seed)
    x = torch.randn(B, C, T, H, W, device=dev)
    w1 = torch.randn(C, C, k, 1, 1, device=dev, requires_grad=True)
    xp = F.time_to_pixel_major(x, dtype)
    CT, Tp = C * T, T - k + 1
    assert xp.shape[-1] % 8 == 0 and xp.shape[-1] >= CT
    xr = x.to(dtype).float()  # the GEMM reads x in the compute dtype
    _check("time->pixel major", xp[..., :CT], xr.permute(0, 3, 4, 1, 2).reshape(B, H, W, CT), 1e-7)
    assert float(xp[..., CT:].float().abs().sum()) == 0.0
    u = F.pretime_conv_gemm(xp, w1, T)
    ur = TF.conv3d(xr, w1).permute(0, 3, 4, 1, 2).reshape(B, H, W, C * Tp)
    tol = _tol(dtype)
    assert u.shape[-1] % 8 == 0 and u.shape[-1] >= C * Tp
    assert float(u[..., C * Tp:].float().abs().sum()) == 0.0  # row padding is zero
    _check("pretime gemm fwd", u[..., :C * Tp], ur, tol)
    g = torch.randn_like(ur)
    gp = torch.zeros_like(u, dtype=torch.float32)
    gp[..., :C * Tp] = g
    _check("pretime gemm wgrad", torch.autograd.grad(u, w1, gp.to(dtype))[0], torch.autograd.grad(ur, w1, g.to(dtype).float())[0], tol)


def final_combine_case(dev, dtype, B, H, W, edge_activation=True, mask_activation=True, seed=0):
    torch.manual_seed(seed)
    hs = [_mk((B, H, W, 3), dev, dtype) for _ in range(3)]
    params = [(torch.rand(1, device=dev) * 0.5 + 0.75).requires_grad_(True) for _ in range(9)]
    params += [torch.randn(1, 1, 1, 1, device=dev, requires_grad=True) for _ in range(3)]
    params += [torch.randn(1, device=dev, requires_grad=True) for _ in range(3)]
    params += [torch.randn(1, device=dev, requires_grad=True)]
    outs = F.final_combine(*hs, params, 1e-2, edge_activation, mask_activation)
    hr = [_f(h) for h in hs]
    refs = []
    for t in range(3):
        z = sum(hr[j][..., t] / params[t * 3 + j] for j in range(3)) * params[9 + t].reshape(()) + params[12 + t]
        if t == 1 and edge_activation:
            z = torch.sigmoid(z / (1e-2 + torch.sigmoid(params[15])))
        elif t == 0 or (t == 2 and mask_activation):
            z = torch.sigmoid(z)
        refs.append(z.unsqueeze(1))
    tol = _tol(dtype)
    gs = [torch.randn_like(r) for r in refs]
    for o, r in zip(outs, refs):
        assert o.shape == r.shape
        _check("combine fwd", o, r, tol)
    gm = torch.autograd.grad(outs, [*hs, *params], gs, allow_unused=True)
    gr = torch.autograd.grad(refs, [*hr, *params], gs, allow_unused=True)
    for i, (a, c) in enumerate(zip(gm, gr)):
        if c is None:
            continue
        _check(f"combine grad {i}", a, c, tol * 3)


def loss_case(dev, B, H, W, y_low=-1, seed=0):
    from cultionet_b200.losses import tower_unet_loss
    from oracle import towerunet_port as port

    torch.manual_seed(seed)
    preds = {k: torch.rand(B, 1, H, W, device=dev, requires_grad=True) for k in ("distance", "edge", "crop")}
    y = torch.randint(y_low, 3, (B, H, W), device=dev)
    bdist = torch.rand(B, H, W, device=dev)
    total, parts = tower_unet_loss(preds, y, bdist)
    want, wparts = port.training_loss(preds, y, bdist)
    assert abs(float(total) - float(want)) < 1e-5 * max(1.0, abs(float(want)))
    assert np.allclose(parts[1:].cpu().numpy(), [float(wparts[k]) for k in ("dloss", "eloss", "closs")], rtol=1e-5)
    gm = torch.autograd.grad(total, list(preds.values()))
    gr = torch.autograd.grad(want, list(preds.values()))
    for a, c in zip(gm, gr):
        _check("loss grad", a, c, 1e-4)


def adamw_case(dev, n=10007, steps=3, clip=1.0, seed=0):
    import ctypes as C

    from cultionet_b200 import _lib

    torch.manual_seed(seed)
    p = torch.randn(n, device=dev)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=0.01, weight_decay=1e-3, eps=1e-4, betas=(0.9, 0.98))
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    hyper = torch.zeros(2, device=dev)
    ws = torch.zeros(1, device=dev)
    for s in range(1, steps + 1):
        g = torch.randn(n, device=dev) * (3.0 if s == 1 else 0.001)
        pr.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([pr], clip)
        opt.step()
        hyper.copy_(torch.tensor([0.01, float(s)]))
        st = _lib.stream_ptr(p)
        _lib.call("cnb_grad_sqnorm", _lib.ptr(g), n, _lib.ptr(ws), st)
        _lib.call("cnb_adamw_step", _lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), n, _lib.ptr(hyper), 0.9, 0.98, 1e-4, 1e-3, 1.0, clip,
                  _lib.ptr(ws), st)
        assert rel_err(p, pr) < 1e-5, f"adamw step {s}"


def model_vs_golden(dev, name, dtype=torch.float32):
    """The product model against the golden vectors the REAL reference produced (tests/golden, oracle/make_golden.py)."""
    from cultionet_b200.losses import tower_unet_loss
    from oracle.make_golden import FULL_GRADS, golden_case, golden_latlon
    from tests.util import MASK_AGREEMENT, TOL_GRAD_FP32, TOL_OUT_BF16, TOL_OUT_FP32, golden_spec, load_golden, mine_from_state_dict

    cfg, z = load_golden(name)
    spec, sd, x, y, bdist = golden_case(cfg, golden_spec(z))
    model = mine_from_state_dict(cfg, sd, dev, dtype).train()
    latlon = golden_latlon(cfg)
    out = model(x.to(dev), latlon_coords=None if latlon is None else latlon.to(dev))
    tol = TOL_OUT_FP32 if dtype == torch.float32 else TOL_OUT_BF16
    for k in ("distance", "edge", "crop"):
        got = out[k][:, :, ::3, ::3]
        want = torch.from_numpy(z["out_" + k])
        assert rel_err(got, want) < tol, (k, rel_err(got, want))
    crop_agree = ((out["crop"][:, :, ::3, ::3].cpu() > 0.5) == (torch.from_numpy(z["out_crop"]) > 0.5)).float().mean()
    if dtype == torch.float32:
        assert float(crop_agree) >= MASK_AGREEMENT
    else:
        from tests.util import MASK_AGREEMENT_BF16_RANDOM_INIT

        # the fixtures store ~100 pixels (stride-3 grid): one pixel is 1 %, so the random-init bar is counted in pixels here
        n_px = z["out_crop"].size
        assert round((1.0 - float(crop_agree)) * n_px) <= max(2, int((1.0 - MASK_AGREEMENT_BF16_RANDOM_INIT) * n_px)), float(crop_agree)
    loss, parts = tower_unet_loss(out, y.to(dev), bdist.to(dev))
    assert np.allclose(parts.detach().cpu().numpy(), z["losses"], rtol=tol, atol=1e-6), (parts, z["losses"])
    loss.backward()
    if dtype == torch.float32:
        grads = dict(model.named_parameters())
        names = [str(n) for n in z["grad_names"]]
        norms = np.array([float(grads[n].grad.double().norm()) for n in names])
        big = z["grad_norms"] > 1e-6
        assert np.allclose(norms[big], z["grad_norms"][big], rtol=TOL_GRAD_FP32), np.abs(norms[big] / z["grad_norms"][big] - 1).max()
        for n in FULL_GRADS:
            if "grad::" + n in z:
                assert rel_err(grads[n].grad, torch.from_numpy(z["grad::" + n])) < TOL_GRAD_FP32, n
        bufs = dict(model.named_buffers())
        rmn = np.array([float(bufs[str(n)].double().norm()) for n in z["bn_names"]])
        rvn = np.array([float(bufs[str(n).replace("running_mean", "running_var")].double().norm()) for n in z["bn_names"]])
        assert np.allclose(rvn, z["bn_var_norms"], rtol=1e-3)
        assert np.allclose(rmn, z["bn_mean_norms"], rtol=1e-3, atol=1e-5)


def model_vs_port(dev, cfg, dtype=torch.float32, training=True, seed=3, check_grads=True):
    """The product model against the oracle port on the same seeded inputs and weights (any size)."""
    from cultionet_b200.losses import tower_unet_loss
    from oracle import towerunet_port as port
    from tests.util import (MASK_AGREEMENT, MASK_AGREEMENT_BF16_RANDOM_INIT, MASK_FLIP_BAND_BF16, TOL_GRAD_BF16_GLOBAL, TOL_GRAD_BF16_WORST,
                            TOL_GRAD_FP32, TOL_OUT_BF16, TOL_OUT_FP32, mine_from_state_dict)

    spec = port.param_spec(cfg["C"], cfg["T"], cfg["hidden"], cfg["dilations"])
    sd = port.synth_state_dict(spec, seed=seed)
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(cfg["B"], cfg["C"], cfg["T"], cfg["H"], cfg["W"], generator=g)
    y = torch.randint(cfg.get("y_low", 0), 3, (cfg["B"], cfg["H"], cfg["W"]), generator=g)
    bdist = torch.rand(cfg["B"], cfg["H"], cfg["W"], generator=g)
    model = mine_from_state_dict(cfg, sd, dev, dtype).train(training)
    odev = dev  # the port runs in fp32 on the same device (TF32 disabled by conftest)
    sdo = {k: v.to(odev) for k, v in sd.items()}
    sdo = {k: (v.requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sdo.items()}
    with torch.set_grad_enabled(bool(training and check_grads)):  # a forward-only comparison keeps no autograd graph (full-size cases)
        want = port.towerunet_forward(sdo, x.to(odev), cfg["dilations"], training=training, natten_params=cfg.get("natten"),
                                      activation_type=cfg.get("activation_type", "SiLU"))
        out = model(x.to(dev))
    tol = TOL_OUT_FP32 if dtype == torch.float32 else TOL_OUT_BF16
    errs = {k: rel_err(out[k], want[k]) for k in ("distance", "edge", "crop")}
    assert all(e < tol for e in errs.values()), errs
    flipped = (out["crop"] > 0.5) != (want["crop"] > 0.5)
    agree = 1.0 - float(flipped.float().mean())
    if dtype == torch.float32:
        assert agree >= MASK_AGREEMENT, agree
    else:  # random-init weights in bf16: see tests/util.py
        assert agree >= MASK_AGREEMENT_BF16_RANDOM_INIT, agree
        decisive = (want["crop"].detach() - 0.5).abs() >= MASK_FLIP_BAND_BF16
        agree_decisive = 1.0 - float((flipped & decisive).float().sum() / decisive.float().sum().clamp_min(1.0))
        assert agree_decisive >= MASK_AGREEMENT, (agree, agree_decisive)
    report = {"out_err": errs, "crop_agreement": agree}
    if dtype != torch.float32:
        report["crop_agreement_decisive"] = agree_decisive
    if training and not check_grads:
        with torch.no_grad():
            loss, _ = tower_unet_loss(out, y.to(dev), bdist.to(dev))
            wloss, _ = port.training_loss(want, y.to(odev), bdist.to(odev))
        assert abs(float(loss) - float(wloss)) < tol * abs(float(wloss)), (float(loss), float(wloss))
        report["loss"] = (float(loss), float(wloss))
    if training and check_grads:
        loss, parts = tower_unet_loss(out, y.to(dev), bdist.to(dev))
        wloss, _ = port.training_loss(want, y.to(odev), bdist.to(odev))
        assert abs(float(loss) - float(wloss)) < tol * abs(float(wloss)), (float(loss), float(wloss))
        report["loss"] = (float(loss), float(wloss))
        loss.backward()
        wloss.backward()
        worst = 0.0
        worst_name = ""
        num = den = 0.0
        for n, p in model.named_parameters():
            gw = sdo[n].grad
            if gw is None or float(gw.norm()) < 1e-7:
                continue
            num += float((p.grad.double() - gw.double()).pow(2).sum())
            den += float(gw.double().pow(2).sum())
            e = rel_err(p.grad, gw)
            # in bf16 the per-parameter bar skips parameters of fewer than 16 elements (a Psi-Net stream bias is ONE number, the sum of a
            # sign-alternating gradient over every pixel: cancellation makes its relative error meaningless; it still counts in the
            # all-parameter error above)
            if e > worst and (dtype == torch.float32 or p.numel() >= 16):
                worst, worst_name = e, n
        report["worst_grad"] = (worst, worst_name)
        report["global_grad_err"] = (num / max(den, 1e-300)) ** 0.5
        if dtype == torch.float32:
            assert worst < TOL_GRAD_FP32, (worst, worst_name)
        else:
            assert report["global_grad_err"] < TOL_GRAD_BF16_GLOBAL, report
            assert worst < TOL_GRAD_BF16_WORST, report
    return report


# ---------------------------------------------------------------------------------------------------------------------
# optional block variants
# ---------------------------------------------------------------------------------------------------------------------
def maxpool_case(dev, dtype, B, H, W, C, seed=0):
    from oracle import towerunet_port as port

    torch.manual_seed(seed)
    x = _mk((B, H, W, C), dev, dtype)
    y = F.adaptive_max_pool2d(x, (H // 2, W // 2))
    xr = _f(x)
    yr = port.adaptive_max_pool_half(xr.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    tol = _tol(dtype)
    _check("maxpool fwd", y, yr, 1e-7)  # a selection: exact
    g = torch.randn_like(yr)
    _check("maxpool grad", torch.autograd.grad(y, x, g.to(dtype))[0], torch.autograd.grad(yr, xr, g.to(dtype).float())[0], tol)


def sca_case(dev, dtype, B, H, W, C, seed=0):
    """SpatialChannelAttention applied to a second tensor, against the oracle's restatement of the reference module."""
    from cultionet_b200.nn.modules.convolution import SpatialChannelAttention
    from oracle import towerunet_port as port

    torch.manual_seed(seed)
    m = SpatialChannelAttention(C, "SiLU").to(dev)
    with torch.no_grad():
        m.gamma.fill_(0.7)
    x = _mk((B, H, W, C), dev, dtype)
    y = _mk((B, H, W, C), dev, dtype)
    out = m.scale(x, y)
    xr, yr = _f(x), _f(y)
    ca, sa = m.channel_attention, m.spatial_attention
    prm = [ca.fc1[0].weight, ca.fc1[2].weight, ca.fc2[0].weight, ca.fc2[2].weight, sa.conv.weight, m.gamma]
    ref_prm = [p.detach().clone().requires_grad_(True) for p in prm]
    att = port.spatial_channel_attention(xr.permute(0, 3, 1, 2), *ref_prm)
    outr = (yr.permute(0, 3, 1, 2) * att).permute(0, 2, 3, 1)
    tol = _tol(dtype)
    _check("sca fwd", out, outr, tol)
    g = torch.randn_like(outr)
    got = torch.autograd.grad(out, [x, y, *prm], g.to(dtype))
    want = torch.autograd.grad(outr, [xr, yr, *ref_prm], g.to(dtype).float())
    names = ["x", "y", "fc1.0", "fc1.2", "fc2.0", "fc2.2", "spatial conv", "gamma"]
    for n, a, c in zip(names, got, want):
        _check(f"sca grad {n}", a, c, tol * 4)


def silu_case(dev, dtype, n=1000, seed=0):
    torch.manual_seed(seed)
    x = _mk((n,), dev, dtype, scale=3.0)
    y = F.silu(x)
    xr = _f(x)
    yr = TF.silu(xr)
    tol = _tol(dtype)
    _check("silu fwd", y, yr, tol)
    g = torch.randn_like(yr)
    _check("silu grad", torch.autograd.grad(y, x, g.to(dtype))[0], torch.autograd.grad(yr, xr, g.to(dtype).float())[0], tol)


def dropout_case(dev, dtype, B=3, H=16, W=20, C=24, p=0.3, seed=0):
    """The generator cannot match torch's: check the contract instead -- the kept fraction, the 1/(1-p) scale, whole-channel masks for
    Dropout2d, a backward that applies the forward's mask, and a new mask after the step counter advances."""
    torch.manual_seed(seed)
    x = _mk((B, H, W, C), dev, dtype, shift=3.0)  # no zeros in x
    site = F.new_rng_site()
    for channelwise in (False, True):
        fn = F.dropout2d if channelwise else F.dropout
        y = fn(x, p, site)
        keep = y.detach().float() != 0
        frac = float(keep.float().mean())
        n = B * C if channelwise else x.numel()
        assert abs(frac - (1 - p)) < 5 * (p * (1 - p) / n) ** 0.5 + 1e-3, (channelwise, frac)
        assert rel_err(y.detach().float()[keep], x.detach().float()[keep] / (1 - p)) < _tol(dtype)
        if channelwise:
            per_channel = keep.reshape(B, H * W, C)
            assert bool((per_channel.all(dim=1) | (~per_channel).all(dim=1)).all())
        g = torch.randn_like(y)
        dx = torch.autograd.grad(y, x, g)[0]
        assert bool(((dx.float() != 0) == (keep & (g.float() != 0))).all())
        assert rel_err(dx.float()[keep], g.float()[keep] / (1 - p)) < _tol(dtype)
        y2 = fn(x, p, site)
        assert torch.equal(y2, y)  # same step, same site: same mask
        F.rng_advance(x.device)
        y3 = fn(x, p, site)
        assert not torch.equal(y3 != 0, y != 0)


def na_dropout_case(dev, B, H, W, heads, hd, k, d, p=0.3, seed=0):
    """Attention dropout: p -> 0 reproduces plain neighbourhood attention; for p > 0 the analytic gradient agrees with a central
    difference of the (deterministic, fixed-mask) forward; the mean over many masks approaches the plain result."""
    torch.manual_seed(seed)
    Cn = heads * hd
    qkv = _mk((B, H, W, 3 * Cn), dev, torch.float32)
    site = F.new_rng_site()
    plain = F.na2d(qkv, heads, k, d, hd ** -0.5)
    tiny = F.na2d(qkv, heads, k, d, hd ** -0.5, attn_drop=1e-7, site=site)
    _check("na dropout p->0", tiny, plain.detach(), 1e-5)
    y = F.na2d(qkv, heads, k, d, hd ** -0.5, attn_drop=p, site=site)
    assert rel_err(y, plain) > 1e-2  # the mask did something
    g = torch.randn_like(y)
    (dq,) = torch.autograd.grad(y, qkv, g)
    dirn = torch.randn_like(qkv)
    eps = 1e-2
    with torch.no_grad():
        yp = F.na2d(qkv + eps * dirn, heads, k, d, hd ** -0.5, attn_drop=p, site=site)
        ym = F.na2d(qkv - eps * dirn, heads, k, d, hd ** -0.5, attn_drop=p, site=site)
    fd = float(((yp - ym) * g).sum() / (2 * eps))
    an = float((dq * dirn).sum())
    assert abs(fd - an) < 2e-2 * max(1.0, abs(an)), (fd, an)
    acc = torch.zeros_like(plain)
    n_masks = 64
    with torch.no_grad():
        for _ in range(n_masks):
            F.rng_advance(qkv.device)
            acc += F.na2d(qkv, heads, k, d, hd ** -0.5, attn_drop=p, site=site)
    assert rel_err(acc / n_masks, plain) < 0.2


# ---------------------------------------------------------------------------------------------------------------------------------
# prediction over a resident tile (cultionet_b200/tile.py; oracle/tile_port.py)
def _random_tile(T, C, H, W, seed):
    rng = np.random.default_rng(seed)
    tile = rng.integers(-200, 12000, size=(T, C, H, W), dtype=np.int64).astype(np.int16)  # below 0 and above 10000: both clips fire
    tile[:, :, : H // 3, : W // 4] = 0  # a no-data corner
    return tile


def window_load_case(dev, T, C, H, W, ws, pad, norm=True, seed=0):
    """cnb_window_load == reference windowing + load arithmetic, bit for bit."""
    from cultionet_b200.tile import WindowLoader, predict_windows
    from oracle import tile_port

    tile = _random_tile(T, C, H, W, seed)
    g = torch.Generator().manual_seed(seed)
    mean = torch.rand(C, generator=g) * 0.3 if norm else None
    std = torch.rand(C, generator=g) * 0.2 + 0.05 if norm else None
    want = tile_port.create_predict_windows(tile, ws, pad)
    win = predict_windows(H, W, ws, pad)
    assert len(win) == len(want)
    for row, wd in zip(win, want):
        assert tuple(row) == (wd["window_row_off"], wd["window_col_off"], wd["window_height"], wd["window_width"])
    loader = WindowLoader(torch.from_numpy(tile).to(dev), ws, pad, (mean.reshape(1, C, 1, 1, 1), std) if norm else None)
    batch = loader.load(torch.from_numpy(win))
    assert batch.x.shape == (len(win), C, T, ws + 2 * pad, ws + 2 * pad) and batch.x.dtype == torch.float32
    ref = torch.cat([tile_port.load_window(wd["x"], mean, std) for wd in want], dim=0)
    assert torch.equal(batch.x.cpu(), ref), float((batch.x.cpu() - ref).abs().max())
    assert batch.window_row_off.tolist() == [wd["window_row_off"] for wd in want]
    return batch


def window_load_all_values_case(dev, seed=0):
    """Every int16 value through cnb_window_load with three (mean, std) pairs: the correctly rounded two-step division of the kernel
    is bit-identical to torch's fp32 division for all 65536 raw values (incl. negative no-data and > 10000 saturated values)."""
    from cultionet_b200.tile import WindowLoader
    from oracle import tile_port

    C = 3
    plane = np.arange(-32768, 32768, dtype=np.int64).astype(np.int16).reshape(256, 256)
    tile = np.broadcast_to(plane, (1, C, 256, 256)).copy()
    g = torch.Generator().manual_seed(seed)
    for trial in range(4):
        mean = torch.rand(C, generator=g) * 0.6
        std = torch.rand(C, generator=g) * 0.5 + 1e-3
        if trial == 3:
            std[0] = float(np.float32(np.nextafter(np.float32(0.25), np.float32(0))))  # all-ones significand: the guarded divisor
        loader = WindowLoader(torch.from_numpy(tile).to(dev), 248, 4, (mean, std))
        x = loader.load(torch.tensor([[4, 4, 248, 248]], dtype=torch.int32)).x.cpu()  # origin (4,4), halo 4: covers the tile exactly
        want = tile_port.load_window(torch.from_numpy(tile.astype("int32")).permute(1, 0, 2, 3)[None], mean, std)
        assert torch.equal(x, want), (trial, int((x != want).sum()))


def predict_pack_case(dev, H, W, ws, pad, crop_channels=1, seed=0):
    """cnb_predict_pack == LightningGTiffWriter's slice / scale / clip / uint16 / windowed write, bit for bit."""
    from cultionet_b200.data import Data
    from cultionet_b200.tile import MosaicWriter, predict_windows

    from oracle import tile_port

    win = predict_windows(H, W, ws, pad)
    B, s = len(win), ws + 2 * pad
    g = torch.Generator().manual_seed(seed)
    pred = {k: torch.rand(B, 1, s, s, generator=g) * 1.2 - 0.1 for k in ("distance", "edge")}  # outside [0, 1]: the clip fires
    pred["crop"] = torch.rand(B, crop_channels, s, s, generator=g) * 1.2 - 0.1
    pred["distance"][0, 0, pad, pad] = 0.99999  # truncation, not rounding
    windows = [dict(window_row_off=r, window_col_off=c, window_height=h, window_width=w, padding=pad) for r, c, h, w in win.tolist()]
    want = np.zeros((3, H, W), dtype=np.uint16)
    tile_port.write_windows(want, pred, windows)
    writer = MosaicWriter(H, W, dev, ws)
    dev_pred = {k: v.to(dev) for k, v in pred.items()}
    half = B // 2  # two batches through the reference-shaped entry point
    for lo, hi in ((0, half), (half, B)):
        if hi > lo:
            batch = Data(x=torch.empty(hi - lo, 1, 1, s, s), padding=[pad] * (hi - lo),
                         window_row_off=torch.from_numpy(win[lo:hi, 0]), window_col_off=torch.from_numpy(win[lo:hi, 1]),
                         window_height=torch.from_numpy(win[lo:hi, 2]), window_width=torch.from_numpy(win[lo:hi, 3]))
            writer.write_on_batch_end({k: v[lo:hi] for k, v in dev_pred.items()}, batch)
    got = writer.mosaic.cpu().numpy()
    assert got.shape == want.shape and np.array_equal(got, want), int((got.astype(np.int32) != want.astype(np.int32)).sum())
    assert int(want[0, 0, 0]) == 9999 or win[0][2] == 0


def tile_predictor_case(dev, H=37, W=50, ws=12, pad=4, batch_windows=5, rank=0, world_size=1, cuda_graph=False, seed=0,
                        dtype=torch.float32, hidden=8, streaming=False):
    """TilePredictor (load -> predict_step -> pack per window batch, windows rank::world) == the reference pipeline run window by
    window on the oracle's windows with the same model: identical uint16 mosaic on this rank's windows, zeros elsewhere."""
    from cultionet_b200.data import Data
    from cultionet_b200.models.lightning import CultionetLitModel
    from cultionet_b200.tile import TilePredictor

    from oracle import tile_port

    T, C = 6, 2
    tile = _random_tile(T, C, H, W, seed)
    torch.manual_seed(seed)
    model = CultionetLitModel(in_channels=C, in_time=T, hidden_channels=hidden, dropout=0.0, compute_dtype=dtype).to(dev)
    model.eval()
    mean, std = torch.tensor([0.1, 0.2]), torch.tensor([0.07, 0.11])
    resident = torch.zeros(tile.shape, dtype=torch.int16, device=dev) if streaming else torch.from_numpy(tile).to(dev)
    tp = TilePredictor(model, resident, (mean, std), ws, pad, batch_windows, cuda_graph=cuda_graph, rank=rank, world_size=world_size)
    if streaming:  # the tile starts in host memory; rows are copied in ahead of the batches, finished mosaic rows copied back
        host_tile = torch.from_numpy(tile)
        host_tile = host_tile.pin_memory() if dev != "cpu" else host_tile
        host_mosaic = tp.run_streaming(host_tile)
        if dev != "cpu":
            torch.cuda.synchronize()
        assert torch.equal(tp.loader.tile.cpu(), torch.from_numpy(tile))
        got = host_mosaic[:, :, :W].numpy().copy()
        assert np.array_equal(got, tp.writer.mosaic.cpu().numpy())
    else:
        got = tp.run().cpu().numpy().copy()  # (a CPU mosaic would alias the writer's storage)
    windows = tile_port.create_predict_windows(tile, ws, pad)[rank::world_size]
    assert tp.num_windows == len(windows)
    want = np.zeros((3, H, W), dtype=np.uint16)
    with torch.no_grad():
        for lo in range(0, len(windows), batch_windows):
            chunk = windows[lo:lo + batch_windows]
            fill = batch_windows - len(chunk)  # the predictor fills a ragged last batch with (0, 0, 0, 0) windows
            x = torch.cat([tile_port.load_window(wd["x"], mean, std) for wd in chunk], dim=0).to(dev)
            if fill:
                origin = tile_port.create_predict_windows(tile, ws, pad)[0]["x"]
                x = torch.cat([x, tile_port.load_window(origin, mean, std).to(dev).expand(fill, -1, -1, -1, -1)], dim=0)
            out = model.predict_step(Data(x=x.contiguous()), 0)
            tile_port.write_windows(want, {k: v[: len(chunk)].float().cpu() for k, v in out.items() if v is not None}, chunk)
    diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
    # eval-mode results are per-sample; the same kernels ran on the same values, so the mosaics agree exactly in fp32
    tol = 0 if dtype == torch.float32 else 200
    assert diff.max() <= tol, (int(diff.max()), int((diff > 0).sum()))
    assert got.any()
    return got


def conv_bn_act_eval_case(dev, B, H, W, cins, cout, k, stride=1, act=True, seed=0):
    """Inference epilogue fusion (cnb_conv_desc::ep_scale/ep_shift/ep_act): ONE tcgen05 launch == Conv2d -> BatchNorm2d(eval) -> [SiLU] of
    torch in fp32 (bf16 tolerance), and == the two-launch path of this library."""
    torch.manual_seed(seed)
    dtype = torch.bfloat16
    xs = [_mk((B, H, W, c), dev, dtype, grad=False) for c in cins]
    w = torch.nn.Parameter(torch.randn(cout, sum(cins), k, k, device=dev) / (sum(cins) * k * k) ** 0.5)
    gamma, beta = torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev) * 0.2
    mean, var = torch.randn(cout, device=dev) * 0.3, torch.rand(cout, device=dev) + 0.3
    pad = k // 2
    with torch.no_grad():
        fused = F.conv2d_bn_act_eval(xs, w, gamma, beta, mean, var, 1e-5, act, ksize=k, stride=stride, pad=pad, dil=1)
        assert fused is not None, "the shape was expected to take the tcgen05 kernel"
        y = F.conv2d(xs, w, None, ksize=k, stride=stride, pad=pad, dil=1)
        two = F.batchnorm_act(y, gamma, beta, mean, var, False, eps=1e-5, act=act)
        xr = torch.cat([_f(x) for x in xs], dim=-1).permute(0, 3, 1, 2)
        ref = TF.batch_norm(TF.conv2d(xr, w, None, stride=stride, padding=pad), mean, var, gamma, beta, False, 0.0, 1e-5)
        code = int(act)  # True = SiLU, or a CNB_ACT_* code
        names = {v: k for k, v in F.ACT_CODES.items()}
        ref = (getattr(torch.nn, names[code])()(ref) if code else ref).permute(0, 2, 3, 1)
    tol = _tol(dtype)
    _check("fused eval epilogue vs torch", fused, ref, tol)
    _check("fused eval epilogue vs two launches", fused, two, tol)


def tile_kernels_vs_reference_golden(dev):
    """cnb_window_load and cnb_predict_pack against tests/golden/tile_reference.npz -- the windows, the normalised input and the uint16
    mosaic that the REAL reference code produced (BatchStore.write_batch, NormValues.transform, LightningGTiffWriter.write_on_batch_end
    run by oracle/make_tile_golden.py) -- bit for bit."""
    from cultionet_b200.data import Data
    from cultionet_b200.tile import MosaicWriter, WindowLoader, predict_windows
    from oracle.make_tile_golden import CASE, prediction_for, tile_golden_inputs
    from tests.util import GOLDEN_DIR

    z = np.load(GOLDEN_DIR / "tile_reference.npz")
    assert {k: int(v) for k, v in zip(z["cfg_keys"], z["cfg_vals"])} == CASE
    tile, mean, std = tile_golden_inputs()
    ws, pad = CASE["window_size"], CASE["padding"]
    H, W = tile.shape[-2:]
    win = predict_windows(H, W, ws, pad)
    assert np.array_equal(win.astype(np.int64), z["fields"][:, :4]) and (z["fields"][:, 4] == pad).all()
    loader = WindowLoader(torch.from_numpy(tile).to(dev), ws, pad, (mean, std))
    x = loader.load(torch.from_numpy(win)).x.cpu()
    assert torch.equal(x, torch.from_numpy(z["x_norm"])), int((x != torch.from_numpy(z["x_norm"])).sum())
    windows = [dict(window_row_off=r, window_col_off=c, window_height=h, window_width=w, padding=pad) for r, c, h, w in win.tolist()]
    pred = prediction_for(windows, ws + 2 * pad, CASE["seed"] + 1)
    writer = MosaicWriter(H, W, dev, ws)
    batch = Data(x=torch.empty(len(win), 1, 1, 1, 1), padding=[pad] * len(win), window_row_off=torch.from_numpy(win[:, 0]),
                 window_col_off=torch.from_numpy(win[:, 1]), window_height=torch.from_numpy(win[:, 2]),
                 window_width=torch.from_numpy(win[:, 3]))
    writer.write_on_batch_end({k: v.to(dev) for k, v in pred.items()}, batch)
    got = writer.mosaic.cpu().numpy()
    assert np.array_equal(got, z["mosaic"]), int((got != z["mosaic"]).sum())


# ---------------------------------------------------------------------------------------------------------------------
# a learnable synthetic task: chips whose labels follow from the pixels (blocky fields brighter than their surroundings, edges on
# the field boundaries), so that a few optimisation steps move the outputs away from the 0.5 threshold
# ---------------------------------------------------------------------------------------------------------------------
def learnable_batch(B, C, T, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(B, 1, max(H // 8, 1), max(W // 8, 1), generator=g)
    field = TF.interpolate(low, size=(H, W), mode="nearest")
    crop = field > 0.5
    pad = TF.pad(crop.float(), (1, 1, 1, 1), mode="replicate")
    nb_min = torch.minimum(torch.minimum(pad[:, :, :-2, 1:-1], pad[:, :, 2:, 1:-1]), torch.minimum(pad[:, :, 1:-1, :-2], pad[:, :, 1:-1, 2:]))
    edge = crop & (nb_min < 0.5)
    y = torch.where(edge, 2, torch.where(crop, 1, 0))[:, 0].long()
    bdist = TF.avg_pool2d(crop.float(), 5, 1, 2)[:, 0] * crop[:, 0].float()
    x = (0.2 + 0.6 * field.unsqueeze(2) * torch.linspace(0.5, 1.0, T).view(1, 1, T, 1, 1)).expand(B, C, T, H, W)
    x = x + 0.05 * torch.rand(B, C, T, H, W, generator=g)
    return x.contiguous(), y, bdist


def trained_mask_agreement_case(dev, dtype=torch.bfloat16, steps=40, cfg=None, cuda_graph=False, min_agreement=None):
    """Train the product model in ``dtype`` for ``steps`` optimisation steps on the learnable task, hand its weights to the fp32
    oracle port, and compare both on a held-out batch (training-mode BatchNorm on that batch in both): outputs within the dtype's
    tolerance and crop masks agreeing on >= 99.9 % of the pixels (the north_star line)."""
    import cultionet_b200 as cb
    from cultionet_b200.engine import TrainStep
    from cultionet_b200.models.lightning import CultionetLitModel
    from oracle import towerunet_port as port
    from tests.util import MASK_AGREEMENT, TOL_OUT_BF16_TRAINED, TOL_OUT_FP32

    cfg = cfg or dict(B=4, C=3, T=8, H=48, W=48, hidden=16)
    torch.manual_seed(7)
    lit = CultionetLitModel(in_channels=cfg["C"], in_time=cfg["T"], hidden_channels=cfg["hidden"], dilations=[1, 2], dropout=0.0,
                            learning_rate=3e-3, compute_dtype=dtype).to(dev)
    step = TrainStep(lit, total_steps=None, cuda_graph=cuda_graph)
    first = last = None
    for i in range(steps):
        x, y, bd = learnable_batch(cfg["B"], cfg["C"], cfg["T"], cfg["H"], cfg["W"], 100 + i)
        loss = float(step(cb.Data(x=x.to(dev), y=y.to(dev), bdist=bd.to(dev))))
        first = loss if first is None else first
        last = loss
    assert last < first, (first, last)
    step.close()
    x, y, bd = learnable_batch(cfg["B"], cfg["C"], cfg["T"], cfg["H"], cfg["W"], 999)
    net = lit.cultionet_model.mask_model
    sd = {k: v.detach().clone().float() for k, v in net.state_dict().items()}
    with torch.no_grad():
        net.train()
        out = net(x.to(dev))
        want = port.towerunet_forward(sd, x.to(dev), [1, 2], training=True)
    tol = TOL_OUT_FP32 if dtype == torch.float32 else TOL_OUT_BF16_TRAINED
    errs = {k: rel_err(out[k], want[k]) for k in ("distance", "edge", "crop")}
    agree = float(((out["crop"] > 0.5) == (want["crop"] > 0.5)).float().mean())
    acc = float(((want["crop"][:, 0] > 0.5).cpu() == (y == 1)).float().mean())
    report = {"loss": (first, last), "out_err": errs, "crop_agreement": agree, "crop_accuracy_vs_labels": acc}
    assert all(e < tol for e in errs.values()), report
    assert agree >= (MASK_AGREEMENT if min_agreement is None else min_agreement), report
    return report
