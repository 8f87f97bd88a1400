"""Pins the oracle (oracle/towerunet_port.py, oracle/natten_ref.py) against the reference's own known answers, the committed
golden vectors produced by the real reference, and -- where /root/reference exists -- the reference itself."""
import numpy as np
import pytest
import torch

from oracle import natten_ref
from oracle import towerunet_port as port
from oracle.make_golden import CASES, FULL_GRADS, golden_case, is_variant
from oracle.ref_loader import reference_available
from tests.util import load_golden, rel_err


def test_natten_window_rule_matches_branchy_form():
    for L in range(1, 70):
        for k in (1, 3, 5, 7, 9):
            for d in (1, 2, 3, 4):
                if k * d > L:
                    continue
                for i in range(L):
                    assert natten_ref.window_start(i, L, k, d) == natten_ref.natten_get_window_start(i, L, k, d)


def _reference_loss_inputs():
    # tests/test_loss.py:14-50 of the reference: one numpy Generator(100) drawn in this exact order
    from einops import rearrange

    rng = np.random.default_rng(100)
    B, H, W = 2, 20, 20
    rng.uniform(low=-3, high=3, size=(B, 2, H, W))
    crop_prob = rearrange(torch.from_numpy(rng.dirichlet((0.5, 0.5), size=(B * H * W))).float(), "(b h w) c -> b c h w", b=B, c=2, h=H, w=W)
    rng.random((B, 1, H, W))
    dist = torch.from_numpy(rng.random((B, 1, H, W))).float()
    discrete = torch.from_numpy(rng.integers(low=0, high=2, size=(B, H, W))).long()
    rng.integers(low=0, high=1, size=(B, H, W))
    dist_targets = torch.from_numpy(rng.random((B, H, W))).float()
    mask = torch.from_numpy(rng.integers(low=0, high=2, size=(B, 1, H, W))).long()
    return crop_prob, dist, discrete, dist_targets, mask


def test_port_loss_reproduces_reference_known_answers():
    crop_prob, dist, discrete, dist_targets, mask = _reference_loss_inputs()
    # reference tests/test_loss.py:118-123, :143-145
    assert round(float(port.tanimoto_complement_loss(crop_prob, discrete)), 3) == 0.824
    assert round(float(port.tanimoto_complement_loss(crop_prob, discrete, mask=mask)), 3) == 0.692
    assert round(float(port.tanimoto_complement_loss(dist, dist_targets, one_hot_targets=False)), 3) == 0.704


@pytest.mark.gpu
def test_kernel_loss_reproduces_reference_known_answers_gpu(dev):
    test_kernel_loss_reproduces_reference_known_answers(dev)


def test_kernel_loss_reproduces_reference_known_answers(dev):
    from cultionet_b200.losses import TanimotoComplementLoss

    crop_prob, dist, discrete, dist_targets, mask = (t.to(dev) for t in _reference_loss_inputs())
    loss = TanimotoComplementLoss()
    assert round(float(loss(crop_prob, discrete)), 3) == 0.824
    assert round(float(loss(crop_prob, discrete, mask=mask)), 3) == 0.692
    assert round(float(TanimotoComplementLoss(one_hot_targets=False)(dist, dist_targets)), 3) == 0.704


@pytest.mark.parametrize("name", [n for n, c in CASES.items() if not is_variant(c)])
def test_port_reproduces_golden(name):
    cfg, z = load_golden(name)
    assert {k: v for k, v in CASES[name].items()} == cfg
    spec, sd, x, y, bdist = golden_case(cfg)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    out = port.towerunet_forward(sd, x, cfg["dilations"], training=True)
    for k in ("distance", "edge", "crop"):
        assert rel_err(out[k][:, :, ::3, ::3], torch.from_numpy(z["out_" + k])) < 1e-5
    loss, parts = port.training_loss(out, y, bdist)
    got = np.array([float(loss), float(parts["dloss"]), float(parts["eloss"]), float(parts["closs"])])
    assert np.allclose(got, z["losses"], rtol=1e-5, atol=1e-6)
    loss.backward()
    names = [str(n) for n in z["grad_names"]]
    norms = np.array([float(sd[n].grad.double().norm()) for n in names])
    assert np.allclose(norms, z["grad_norms"], rtol=2e-3, atol=1e-7)
    for n in FULL_GRADS:
        if "grad::" + n in z:
            assert rel_err(sd[n].grad, torch.from_numpy(z["grad::" + n])) < 2e-4


@pytest.mark.skipif(not reference_available(), reason="/root/reference is only present in the authoring container")
@pytest.mark.parametrize("activation_type", ["SiLU", "ReLU", "GELU", "LeakyReLU"])
def test_port_matches_reference_module(activation_type):
    from oracle.ref_loader import load_reference

    ref = load_reference()
    torch.manual_seed(5)
    B, C, T, H, W, hid = 1, 4, 9, 26, 22, 8
    model = ref.TowerUNet(in_channels=C, in_time=T, hidden_channels=hid, dilations=[1, 2], activation_type=activation_type)
    spec = port.param_spec(C, T, hid, [1, 2])
    sd_ref = model.state_dict()
    assert sorted(n for n, _ in spec) == sorted(sd_ref.keys())
    assert all(tuple(sd_ref[n].shape) == tuple(s) for n, s in spec)
    x = torch.rand(B, C, T, H, W)
    for training in (True, False):
        model.train(training)
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        got = port.towerunet_forward(sd, x, [1, 2], training=training, activation_type=activation_type)
        want = model(x)
        for k in ("distance", "edge", "crop"):
            assert rel_err(got[k], want[k]) < 1e-5


@pytest.mark.skipif(not reference_available(), reason="/root/reference is only present in the authoring container")
def test_variant_restatements_match_reference_modules():
    """The oracle's SpatialChannelAttention / pool_by_max restatements against the real reference modules."""
    import importlib

    from oracle.ref_loader import load_reference

    load_reference()
    att = importlib.import_module("cultionet.nn.modules.attention")
    torch.manual_seed(5)
    m = att.SpatialChannelAttention(in_channels=12, activation_type="SiLU")
    with torch.no_grad():
        m.gamma.fill_(0.6)
    x = torch.randn(2, 12, 9, 7)
    ca, sa = m.channel_attention, m.spatial_attention
    got = port.spatial_channel_attention(x, ca.fc1[0].weight, ca.fc1[2].weight, ca.fc2[0].weight, ca.fc2[2].weight, sa.conv.weight, m.gamma)
    assert rel_err(got, m(x)) < 1e-6
    conv = importlib.import_module("cultionet.nn.modules.convolution")
    blk = conv.PoolResidualConv(4, 8, pool_by_max=True, dilations=[1]).eval()
    x = torch.randn(1, 4, 11, 9)
    pooled = port.adaptive_max_pool_half(x)
    assert rel_err(blk.res_conv(pooled), blk(x)) < 1e-6


# ---------------------------------------------------------------------------------------------------------------------
# the tile path (prediction windowing + GeoTIFF writing): oracle/tile_port.py pinned to the REAL reference code
# ---------------------------------------------------------------------------------------------------------------------
def _tile_port_outputs():
    from oracle import tile_port
    from oracle.make_tile_golden import CASE, prediction_for, tile_golden_inputs

    tile, mean, std = tile_golden_inputs()
    ws, pad = CASE["window_size"], CASE["padding"]
    windows = tile_port.create_predict_windows(tile, ws, pad)
    x_int = torch.cat([w["x"] for w in windows], dim=0)
    fields = np.array([[w["window_row_off"], w["window_col_off"], w["window_height"], w["window_width"], w["padding"]] for w in windows])
    x_norm = torch.cat([tile_port.load_window(w["x"], mean, std) for w in windows], dim=0)
    mosaic = np.zeros((3, tile.shape[-2], tile.shape[-1]), dtype=np.uint16)
    tile_port.write_windows(mosaic, prediction_for(windows, ws + 2 * pad, CASE["seed"] + 1), windows)
    return x_int, fields, x_norm, mosaic


def test_tile_port_reproduces_reference_golden():
    """tests/golden/tile_reference.npz was written by the real BatchStore / NormValues / LightningGTiffWriter code
    (oracle/make_tile_golden.py): the port reproduces every stored window, field, normalised value and mosaic pixel exactly."""
    from tests.util import GOLDEN_DIR

    z = np.load(GOLDEN_DIR / "tile_reference.npz")
    x_int, fields, x_norm, mosaic = _tile_port_outputs()
    assert np.array_equal(x_int.numpy(), z["x_int"].astype(np.int32))
    assert np.array_equal(fields, z["fields"])
    assert torch.equal(x_norm, torch.from_numpy(z["x_norm"]))
    assert np.array_equal(mosaic, z["mosaic"])
    assert int(z["normalize_from_reference"]) == 1


def test_tile_kernels_reproduce_reference_golden(dev):
    from tests import cases

    cases.tile_kernels_vs_reference_golden(dev)


@pytest.mark.skipif(not reference_available(), reason="/root/reference is only present in the authoring container")
def test_tile_port_matches_reference_code_live():
    """The same comparison against the reference modules imported here, on a second geometry (ragged ends smaller than the padding)."""
    from oracle import tile_port
    from oracle.make_tile_golden import prediction_for
    from oracle.ref_tile_loader import load_tile_reference, reference_store_window, reference_write_windows

    ns = load_tile_reference()
    rng = np.random.default_rng(5)
    T, C, H, W, ws, pad = 2, 3, 35, 41, 10, 6
    tile = rng.integers(-50, 10500, size=(T, C, H, W)).astype(np.int16)
    windows = tile_port.create_predict_windows(tile, ws, pad)
    padded = np.pad(tile, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    for wd in windows:
        r0, c0, h, w = wd["window_row_off"], wd["window_col_off"], wd["window_height"], wd["window_width"]
        b = reference_store_window(ns, padded[:, :, r0:r0 + h + 2 * pad, c0:c0 + w + 2 * pad], slice(r0, r0 + h), slice(c0, c0 + w), ws, pad)
        assert torch.equal(b.x, wd["x"])
        assert (b.window_row_off[0], b.window_col_off[0], b.window_height[0], b.window_width[0], b.padding[0]) == (r0, c0, h, w, pad)
    pred = prediction_for(windows, ws + 2 * pad, 9)
    f = np.array([[w["window_row_off"], w["window_col_off"], w["window_height"], w["window_width"], pad] for w in windows])
    batch = ns.TileData(x=torch.zeros(len(windows), 1), window_row_off=f[:, 0].tolist(), window_col_off=f[:, 1].tolist(),
                        window_height=f[:, 2].tolist(), window_width=f[:, 3].tolist(), padding=f[:, 4].tolist())
    want = reference_write_windows(ns, (3, H, W), pred, batch)
    got = np.zeros((3, H, W), dtype=np.uint16)
    tile_port.write_windows(got, pred, windows)
    assert np.array_equal(got, want)


# ---------------------------------------------------------------------------------------------------------------------
# the other Tanimoto losses of LOSS_DICT and the learning-rate schedulers
# ---------------------------------------------------------------------------------------------------------------------
def test_kernel_tanimoto_dist_and_combined_reproduce_reference_known_answers(dev):
    """tests/test_loss.py:110-116, :125-135, :139-141 of the reference: TanimotoDistLoss 0.611 / 0.431 / 0.417, CombinedLoss 0.717 / 0.561."""
    from cultionet_b200.losses import CombinedLoss, TanimotoComplementLoss, TanimotoDistLoss

    crop_prob, dist, discrete, dist_targets, mask = (t.to(dev) for t in _reference_loss_inputs())
    loss = TanimotoDistLoss()
    assert round(float(loss(crop_prob, discrete)), 3) == 0.611
    assert round(float(loss(crop_prob, discrete, mask=mask)), 3) == 0.431
    assert round(float(TanimotoDistLoss(one_hot_targets=False)(dist, dist_targets)), 3) == 0.417
    comb = CombinedLoss(losses=[TanimotoDistLoss(), TanimotoComplementLoss()])
    assert round(float(comb(crop_prob, discrete)), 3) == 0.717
    assert round(float(comb(crop_prob, discrete, mask=mask)), 3) == 0.561


def test_kernel_loss_variants_match_torch_restatement_with_gradients(dev):
    """Values and gradients of the three variants against plain-torch restatements of losses.py:62-100, :221-340."""
    from cultionet_b200.losses import CombinedLoss, TanimotoComplementLoss, TanimotoDistLoss

    def dist_loss(p, t, smooth=1e-5):  # tanimoto_dist on the pair and on its complement (losses.py:221-243, :318-340)
        def td(a, b):
            tpl = (a * b).sum((1, 2, 3))
            sq = (a ** 2 + b ** 2).sum((1, 2, 3))
            return 1.0 - (tpl + smooth) / ((sq - tpl) + smooth)
        return ((td(p, t) + td(1.0 - p, 1.0 - t)) * 0.5).mean()

    torch.manual_seed(3)
    p = torch.rand(3, 1, 12, 10, device=dev, requires_grad=True)
    t = (torch.rand(3, 12, 10, device=dev) > 0.5).long()
    tf = t.unsqueeze(1).float()
    for mod, want_fn in (
        (TanimotoDistLoss(), lambda q: dist_loss(q, tf)),
        (CombinedLoss([TanimotoDistLoss(), TanimotoComplementLoss()]),
         lambda q: 0.5 * (dist_loss(q, tf) + port.tanimoto_complement_loss(q, t))),
    ):
        got, want = mod(p, t), want_fn(p)
        assert abs(float(got) - float(want)) < 1e-6
        (gg,) = torch.autograd.grad(got, p)
        (gw,) = torch.autograd.grad(want, p)
        assert rel_err(gg, gw) < 1e-4


def test_one_cycle_lr_matches_torch_scheduler():
    from cultionet_b200.optim import make_lr_schedule, one_cycle_lr

    for total, lr in ((10, 0.01), (37, 0.003), (1000, 0.01)):
        opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr)
        sch = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=lr, total_steps=total)
        f = make_lr_schedule("OneCycleLR", lr, total_steps=total)
        for step in range(total):
            want = opt.param_groups[0]["lr"]
            assert abs(one_cycle_lr(step, total, lr) - want) <= 1e-12 + 1e-9 * want, (total, step)
            assert f(step) == one_cycle_lr(step, total, lr)
            opt.step()
            if step < total - 1:
                sch.step()


def test_epoch_schedulers_match_torch_schedulers():
    """CosineAnnealingLR(T_max=20, eta_min=1e-5), ExponentialLR(0.5), StepLR(step, 0.5) as configure_optimizers builds them
    (models/lightning.py:650-672), per epoch."""
    from cultionet_b200.optim import make_lr_schedule

    lr = 0.01
    makers = {
        "CosineAnnealingLR": lambda o: torch.optim.lr_scheduler.CosineAnnealingLR(o, T_max=20, eta_min=1e-5, last_epoch=-1),
        "ExponentialLR": lambda o: torch.optim.lr_scheduler.ExponentialLR(o, gamma=0.5),
        "StepLR": lambda o: torch.optim.lr_scheduler.StepLR(o, step_size=3, gamma=0.5),
    }
    for name, mk in makers.items():
        opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr)
        sch = mk(opt)
        f = make_lr_schedule(name, lr, steps_per_epoch=7, steplr_step_size=3)
        for epoch in range(20):  # torch's recursive cosine form follows the closed form for the first half period
            want = opt.param_groups[0]["lr"]
            assert abs(f(epoch * 7 + 3) - want) <= 1e-9 + 1e-6 * want, (name, epoch, f(epoch * 7 + 3), want)
            assert f(0, epoch) == f(epoch * 7)
            opt.step()
            sch.step()
    with pytest.raises(NameError):
        make_lr_schedule("ReduceLROnPlateau", lr)
