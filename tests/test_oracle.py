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
def test_port_matches_reference_module():
    from oracle.ref_loader import load_reference

    ref = load_reference()
    torch.manual_seed(5)
    B, C, T, H, W, hid = 1, 4, 9, 26, 22, 8
    model = ref.TowerUNet(in_channels=C, in_time=T, hidden_channels=hid, dilations=[1, 2])
    spec = port.param_spec(C, T, hid, [1, 2])
    sd_ref = model.state_dict()
    assert sorted(n for n, _ in spec) == sorted(sd_ref.keys())
    assert all(tuple(sd_ref[n].shape) == tuple(s) for n, s in spec)
    x = torch.rand(B, C, T, H, W)
    for training in (True, False):
        model.train(training)
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        got = port.towerunet_forward(sd, x, [1, 2], training=training)
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
