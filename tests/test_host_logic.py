"""Host-side behaviour of the reference-facing surface (no kernels beyond the CPU interpreter)."""
import pytest
import torch

import cultionet_b200 as cb
from cultionet_b200.enums import InferenceNames


def test_state_dict_keys_match_reference_inventory():
    from oracle import towerunet_port as port

    m = cb.TowerUNet(in_channels=3, in_time=8, hidden_channels=8, dilations=[1, 2, 3])
    spec = dict(port.param_spec(3, 8, 8, [1, 2, 3]))
    sd = m.state_dict()
    assert sorted(sd) == sorted(spec)
    assert all(tuple(sd[k].shape) == tuple(spec[k]) for k in sd)


def test_loads_torch_compile_prefixed_checkpoints():
    # the reference wraps pre_unet in torch.compile (nunet.py:141): keys may be `pre_unet._orig_mod.*`
    m = cb.TowerUNet(in_channels=2, in_time=6, hidden_channels=8)
    sd = {k.replace("pre_unet.", "pre_unet._orig_mod."): v for k, v in m.state_dict().items()}
    m2 = cb.TowerUNet(in_channels=2, in_time=6, hidden_channels=8)
    m2.load_state_dict(sd, strict=True)
    assert torch.equal(m2.pre_unet.conv3.seq[0].weight, m.pre_unet.conv3.seq[0].weight)


def test_data_container_contract():
    d = cb.Data(x=torch.zeros(2, 3, 4, 5, 6), y=torch.zeros(2, 5, 6), bdist=torch.zeros(2, 5, 6), lon=torch.zeros(2), lat=torch.zeros(2))
    assert (d.num_samples, d.num_channels, d.num_time, d.height, d.width) == (2, 3, 4, 5, 6)
    c = d.copy()
    c.x += 1
    assert float(d.x.sum()) == 0.0
    with pytest.raises(AssertionError):
        cb.Data(x=torch.zeros(1), bad="string")


def test_option_errors_mirror_the_reference():
    with pytest.raises(AssertionError):
        cb.CultioNet(in_channels=2, in_time=6, model_type="UNet3")
    with pytest.raises(AssertionError):  # ResidualConv only knows spatial_channel (convolution.py:197-200)
        cb.TowerUNet(in_channels=2, in_time=6, hidden_channels=8, res_block_type="res", attention_weights="natten")
    with pytest.raises(TypeError):  # ... and cannot build it either (SpatialChannelAttention(out_channels=...), convolution.py:203-205)
        cb.TowerUNet(in_channels=2, in_time=6, hidden_channels=8, res_block_type="res", attention_weights="spatial_channel")
    with pytest.raises(AssertionError):
        cb.TowerUNet(in_channels=2, in_time=6, hidden_channels=8, res_block_type="resx")
    with pytest.raises(NotImplementedError):  # the Tanimoto family of LOSS_DICT is built, the rest is not
        cb.CultionetLitModel(in_channels=2, in_time=6, hidden_channels=8, loss_name="TverskyLoss")


def test_variant_state_dict_keys_match_the_reference_inventory():
    # inventories recorded from the real reference models by oracle/make_golden.py
    from tests.util import golden_spec, load_golden, mine_from_state_dict
    from oracle.make_golden import golden_case

    for name in ("sca_maxpool", "res_bnfirst", "bnfirst_maxpool_odd", "latlon"):
        cfg, z = load_golden(name)
        spec, sd, *_ = golden_case(cfg, golden_spec(z))
        m = mine_from_state_dict(cfg, sd, "cpu")
        mine = m.state_dict()
        assert sorted(mine) == sorted(dict(spec)), name
        assert all(tuple(mine[k].shape) == tuple(dict(spec)[k]) for k in mine), name


def test_reference_test_cultionet_configuration(dev):
    """tests/test_cultionet.py:58-120 of the reference: spatial_channel attention, pool_by_max, dropout 0.2, train mode, shapes only."""
    torch.manual_seed(0)
    model = cb.CultioNet(in_channels=5, in_time=13, hidden_channels=8, model_type="TowerUNet", activation_type="SiLU", dilations=[1, 2],
                         dropout=0.2, res_block_type="resa", attention_weights="spatial_channel", pool_by_max=True)
    batch = cb.Data(x=torch.rand(2, 5, 13, 20, 20), lon=torch.zeros(2), lat=torch.zeros(2))
    out = model(batch)
    for k in (InferenceNames.DISTANCE, InferenceNames.EDGE, InferenceNames.CROP):
        assert out[k].shape == (2, 1, 20, 20)
    (out[InferenceNames.DISTANCE].mean() + out[InferenceNames.CROP].mean()).backward()
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in model.parameters() if p.requires_grad)


def test_default_dropout_trains_and_eval_ignores_it(dev):
    """CultionetLitModel defaults to dropout=0.2 (lightning.py:828): Dropout2d in the encoder, attn_drop / proj_drop in natten."""
    torch.manual_seed(1)
    m = cb.TowerUNet(in_channels=2, in_time=6, hidden_channels=8, dropout=0.2)
    m0 = cb.TowerUNet(in_channels=2, in_time=6, hidden_channels=8, dropout=0.0)
    m0.load_state_dict(m.state_dict())
    x = torch.rand(1, 2, 6, 16, 16)
    assert torch.equal(m.eval()(x)["crop"], m0.eval()(x)["crop"])
    a = m.train()(x)["crop"]
    b = m.train()(x)["crop"]
    assert not torch.equal(a, b)  # a new mask every forward
    a.mean().backward()
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in m.parameters())


def test_cultionet_forward_adds_reference_keys(dev):
    model = cb.CultioNet(in_channels=2, in_time=6, hidden_channels=8, dropout=0.0)
    batch = cb.Data(x=torch.rand(1, 2, 6, 16, 16), lon=torch.zeros(1), lat=torch.zeros(1))
    out = model(batch)
    for k in (InferenceNames.DISTANCE, InferenceNames.EDGE, InferenceNames.CROP):
        assert out[k].shape == (1, 1, 16, 16)
        assert float(out[k].min()) >= 0.0 and float(out[k].max()) <= 1.0
    assert out[InferenceNames.CROP_TYPE] is None and out["classes_l2"] is None and out["classes_l3"] is None


def test_bad_input_shape_raises(dev):
    m = cb.TowerUNet(in_channels=2, in_time=6, hidden_channels=8)
    with pytest.raises(ValueError):
        m(torch.rand(1, 3, 6, 16, 16))


def test_direct_parameter_gradients_match_autograd_accumulation(dev):
    """engine.TrainStep lets the weight / bias gradient kernels write into the flat gradient buffer (functional.direct_param_grads);
    the result must equal what autograd's AccumulateGrad produces."""
    from cultionet_b200 import functional as F
    from cultionet_b200.models.lightning import CultionetLitModel

    torch.manual_seed(3)
    model = CultionetLitModel(in_channels=2, in_time=6, hidden_channels=8, dropout=0.0, compute_dtype=torch.float32).to(dev)
    opt = model.configure_optimizers(total_steps=10)
    batch = cb.Data(x=torch.rand(2, 2, 6, 16, 16, device=dev), y=torch.randint(-1, 3, (2, 16, 16), device=dev),
                    bdist=torch.rand(2, 16, 16, device=dev))
    grads = []
    for direct in (False, True, True):  # twice direct: the persistent split-K accumulators must come back cleared
        opt.zero_grad()
        loss = model.training_step(batch, 0)
        with F.direct_param_grads(direct):
            loss.backward()
        grads.append(opt.flat_grad.clone())
    assert float(grads[0].norm()) > 0
    assert float((grads[0] - grads[1]).norm() / grads[0].norm()) < 1e-5
    assert float((grads[0] - grads[2]).norm() / grads[0].norm()) < 1e-5


def test_fit_checkpoint_roundtrip_and_resume(dev, tmp_path):
    """model.fit writes a Lightning-layout checkpoint (best val_score), load_from_checkpoint rebuilds an identical module from its
    hyper_parameters + state_dict, and a second fit() call resumes from the file (model.py:308-314, callbacks.py:238-249)."""
    from cultionet_b200 import model as M
    from cultionet_b200.models.lightning import CultionetLitModel

    torch.manual_seed(5)
    lit = CultionetLitModel(in_channels=2, in_time=6, hidden_channels=8, dropout=0.0, compute_dtype=torch.float32).to(dev)
    g = torch.Generator().manual_seed(1)
    batches = [cb.Data(x=torch.rand(2, 2, 6, 16, 16, generator=g), y=torch.randint(-1, 3, (2, 16, 16), generator=g),
                       bdist=torch.rand(2, 16, 16, generator=g)) for _ in range(2)]
    ckpt = tmp_path / "ckpt" / "last.ckpt"
    hist = M.fit(lit, batches, val_batches=batches[:1], epochs=2, ckpt_file=ckpt, device=dev, cuda_graph=False)
    assert len(hist["loss"]) == 2 and ckpt.is_file() and hist["checkpoint"] == str(ckpt)
    raw = torch.load(ckpt, weights_only=False)
    assert {"state_dict", "hyper_parameters", "epoch", "global_step", "optimizer_states", "pytorch-lightning_version"} <= set(raw)
    assert all(k.startswith("cultionet_TowerUNet.") for k in raw["state_dict"])  # the reference's attribute prefix (lightning.py:874)
    assert raw["hyper_parameters"]["in_time"] == 6 and raw["hyper_parameters"]["hidden_channels"] == 8

    again = CultionetLitModel.load_from_checkpoint(ckpt, map_location=dev)
    again.freeze()
    want = raw["state_dict"]
    for k, v in again.state_dict().items():
        assert torch.equal(v.cpu(), want[k]), k
    assert not any(p.requires_grad for p in again.parameters()) and not again.training

    # resume: the saved epoch was the best of {0, 1}; a 3-epoch fit continues after it with the saved AdamW moments
    # optimizer_states[0] is laid out like torch.optim.AdamW.state_dict(): {"state": {i: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups"}
    saved_epoch, saved_step = raw["epoch"], int(raw["optimizer_states"][0]["state"][0]["step"])
    lit2 = CultionetLitModel(**{k: v for k, v in raw["hyper_parameters"].items()}).to(dev)
    hist2 = M.fit(lit2, batches, val_batches=batches[:1], epochs=3, ckpt_file=ckpt, device=dev, cuda_graph=False)
    assert len(hist2["loss"]) == 3 - (saved_epoch + 1)
    assert saved_step == 2 * (saved_epoch + 1)


def test_reference_layout_checkpoint_loads_and_reproduces_golden_outputs(dev, tmp_path):
    """A checkpoint laid out as the reference's Lightning run writes it -- ``state_dict`` keys ``cultionet_TowerUNet.mask_model.<TowerUNet key>``
    (lightning.py:874, cultionet.py:70-78) with the ``pre_unet._orig_mod.`` infix of torch.compile (nunet.py:141) and the reference's
    hyper-parameter names -- loads through ``CultionetLitModel.load_from_checkpoint`` and reproduces the golden outputs the real
    reference produced from the same weights."""
    from cultionet_b200.models.lightning import CultionetLitModel
    from oracle.make_golden import golden_case
    from tests.util import TOL_OUT_FP32, golden_spec, load_golden, rel_err

    cfg, z = load_golden("small_masked")
    _, sd, x, _, _ = golden_case(cfg, golden_spec(z))
    ref_sd = {}
    for k, v in sd.items():
        k = k.replace("pre_unet.", "pre_unet._orig_mod.", 1) if k.startswith("pre_unet.") and "_orig_mod" not in k else k
        ref_sd["cultionet_TowerUNet.mask_model." + k] = v
    hp = dict(in_channels=cfg["C"], in_time=cfg["T"], hidden_channels=cfg["hidden"], model_type="TowerUNet", dropout=0.0,
              activation_type="SiLU", dilations=cfg["dilations"], res_block_type="resa", attention_weights="natten", optimizer="AdamW",
              loss_name="TanimotoComplementLoss", learning_rate=0.01, lr_scheduler="OneCycleLR", steplr_step_size=5, weight_decay=1e-3,
              eps=1e-4, ckpt_name="last", model_name="cultionet", pool_by_max=False, batchnorm_first=False, class_counts=None,
              edge_class=2, scale_pos_weight=False, save_batch_val_metrics=False)
    path = tmp_path / "last.ckpt"
    torch.save({"epoch": 3, "global_step": 40, "pytorch-lightning_version": "2.1.0", "state_dict": ref_sd, "hyper_parameters": hp,
                "optimizer_states": [], "lr_schedulers": []}, path)
    lit = CultionetLitModel.load_from_checkpoint(path, map_location=dev, compute_dtype=torch.float32)
    assert lit.loaded_checkpoint["epoch"] == 3
    lit.train()  # the golden vectors are train-mode (batch statistics) outputs
    out = lit.cultionet_model.mask_model(x.to(dev))
    for k in ("distance", "edge", "crop"):
        assert rel_err(out[k][:, :, ::3, ::3], torch.from_numpy(z["out_" + k])) < TOL_OUT_FP32, k


def test_predict_windows_geometry_of_a_sentinel2_tile():
    """BASELINE configs[4]: a 10980 x 10980 tile in 100 px windows = 110 x 110 = 12 100 windows, the last row / column 80 px."""
    from cultionet_b200.tile import predict_windows

    win = predict_windows(10980, 10980, 100, 20)
    assert win.shape == (12100, 4) and win.dtype.name == "int32"
    assert tuple(win[0]) == (0, 0, 100, 100) and tuple(win[1]) == (0, 100, 100, 100)  # chunk order: x fastest
    assert tuple(win[109]) == (0, 10900, 100, 80) and tuple(win[-1]) == (10900, 10900, 80, 80)
    assert int((win[:, 2].astype("int64") * win[:, 3]).sum()) == 10980 * 10980  # the windows tile the image exactly
    with pytest.raises(ValueError):
        predict_windows(100, 100, 0, 20)


def test_tile_and_checkpoint_argument_errors(dev, tmp_path):
    from cultionet_b200 import model as M
    from cultionet_b200.tile import WindowLoader

    with pytest.raises(TypeError):
        WindowLoader(torch.zeros(2, 2, 8, 8, dtype=torch.int32, device=dev), 4, 2)  # the tile is int16 (data/create.py:70-79)
    with pytest.raises(ValueError):
        WindowLoader(torch.zeros(2, 2, 8, 8, dtype=torch.int16, device=dev), 5, 2)  # window + halo must be whole 16-byte rows
    with pytest.raises(ValueError):
        WindowLoader(torch.zeros(2, 2, 8, 8, dtype=torch.int16, device=dev), 4, 2, (torch.zeros(3), torch.ones(3)))  # one mean / std per band
    bad = tmp_path / "not_a_checkpoint.ckpt"
    torch.save({"weights": {}}, bad)
    with pytest.raises(KeyError):
        M.load_from_checkpoint(bad)
    torch.save({"state_dict": {}, "hyper_parameters": {}}, bad)
    with pytest.raises(KeyError):
        M.load_from_checkpoint(bad)  # no in_channels / in_time: must be passed as keyword arguments


def test_loss_reader_returns_every_value_in_order():
    """engine.LossReader: pipelined read-back of per-step scalars (each value exactly once, in push order, for any depth)."""
    import torch

    from cultionet_b200.engine import LossReader

    for depth in (1, 2, 3):
        r = LossReader(depth)
        got = []
        for i in range(7):
            v = r.push(torch.tensor([float(i)]) if i % 2 else torch.tensor(float(i)))
            if v is not None:
                got.append(v)
            assert len(r.pending) <= depth
        got += r.drain()
        assert got == [float(i) for i in range(7)], (depth, got)
        assert r.drain() == []
