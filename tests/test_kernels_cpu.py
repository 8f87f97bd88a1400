"""`-m "not gpu"`: kernel arithmetic and module wiring checked on the CPU kernel interpreter (tests/emu) at small sizes."""
import pytest
import torch

from tests import cases

F32 = torch.float32
BF16 = torch.bfloat16


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("geom", [(3, 1, 1, 1), (3, 2, 1, 1), (3, 1, 2, 2), (1, 1, 0, 1)])
def test_conv(dev, dtype, geom):
    k, s, p, d = geom
    cases.conv_case(dev, dtype, 2, 9, 11, [5, 7], 6, k, s, p, d)


def test_conv_wide_tiles(dev):
    cases.conv_case(dev, F32, 1, 10, 13, [70, 3], 67, 3, 1, 1, 1)


@pytest.mark.parametrize("stride", [2, 4])
def test_conv_transpose(dev, stride):
    cases.convT_case(dev, F32, 2, 7, 6, 5, 4, stride)


@pytest.mark.parametrize("direct", [False, True])
def test_attention_module_against_torch(dev, direct):
    cases.na_module_case(dev, F32, 1, 9, 10, 16, 2, 3, 1, direct)      # generic kernels
    if direct:
        cases.na_module_case(dev, BF16, 1, 10, 18, 64, 2, 3, 2, direct)   # specialised bf16 kernels (head_dim 32)
    else:
        cases.na_module_case(dev, BF16, 1, 9, 17, 128, 2, 3, 1, direct)   # head_dim 64: two vectors per lane


@pytest.mark.parametrize("direct", [False, True])
def test_upconv_bias_gradient_from_resize_backward(dev, direct):
    cases.upconv_case(dev, F32, 2, 5, 6, 8, direct)
    cases.upconv_case(dev, BF16, 1, 7, 5, 16, direct)


def test_linear(dev):
    cases.linear_case(dev, F32, 2, 5, 7, 10, 21)


@pytest.mark.parametrize("training,act,res", [(True, True, True), (True, False, False), (False, True, False)])
def test_batchnorm(dev, training, act, res):
    cases.batchnorm_case(dev, F32, 2, 9, 11, 6, training, act, res)


def test_batchnorm_bf16(dev):
    cases.batchnorm_case(dev, BF16, 2, 9, 11, 6)


def test_batchnorm3d_channel_map(dev):
    cases.batchnorm3d_case(dev, F32, 2, 5, 6, 3, 4)


@pytest.mark.parametrize("C", [8, 40, 100])
def test_layernorm(dev, C):
    cases.layernorm_case(dev, F32, 2, 5, 7, C)


@pytest.mark.parametrize("cfg", [(2, 8, 3, 1, 7, 9), (2, 16, 3, 2, 9, 8), (1, 40, 7, 2, 15, 14), (4, 4, 5, 1, 6, 6), (8, 2, 3, 1, 3, 3)])
def test_neighborhood_attention(dev, cfg):
    heads, hd, k, d, H, W = cfg
    cases.na_case(dev, F32, 2, H, W, heads, hd, k, d)


@pytest.mark.parametrize("cfg", [(2, 8, 3, 2, 30, 41), (1, 16, 5, 1, 21, 37), (2, 4, 3, 1, 17, 50), (1, 8, 7, 2, 33, 35)])
def test_neighborhood_attention_multi_tile(dev, cfg):
    # images larger than one 8x16 tile + halo: interior tiles, slid border regions, border queries reaching across tiles
    heads, hd, k, d, H, W = cfg
    cases.na_case(dev, F32, 1, H, W, heads, hd, k, d)


def test_neighborhood_attention_multi_tile_bf16(dev):
    cases.na_case(dev, BF16, 1, 20, 40, 2, 16, 3, 2)


@pytest.mark.parametrize("cfg", [(2, 32, 3, 1, 19, 35), (1, 64, 3, 2, 21, 37), (2, 32, 7, 1, 17, 20), (1, 64, 7, 2, 30, 33), (1, 32, 3, 2, 6, 7), (1, 32, 3, 2, 33, 21), (1, 32, 7, 2, 33, 35)])
def test_neighborhood_attention_specialised_bf16(dev, cfg):
    # the template-specialised kernels (k_na_fast.cuh): bf16, k in {3,7}, dilation in {1,2}, head_dim in {32,64}; odd sizes put
    # clamped windows, slid regions, partial tiles and cross-tile border queries on every path
    heads, hd, k, d, H, W = cfg
    cases.na_case(dev, BF16, 1, H, W, heads, hd, k, d)


def test_neighborhood_attention_rejects_small_input(dev):
    from cultionet_b200 import _lib
    from cultionet_b200 import functional as F

    with pytest.raises(_lib.CnbError):
        F.na2d(torch.zeros(1, 4, 4, 3 * 8), 2, 3, 2, 0.5)  # k*d = 6 > 4, natten raises too


@pytest.mark.parametrize("sizes", [(7, 7, 8, 8), (13, 13, 25, 25), (5, 9, 12, 10), (10, 10, 7, 7), (1, 1, 4, 4)])
def test_resize_bilinear(dev, sizes):
    cases.resize_case(dev, F32, 2, *sizes, 3)


@pytest.mark.parametrize("sizes", [(7, 7, 8, 8), (13, 13, 25, 25), (5, 9, 12, 10), (10, 10, 7, 7), (1, 1, 4, 4), (31, 17, 32, 33), (8, 8, 11, 11), (15, 15, 21, 22)])
def test_resize_bilinear_vector_kernels(dev, sizes):
    # channel counts that are whole 16-byte vectors: the row-per-CTA forward and the table-form / gather-form backward kernels
    cases.resize_case(dev, F32, 2, *sizes, 8)
    cases.resize_case(dev, BF16, 1, *sizes, 16)


def test_conv_skinny_gemm_plus_shift_add(dev):
    cases.conv_skinny_case(dev, F32, 2, 9, 11, 16, 9)
    cases.conv_skinny_case(dev, F32, 1, 7, 6, 8, 3, k=3, pad=2, dil=2)
    cases.conv_skinny_case(dev, BF16, 1, 10, 12, 32, 9)
    cases.conv_skinny_case(dev, BF16, 1, 5, 70, 16, 9)   # two 64-pixel strips per row (tiled shift-add / gather kernels), ragged second strip


@pytest.mark.parametrize("k", [3, 5])
def test_pretime_conv(dev, k):
    cases.pretime_case(dev, F32, 2, 3, 12, 5, 6, k)


@pytest.mark.parametrize("k", [3, 5])
def test_pretime_conv_as_banded_gemm(dev, k):
    cases.pretime_gemm_case(dev, F32, 2, 3, 12, 5, 6, k)
    cases.pretime_gemm_case(dev, torch.bfloat16, 1, 5, 10, 9, 8, k)  # K = 50 -> pitch 56; 72 pixels = one full + one ragged tile
    cases.pretime_gemm_case(dev, torch.bfloat16, 2, 3, 8, 8, 16, k)  # 128 pixels per image = whole 64-pixel tiles: the 16-byte load path


@pytest.mark.parametrize("flags", [(True, True), (False, False)])
def test_final_combine(dev, flags):
    cases.final_combine_case(dev, F32, 2, 6, 7, *flags)


@pytest.mark.parametrize("y_low", [-1, 0])
def test_training_loss(dev, y_low):
    cases.loss_case(dev, 3, 10, 12, y_low)


def test_adamw_and_clipping(dev):
    cases.adamw_case(dev)


@pytest.mark.parametrize("name", ["small_masked", "odd_dil3", "sca_maxpool", "res_bnfirst", "bnfirst_maxpool_odd", "latlon"])
def test_model_reproduces_reference_golden(dev, name):
    cases.model_vs_golden(dev, name)


@pytest.mark.parametrize("name", ["ReLU", "LeakyReLU", "GELU", "Mish", "ELU", "Tanh", "Sigmoid", "Hardswish"])
def test_activation_codes(dev, name):
    cases.activation_case(dev, F32, name)
    cases.activation_case(dev, BF16, name)


@pytest.mark.parametrize("name", ["ReLU", "GELU"])
def test_model_with_another_activation_type(dev, name):
    """``activation_type`` other than the default SiLU (reference ``model.py:58``, ``activations.py``): training-mode forward, loss and
    every gradient against the oracle port (which ``test_port_matches_reference_module`` pins to the real reference for these names)."""
    cfg = dict(B=1, C=2, T=6, H=16, W=16, hidden=8, dilations=[1, 2], activation_type=name)
    cases.model_vs_port(dev, cfg, training=True)


def test_model_eval_mode_matches_port(dev):
    cfg = dict(B=1, C=2, T=6, H=16, W=16, hidden=8, dilations=[1, 2])
    cases.model_vs_port(dev, cfg, training=False)


# ---- 128-bit ("vec") variants of the bandwidth kernels: channel counts that are multiples of the vector width take them ----
@pytest.mark.parametrize("dtype,C", [(F32, 8), (F32, 12), (BF16, 8), (BF16, 24)])
@pytest.mark.parametrize("training,act,res", [(True, True, True), (False, False, False)])
def test_batchnorm_vector_path(dev, dtype, C, training, act, res):
    cases.batchnorm_case(dev, dtype, 2, 9, 11, C, training, act, res)


@pytest.mark.parametrize("dtype,C", [(F32, 132), (F32, 300), (BF16, 64), (BF16, 264)])
def test_layernorm_vector_path(dev, dtype, C):
    cases.layernorm_case(dev, dtype, 2, 3, 5, C)


@pytest.mark.parametrize("sizes", [(7, 7, 8, 8), (13, 13, 25, 25), (10, 10, 7, 7)])
def test_resize_bilinear_vector_path(dev, sizes):
    cases.resize_case(dev, F32, 2, *sizes, 8)
    cases.resize_case(dev, BF16, 2, *sizes, 16)


def test_conv_bias_grad_vector_path(dev):
    cases.conv_case(dev, F32, 2, 9, 11, [5, 7], 8, 3, 1, 1, 1)
    cases.conv_case(dev, BF16, 2, 9, 11, [8], 16, 1, 1, 0, 1)


def test_tiny_channel_convs(dev):
    # Psi-Net second-stage convolutions: 3 -> 1 (bias) and the 3 x [1] -> 3 fuse convolution
    cases.conv_case(dev, F32, 2, 9, 11, [3], 1, 3, 1, 1, 1)
    cases.conv_case(dev, BF16, 2, 9, 11, [1, 1, 1], 3, 3, 1, 1, 1)
    cases.conv_case(dev, F32, 1, 9, 11, [4], 2, 3, 2, 1, 1)
    cases.convT_case(dev, F32, 1, 5, 6, 3, 2, 2)
    # the head shapes with compile-time channel counts (forward, data gradient and the register-blocked weight gradient)
    for dtype in (F32, BF16):
        cases.conv_case(dev, dtype, 2, 9, 11, [9], 3, 3, 1, 1, 1)
        cases.conv_case(dev, dtype, 2, 9, 11, [3], 3, 3, 1, 1, 1)


def test_packed_weight_cache_tracks_parameter_updates(dev):
    """Packed bf16/fp32 copies of nn.Parameters are cached between calls: torch in-place updates, the flat AdamW kernel and a freed
    and re-created parameter must all miss the cache."""
    from cultionet_b200 import functional as F
    from cultionet_b200.optim import FlatAdamW

    torch.manual_seed(0)
    x = torch.randn(1, 6, 5, 8, device=dev)
    w = torch.nn.Parameter(torch.randn(4, 8, 3, 3, device=dev))
    y1 = F.conv2d([x], w, None, 3, 1, 1, 1)
    assert F.packed_weights(w, F.W_CONV, 4, 8, 9, torch.float32, False)[0] is F.packed_weights(w, F.W_CONV, 4, 8, 9, torch.float32, False)[0]
    with torch.no_grad():
        w.mul_(2.0)
    y2 = F.conv2d([x], w, None, 3, 1, 1, 1)
    assert torch.allclose(y2, 2 * y1, rtol=1e-5, atol=1e-6)
    opt = FlatAdamW([w], lr=0.1, clip_norm=0.0)
    opt.zero_grad()
    F.conv2d([x], w, None, 3, 1, 1, 1).sum().backward()
    opt.step()
    y3 = F.conv2d([x], w, None, 3, 1, 1, 1)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.detach(), padding=1).permute(0, 2, 3, 1)
    assert torch.allclose(y3, ref, rtol=1e-4, atol=1e-5) and not torch.allclose(y3, y2)
    del w
    w2 = torch.nn.Parameter(torch.randn(4, 8, 3, 3, device=dev))
    ref2 = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w2.detach(), padding=1).permute(0, 2, 3, 1)
    assert torch.allclose(F.conv2d([x], w2, None, 3, 1, 1, 1), ref2, rtol=1e-4, atol=1e-5)


# --- optional block variants -----------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("shape", [(2, 12, 10, 8), (1, 25, 25, 5), (2, 7, 9, 16)])
def test_adaptive_max_pool(dev, dtype, shape):
    cases.maxpool_case(dev, dtype, *shape)


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("shape", [(2, 6, 7, 16), (1, 5, 4, 6), (2, 9, 8, 64)])
def test_spatial_channel_attention(dev, dtype, shape):
    cases.sca_case(dev, dtype, *shape)


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_silu(dev, dtype):
    cases.silu_case(dev, dtype)


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_dropout(dev, dtype):
    cases.dropout_case(dev, dtype)


def test_attention_dropout(dev):
    cases.na_dropout_case(dev, 1, 9, 8, 2, 8, 3, 1)


@pytest.mark.parametrize("geom", [(3, 2, 37, 50, 12, 4, True), (2, 3, 24, 24, 12, 2, False), (1, 1, 5, 9, 8, 2, True), (2, 2, 45, 64, 16, 4, True)])
def test_window_load_matches_reference_windowing(dev, geom):
    cases.window_load_case(dev, *geom)


@pytest.mark.parametrize("geom", [(37, 50, 12, 4, 1), (24, 24, 12, 2, 2), (5, 9, 8, 2, 1)])
def test_predict_pack_matches_reference_writer(dev, geom):
    cases.predict_pack_case(dev, *geom)


@pytest.mark.parametrize("streaming", [False, True])
def test_tile_predictor_matches_window_by_window_reference_pipeline(dev, streaming):
    cases.tile_predictor_case(dev, streaming=streaming)


def test_window_load_division_is_correctly_rounded_for_every_int16_value(dev):
    cases.window_load_all_values_case(dev)


def test_neighborhood_attention_key_side_pass_both_tilings(dev):
    """The key-side backward pass has an image-tile and a sub-image-tile kernel (CNB_NA_DKV=img|group, read once per process); the
    default picks per shape, so the other one is exercised here in a child process."""
    import os
    import subprocess
    import sys

    code = (
        "import sys, torch; sys.path.insert(0, %r)\n"
        "from cultionet_b200 import _lib\n"
        "from tests.emu.build_emu import build\n"
        "_lib.use_library(build())\n"
        "from tests import cases\n"
        "for cfg in [(1, 64, 3, 2, 21, 37), (1, 32, 7, 2, 33, 35), (2, 32, 3, 1, 19, 35)]:\n"
        "    heads, hd, k, d, H, W = cfg\n"
        "    cases.na_case('cpu', torch.bfloat16, 1, H, W, heads, hd, k, d)\n"
        "print('ok')\n" % str(cases.__file__.rsplit('/tests/', 1)[0])
    )
    for variant in ("group", "img"):
        env = dict(os.environ, CNB_NA_DKV=variant)
        res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert res.returncode == 0 and "ok" in res.stdout, (variant, res.stderr[-2000:])
