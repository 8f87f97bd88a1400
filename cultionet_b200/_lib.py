"""ctypes binding of the C ABI declared in ``include/cultionet_b200.h``.

The product path loads exactly one library -- ``cultionet_b200/libcultionet_b200.so`` built by nvcc for sm_100a
(``python -m cultionet_b200.build``) -- and raises when it is missing: there is no CPU or PyTorch fallback.
``use_library(path)`` exists for the CPU test interpreter build (``tests/emu``), which the ``-m "not gpu"`` tests
load explicitly; the package never selects it on its own.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

PKG_DIR = Path(__file__).resolve().parent
DEFAULT_LIB = PKG_DIR / "libcultionet_b200.so"

CNB_F32, CNB_BF16 = 0, 1
CNB_MAX_SRC = 6
TN_MAX_TERMS = 4

_lib = None
_lib_path = None
_is_emulator = False


class CnbError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("src", C.c_void_p * CNB_MAX_SRC),
        ("src_c", C.c_int32 * CNB_MAX_SRC),
        ("src_stride", C.c_int32 * CNB_MAX_SRC),
        ("nsrc", C.c_int32),
        ("B", C.c_int32), ("Hin", C.c_int32), ("Win", C.c_int32), ("Hout", C.c_int32), ("Wout", C.c_int32),
        ("KH", C.c_int32), ("KW", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32), ("dil", C.c_int32),
        ("transposed", C.c_int32),
        ("w_packed", C.c_void_p),
        ("w_tap_stride", C.c_int64),
        ("w_row_stride", C.c_int32),
        ("N", C.c_int32),
        ("bias", C.c_void_p),
        ("out", C.c_void_p),
        ("out_stride", C.c_int32),
        ("stats", C.c_void_p),
        ("nout", C.c_int32),
        ("out_seg", C.c_void_p * CNB_MAX_SRC),
        ("out_seg_c", C.c_int32 * CNB_MAX_SRC),
        ("out_seg_stride", C.c_int32 * CNB_MAX_SRC),
        ("ep_scale", C.c_void_p), ("ep_shift", C.c_void_p), ("ep_act", C.c_int32),
    ]


class PackDesc(C.Structure):
    _fields_ = [
        ("w", C.c_void_p), ("wp", C.c_void_p), ("wd", C.c_void_p),
        ("taps", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("pitch_k", C.c_int32), ("pitch_n", C.c_int32),
        ("tile0", C.c_int32), ("tiles_x", C.c_int32), ("reserved", C.c_int32),
        ("s_n", C.c_int64), ("s_k", C.c_int64), ("s_tap", C.c_int64),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("src_c", C.c_int32), ("src_stride", C.c_int32),
        ("k_off", C.c_int32), ("Ctot", C.c_int32),
        ("B", C.c_int32), ("Hin", C.c_int32), ("Win", C.c_int32), ("Hout", C.c_int32), ("Wout", C.c_int32),
        ("KH", C.c_int32), ("KW", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32), ("dil", C.c_int32),
        ("transposed", C.c_int32),
        ("dy", C.c_void_p), ("dy_stride", C.c_int32), ("N", C.c_int32),
        ("dwp", C.c_void_p),
    ]


class MultiCopyEntry(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("n", C.c_int32), ("mode", C.c_int32)]


class TanimotoTerm(C.Structure):
    _fields_ = [
        ("pred", C.c_void_p), ("target", C.c_void_p), ("mask", C.c_void_p), ("dpred", C.c_void_p),
        ("C", C.c_int32), ("tgt_c", C.c_int32),
        ("target_mode", C.c_int32), ("mask_mode", C.c_int32), ("edge_class", C.c_int32),
        ("weight", C.c_float),
    ]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes; every function returns int
_PROTOS = {
    "cnb_conv2d_fwd": [C.POINTER(ConvDesc), _i, _vp],
    "cnb_conv2d_fwd_generic": [C.POINTER(ConvDesc), _i, _vp],
    "cnb_conv2d_fwd_tc": [C.POINTER(ConvDesc), _i, _vp],
    "cnb_conv2d_tc_eligible": [C.POINTER(ConvDesc), _i],
    "cnb_conv2d_fwd_tiny": [C.POINTER(ConvDesc), _i, _vp],
    "cnb_conv2d_wgrad": [C.POINTER(WgradDesc), _i, _vp],
    "cnb_conv2d_wgrad_generic": [C.POINTER(WgradDesc), _i, _vp],
    "cnb_conv2d_wgrad_tc": [C.POINTER(WgradDesc), _i, _vp],
    "cnb_conv2d_wgrad_tc_eligible": [C.POINTER(WgradDesc), _i],
    "cnb_conv2d_wgrad_tiny": [C.POINTER(WgradDesc), _i, _vp],
    "cnb_repitch": [_vp, _i, _vp, _i, _i64, _i, _i, _vp],
    "cnb_multi_copy": [C.POINTER(MultiCopyEntry), _i, _i, _vp],
    "cnb_broadcast_pixels": [_vp, _vp, _i, _i64, _i, _i, _vp],
    "cnb_pack_weight": [_vp, _vp, _i, _i, _i, _i, _i, _i64, _i64, _i64, _vp],
    "cnb_pack_weight2": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i64, _i64, _i64, _vp],
    "cnb_pack_weights_batched": [_vp, _i, _i, _i, _i, _vp],
    "cnb_unpack_wgrad": [_vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i, _vp],
    "cnb_unpack_wgrads_batched": [_vp, _i, _i, _i, _vp],
    "cnb_bias_grad": [_vp, _i, _i64, _i, _vp, _i, _i, _vp],
    "cnb_bn_stats": [_vp, _i64, _i, _i, _i, _vp, _i, _vp],
    "cnb_bn_finalize": [_vp, _i64, _i, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "cnb_bn_act_fwd": [_vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp],
    "cnb_bn_train_fwd": [_vp, _vp, _i64, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp],
    "cnb_bn_act_bwd_reduce": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _i, _vp],
    "cnb_bn_act_bwd_reduce_acc": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _i, _vp],
    "cnb_bn_act_bwd_apply": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp],
    "cnb_add_n": [_vp, _vp, _vp, _vp, _vp, _i64, _i, _vp],
    "cnb_layernorm_fwd": [_vp, _vp, _vp, _f, _vp, _vp, _vp, _i64, _i, _i, _vp],
    "cnb_layernorm_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp],
    "cnb_na2d_tiled_eligible": [_i, _i, _i, _i, _i, _i, _i, _i],
    "cnb_na2d_fwd": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp],
    "cnb_na2d_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp],
    "cnb_na2d_bwd_workspace_floats": [_i, _i, _i, _i, _i, _i, _i, _i],
    "cnb_resize_bilinear_fwd": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "cnb_resize_bilinear_bwd": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "cnb_resize_bilinear_bwd_colsum": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _vp],
    "cnb_pretime_conv_fwd": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "cnb_pretime_conv_wgrad": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "cnb_time_to_pixel_major": [_vp, _vp, _i, _i, _i64, _i, _i, _vp],
    "cnb_toeplitz_expand": [_vp, _vp, _i, _i, _i, _i, _vp],
    "cnb_toeplitz_fold": [_vp, _vp, _i, _i, _i, _vp],
    "cnb_tap_shift_add": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "cnb_tap_shift_gather": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "cnb_final_combine_fwd": [_vp, _vp, _vp, _vp, _f, _i, _vp, _vp, _vp, _i64, _i, _vp],
    "cnb_final_combine_bwd": [_vp, _vp, _vp, _vp, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _vp],
    "cnb_tanimoto_fwd": [C.POINTER(TanimotoTerm), _i, _i, _i64, _f, _i, _i, _vp, _vp, _vp, _vp],
    "cnb_val_counts": [_vp, _vp, _vp, _vp, _vp, _i64, _i, _f, _vp, _vp],
    "cnb_tanimoto_bwd": [C.POINTER(TanimotoTerm), _i, _i, _i64, _vp, _vp, _vp],
    "cnb_grad_sqnorm": [_vp, _i64, _vp, _vp],
    "cnb_adamw_step": [_vp, _vp, _vp, _vp, _i64, _vp, _f, _f, _f, _f, _f, _f, _vp, _vp],
    "cnb_adaptive_maxpool_fwd": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "cnb_adaptive_maxpool_bwd": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "cnb_silu_fwd": [_vp, _vp, _i64, _i, _vp],
    "cnb_silu_bwd": [_vp, _vp, _vp, _i64, _i, _vp],
    "cnb_act_fwd": [_vp, _vp, _i64, _i, _i, _vp],
    "cnb_act_bwd": [_vp, _vp, _vp, _i64, _i, _i, _vp],
    "cnb_sca_slices": [_i, _i, _i, _i],
    "cnb_sca_pool_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "cnb_sca_pool_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "cnb_sca_apply_fwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "cnb_sca_apply_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "cnb_rng_advance": [_vp, _vp],
    "cnb_na2d_dropout_fwd": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _i, _f, _i, _vp],
    "cnb_na2d_dropout_bwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _i, _f, _i, _vp],
    "cnb_dropout": [_vp, _vp, _i64, _vp, _i, _f, _i, _vp],
    "cnb_dropout2d": [_vp, _vp, _i, _i, _i, _vp, _i, _f, _i, _vp],
    "cnb_window_load": [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _f, _f, _f, _vp, _vp, _vp, _vp],
    "cnb_predict_pack": [_vp, _vp, _vp, _i64, _i, _i, _i, _vp, _i, _i, _f, _vp, _i, _i, _i, _vp],
}

EXPORTED_SYMBOLS = ["cnb_version", "cnb_sm_arch", "cnb_last_error", "cnb_launch_count", *_PROTOS.keys()]


def _bind(lib) -> None:
    lib.cnb_version.restype = C.c_int
    lib.cnb_sm_arch.restype = C.c_int
    lib.cnb_last_error.restype = C.c_char_p
    lib.cnb_launch_count.restype = C.c_int64
    for name, argtypes in _PROTOS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int64 if name.endswith("_workspace_floats") else C.c_int


def use_library(path) -> None:
    """Bind a specific build of the C ABI (tests use this for the CPU interpreter build)."""
    global _lib, _lib_path, _is_emulator
    path = Path(path)
    if not path.is_file():
        raise CnbError(f"cultionet_b200: native library not found at {path}")
    lib = C.CDLL(str(path))
    _bind(lib)
    _lib, _lib_path = lib, path
    _is_emulator = lib.cnb_sm_arch() == 0


def lib():
    """The bound library; loads the nvcc build on first use and fails loudly when it is absent."""
    if _lib is None:
        if not DEFAULT_LIB.is_file():
            raise CnbError(
                f"cultionet_b200: {DEFAULT_LIB.name} is not built (run `python -m cultionet_b200.build`); "
                "there is no CPU or PyTorch fallback for the TowerUNet hot path."
            )
        use_library(DEFAULT_LIB)
    return _lib


def library_path():
    lib()
    return _lib_path


def is_emulator() -> bool:
    lib()
    return _is_emulator


def has_symbol(name: str) -> bool:
    return hasattr(lib(), name)


class KernelTimer:
    """Optional per-call CUDA-event timing (bench.py's roofline pass): records (name, flops, bytes, start, end) per C-ABI call on
    the stream the kernels are launched on."""

    def __init__(self):
        self.records = []

    def summary(self) -> dict:
        torch.cuda.synchronize()
        out: dict = {}
        for name, flops, nbytes, e0, e1, _detail in self.records:
            d = out.setdefault(name, {"calls": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["calls"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += flops or 0.0
            d["bytes"] += nbytes or 0.0
        return out


    def by_detail(self) -> dict:
        """Same records keyed by (tag, detail), e.g. one entry per convolution shape."""
        torch.cuda.synchronize()
        out: dict = {}
        for name, flops, nbytes, e0, e1, detail in self.records:
            d = out.setdefault(f"{name} {detail or ''}".strip(), {"calls": 0, "ms": 0.0, "flops": 0.0})
            d["calls"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += flops or 0.0
        return out


TIMER: "KernelTimer | None" = None


def _es(code: int) -> int:
    return 2 if code == CNB_BF16 else 4


def _nn(*ptrs) -> int:
    return sum(1 for p in ptrs if p)


# Algorithmic HBM bytes of the bandwidth-bound entry points, from the C-ABI arguments by position; evaluated only while
# a KernelTimer is recording (bench.py's roofline pass).
ALG_BYTES = {
    # int16 in + fp32 out per element of the window batch; fp32 in + uint16 out per kept pixel of three bands
    "cnb_window_load": lambda a: 6 * a[7] * a[2] * a[1] * (a[8] + 2 * a[9]) ** 2,
    "cnb_predict_pack": lambda a: 6 * 3 * a[8] * a[9] * a[9],
    "cnb_bn_stats": lambda a: a[1] * a[2] * _es(a[6]),
    "cnb_bn_train_fwd": lambda a: (2 + _nn(a[13])) * a[15] * a[16] * _es(a[20]),
    "cnb_bn_act_fwd": lambda a: (2 + _nn(a[3])) * a[5] * a[6] * _es(a[10]),
    "cnb_bn_act_bwd_reduce": lambda a: 2 * a[6] * a[7] * _es(a[12]),
    "cnb_bn_act_bwd_reduce_acc": lambda a: 2 * a[6] * a[7] * _es(a[12]),
    "cnb_bn_act_bwd_apply": lambda a: 3 * a[9] * a[10] * _es(a[15]),
    "cnb_add_n": lambda a: (_nn(a[0], a[1], a[2], a[3]) + 1) * a[5] * _es(a[6]),
    "cnb_layernorm_fwd": lambda a: 2 * a[7] * a[8] * _es(a[9]),
    "cnb_layernorm_bwd": lambda a: 3 * a[8] * a[9] * _es(a[10]),
    "cnb_na2d_fwd": lambda a: 4 * a[3] * a[4] * a[5] * a[6] * a[7] * _es(a[11]),
    "cnb_na2d_bwd": lambda a: 8 * a[8] * a[9] * a[10] * a[11] * a[12] * _es(a[16]),
    "cnb_resize_bilinear_fwd": lambda a: a[2] * a[7] * _es(a[8]) * (a[3] * a[4] + a[5] * a[6]),
    "cnb_resize_bilinear_bwd": lambda a: a[2] * a[7] * _es(a[8]) * (a[3] * a[4] + a[5] * a[6]),
    "cnb_resize_bilinear_bwd_colsum": lambda a: a[2] * a[7] * _es(a[10]) * (a[3] * a[4] + a[5] * a[6]),
    "cnb_time_to_pixel_major": lambda a: a[2] * a[4] * (a[3] * 4 + a[5] * _es(a[6])),
    "cnb_bias_grad": lambda a: a[2] * a[3] * _es(a[6]),
    "cnb_adamw_step": lambda a: 28 * a[4],
    "cnb_grad_sqnorm": lambda a: 4 * a[1],
    "cnb_tanimoto_fwd": lambda a: 24 * a[2] * a[3],
    "cnb_tanimoto_bwd": lambda a: 36 * a[2] * a[3],
}


def launch_count() -> int:
    return int(lib().cnb_launch_count())


def call(name: str, *args, flops: float = None, nbytes: float = None, tag: str = None, detail: str = None) -> None:
    fn = getattr(lib(), name)
    if TIMER is not None and not _is_emulator:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        if nbytes is None and name in ALG_BYTES:
            try:
                nbytes = float(ALG_BYTES[name](args))
            except Exception:  # noqa: BLE001 - accounting must never break a launch
                nbytes = None
        TIMER.records.append((tag or name, flops, nbytes, e0, e1, detail))
    else:
        rc = fn(*args)
    if rc != 0:
        msg = lib().cnb_last_error().decode("utf-8", "replace")
        raise CnbError(f"{name} failed (code {rc}): {msg}")


def dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return CNB_F32
    if dtype == torch.bfloat16:
        return CNB_BF16
    raise CnbError(f"cultionet_b200: unsupported activation dtype {dtype} (float32 or bfloat16)")


def check_device(*tensors) -> None:
    """Tensors must live on a CUDA device (or on the CPU when the test interpreter build is bound)."""
    emu = is_emulator()
    for t in tensors:
        if t is None:
            continue
        if emu:
            if t.device.type != "cpu":
                raise CnbError("the CPU test interpreter build takes CPU tensors")
        elif t.device.type != "cuda":
            raise CnbError(
                "cultionet_b200 kernels need CUDA tensors (sm_100a); there is no CPU fallback -- move the model and batch to cuda"
            )


def stream_ptr(ref: torch.Tensor):
    if ref.device.type == "cuda":
        return C.c_void_p(torch.cuda.current_stream(ref.device).cuda_stream)
    return C.c_void_p(0)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
