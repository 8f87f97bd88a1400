"""Autograd operators of the TowerUNet hot path, each a thin wrapper over the C ABI (``include/cultionet_b200.h``).

Activations are pixel-major ``[B, H, W, C]`` tensors (float32 in parity mode, bfloat16 in throughput mode); parameters
stay fp32 in the reference's own layouts so reference checkpoints load unchanged, and are repacked on the fly for the
kernels.  Nothing here computes with torch ops: torch only owns memory, streams and the autograd tape.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ConvDesc, TanimotoTerm, WgradDesc, call, check_device, dtype_code, ptr, stream_ptr

# weight layouts (of the fp32 parameter)
W_CONV = "conv"      # nn.Conv2d            [N, Ctot, KH, KW]
W_CONVT = "convT"    # nn.ConvTranspose2d   [Ctot, N, KH, KW]
W_LINEAR = "linear"  # nn.Linear            [N, Ctot]


def _contig(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _weight_strides(kind: str, N: int, K: int, taps: int, for_dgrad: bool):
    """(rows, cols, s_n, s_k, s_tap) of the packed [taps][rows][cols] view of a parameter; see cnb_pack_weight."""
    if kind in (W_CONV, W_LINEAR):  # parameter [N][K][taps]
        s_out, s_in = K * taps, taps
    elif kind == W_CONVT:  # parameter [K][N][taps]
        s_out, s_in = taps, N * taps
    else:
        raise ValueError(kind)
    if not for_dgrad:
        return N, K, s_out, s_in, 1
    return K, N, s_in, s_out, 1


def pack_weight(weight: torch.Tensor, kind: str, N: int, K: int, taps: int, dtype: torch.dtype, for_dgrad: bool) -> torch.Tensor:
    """[taps][rows][pitch] packed copy of a parameter; the row pitch is rounded up to 8 elements in bf16 (16-byte rows for TMA),
    the padding columns are zero."""
    rows, cols, s_n, s_k, s_tap = _weight_strides(kind, N, K, taps, for_dgrad)
    w = _contig(weight)
    pitch = (cols + 7) // 8 * 8 if dtype == torch.bfloat16 else cols
    out = torch.empty((taps, rows, pitch), dtype=dtype, device=w.device)
    call("cnb_pack_weight", ptr(w), ptr(out), dtype_code(dtype), taps, rows, cols, pitch, s_n, s_k, s_tap, stream_ptr(w))
    return out


# Packed copies of nn.Parameters, valid until the parameter changes: torch in-place updates bump `_version`; the flat AdamW kernel
# writes through raw pointers and calls invalidate_packed_weights() instead.  Inference therefore packs each weight once, and a
# training step reads each parameter once (forward and data-gradient layouts come out of one launch).
_PACK_CACHE: dict = {}
_PACK_CACHE_MAX = 4096


def invalidate_packed_weights() -> None:
    _PACK_CACHE.clear()


def _pack_geometry(kind: str, N: int, K: int, taps: int, dtype: torch.dtype):
    _, _, s_n, s_k, s_tap = _weight_strides(kind, N, K, taps, False)
    bf = dtype == torch.bfloat16
    return s_n, s_k, s_tap, ((K + 7) // 8 * 8 if bf else K), ((N + 7) // 8 * 8 if bf else N)


# Every Parameter that went through packed_weights() is remembered here ("the plan"), so that a training loop can pack ALL of them
# with one launch at the top of a step (prepack_parameters) instead of one launch per layer as the forward reaches it.
_PACK_PLAN: dict = {}       # key -> [weakref, kind, N, K, taps, dtype, need_dgrad]
_PACK_BATCH: dict = {}      # dtype -> persistent state of the batched launch (buffers + device descriptor table)
_PACK_PLAN_DIRTY = [True]


def packed_weights(weight: torch.Tensor, kind: str, N: int, K: int, taps: int, dtype: torch.dtype, need_dgrad: bool):
    """(wp, wd): forward layout [taps][N][K-pitch] and, if ``need_dgrad``, the data-gradient layout [taps][K][N-pitch] (else None)."""
    key = None
    if isinstance(weight, torch.nn.Parameter):
        # keyed by the Parameter OBJECT (a weak reference guards against id / address reuse after a model is freed)
        key = (id(weight), kind, dtype)
        stamp = (weight.data_ptr(), weight._version, tuple(weight.shape))
        hit = _PACK_CACHE.get(key)
        if hit is not None and hit[0]() is weight and hit[1] == stamp and (hit[3] is not None or not need_dgrad):
            return hit[2], hit[3]
        plan = _PACK_PLAN.get(key)
        if plan is None or plan[0]() is not weight or (need_dgrad and not plan[6]):
            _PACK_PLAN[key] = [weakref.ref(weight), kind, N, K, taps, dtype, bool(need_dgrad or (plan is not None and plan[0]() is weight and plan[6]))]
            _PACK_PLAN_DIRTY[0] = True
    s_n, s_k, s_tap, pitch_k, pitch_n = _pack_geometry(kind, N, K, taps, dtype)
    w = _contig(weight.detach())
    wp = torch.empty((taps, N, pitch_k), dtype=dtype, device=w.device)
    wd = torch.empty((taps, K, pitch_n), dtype=dtype, device=w.device) if need_dgrad else None
    call("cnb_pack_weight2", ptr(w), ptr(wp), ptr(wd), dtype_code(dtype), taps, N, K, pitch_k, pitch_n, s_n, s_k, s_tap, stream_ptr(w))
    if key is not None:
        if len(_PACK_CACHE) >= _PACK_CACHE_MAX:
            _PACK_CACHE.clear()
        _PACK_CACHE[key] = (weakref.ref(weight), stamp, wp, wd)
    return wp, wd


def prepack_parameters() -> int:
    """Packs every planned Parameter with ONE launch per compute dtype and refreshes the cache; returns the number packed.  Call at
    the top of a training step (after the optimizer has changed the parameters).  The packed buffers and the device-resident
    descriptor table persist between calls, so under CUDA-graph capture this is a single kernel node."""
    if _PACK_PLAN_DIRTY[0]:
        for k in [k for k, e in _PACK_PLAN.items() if e[0]() is None]:
            del _PACK_PLAN[k]
        _PACK_BATCH.clear()
        by_dtype: dict = {}
        for key, e in _PACK_PLAN.items():
            w = e[0]()
            if w.is_contiguous():
                by_dtype.setdefault((e[5], w.device), []).append((key, e))
        for (dtype, dev), entries in by_dtype.items():
            table = (_lib.PackDesc * len(entries))()
            bufs, tile0, max_taps = [], 0, 1
            for i, (key, (ref, kind, N, K, taps, _, need_dgrad)) in enumerate(entries):
                w = ref()
                s_n, s_k, s_tap, pitch_k, pitch_n = _pack_geometry(kind, N, K, taps, dtype)
                wp = torch.empty((taps, N, pitch_k), dtype=dtype, device=dev)
                wd = torch.empty((taps, K, pitch_n), dtype=dtype, device=dev) if need_dgrad else None
                d = table[i]
                d.w, d.wp, d.wd = w.data_ptr(), wp.data_ptr(), (wd.data_ptr() if wd is not None else None)
                d.taps, d.N, d.K, d.pitch_k, d.pitch_n = taps, N, K, pitch_k, pitch_n
                d.s_n, d.s_k, d.s_tap = s_n, s_k, s_tap
                tiles_x = (max(K, pitch_k) + 31) // 32
                tiles_y = (max(N, pitch_n if need_dgrad else N) + 31) // 32
                d.tile0, d.tiles_x = tile0, tiles_x
                tile0 += tiles_x * tiles_y
                max_taps = max(max_taps, taps)
                bufs.append((key, ref, w.data_ptr(), wp, wd))
            raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).clone()
            _PACK_BATCH[(dtype, dev)] = (raw.to(dev), len(entries), tile0, max_taps, bufs)
        _PACK_PLAN_DIRTY[0] = False
    n = 0
    for (dtype, dev), (table_dev, ndesc, total_tiles, max_taps, bufs) in _PACK_BATCH.items():
        # a parameter whose storage moved since the table was built (e.g. FlatAdamW re-homed it) invalidates the table
        if any(ref() is None or ref().data_ptr() != p0 for _, ref, p0, _, _ in bufs):
            _PACK_PLAN_DIRTY[0] = True
            return prepack_parameters()
        call("cnb_pack_weights_batched", ptr(table_dev), ndesc, total_tiles, max_taps, dtype_code(dtype), stream_ptr(table_dev))
        for key, ref, _, wp, wd in bufs:
            w = ref()
            _PACK_CACHE[key] = (ref, (w.data_ptr(), w._version, tuple(w.shape)), wp, wd)
        n += ndesc
    return n


def _conv_out_size(n: int, k: int, stride: int, pad: int, dil: int) -> int:
    return (n + 2 * pad - dil * (k - 1) - 1) // stride + 1


def _convT_out_size(n: int, k: int, stride: int, pad: int, dil: int) -> int:
    return (n - 1) * stride - 2 * pad + dil * (k - 1) + 1


TINY_MAX_C = 16  # csrc/k_conv_tiny.cuh

# "auto": tcgen05 kernel when eligible, CUDA-core kernel otherwise; "generic" / "tc" force one (tests, A/B timing)
CONV_BACKEND = "auto"
_CONV_ENTRY = {"auto": "cnb_conv2d_fwd", "generic": "cnb_conv2d_fwd_generic", "tc": "cnb_conv2d_fwd_tc"}
_WGRAD_ENTRY = {"auto": "cnb_conv2d_wgrad", "generic": "cnb_conv2d_wgrad_generic", "tc": "cnb_conv2d_wgrad_tc"}


STATS_MAX_N = 1024  # csrc/k_conv_tc.cuh
MERGE_SOURCE_DGRADS = os.environ.get("CNB_MERGE_DGRAD", "1") != "0"  # data gradient of a multi-source convolution as one split-output launch (A/B switch for bench runs)


# Zero-initialised [2, N] slices for the BatchNorm sums the tcgen05 epilogue accumulates: one arena per device, cleared by ONE fill at the
# top of a forward (reset_stats_arena, called by TowerUNet.forward in training mode) instead of one torch.zeros per convolution.  Slices
# are only handed out once between two resets, so a slice is always clean; when the arena runs out the caller gets a fresh tensor.
_STATS_ARENA: dict = {}
_STATS_ARENA_FLOATS = 1 << 17


def reset_stats_arena(dev) -> None:
    dev = torch.device(dev)
    if dev.type != "cuda":
        return
    ent = _STATS_ARENA.get(dev)
    if ent is None:
        _STATS_ARENA[dev] = [torch.zeros(_STATS_ARENA_FLOATS, dtype=torch.float32, device=dev), 0]
        return
    if ent[1] > 0:
        ent[0][:ent[1]].zero_()
        ent[1] = 0


def _stats_slice(N: int, dev) -> torch.Tensor:
    ent = _STATS_ARENA.get(torch.device(dev))
    n = 2 * N
    if ent is None or ent[1] + n > _STATS_ARENA_FLOATS:
        return torch.zeros((2, N), dtype=torch.float32, device=dev)
    off = ent[1]
    ent[1] = off + (n + 31) // 32 * 32
    return ent[0][off:off + n].view(2, N)


def _launch_conv(sources, src_channels, wp, w_row_off, w_row_stride, w_tap_stride, N, bias, out, geom, transposed, dtype,
                 want_stats=False, out_segments=None, epilogue=None):
    """Launches the convolution; with ``want_stats`` returns the [2, N] fp32 BatchNorm partial sums the tcgen05 epilogue produced
    (None when the shape takes another kernel and the caller has to run ``cnb_bn_stats``)."""
    d = ConvDesc()
    B, Hin, Win, Hout, Wout, KH, KW, stride, pad, dil = geom
    esize = wp.element_size()
    for s, (t, c) in enumerate(zip(sources, src_channels)):
        d.src[s] = t.data_ptr()
        d.src_c[s] = c
        d.src_stride[s] = t.shape[-1]
    d.nsrc = len(sources)
    d.B, d.Hin, d.Win, d.Hout, d.Wout = B, Hin, Win, Hout, Wout
    d.KH, d.KW, d.stride, d.pad, d.dil, d.transposed = KH, KW, stride, pad, dil, int(transposed)
    d.w_packed = wp.data_ptr() + w_row_off * w_row_stride * esize
    d.w_tap_stride = w_tap_stride
    d.w_row_stride = w_row_stride
    d.N = N
    d.bias = bias.data_ptr() if bias is not None else None
    if out_segments is None:
        d.out = out.data_ptr()
        d.out_stride = out.shape[-1]
        d.nout = 0
    else:  # split output: [(tensor, channels), ...] covering the N output channels in order
        d.out, d.out_stride = None, 0
        d.nout = len(out_segments)
        for i, (t, c) in enumerate(out_segments):
            d.out_seg[i] = t.data_ptr()
            d.out_seg_c[i] = c
            d.out_seg_stride[i] = t.shape[-1]
        out = out_segments[0][0]
        if not (CONV_BACKEND != "generic" and not _lib.is_emulator() and _lib.lib().cnb_conv2d_tc_eligible(C.byref(d), dtype_code(dtype))):
            return False
    if epilogue is not None:  # (scale, shift, act): eval-mode BatchNorm + activation in the tcgen05 epilogue; False when not available
        if CONV_BACKEND == "generic" or _lib.is_emulator() or bias is not None or out_segments is not None:
            return False
        d.ep_scale, d.ep_shift, d.ep_act = epilogue[0].data_ptr(), epilogue[1].data_ptr(), int(epilogue[2])
        d.stats = None
        if not _lib.lib().cnb_conv2d_tc_eligible(C.byref(d), dtype_code(dtype)):
            return False
    d.stats = None
    stats = None
    if want_stats and N <= STATS_MAX_N and CONV_BACKEND != "generic" and not _lib.is_emulator():
        if _lib.lib().cnb_conv2d_tc_eligible(C.byref(d), dtype_code(dtype)) == 2:
            stats = _stats_slice(N, out.device)
            d.stats = stats.data_ptr()
    flops = 2.0 * B * (Hin * Win if transposed else Hout * Wout) * N * sum(src_channels) * KH * KW  # algorithmic
    detail = None
    if _lib.TIMER is not None:
        detail = f"B{B} {Hin}x{Win}->{Hout}x{Wout} {'+'.join(map(str, src_channels))}->{N} k{KH} s{stride}{'T' if transposed else ''}"
    if _lib.TIMER is not None and out_segments is not None:
        detail = detail.replace(f"->{N} ", "->" + "+".join(str(c) for _, c in out_segments) + " ")
    call(_CONV_ENTRY[CONV_BACKEND], C.byref(d), dtype_code(dtype), stream_ptr(out), flops=flops, tag="conv_fwd/dgrad", detail=detail)
    return True if (out_segments is not None or epilogue is not None) else stats


# With DIRECT_PARAM_GRAD the weight / bias gradient kernels accumulate straight into ``param.grad`` (the flat gradient buffer of
# optim.FlatAdamW) and the backward returns None for them: no temporary gradient tensor and no AccumulateGrad add per parameter.
# engine.TrainStep switches it on when no per-parameter gradient hook has to fire (no overlapped bucketed all-reduce).
DIRECT_PARAM_GRAD = [False]
_DIRECT_ZEROED = [False]
_DIRECT_WRITTEN: set = set()  # parameters whose gradient buffer already holds a contribution of the running backward


class direct_param_grads:
    """``with direct_param_grads(): loss.backward()`` -- the gradient buffers must be ZERO (or stale) on entry: the first contribution
    to a parameter overwrites its buffer (a plain store instead of a read-modify-write), later ones (a shared parameter) accumulate."""

    def __init__(self, enabled: bool = True, zeroed: bool = False):
        self.enabled = enabled
        self.zeroed = zeroed  # the caller has just cleared every gradient buffer: every contribution accumulates, no kernel clears

    def __enter__(self):
        self.prev = DIRECT_PARAM_GRAD[0]
        self.prev_zeroed = _DIRECT_ZEROED[0]
        _DIRECT_ZEROED[0] = bool(self.zeroed)
        DIRECT_PARAM_GRAD[0] = bool(self.enabled)
        _DIRECT_WRITTEN.clear()
        _PENDING_SMALL.clear()
        return self

    def __exit__(self, *exc):
        DIRECT_PARAM_GRAD[0] = self.prev
        _DIRECT_ZEROED[0] = self.prev_zeroed
        _DIRECT_WRITTEN.clear()
        if exc and exc[0] is not None:  # a failed backward: drop the deferred work, leave the accumulators clean
            for _, dwp, *_ in _PENDING_UNPACK.values():
                dwp.zero_()
            _PENDING_UNPACK.clear()
            _PENDING_SMALL.clear()
        else:
            flush_weight_gradients()
        return False


def _direct_grad_target(param, like_shape, full_overwrite: bool = False):
    """(gradient buffer, accumulate flag) when the kernel may write ``param.grad`` itself, else (None, 0).  ``full_overwrite``: the
    kernel stores every element (the weight-gradient unpack), so a first contribution overwrites even when the buffer is known to be
    zero -- no read of the buffer; kernels that ATOMICALLY add (bias / column sums) accumulate into known-zero buffers instead of
    clearing them first."""
    if not DIRECT_PARAM_GRAD[0] or not isinstance(param, torch.nn.Parameter):
        return None, 0
    g = param.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or tuple(g.shape) != tuple(like_shape) or g.requires_grad:
        return None, 0
    first = id(param) not in _DIRECT_WRITTEN and (full_overwrite or not _DIRECT_ZEROED[0])
    _DIRECT_WRITTEN.add(id(param))
    return g, 0 if first else 1


# Persistent fp32 accumulators [taps][N][Ctot] of the split-K weight-gradient kernels, one per Parameter taking its gradient
# directly: ``cnb_unpack_wgrad`` (mode bit 1) clears the accumulator while reading it, so a step needs no fill launch per weight.
_DWP_ACC: dict = {}


def _wgrad_accumulator(param, taps, N, Ctot, dev):
    key = id(param)
    hit = _DWP_ACC.get(key)
    if hit is not None and hit[0]() is param and hit[1].shape == (taps, N, Ctot) and hit[1].device == dev:
        return hit[1]
    buf = torch.zeros((taps, N, Ctot), dtype=torch.float32, device=dev)
    _DWP_ACC[key] = (weakref.ref(param, lambda _r, k=key: _DWP_ACC.pop(k, None)), buf)
    return buf


_DERIVED_ACC: dict = {}


def _derived_accumulator(key, taps, N, Ctot, dev):
    hit = _DERIVED_ACC.get(key)
    if hit is not None and hit.shape == (taps, N, Ctot) and hit.device == dev:
        return hit
    if len(_DERIVED_ACC) > 256:
        _DERIVED_ACC.clear()
    buf = _DERIVED_ACC[key] = torch.zeros((taps, N, Ctot), dtype=torch.float32, device=dev)
    return buf


# Weight gradients taken directly are unpacked from their accumulators into ``param.grad`` by ONE launch when the backward pass
# ends (``direct_param_grads.__exit__``): id(param) -> (param ref, accumulator, gradient buffer, taps, rows, cols, strides, mode).
_PENDING_UNPACK: dict = {}
_UNPACK_TABLES: dict = {}  # signature of a pending set -> (device table, ndesc, total_tiles, max_taps)
BATCH_UNPACK = os.environ.get("CNB_BATCH_UNPACK", "1") != "0"


def _defer_unpack(param, dwp, target, taps, rows, cols, s_n, s_k, s_tap, acc_flag) -> bool:
    if not BATCH_UNPACK or taps > 32:
        return False
    key = id(param)
    if key not in _PENDING_UNPACK:  # a shared weight: its launches accumulate in the same accumulator, one unpack serves them all
        _PENDING_UNPACK[key] = (param, dwp, target, taps, rows, cols, s_n, s_k, s_tap, acc_flag | 2)
    return True


def flush_small_gradients() -> int:
    if not _PENDING_SMALL:
        return 0
    entries = [(src.data_ptr(), dst.data_ptr(), src.numel(), mode) for src, dst, mode in _PENDING_SMALL]
    ref = _PENDING_SMALL[0][1]
    keep = list(_PENDING_SMALL)  # the sources stay referenced until the launch is enqueued
    _PENDING_SMALL.clear()
    multi_copy(entries, ref)
    del keep
    return len(entries)


def flush_weight_gradients() -> int:
    """Unpack every deferred weight-gradient accumulator into its gradient buffer (and clear it) with one launch; returns how many."""
    flush_small_gradients()
    if not _PENDING_UNPACK:
        return 0
    entries = list(_PENDING_UNPACK.values())
    _PENDING_UNPACK.clear()
    sig = tuple((e[1].data_ptr(), e[2].data_ptr(), e[3], e[4], e[5], e[9]) for e in entries)
    plan = _UNPACK_TABLES.get(sig)
    dev = entries[0][1].device
    if plan is None:
        capturing = dev.type == "cuda" and torch.cuda.is_current_stream_capturing()
        if capturing:  # the table upload is a host-to-device copy: build it in an eager step; here fall back to single launches
            for _, dwp, target, taps, rows, cols, s_n, s_k, s_tap, mode in entries:
                call("cnb_unpack_wgrad", ptr(dwp), ptr(target), taps, rows, cols, s_n, s_k, s_tap, mode, stream_ptr(dwp))
            return len(entries)
        table = (_lib.PackDesc * len(entries))()
        tile0, max_taps = 0, 1
        for d, (_, dwp, target, taps, rows, cols, s_n, s_k, s_tap, mode) in zip(table, entries):
            d.w, d.wp, d.wd = target.data_ptr(), dwp.data_ptr(), None
            d.taps, d.N, d.K, d.pitch_k, d.pitch_n = taps, rows, cols, cols, rows
            d.s_n, d.s_k, d.s_tap, d.reserved = s_n, s_k, s_tap, mode
            d.tiles_x = (cols + 31) // 32
            d.tile0 = tile0
            tile0 += d.tiles_x * ((rows + 31) // 32)
            max_taps = max(max_taps, taps)
        raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).clone().to(dev)
        if len(_UNPACK_TABLES) > 8:
            _UNPACK_TABLES.clear()
        plan = _UNPACK_TABLES[sig] = (raw, len(entries), tile0, max_taps)
    raw, ndesc, total_tiles, max_taps = plan
    call("cnb_unpack_wgrads_batched", ptr(raw), ndesc, total_tiles, max_taps, stream_ptr(raw))
    return ndesc


# ----------------------------------------------------------------------------------------------------------------
# Multi-tensor copies of small fp32 segments (``cnb_multi_copy``): ONE launch for parameter plumbing that would otherwise be one tiny
# torch kernel per tensor -- stacking the Psi-Net stream parameters, scattering the BatchNorm / LayerNorm / scalar parameter gradients
# of a backward pass into the flat gradient buffer.
# ----------------------------------------------------------------------------------------------------------------
def multi_copy(entries, ref: torch.Tensor) -> None:
    """``entries``: ``[(src data_ptr, dst data_ptr, n, mode)]`` over fp32 memory on ``ref``'s device (mode 0 copies, 1 accumulates).  The
    table is a host array handed to ``cnb_multi_copy``, which passes it to the kernel by value."""
    if not entries:
        return
    table = (_lib.MultiCopyEntry * len(entries))()
    max_len = 1
    for e, (src, dst, n, mode) in zip(table, entries):
        e.src, e.dst, e.n, e.mode = src, dst, n, mode
        max_len = max(max_len, n)
    call("cnb_multi_copy", table, len(entries), max_len, stream_ptr(ref))


# Small parameter gradients (BatchNorm / LayerNorm scale and shift, the final-combine scalars, stacked Psi-Net parameters) taken
# directly: (source tensor, gradient buffer, mode) collected during backward, scattered by ONE launch in flush_weight_gradients().
_PENDING_SMALL: list = []


def _defer_small_grad(param, src: torch.Tensor) -> bool:
    """True when ``src`` (fp32, contiguous, ``param.numel()`` elements) will be written into ``param.grad`` at the end of backward; the
    backward then returns None for that parameter."""
    if src.dtype != torch.float32 or not src.is_contiguous():
        return False
    target, acc = _direct_grad_target(param, param.shape)
    if target is None:
        return False
    _PENDING_SMALL.append((src, target, acc))
    return True


def _small_grad(param, src: torch.Tensor):
    """``None`` when the gradient was taken directly, else ``src`` shaped like the parameter (autograd accumulates it)."""
    return None if _defer_small_grad(param, src) else src.view(param.shape)


def _zeros_like_cached(n: int, dev) -> torch.Tensor:
    z = _ZEROS.get(dev)
    if z is None or z.numel() < n:
        z = _ZEROS[dev] = torch.zeros(max(n, 4096), dtype=torch.float32, device=dev)
    return z


_ZEROS: dict = {}


class _StackParamsFn(torch.autograd.Function):
    """``total`` fp32 elements holding parameter ``i`` at ``offsets[i]`` and zeros elsewhere -- ``torch.cat`` / ``F.pad`` of small
    parameters as ONE launch; the backward hands each parameter its slice (directly into its gradient buffer when possible)."""

    @staticmethod
    def forward(ctx, total, offsets, *params):
        check_device(*params)
        ps = [_contig(p.detach()) for p in params]
        out = torch.empty((total,), dtype=torch.float32, device=ps[0].device)
        zeros = _zeros_like_cached(total, out.device)
        entries, pos = [], 0
        for p, off in sorted(zip(ps, offsets), key=lambda t: t[1]):
            if off > pos:
                entries.append((zeros.data_ptr(), out.data_ptr() + 4 * pos, off - pos, 0))
            entries.append((p.data_ptr(), out.data_ptr() + 4 * off, p.numel(), 0))
            pos = off + p.numel()
        if pos < total:
            entries.append((zeros.data_ptr(), out.data_ptr() + 4 * pos, total - pos, 0))
        multi_copy(entries, out)
        ctx.params, ctx.offsets = params, offsets
        return out

    @staticmethod
    def backward(ctx, g):
        g = _contig(g.float())
        grads = []
        for i, (p, off) in enumerate(zip(ctx.params, ctx.offsets)):
            if not ctx.needs_input_grad[2 + i]:
                grads.append(None)
                continue
            grads.append(_small_grad(p, g[off:off + p.numel()]))
        ctx.keep = g
        return (None, None, *grads)


def stack_params(params: Sequence[torch.Tensor], offsets: Sequence[int], total: int) -> torch.Tensor:
    out = _StackParamsFn.apply(int(total), tuple(int(o) for o in offsets), *params)
    out._cnb_acc_key = ("stack",) + tuple(id(p) for p in params)
    return out


def tag_derived(t: torch.Tensor, like: torch.Tensor, *extra) -> torch.Tensor:
    """Carry the accumulator key of a derived weight over a reshape / permute / pad of it."""
    key = getattr(like, "_cnb_acc_key", None)
    if key is None and isinstance(like, torch.nn.Parameter):
        key = ("param", id(like))
    if key is not None:
        t._cnb_acc_key = key + tuple(extra)
    return t


class _Conv2dFn(torch.autograd.Function):
    """y = conv(cat(sources, channel), weight) + bias over pixel-major tensors; see ``cnb_conv2d_fwd``."""

    @staticmethod
    def forward(ctx, weight, bias, kind, ksize, stride, pad, dil, transposed, out_hw, want_stats, src_c, *sources):
        check_device(weight, bias, *sources)
        sources = [_contig(s) for s in sources]
        x0 = sources[0]
        dtype = x0.dtype
        B, Hin, Win = x0.shape[0], x0.shape[1], x0.shape[2]
        # logical channels per source; a source may carry row padding (last dim = pixel pitch > channels)
        src_channels = list(src_c) if src_c is not None else [s.shape[-1] for s in sources]
        assert all(c <= s.shape[-1] for c, s in zip(src_channels, sources))
        Ctot = sum(src_channels)
        KH = KW = ksize
        taps = KH * KW
        if kind == W_CONVT:
            assert transposed
            N = weight.shape[1]
            assert weight.shape[0] == Ctot, (weight.shape, Ctot)
        else:
            N = weight.shape[0]
            assert weight.shape[1] == Ctot, (weight.shape, Ctot)
        if transposed:
            Hout, Wout = _convT_out_size(Hin, KH, stride, pad, dil), _convT_out_size(Win, KW, stride, pad, dil)
        else:
            Hout, Wout = _conv_out_size(Hin, KH, stride, pad, dil), _conv_out_size(Win, KW, stride, pad, dil)
        if out_hw is not None:
            assert tuple(out_hw) == (Hout, Wout), (out_hw, Hout, Wout)
        need_dgrad = any(ctx.needs_input_grad[11 + i] for i in range(len(sources)))
        wp, ctx.wd = packed_weights(weight, kind, N, Ctot, taps, dtype, need_dgrad)
        out = torch.empty((B, Hout, Wout, N), dtype=dtype, device=x0.device)
        geom = (B, Hin, Win, Hout, Wout, KH, KW, stride, pad, dil)
        bias_c = _contig(bias) if bias is not None else None
        stats = _launch_conv(sources, src_channels, wp, 0, wp.shape[2], N * wp.shape[2], N, bias_c, out, geom, transposed, dtype,
                             want_stats=want_stats)
        ctx.save_for_backward(weight, *sources)
        ctx.set_materialize_grads(False)  # the statistics output never has a gradient: no zero tensor (a fill launch) per convolution
        ctx.params = (weight, bias)  # the Parameter objects themselves (their .grad may take the gradient directly)
        ctx.meta = (kind, geom, transposed, src_channels, N, Ctot, bias is not None)
        if not want_stats:
            return out
        if stats is None:
            stats = torch.empty((0,), dtype=torch.float32, device=out.device)  # "not produced": the caller runs cnb_bn_stats
        ctx.mark_non_differentiable(stats)
        return out, stats

    @staticmethod
    def backward(ctx, dy, _dstats=None):
        weight, *sources = ctx.saved_tensors
        kind, geom, transposed, src_channels, N, Ctot, has_bias = ctx.meta
        B, Hin, Win, Hout, Wout, KH, KW, stride, pad, dil = geom
        taps = KH * KW
        dy = _contig(dy)
        dtype = dy.dtype
        dev = dy.device
        dy_pitch = N
        if dtype == torch.bfloat16 and N % 8 != 0 and Ctot > TINY_MAX_C:
            # a skinny gradient (Psi-Net streams: N = 3) gets the 16-byte pixel pitch the TMA-fed dgrad / wgrad kernels need
            dy_pitch = (N + 7) // 8 * 8
            dyp = torch.empty((*dy.shape[:-1], dy_pitch), dtype=dtype, device=dev)
            call("cnb_repitch", ptr(dy), N, ptr(dyp), dy_pitch, B * Hout * Wout, N, dtype_code(dtype), stream_ptr(dy))
            dy = dyp
        need_w = ctx.needs_input_grad[0]
        need_b = has_bias and ctx.needs_input_grad[1]
        src_grads = [None] * len(sources)
        need_src = [ctx.needs_input_grad[11 + i] for i in range(len(sources))]

        if any(need_src):
            # dgrad: the adjoint gather with the per-tap transposed weights [taps][Ctot][N]; one launch per source slice
            wd = ctx.wd if ctx.wd is not None else pack_weight(weight, kind, N, Ctot, taps, dtype, for_dgrad=True)
            dgeom = (B, Hout, Wout, Hin, Win, KH, KW, stride, pad, dil)
            merged = False
            if len(sources) > 1 and all(need_src) and MERGE_SOURCE_DGRADS:
                # ONE GEMM with N = Ctot whose column ranges land in the sources' gradient tensors (tcgen05 kernel only): dY is
                # staged once per 256 output columns instead of once per source, and narrow sources share full-width tiles
                dxs = [torch.empty_like(s) if s.shape[-1] == c else torch.zeros_like(s) for s, c in zip(sources, src_channels)]
                merged = _launch_conv([dy], [N], wd, 0, wd.shape[2], Ctot * wd.shape[2], Ctot, None, None, dgeom, not transposed, dtype,
                                      out_segments=list(zip(dxs, src_channels)))
                if merged:
                    src_grads = dxs
            coff = 0
            for i, (s, c) in enumerate(zip(sources, src_channels)):
                if merged:
                    break
                if need_src[i]:
                    # a padded source gets zeros in its padding columns (the kernels only write the logical channels)
                    dx = torch.empty_like(s) if s.shape[-1] == c else torch.zeros_like(s)
                    _launch_conv([dy], [N], wd, coff, wd.shape[2], Ctot * wd.shape[2], c, None, dx, dgeom, not transposed, dtype)
                    src_grads[i] = dx
                coff += c

        dw = None
        if need_w:
            target, acc_flag = _direct_grad_target(ctx.params[0], weight.shape, full_overwrite=True)
            derived = None
            if target is not None:
                dwp = _wgrad_accumulator(ctx.params[0], taps, N, Ctot, dev)  # zero on entry: cleared by the previous unpack
            else:
                # a weight DERIVED from parameters (stacked Psi-Net filters, the Toeplitz matrix, a view): its accumulator persists
                # under the key the producer attached, and the unpack below clears it while reading -- no zero fill per step
                derived = getattr(ctx.params[0], "_cnb_acc_key", None) if DIRECT_PARAM_GRAD[0] else None
                dwp = _derived_accumulator(derived, taps, N, Ctot, dev) if derived is not None else torch.zeros(
                    (taps, N, Ctot), dtype=torch.float32, device=dev)
            coff = 0
            for s, c in zip(sources, src_channels):
                d = WgradDesc()
                d.src, d.src_c, d.src_stride = s.data_ptr(), c, s.shape[-1]
                d.k_off, d.Ctot = coff, Ctot
                d.B, d.Hin, d.Win, d.Hout, d.Wout = B, Hin, Win, Hout, Wout
                d.KH, d.KW, d.stride, d.pad, d.dil, d.transposed = KH, KW, stride, pad, dil, int(transposed)
                d.dy, d.dy_stride, d.N = dy.data_ptr(), dy_pitch, N
                d.dwp = dwp.data_ptr()
                detail = None
                if _lib.TIMER is not None:
                    detail = f"B{B} {Hin}x{Win}->{Hout}x{Wout} {c}(of {Ctot})->{N} k{KH} s{stride}{'T' if transposed else ''}"
                call(_WGRAD_ENTRY[CONV_BACKEND], C.byref(d), dtype_code(dtype), stream_ptr(dy), flops=2.0 * B * (Hin * Win if transposed else Hout * Wout) * N * c * taps,
                     tag="conv_wgrad", detail=detail)
                coff += c
            rows, cols, s_n, s_k, s_tap = _weight_strides(kind, N, Ctot, taps, for_dgrad=False)
            if target is not None:
                if not _defer_unpack(ctx.params[0], dwp, target, taps, rows, cols, s_n, s_k, s_tap, acc_flag):
                    call("cnb_unpack_wgrad", ptr(dwp), ptr(target), taps, rows, cols, s_n, s_k, s_tap, acc_flag | 2, stream_ptr(dy))
            else:
                dw = torch.empty_like(weight, dtype=torch.float32, memory_format=torch.contiguous_format)
                call("cnb_unpack_wgrad", ptr(dwp), ptr(dw), taps, rows, cols, s_n, s_k, s_tap, 2 if derived is not None else 0, stream_ptr(dy))

        db = None
        if need_b:
            target, acc_flag = _direct_grad_target(ctx.params[1], (N,))
            if target is not None:
                call("cnb_bias_grad", ptr(dy), dy_pitch, B * Hout * Wout, N, ptr(target), acc_flag, dtype_code(dtype), stream_ptr(dy))
            else:
                db = torch.empty((N,), dtype=torch.float32, device=dev)
                call("cnb_bias_grad", ptr(dy), dy_pitch, B * Hout * Wout, N, ptr(db), 0, dtype_code(dtype), stream_ptr(dy))
        return (dw, db, None, None, None, None, None, None, None, None, None, *src_grads)


def conv2d(sources: Sequence[torch.Tensor], weight, bias=None, ksize=3, stride=1, pad=1, dil=1, want_stats: bool = False):
    """nn.Conv2d over the channel concatenation of ``sources`` (each ``[B,H,W,Cs]``).  With ``want_stats`` returns
    ``(y, sums)`` where ``sums`` is the ``[2, N]`` per-channel (sum, sum of squares) of ``y`` computed in the convolution epilogue,
    or an empty tensor when the shape took a kernel without that epilogue."""
    return _Conv2dFn.apply(weight, bias, W_CONV, ksize, stride, pad, dil, False, None, want_stats, None, *sources)


FUSE_EVAL_EPILOGUE = os.environ.get("CNB_FUSE_EVAL", "1") != "0"
_EVAL_FUSE_REJECTED: set = set()


def conv2d_bn_act_eval(sources: Sequence[torch.Tensor], weight, gamma, beta, running_mean, running_var, eps: float, act: bool, ksize=3,
                       stride=1, pad=1, dil=1) -> Optional[torch.Tensor]:
    """Inference form of ``Conv2d(bias=False) -> BatchNorm2d(running statistics) -> [SiLU]`` (reference ``convolution.py:88-116``) as ONE
    launch: the tcgen05 epilogue applies ``y = act(acc * scale + shift)`` to the fp32 accumulator.  No autograd (call under
    ``torch.no_grad()``).  Returns ``None`` when the shape does not take the tcgen05 kernel (the caller runs the two-launch path)."""
    if not FUSE_EVAL_EPILOGUE or torch.is_grad_enabled() or _lib.is_emulator():
        return None
    sources = [_contig(s) for s in sources]
    x0 = sources[0]
    if x0.dtype != torch.bfloat16:
        return None
    check_device(weight, *sources)
    B, Hin, Win = x0.shape[0], x0.shape[1], x0.shape[2]
    src_channels = [s.shape[-1] for s in sources]
    Ctot, N, taps = sum(src_channels), weight.shape[0], ksize * ksize
    Hout, Wout = _conv_out_size(Hin, ksize, stride, pad, dil), _conv_out_size(Win, ksize, stride, pad, dil)
    shape_key = (B, Hin, Win, tuple(s.shape[-1] for s in sources), N, ksize, stride, pad, dil)
    if shape_key in _EVAL_FUSE_REJECTED:
        return None
    wp, _ = packed_weights(weight, W_CONV, N, Ctot, taps, x0.dtype, False)
    stats = torch.empty((6, N), dtype=torch.float32, device=x0.device)
    call("cnb_bn_finalize", None, 1, N, ptr(gamma), ptr(beta), eps, 0.0, ptr(running_mean), ptr(running_var), ptr(stats[2]),
         ptr(stats[3]), ptr(stats[4]), ptr(stats[5]), stream_ptr(x0))
    out = torch.empty((B, Hout, Wout, N), dtype=x0.dtype, device=x0.device)
    geom = (B, Hin, Win, Hout, Wout, ksize, ksize, stride, pad, dil)
    ok = _launch_conv(sources, src_channels, wp, 0, wp.shape[2], N * wp.shape[2], N, None, out, geom, False, x0.dtype,
                      epilogue=(stats[4], stats[5], act))
    if not ok:
        _EVAL_FUSE_REJECTED.add(shape_key)  # this shape takes another kernel: do not try again
        return None
    return out


def conv_transpose2d(x: torch.Tensor, weight, bias=None, ksize=3, stride=2, pad=1, dil=1) -> torch.Tensor:
    """nn.ConvTranspose2d (output_padding=0): ``[B,H,W,C] -> [B,(H-1)s-2p+d(k-1)+1, ..., N]``."""
    return _Conv2dFn.apply(weight, bias, W_CONVT, ksize, stride, pad, dil, True, None, False, None, x)


class _TapShiftAddFn(torch.autograd.Function):
    """out[p][n] = sum_tap t[p + off(tap)][tap*N + n]; backward = the gather dt[q][tap*N + n] = dout[q - off(tap)][n]."""

    @staticmethod
    def forward(ctx, t, N, ksize, pad, dil):
        check_device(t)
        t = _contig(t)
        B, H, W, pitch = t.shape
        out = torch.empty((B, H, W, N), dtype=t.dtype, device=t.device)
        call("cnb_tap_shift_add", ptr(t), ptr(out), B, H, W, N, ksize, ksize, pad, dil, pitch, N, dtype_code(t.dtype), stream_ptr(t))
        ctx.meta = (B, H, W, pitch, N, ksize, pad, dil)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, H, W, pitch, N, ksize, pad, dil = ctx.meta
        dout = _contig(dout)
        dt = torch.empty((B, H, W, pitch), dtype=dout.dtype, device=dout.device)
        call("cnb_tap_shift_gather", ptr(dout), ptr(dt), B, H, W, N, ksize, ksize, pad, dil, pitch, N, dtype_code(dout.dtype), stream_ptr(dout))
        return dt, None, None, None, None


SKINNY_MAX_COLS = 128  # taps * N of the intermediate tensor


def conv2d_skinny(x: torch.Tensor, weight: torch.Tensor, ksize: int = 3, pad: int = 1, dil: int = 1) -> torch.Tensor:
    """Unit-stride ``nn.Conv2d(C -> N, k, bias=False)`` with a handful of output channels (taps*N <= 128) as a 1x1 GEMM with
    taps*N outputs followed by a shift-and-add: the wide input is read once instead of once per tap (see csrc/k_misc.cuh).  The
    per-tap partial sums pass through the activation dtype before they are added."""
    N, Cin = weight.shape[0], weight.shape[1]
    taps = ksize * ksize
    assert taps * N <= SKINNY_MAX_COLS and N <= 16
    w2 = weight.permute(2, 3, 0, 1).reshape(taps * N, Cin)  # rows (tap, n)
    # a whole number of 32-column epilogue chunks (81 -> 96): every chunk takes the 128-bit store path of the tcgen05 epilogue
    rows = (taps * N + 31) // 32 * 32
    if rows > taps * N:
        w2 = torch.nn.functional.pad(w2, (0, 0, 0, rows - taps * N))
    tag_derived(w2, weight, "skinny")
    t = linear(x, w2, None, in_features=Cin)
    return _TapShiftAddFn.apply(t, N, ksize, pad, dil)


def linear(x: torch.Tensor, weight, bias=None, in_features: Optional[int] = None) -> torch.Tensor:
    """nn.Linear over the channel axis of a pixel-major tensor (``in_features`` < last dim: the rest is row padding)."""
    return _Conv2dFn.apply(weight, bias, W_LINEAR, 1, 1, 0, 1, False, None, False, None if in_features is None else (in_features,), x)


# ----------------------------------------------------------------------------------------------------------------
class _BatchNormActFn(torch.autograd.Function):
    """BatchNorm (batch statistics in training, running statistics in eval) + optional SiLU (+ residual)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training, momentum, eps, act, ch_div, residual, sums=None):
        check_device(x, gamma, beta, running_mean, running_var, residual)
        x = _contig(x)
        dev, dtype = x.device, x.dtype
        L = x.shape[-1]
        P = x.numel() // L
        Cn = gamma.shape[0]
        st = stream_ptr(x)
        stats = torch.empty((6, Cn), dtype=torch.float32, device=dev)  # sum, sumsq, mean, rstd, scale, shift
        count = P * ch_div  # elements per channel: ch_div columns of every row
        y = torch.empty_like(x)
        res = _contig(residual) if residual is not None else None
        if training:
            if sums is not None and sums.numel() == 2 * Cn:
                sums_ptr = ptr(_contig(sums))  # batch statistics came out of the convolution epilogue
            else:
                sums_ptr = ptr(stats[0])
                call("cnb_bn_stats", ptr(x), P, L, Cn, ch_div, sums_ptr, dtype_code(dtype), st)
            call("cnb_bn_train_fwd", ptr(x), sums_ptr, count, ptr(gamma), ptr(beta), eps, momentum, ptr(running_mean), ptr(running_var),
                 ptr(stats[2]), ptr(stats[3]), ptr(stats[4]), ptr(stats[5]), ptr(res), ptr(y), P, L, Cn, ch_div, int(act),
                 dtype_code(dtype), st)
        else:
            call("cnb_bn_finalize", None, count, Cn, ptr(gamma), ptr(beta), eps, momentum, ptr(running_mean),
                 ptr(running_var), ptr(stats[2]), ptr(stats[3]), ptr(stats[4]), ptr(stats[5]), st)
            call("cnb_bn_act_fwd", ptr(x), ptr(stats[4]), ptr(stats[5]), ptr(res), ptr(y), P, L, Cn, ch_div, int(act),
                 dtype_code(dtype), st)
        ctx.save_for_backward(x, gamma, beta, stats)
        ctx.params = (gamma, beta)
        ctx.meta = (P, L, Cn, ch_div, int(act), bool(training), count, residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, stats = ctx.saved_tensors
        P, L, Cn, ch_div, act, training, count, has_res = ctx.meta
        dy = _contig(dy)
        dtype = x.dtype
        st = stream_ptr(x)
        dsums = _stats_slice(Cn, x.device)  # zero on entry (pre-cleared arena, or a fresh torch.zeros): no memset node per layer
        call("cnb_bn_act_bwd_reduce_acc", ptr(x), ptr(dy), ptr(stats[2]), ptr(stats[3]), ptr(gamma), ptr(beta), P, L, Cn, ch_div, act,
             ptr(dsums), dtype_code(dtype), st)
        dx = torch.empty_like(x)
        call("cnb_bn_act_bwd_apply", ptr(x), ptr(dy), ptr(stats[2]), ptr(stats[3]), ptr(gamma), ptr(beta), ptr(dsums), count, ptr(dx),
             P, L, Cn, ch_div, act, int(training), dtype_code(dtype), st)
        dres = dy if has_res else None
        dgamma = _small_grad(ctx.params[0], dsums[1]) if ctx.needs_input_grad[1] else None
        dbeta = _small_grad(ctx.params[1], dsums[0]) if ctx.needs_input_grad[2] else None
        return dx, dgamma, dbeta, None, None, None, None, None, None, None, dres, None


def batchnorm_act(x, gamma, beta, running_mean, running_var, training: bool, momentum: float = 0.1, eps: float = 1e-5,
                  act: bool = True, ch_div: int = 1, residual: Optional[torch.Tensor] = None,
                  sums: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _BatchNormActFn.apply(x, gamma, beta, running_mean, running_var, training, momentum, eps, act, ch_div, residual, sums)


class _AddNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *xs):
        check_device(*xs)
        xs = [_contig(t) for t in xs]
        out = torch.empty_like(xs[0])
        p = [ptr(t) for t in xs] + [ptr(None)] * (4 - len(xs))
        call("cnb_add_n", p[0], p[1], p[2], p[3], ptr(out), out.numel(), dtype_code(out.dtype), stream_ptr(out))
        ctx.n = len(xs)
        return out

    @staticmethod
    def backward(ctx, dy):
        return (dy,) * ctx.n


def add_n(*xs: torch.Tensor) -> torch.Tensor:
    xs = [t for t in xs if t is not None]
    assert 2 <= len(xs) <= 4
    return _AddNFn.apply(*xs)


class _FanoutFn(torch.autograd.Function):
    """n aliases of one tensor whose gradients are summed by ONE n-ary add kernel.  Without it autograd accumulates the gradients of
    a tensor with n consumers pairwise with torch's own add kernels: n-1 launches and 3(n-1) passes over the tensor instead of n+1."""

    @staticmethod
    def forward(ctx, x, n):
        return tuple(x.view_as(x) for _ in range(n))

    @staticmethod
    def backward(ctx, *grads):
        gs = [_contig(g) for g in grads if g is not None]
        if not gs:
            return None, None
        while len(gs) > 1:
            head, gs = gs[:4], gs[4:]
            out = torch.empty_like(head[0])
            p = [ptr(t) for t in head] + [ptr(None)] * (4 - len(head))
            call("cnb_add_n", p[0], p[1], p[2], p[3], ptr(out), out.numel(), dtype_code(out.dtype), stream_ptr(out))
            gs.insert(0, out)
        return gs[0], None


def fanout(x: torch.Tensor, n: int):
    """``n`` references to ``x`` for ``n`` different consumers (a no-op outside autograd)."""
    if n <= 1 or not (torch.is_grad_enabled() and x.requires_grad):
        return (x,) * max(n, 1)
    check_device(x)
    return _FanoutFn.apply(x, n)


# ----------------------------------------------------------------------------------------------------------------
class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        check_device(x, gamma, beta)
        x = _contig(x)
        Cn = x.shape[-1]
        P = x.numel() // Cn
        y = torch.empty_like(x)
        ms = torch.empty((2, P), dtype=torch.float32, device=x.device)
        call("cnb_layernorm_fwd", ptr(x), ptr(gamma), ptr(beta), eps, ptr(y), ptr(ms[0]), ptr(ms[1]), P, Cn, dtype_code(x.dtype),
             stream_ptr(x))
        ctx.save_for_backward(x, gamma, ms)
        ctx.params = (gamma, beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, ms = ctx.saved_tensors
        dy = _contig(dy)
        Cn = x.shape[-1]
        P = x.numel() // Cn
        dx = torch.empty_like(x)
        # the kernel ACCUMULATES dgamma / dbeta atomically: when the parameters' gradient buffers may take them directly (zeroed at the
        # top of the step) they are the target -- no zero-filled temporary, no accumulation kernel afterwards
        tg, _ = _direct_grad_target(ctx.params[0], (Cn,))
        tb, _ = _direct_grad_target(ctx.params[1], (Cn,)) if tg is not None else (None, 0)
        if tg is not None and tb is not None:
            call("cnb_layernorm_bwd", ptr(x), ptr(dy), ptr(gamma), ptr(ms[0]), ptr(ms[1]), ptr(dx), ptr(tg), ptr(tb), P, Cn,
                 dtype_code(x.dtype), stream_ptr(x))
            return dx, None, None, None
        dgb = torch.zeros((2, Cn), dtype=torch.float32, device=x.device)
        call("cnb_layernorm_bwd", ptr(x), ptr(dy), ptr(gamma), ptr(ms[0]), ptr(ms[1]), ptr(dx), ptr(dgb[0]), ptr(dgb[1]), P, Cn,
             dtype_code(x.dtype), stream_ptr(x))
        if tg is not None:  # gamma's buffer was claimed above but beta's was not available: hand gamma's gradient over by the table
            _PENDING_SMALL.append((dgb[0], tg, 1))
            return dx, None, dgb[1], None
        return dx, dgb[0], dgb[1], None


def layernorm(x, gamma, beta, eps: float = 1e-5) -> torch.Tensor:
    return _LayerNormFn.apply(x, gamma, beta, eps)


# ----------------------------------------------------------------------------------------------------------------
class _NA2DFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, heads, ksize, dilation, scale):
        check_device(qkv)
        qkv = _contig(qkv)
        B, H, W, C3 = qkv.shape
        Cn = C3 // 3
        hd = Cn // heads
        out = torch.empty((B, H, W, Cn), dtype=qkv.dtype, device=qkv.device)
        # the eligibility query returns 1/0 directly (every other entry point returns an error code)
        tiled = bool(_lib.lib().cnb_na2d_tiled_eligible(B, H, W, heads, hd, ksize, dilation, dtype_code(qkv.dtype)))
        lse = torch.empty((B, H, W, heads), dtype=torch.float32, device=qkv.device) if tiled else None
        call("cnb_na2d_fwd", ptr(qkv), ptr(out), ptr(lse), B, H, W, heads, hd, ksize, dilation, scale, dtype_code(qkv.dtype), stream_ptr(qkv))
        if tiled:
            ctx.save_for_backward(qkv, out, lse)
        else:
            ctx.save_for_backward(qkv)
        ctx.meta = (heads, hd, ksize, dilation, scale, tiled)
        return out

    @staticmethod
    def backward(ctx, dout):
        heads, hd, ksize, dilation, scale, tiled = ctx.meta
        dout = _contig(dout)
        if tiled:
            qkv, out, lse = ctx.saved_tensors
            acc = None
            dvec = torch.empty_like(lse)
        else:
            (qkv,) = ctx.saved_tensors
            out = lse = dvec = None
            acc = torch.empty(qkv.shape, dtype=torch.float32, device=qkv.device)
        B, H, W, _ = qkv.shape
        dqkv = torch.empty_like(qkv)
        # specialised shapes keep (p, scaled dlogit) per (pixel, head, neighbour) between the query-side and key-side passes
        nws = int(_lib.lib().cnb_na2d_bwd_workspace_floats(B, H, W, heads, hd, ksize, dilation, dtype_code(qkv.dtype))) if tiled else 0
        pds = torch.empty((nws,), dtype=torch.float32, device=qkv.device) if nws > 0 else None
        call("cnb_na2d_bwd", ptr(qkv), ptr(dout), ptr(out), ptr(lse), ptr(dvec), ptr(acc), ptr(pds), ptr(dqkv), B, H, W, heads, hd, ksize,
             dilation, scale, dtype_code(qkv.dtype), stream_ptr(qkv))
        return dqkv, None, None, None, None


class _NA2DDropoutFn(torch.autograd.Function):
    """Neighbourhood attention with dropout on the attention probabilities (natten ``attn_drop`` in training mode)."""

    @staticmethod
    def forward(ctx, qkv, heads, ksize, dilation, scale, p, site):
        check_device(qkv)
        qkv = _contig(qkv)
        B, H, W, C3 = qkv.shape
        Cn = C3 // 3
        st = rng_state(qkv.device)
        out = torch.empty((B, H, W, Cn), dtype=qkv.dtype, device=qkv.device)
        call("cnb_na2d_dropout_fwd", ptr(qkv), ptr(out), B, H, W, heads, Cn // heads, ksize, dilation, scale, ptr(st), site, p,
             dtype_code(qkv.dtype), stream_ptr(qkv))
        ctx.save_for_backward(qkv)
        ctx.meta = (heads, ksize, dilation, scale, p, site, st)
        return out

    @staticmethod
    def backward(ctx, dout):
        (qkv,) = ctx.saved_tensors
        heads, ksize, dilation, scale, p, site, st = ctx.meta
        dout = _contig(dout)
        B, H, W, C3 = qkv.shape
        acc = torch.empty(qkv.shape, dtype=torch.float32, device=qkv.device)
        dqkv = torch.empty_like(qkv)
        call("cnb_na2d_dropout_bwd", ptr(qkv), ptr(dout), ptr(acc), ptr(dqkv), B, H, W, heads, C3 // 3 // heads, ksize, dilation, scale,
             ptr(st), site, p, dtype_code(qkv.dtype), stream_ptr(qkv))
        return dqkv, None, None, None, None, None, None


def na2d(qkv: torch.Tensor, heads: int, ksize: int, dilation: int, scale: float, attn_drop: float = 0.0, site: int = 0) -> torch.Tensor:
    """Neighbourhood attention core over a packed ``[B,H,W,3*heads*hd]`` qkv tensor (``attn_drop`` > 0: training-mode dropout on the
    attention probabilities, drawn from the device generator state at call site ``site``)."""
    if attn_drop > 0:
        return _NA2DDropoutFn.apply(qkv, heads, ksize, dilation, scale, float(attn_drop), int(site))
    return _NA2DFn.apply(qkv, heads, ksize, dilation, scale)


# ----------------------------------------------------------------------------------------------------------------
class _ResizeBilinearFn(torch.autograd.Function):
    """``bias`` (optional): the bias Parameter of the convolution that produced ``x``.  Its gradient is the per-channel sum of THIS
    function's input gradient, which the backward kernel accumulates while it writes that gradient (``cnb_resize_bilinear_bwd_colsum``);
    the convolution is then called with the bias detached, so it launches no column-sum pass of its own."""

    @staticmethod
    def forward(ctx, x, Hout, Wout, bias=None):
        check_device(x)
        x = _contig(x)
        B, Hin, Win, Cn = x.shape
        y = torch.empty((B, Hout, Wout, Cn), dtype=x.dtype, device=x.device)
        call("cnb_resize_bilinear_fwd", ptr(x), ptr(y), B, Hin, Win, Hout, Wout, Cn, dtype_code(x.dtype), stream_ptr(x))
        ctx.meta = (B, Hin, Win, Hout, Wout, Cn)
        ctx.bias = bias
        return y

    @staticmethod
    def backward(ctx, dy):
        B, Hin, Win, Hout, Wout, Cn = ctx.meta
        dy = _contig(dy)
        dx = torch.empty((B, Hin, Win, Cn), dtype=dy.dtype, device=dy.device)
        db = None
        if ctx.bias is not None and ctx.needs_input_grad[3]:
            target, acc_flag = _direct_grad_target(ctx.bias, (Cn,))
            if target is None:
                db = torch.empty((Cn,), dtype=torch.float32, device=dy.device)
                target, acc_flag = db, 0
            call("cnb_resize_bilinear_bwd_colsum", ptr(dy), ptr(dx), B, Hin, Win, Hout, Wout, Cn, ptr(target), acc_flag, dtype_code(dy.dtype),
                 stream_ptr(dy))
        else:
            call("cnb_resize_bilinear_bwd", ptr(dy), ptr(dx), B, Hin, Win, Hout, Wout, Cn, dtype_code(dy.dtype), stream_ptr(dy))
        return dx, None, None, db


def resize_bilinear(x: torch.Tensor, size, producer_bias=None) -> torch.Tensor:
    """``check_upsample``: bilinear, align_corners=True, only when the spatial size differs.  ``producer_bias``: see
    ``_ResizeBilinearFn`` (the caller must have given the producing convolution ``producer_bias.detach()``)."""
    if tuple(x.shape[1:3]) == tuple(size):
        assert producer_bias is None, "no resize: the producing convolution has to compute its own bias gradient"
        return x
    if producer_bias is not None and producer_bias.requires_grad and torch.is_grad_enabled():
        return _ResizeBilinearFn.apply(x, int(size[0]), int(size[1]), producer_bias)
    return _ResizeBilinearFn.apply(x, int(size[0]), int(size[1]), None)


class _BroadcastPixelsFn(torch.autograd.Function):
    """``[B,1,1,C] -> [B,H,W,C]`` (the per-sample GeoEmbeddings vector over a level's pixels, reference ``unet_parts.py:742-750``);
    backward = the per-sample column sum of the gradient."""

    @staticmethod
    def forward(ctx, e, H, W):
        check_device(e)
        e = _contig(e)
        B, Cn = e.shape[0], e.shape[-1]
        out = torch.empty((B, H, W, Cn), dtype=e.dtype, device=e.device)
        call("cnb_broadcast_pixels", ptr(e), ptr(out), B, H * W, Cn, dtype_code(e.dtype), stream_ptr(e))
        ctx.meta = (B, H, W, Cn)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, H, W, Cn = ctx.meta
        dout = _contig(dout)
        de32 = torch.empty((B, Cn), dtype=torch.float32, device=dout.device)
        for b in range(B):  # one column sum per sample (an option that is off by default: B small launches)
            call("cnb_bias_grad", ptr(dout[b]), Cn, H * W, Cn, ptr(de32[b]), 0, dtype_code(dout.dtype), stream_ptr(dout))
        return de32.view(B, 1, 1, Cn).to(dout.dtype), None, None  # [B, C] values: dtype plumbing, not compute


def broadcast_pixels(e: torch.Tensor, size) -> torch.Tensor:
    return _BroadcastPixelsFn.apply(e, int(size[0]), int(size[1]))


# ----------------------------------------------------------------------------------------------------------------
class _PreTimeConvFn(torch.autograd.Function):
    """Conv3d(C->C, (k,1,1), bias=False) over x[B,C,T,H,W] -> pixel-major u[B,H,W,pitch] (column = c*T' + t'; pitch = C*T'
    rounded up to 8, padding columns zero)."""

    @staticmethod
    def forward(ctx, x, w1, dtype):
        check_device(x, w1)
        if x.dtype != torch.float32:
            raise _lib.CnbError("PreTimeReduction takes the float32 [B,C,T,H,W] input of the reference")
        x = _contig(x)
        w1c = _contig(w1)
        B, Cn, Tn, H, W = x.shape
        k = w1.shape[2]
        Tp = Tn - k + 1
        pitch = (Cn * Tp + 7) // 8 * 8  # 16-byte pixel rows in either dtype; the padding columns are written as zero
        u = torch.empty((B, H, W, pitch), dtype=dtype, device=x.device)
        call("cnb_pretime_conv_fwd", ptr(x), ptr(w1c), ptr(u), B, Cn, Tn, H, W, k, pitch, dtype_code(dtype), stream_ptr(x))
        ctx.save_for_backward(x, w1)
        return u

    @staticmethod
    def backward(ctx, du):
        x, w1 = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("cultionet_b200: the network input x does not receive a gradient")
        du = _contig(du)
        B, Cn, Tn, H, W = x.shape
        k = w1.shape[2]
        dw = torch.zeros(w1.shape, dtype=torch.float32, device=x.device)
        call("cnb_pretime_conv_wgrad", ptr(x), ptr(du), ptr(dw), B, Cn, Tn, H, W, k, du.shape[-1], dtype_code(du.dtype), stream_ptr(x))
        return None, dw, None


def pretime_conv(x: torch.Tensor, w1: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    return _PreTimeConvFn.apply(x, w1, dtype)


def time_to_pixel_major(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """x[B,C,T,H,W] float32 -> pixel-major xp[B,H,W,pitch] (column = c*T + t; pitch = C*T rounded up to 8, padding zero).  The network
    input carries no gradient, so this is a plain function."""
    check_device(x)
    if x.dtype != torch.float32:
        raise _lib.CnbError("PreTimeReduction takes the float32 [B,C,T,H,W] input of the reference")
    if x.requires_grad:
        raise NotImplementedError("cultionet_b200: the network input x does not receive a gradient")
    x = _contig(x)
    B, Cn, Tn, H, W = x.shape
    pitch = (Cn * Tn + 7) // 8 * 8
    xp = torch.empty((B, H, W, pitch), dtype=dtype, device=x.device)
    call("cnb_time_to_pixel_major", ptr(x), ptr(xp), B, Cn * Tn, H * W, pitch, dtype_code(dtype), stream_ptr(x))
    return xp


class _ToeplitzFn(torch.autograd.Function):
    """w1[C,C,k,1,1] -> Wt[rows, C*T] with Wt[c2*T'+t'][c*T+t] = w1[c2,c,t-t'] (rows = C*T' rounded up to 8, padding rows zero)."""

    @staticmethod
    def forward(ctx, w1, Tn):
        check_device(w1)
        Cn, k = w1.shape[0], w1.shape[2]
        rows = (Cn * (Tn - k + 1) + 7) // 8 * 8
        wt = torch.empty((rows, Cn * Tn), dtype=torch.float32, device=w1.device)
        call("cnb_toeplitz_expand", ptr(_contig(w1)), ptr(wt), Cn, Tn, k, rows, stream_ptr(w1))
        ctx.meta = (Cn, Tn, k, tuple(w1.shape))
        return wt

    @staticmethod
    def backward(ctx, dwt):
        Cn, Tn, k, shape = ctx.meta
        dwt = _contig(dwt.float())
        dw = torch.empty(shape, dtype=torch.float32, device=dwt.device)
        call("cnb_toeplitz_fold", ptr(dwt), ptr(dw), Cn, Tn, k, stream_ptr(dwt))
        return dw, None


def pretime_conv_gemm(xp: torch.Tensor, w1: torch.Tensor, in_time: int) -> torch.Tensor:
    """The temporal convolution of ``pretime_conv`` over the pixel-major copy ``xp`` of x, as a 1x1 GEMM on the tensor cores:
    returns u[B,H,W,rows] with the same column order and zero row padding as ``pretime_conv``."""
    wt = tag_derived(_ToeplitzFn.apply(w1, in_time), w1, "toeplitz")
    return linear(xp, wt, None, in_features=w1.shape[0] * in_time)


# ----------------------------------------------------------------------------------------------------------------
class _FinalCombineFn(torch.autograd.Function):
    """TowerUNetFinalCombine + SigmoidCrisp over the three towers' fused [B,H,W,3] streams; ``prm`` = the 16 scalar parameters
    stacked by ``stack_params`` (its backward scatters their gradients)."""

    @staticmethod
    def forward(ctx, ha, hb, hc, smooth, flags, prm):
        check_device(ha, hb, hc, prm)
        ha, hb, hc = _contig(ha), _contig(hb), _contig(hc)
        B, H, W, _ = ha.shape
        P = B * H * W
        prm = _contig(prm.float())
        assert prm.numel() == 16
        out = [torch.empty((B, 1, H, W), dtype=torch.float32, device=ha.device) for _ in range(3)]
        call("cnb_final_combine_fwd", ptr(ha), ptr(hb), ptr(hc), ptr(prm), smooth, flags, ptr(out[0]), ptr(out[1]), ptr(out[2]), P,
             dtype_code(ha.dtype), stream_ptr(ha))
        ctx.save_for_backward(ha, hb, hc, prm)
        ctx.meta = (smooth, flags, P)
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, d_dist, d_edge, d_crop):
        ha, hb, hc, prm = ctx.saved_tensors
        smooth, flags, P = ctx.meta
        dev = ha.device

        def g(t):
            return _zeros_like_cached(P, dev) if t is None else _contig(t.float())

        d_dist, d_edge, d_crop = g(d_dist), g(d_edge), g(d_crop)
        dha, dhb, dhc = torch.empty_like(ha), torch.empty_like(hb), torch.empty_like(hc)
        dprm = torch.empty((16,), dtype=torch.float32, device=dev)
        ws = torch.empty((32,), dtype=torch.float32, device=dev)
        call("cnb_final_combine_bwd", ptr(ha), ptr(hb), ptr(hc), ptr(prm), smooth, flags, ptr(d_dist), ptr(d_edge), ptr(d_crop),
             ptr(dha), ptr(dhb), ptr(dhc), ptr(dprm), ptr(ws), P, dtype_code(ha.dtype), stream_ptr(ha))
        return dha, dhb, dhc, None, None, dprm


def final_combine(ha, hb, hc, params: Sequence[torch.Tensor], smooth: float = 1e-2, edge_activation: bool = True,
                  mask_activation: bool = True):
    """params: 9 gammas (dist 1-3, edge 1-3, crop 1-3), 3 conv weights, 3 conv biases, crisp gamma."""
    flags = (1 if edge_activation else 0) | (2 if mask_activation else 0)
    prm = stack_params(list(params), list(range(16)), 16)  # one launch instead of a torch.cat of 16 scalars (+ 16 gradient adds back)
    return _FinalCombineFn.apply(ha, hb, hc, smooth, flags, prm)


# ----------------------------------------------------------------------------------------------------------------
TARGET_FLOAT, TARGET_ONEHOT, TARGET_EDGE, TARGET_CROP = 0, 1, 2, 3
MASK_NONE, MASK_FLOAT, MASK_INT64, MASK_FROM_LABELS = 0, 1, 2, 3


class TanimotoTermSpec:
    """One (prediction, target) pair of the Tanimoto-complement loss; see ``cnb_tanimoto_term``."""

    def __init__(self, target, target_mode, mask=None, mask_mode=MASK_NONE, edge_class=2, weight=1.0):
        self.target, self.target_mode = target, target_mode
        self.mask, self.mask_mode = mask, mask_mode
        self.edge_class, self.weight = edge_class, weight


class _TanimotoFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, smooth, depth, variant, specs, *preds):
        nterms = len(preds)
        assert 1 <= nterms <= _lib.TN_MAX_TERMS and len(specs) == nterms
        preds = [_contig(p.float()) for p in preds]
        B = preds[0].shape[0]
        HW = preds[0].shape[-2] * preds[0].shape[-1]
        dev = preds[0].device
        keep = []
        terms = (TanimotoTerm * nterms)()
        for i, (p, s) in enumerate(zip(preds, specs)):
            assert p.dim() == 4 and p.shape[0] == B and p.shape[-2] * p.shape[-1] == HW
            tgt = s.target
            if s.target_mode == TARGET_FLOAT:
                tgt = _contig(tgt.float())
                tgt_c = 1 if tgt.dim() == 3 else tgt.shape[1]
            else:
                tgt = _contig(tgt.long())
                tgt_c = 1
            msk = s.mask
            if s.mask_mode == MASK_FLOAT:
                msk = _contig(msk.float())
            elif s.mask_mode in (MASK_INT64, MASK_FROM_LABELS):
                msk = _contig(msk.long())
            check_device(p, tgt, msk)
            keep += [tgt, msk]
            terms[i].pred = p.data_ptr()
            terms[i].target = tgt.data_ptr()
            terms[i].mask = msk.data_ptr() if msk is not None else None
            terms[i].dpred = None
            terms[i].C, terms[i].tgt_c = p.shape[1], tgt_c
            terms[i].target_mode, terms[i].mask_mode, terms[i].edge_class = s.target_mode, s.mask_mode, s.edge_class
            terms[i].weight = s.weight
        sums = torch.empty((nterms, B, 4), dtype=torch.float64, device=dev)
        coef = torch.empty((nterms, B, 4), dtype=torch.float32, device=dev)
        loss = torch.empty((1 + nterms,), dtype=torch.float32, device=dev)
        call("cnb_tanimoto_fwd", terms, nterms, B, HW, smooth, depth, int(variant), ptr(sums), ptr(coef), ptr(loss), stream_ptr(preds[0]))
        ctx.terms, ctx.keep, ctx.preds, ctx.coef = terms, keep, preds, coef
        ctx.meta = (nterms, B, HW)
        ctx.mark_non_differentiable(loss)
        total = loss[0].clone()
        return total, loss

    @staticmethod
    def backward(ctx, gtotal, _gparts):
        nterms, B, HW = ctx.meta
        terms = ctx.terms
        grads = []
        for i, p in enumerate(ctx.preds):
            g = torch.empty_like(p)
            terms[i].dpred = g.data_ptr()
            grads.append(g)
        gs = _contig(gtotal.float().reshape(1))
        call("cnb_tanimoto_bwd", terms, nterms, B, HW, ptr(ctx.coef), ptr(gs), stream_ptr(gs))
        return (None, None, None, None, *grads)


TANIMOTO_COMPLEMENT, TANIMOTO_DIST, TANIMOTO_COMBINED = 0, 1, 2  # `variant` of cnb_tanimoto_fwd


def tanimoto_complement(preds: Sequence[torch.Tensor], specs: Sequence[TanimotoTermSpec], smooth: float = 1e-5, depth: int = 5,
                        variant: int = TANIMOTO_COMPLEMENT):
    """Returns (sum_t weight_t * loss_t, tensor[1 + nterms] of total and per-term losses).  ``variant`` picks the reference's
    TanimotoComplementLoss (0), TanimotoDistLoss (1) or the CombinedLoss of the two (2)."""
    return _TanimotoFn.apply(smooth, depth, variant, list(specs), *preds)


def validation_counts(dist: torch.Tensor, edge: torch.Tensor, crop: torch.Tensor, y: torch.Tensor, bdist: torch.Tensor,
                      edge_class: int = 2, thresh: float = 0.5) -> torch.Tensor:
    """fp64[12] = {valid pixels, sum |dist - bdist|, sum (dist - bdist)^2, edge tp/fp/fn/tn, crop tp/fp/fn/tn, 0} over the labelled
    (``y != -1``) pixels of a batch in one pass (``cnb_val_counts``): the inputs of every scorer of ``_shared_eval_step``."""
    dist, edge, crop = (_contig(t.detach().float()) for t in (dist, edge, crop))
    y, bdist = _contig(y.long()), _contig(bdist.float())
    check_device(dist, edge, crop, y, bdist)
    n = y.numel()
    assert dist.numel() == n and edge.numel() == n and crop.numel() == n and bdist.numel() == n, "validation_counts: 1-channel predictions"
    out = torch.empty((12,), dtype=torch.float64, device=y.device)
    call("cnb_val_counts", ptr(dist), ptr(edge), ptr(crop), ptr(y), ptr(bdist), n, int(edge_class), float(thresh), ptr(out), stream_ptr(y))
    return out


# ----------------------------------------------------------------------------------------------------------------
# optional ResUNet-a block variants: adaptive max pooling, spatial-channel attention, dropout
# ----------------------------------------------------------------------------------------------------------------
class _AdaptiveMaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Hout, Wout):
        check_device(x)
        x = _contig(x)
        B, Hin, Win, Cn = x.shape
        y = torch.empty((B, Hout, Wout, Cn), dtype=x.dtype, device=x.device)
        idx = torch.empty((B, Hout, Wout, Cn), dtype=torch.uint8, device=x.device)
        call("cnb_adaptive_maxpool_fwd", ptr(x), ptr(y), ptr(idx), B, Hin, Win, Hout, Wout, Cn, dtype_code(x.dtype), stream_ptr(x))
        ctx.save_for_backward(idx)
        ctx.meta = (B, Hin, Win, Hout, Wout, Cn)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        B, Hin, Win, Hout, Wout, Cn = ctx.meta
        dy = _contig(dy)
        dx = torch.empty((B, Hin, Win, Cn), dtype=dy.dtype, device=dy.device)
        call("cnb_adaptive_maxpool_bwd", ptr(dy), ptr(idx), ptr(dx), B, Hin, Win, Hout, Wout, Cn, dtype_code(dy.dtype), stream_ptr(dy))
        return dx, None, None


def adaptive_max_pool2d(x: torch.Tensor, size) -> torch.Tensor:
    """``F.adaptive_max_pool2d`` over a pixel-major ``[B,H,W,C]`` tensor."""
    return _AdaptiveMaxPoolFn.apply(x, int(size[0]), int(size[1]))


# Activation codes of the C ABI (include/cultionet_b200.h CNB_ACT_*).  The reference builds ``getattr(torch.nn, activation_type)()``
# (nn/modules/activations.py:5-24) with default arguments: these are the torch.nn classes whose default form the kernels implement.
ACT_CODES = {"Identity": 0, "SiLU": 1, "ReLU": 2, "LeakyReLU": 3, "GELU": 4, "Mish": 5, "ELU": 6, "Tanh": 7, "Sigmoid": 8, "Hardswish": 9}


def act_code(activation_type) -> int:
    """``activation_type`` (a torch.nn class name, as the reference's ``SetActivation`` takes it) -> CNB_ACT_* code."""
    name = str(getattr(activation_type, "value", activation_type))
    if name not in ACT_CODES:
        raise NotImplementedError(f"cultionet_b200: activation_type={name!r} is not built (available: {', '.join(ACT_CODES)})")
    return ACT_CODES[name]


class _ActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        check_device(x)
        x = _contig(x)
        y = torch.empty_like(x)
        call("cnb_act_fwd", ptr(x), ptr(y), x.numel(), int(act), dtype_code(x.dtype), stream_ptr(x))
        ctx.save_for_backward(x)
        ctx.act = int(act)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = _contig(dy)
        dx = torch.empty_like(x)
        call("cnb_act_bwd", ptr(x), ptr(dy), ptr(dx), x.numel(), ctx.act, dtype_code(x.dtype), stream_ptr(x))
        return dx, None


def activation(x: torch.Tensor, act: int = 1) -> torch.Tensor:
    """Stand-alone activation by CNB_ACT_* code (``act_code(name)``)."""
    return x if int(act) == 0 else _ActFn.apply(x, int(act))


def silu(x: torch.Tensor) -> torch.Tensor:
    return _ActFn.apply(x, 1)


class _ScaPoolFn(torch.autograd.Function):
    """x[B,H,W,C] -> sp[B,H,W,2] (per-pixel mean, max over channels), ch_avg[B,1,1,C], ch_max[B,1,1,C] (per-channel mean / max over
    pixels); all three fp32."""

    @staticmethod
    def forward(ctx, x):
        check_device(x)
        x = _contig(x)
        B, H, W, Cn = x.shape
        HW = H * W
        dev = x.device
        sp = torch.empty((B, H, W, 2), dtype=torch.float32, device=dev)
        ties = torch.empty((B, HW), dtype=torch.float32, device=dev)
        ch = torch.empty((2, B, 1, 1, Cn), dtype=torch.float32, device=dev)
        arg = torch.empty((B, Cn), dtype=torch.int32, device=dev)
        S = int(_lib.lib().cnb_sca_slices(B, HW, Cn, dtype_code(x.dtype)))
        ws = torch.empty((3, B, S, Cn), dtype=torch.float32, device=dev)  # the third plane holds int32 indices
        call("cnb_sca_pool_fwd", ptr(x), ptr(sp), ptr(ties), ptr(ch[0]), ptr(ch[1]), ptr(arg), ptr(ws[0]), ptr(ws[1]), ptr(ws[2]), B, HW, Cn,
             dtype_code(x.dtype), stream_ptr(x))
        ctx.save_for_backward(x, sp, ties, arg)
        return sp, ch[0], ch[1]

    @staticmethod
    def backward(ctx, dsp, davg, dmax):
        x, sp, ties, arg = ctx.saved_tensors
        B, H, W, Cn = x.shape
        dev = x.device

        def g(t, shape):
            return _contig(t.float()) if t is not None else torch.zeros(shape, dtype=torch.float32, device=dev)

        dsp, davg, dmax = g(dsp, sp.shape), g(davg, (B, Cn)), g(dmax, (B, Cn))
        dx = torch.empty_like(x)
        call("cnb_sca_pool_bwd", ptr(x), ptr(sp), ptr(ties), ptr(dsp), ptr(davg), ptr(dmax), ptr(arg), ptr(dx), B, H * W, Cn,
             dtype_code(x.dtype), stream_ptr(x))
        return dx


def sca_pool(x: torch.Tensor):
    return _ScaPoolFn.apply(x)


class _ScaApplyFn(torch.autograd.Function):
    """out = y * (1 + gamma * 0.5 * (sigmoid(cl[b,c]) + sigmoid(sl[b,h,w])))."""

    @staticmethod
    def forward(ctx, y, cl, sl, gamma):
        check_device(y, cl, sl, gamma)
        y, cl, sl = _contig(y), _contig(cl.float()), _contig(sl.float())
        gamma_c = _contig(gamma.float())
        B, H, W, Cn = y.shape
        assert cl.numel() == B * Cn and sl.numel() == B * H * W and gamma_c.numel() == 1
        out = torch.empty_like(y)
        call("cnb_sca_apply_fwd", ptr(y), ptr(cl), ptr(sl), ptr(gamma_c), ptr(out), B, H * W, Cn, dtype_code(y.dtype), stream_ptr(y))
        ctx.save_for_backward(y, cl, sl, gamma_c)
        ctx.shapes = (cl.shape, sl.shape, gamma.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        y, cl, sl, gamma_c = ctx.saved_tensors
        B, H, W, Cn = y.shape
        dout = _contig(dout)
        dy = torch.empty_like(y)
        dcl = torch.empty(ctx.shapes[0], dtype=torch.float32, device=y.device)
        dsl = torch.empty(ctx.shapes[1], dtype=torch.float32, device=y.device)
        dgamma = torch.empty(ctx.shapes[2], dtype=torch.float32, device=y.device)
        call("cnb_sca_apply_bwd", ptr(y), ptr(dout), ptr(cl), ptr(sl), ptr(gamma_c), ptr(dy), ptr(dcl), ptr(dsl), ptr(dgamma), B, H * W, Cn,
             dtype_code(y.dtype), stream_ptr(y))
        return dy, dcl, dsl, dgamma


def sca_apply(y: torch.Tensor, cl: torch.Tensor, sl: torch.Tensor, gamma: torch.Tensor) -> torch.Tensor:
    return _ScaApplyFn.apply(y, cl, sl, gamma)


# dropout: one device-resident generator state {seed, step counter} per device; the seed comes from torch's default CPU generator, so
# torch.manual_seed() makes a run repeatable.  rng_advance() is called once per training forward (TowerUNet.forward).
_RNG_STATE: dict = {}
_RNG_SITES = [0]


def new_rng_site() -> int:
    """A unique id per dropout module instance (separates the random streams of the call sites of one step)."""
    _RNG_SITES[0] += 1
    return _RNG_SITES[0]


def rng_state(device: torch.device) -> torch.Tensor:
    key = (device.type, device.index)
    st = _RNG_STATE.get(key)
    if st is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        st = torch.tensor([seed, 0], dtype=torch.int64).to(device)
        _RNG_STATE[key] = st
    return st


def rng_advance(device: torch.device) -> None:
    st = rng_state(device)
    call("cnb_rng_advance", ptr(st), stream_ptr(st))


class _DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, site, channelwise):
        check_device(x)
        x = _contig(x)
        st = rng_state(x.device)
        out = torch.empty_like(x)
        ctx.meta = (p, site, channelwise, st)
        _DropoutFn._launch(x, out, ctx.meta)
        return out

    @staticmethod
    def _launch(x, out, meta):
        p, site, channelwise, st = meta
        if channelwise:
            B, Cn = x.shape[0], x.shape[-1]
            call("cnb_dropout2d", ptr(x), ptr(out), B, x.numel() // (B * Cn), Cn, ptr(st), site, p, dtype_code(x.dtype), stream_ptr(x))
        else:
            call("cnb_dropout", ptr(x), ptr(out), x.numel(), ptr(st), site, p, dtype_code(x.dtype), stream_ptr(x))

    @staticmethod
    def backward(ctx, dy):
        dy = _contig(dy)
        dx = torch.empty_like(dy)
        _DropoutFn._launch(dy, dx, ctx.meta)  # same state, same site => the same mask
        return dx, None, None, None


def dropout(x: torch.Tensor, p: float, site: int) -> torch.Tensor:
    """nn.Dropout in training mode (elementwise)."""
    return x if p <= 0 else _DropoutFn.apply(x, float(p), int(site), False)


def dropout2d(x: torch.Tensor, p: float, site: int) -> torch.Tensor:
    """nn.Dropout2d in training mode over a pixel-major tensor: whole channels of a sample are zeroed."""
    return x if p <= 0 else _DropoutFn.apply(x, float(p), int(site), True)
