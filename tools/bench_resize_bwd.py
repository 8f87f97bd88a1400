"""CUDA-event timing of the ConvTranspose2d fix-up's backward (127 x 127 <- 128 x 128, 256 channels, batch 32) with and without the fused
bias-gradient column sum, against the separate column-sum pass.  `CNB_RESIZE_BWD=persistent` selects the row-looping kernel."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from cultionet_b200 import _lib
from cultionet_b200._lib import call, dtype_code, ptr, stream_ptr

dev, dt = "cuda", torch.bfloat16
B, Hi, Wi, Ho, Wo, C = 32, 127, 127, 128, 128, 256
if len(sys.argv) > 1:  # output size of the fix-up: 128 (level a), 64 (level b), 32 (level c)
    Ho = Wo = int(sys.argv[1])
    Hi = Wi = Ho - 1
print("resize backward", Hi, "<-", Ho)
dy = torch.randn(B, Ho, Wo, C, device=dev).to(dt)
dx = torch.empty(B, Hi, Wi, C, device=dev, dtype=dt)
cs = torch.zeros(C, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return round(1e3 * ms / reps, 1)


st = stream_ptr(dy)
print("resize_bwd              us", t(lambda: call("cnb_resize_bilinear_bwd", ptr(dy), ptr(dx), B, Hi, Wi, Ho, Wo, C, dtype_code(dt), st)))
print("resize_bwd + colsum     us", t(lambda: call("cnb_resize_bilinear_bwd_colsum", ptr(dy), ptr(dx), B, Hi, Wi, Ho, Wo, C, ptr(cs), 1, dtype_code(dt), st)))
print("separate bias_grad pass us", t(lambda: call("cnb_bias_grad", ptr(dx), C, B * Hi * Wi, C, ptr(cs), 1, dtype_code(dt), st)))
want = dx.float().sum((0, 1, 2))
cs.zero_()
call("cnb_resize_bilinear_bwd_colsum", ptr(dy), ptr(dx), B, Hi, Wi, Ho, Wo, C, ptr(cs), 0, dtype_code(dt), st)
print("colsum rel err", float((cs - want).norm() / want.norm()))
