"""Is a read-only HBM stream on this B200 slower than a copy?  Times library kernels (no code of this repo) over tensors far larger
than L2: a pure read (sum, amax), a copy, a two-read one-write add -- the answer bounds what the BatchNorm statistics / backward-reduce
kernels (read-only streams) can reach.  Usage: python tools/probe_read_bw.py > gpurun_out/read_bw_probe.json"""
import json

import torch


def timed(fn, n=10):
    for _ in range(3):
        fn()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    dev = torch.device("cuda")
    n = 1 << 30  # 1 Gi bf16 = 2 GiB per tensor
    a = torch.randn(n // 2, device=dev).to(torch.bfloat16).repeat(2)
    b = torch.empty_like(a)
    c = torch.empty_like(a)
    f = a[: n // 2].float()  # 2 GiB fp32
    out = {}
    gb = a.numel() * 2 / 1e9
    out["sum_bf16_read_only"] = gb / (timed(lambda: a.sum(dtype=torch.float32)) / 1e3)
    out["sum_fp32_read_only"] = f.numel() * 4 / 1e9 / (timed(lambda: f.sum()) / 1e3)
    out["amax_bf16_read_only"] = gb / (timed(lambda: a.amax()) / 1e3)
    out["copy_read_plus_write"] = 2 * gb / (timed(lambda: b.copy_(a)) / 1e3)
    out["add_two_reads_one_write"] = 3 * gb / (timed(lambda: torch.add(a, b, out=c)) / 1e3)
    out["fill_write_only"] = gb / (timed(lambda: b.zero_()) / 1e3)
    out = {k: round(v, 1) for k, v in out.items()}
    out["unit"] = "GB/s"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
