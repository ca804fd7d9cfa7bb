"""HBM bandwidth by read/write mix on this GPU (torch library kernels, CUDA-event timed): pure write (fill), pure read
(sum), 1:1 copy, and a 1-read : 6-write pattern like the b1 expand layer.  Context for roofline.frac of write-heavy
kernels: MEASURED_PEAKS.json's HBM number is a copy (1:1) figure."""
import json

import torch


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def main():
    n = 1 << 30  # 4 GiB of fp32
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    y = torch.empty(n, dtype=torch.float32, device="cuda")
    x.fill_(1.0)
    out = {}
    out["write_only_GBps"] = 4 * n / timed(lambda: y.fill_(2.0)) / 1e9
    out["read_only_GBps"] = 4 * n / timed(lambda: x.sum()) / 1e9
    out["copy_1r1w_GBps"] = 8 * n / timed(lambda: y.copy_(x)) / 1e9
    # 1 read : 6 writes -- expand a (n/6) vector into 6 copies (broadcast store), like 16 -> 96 channels
    m = n // 6
    src = x[:m]
    dst = y[: 6 * m].view(6, m)
    out["expand_1r6w_GBps"] = 4 * 7 * m / timed(lambda: dst.copy_(src.expand(6, m))) / 1e9
    # 6 reads : 1 write -- stride-2-like reduction
    out["reduce_6r1w_GBps"] = 4 * 7 * m / timed(lambda: torch.sum(x[: 6 * m].view(6, m), dim=0, out=y[:m])) / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
