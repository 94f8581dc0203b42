#!/usr/bin/env python
"""Achievable pure-read HBM bandwidth on this GPU (torch reductions over large buffers), to put the K1 roofline
fraction (denominator = copy bandwidth, read+write bytes) into context."""
import torch
dev = torch.device("cuda", 0)
for nbytes in (1 << 28, 1 << 30):
    x = torch.ones(nbytes // 4, dtype=torch.float32, device=dev)
    y = torch.empty_like(x)
    for name, fn, traffic in (("sum (read only)", lambda: x.sum(), nbytes), ("max (read only)", lambda: x.max(), nbytes),
                              ("copy (read+write)", lambda: y.copy_(x), 2 * nbytes)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{nbytes >> 20} MiB {name}: {ms:.3f} ms  {traffic / ms / 1e6:.0f} GB/s")
