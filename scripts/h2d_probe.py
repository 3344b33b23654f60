#!/usr/bin/env python
"""Host-to-device rate of the pinned buffers the end-to-end entry copies (1.12 GB per 262 144 fits), alone and in 16
pieces: the floor of the end-to-end time."""
import sys

import torch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
n = B * 512
h = torch.empty(n, dtype=torch.float64).pin_memory()
h.normal_()
d = torch.empty(n, dtype=torch.float64, device="cuda")
out = torch.empty(B * 30, dtype=torch.float64, device="cuda")
h_out = torch.empty(B * 30, dtype=torch.float64).pin_memory()
for pieces in (1, 16):
    ts = []
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step = n // pieces
        for k in range(pieces):
            d[k * step:(k + 1) * step].copy_(h[k * step:(k + 1) * step], non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts[1:])
    print(f"H2D {n * 8 / 1e9:.2f} GB in {pieces} piece(s): {ms:.2f} ms = {n * 8 / ms / 1e6:.1f} GB/s")
for it in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h_out.copy_(out, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"D2H {B * 30 * 8 / 1e6:.1f} MB: {e0.elapsed_time(e1):.2f} ms")
