#!/usr/bin/env python
"""Where the end-to-end time of cps_dlevmar_batched_pix goes: wall time of the call against the CUDA-event time of its
kernels (cps_lm_set_profiling) and the round count, for the default piece schedule and for the batch already resident."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curvepointslam_b200 import _lib, lm, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
b = synth.make_stereo19_batch_fast(B, K=64, seed=2026)
pix = np.rint(b["meas"]).astype(np.int16)
h_pix = torch.from_numpy(pix).pin_memory()
h_p0 = torch.from_numpy(b["p0"]).pin_memory()
h_off = torch.from_numpy(b["off"]).pin_memory()
h_cnt = torch.from_numpy(b["counts"]).pin_memory()
h_p = torch.empty_like(h_p0).pin_memory()
h_info = torch.zeros((B, 10), dtype=torch.float64).pin_memory()
h_ret = torch.zeros(B, dtype=torch.int32).pin_memory()
opts = np.array(lm.REFERENCE_OPTS)
ctx = lm.LmContext(0)
ctx.set_profiling(True)
fn = _lib.lib().cps_dlevmar_batched_pix
for it in range(6):
    h_p.copy_(h_p0)
    torch.cuda.synchronize()
    l0 = ctx.launch_stats()["launches"]
    t0 = time.perf_counter()
    rc = fn(ctx._h, 19, 0, B, h_p.data_ptr(), h_pix.data_ptr(), 1, h_off.data_ptr(), h_cnt.data_ptr(), 1000,
            opts.ctypes.data, h_info.data_ptr(), h_ret.data_ptr(), None)
    dt = (time.perf_counter() - t0) * 1e3
    assert rc == 0
    prof = ctx.get_profile(reset=True)
    print(f"call {it}: e2e {dt:.2f} ms; kernel events: " +
          ", ".join(f"{k} {v[0]:.2f} ms / {v[1]}" for k, v in prof.items()) +
          f"; sum {sum(v[0] for v in prof.values()):.2f} ms; launches {ctx.launch_stats()['launches'] - l0}")
# H2D and D2H alone
d = torch.empty(pix.size, dtype=torch.int16, device="cuda")
dp = torch.empty_like(h_p0, device="cuda")
di = torch.empty((B, 10), dtype=torch.float64, device="cuda")
for name, f in (("H2D pix+p0", lambda: (d.copy_(h_pix, non_blocking=True), dp.copy_(h_p0, non_blocking=True))),
                ("D2H p+info", lambda: (h_p.copy_(dp, non_blocking=True), h_info.copy_(di, non_blocking=True)))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    f()
    torch.cuda.synchronize()
    print(name, f"{(time.perf_counter() - t0) * 1e3:.2f} ms")
