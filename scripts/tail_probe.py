#!/usr/bin/env python
"""Staged der at 262 144 resident fits with the late fits handed to the warp-per-fit kernel at different thresholds
(CPS_LM_STAGED_TAIL): time, kernel-time profile, and bit-identity against the threshold-0 run."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curvepointslam_b200 import lm, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
ths = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 2048, 4096, 8192, 16384]
dev = torch.device("cuda", 0)
b = synth.make_stereo19_batch_fast(B, K=64, seed=2026)
d_meas, d_p0, d_off, d_cnt = (torch.from_numpy(b[k]).to(dev) for k in ("meas", "p0", "off", "counts"))
d_p = torch.empty_like(d_p0)
d_info = torch.zeros((B, 10), dtype=torch.float64, device=dev)
d_ret = torch.zeros(B, dtype=torch.int32, device=dev)
d_extra = torch.zeros((B, 2), dtype=torch.int32, device=dev)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ref = None
for th in ths:
    os.environ["CPS_LM_STAGED_TAIL"] = str(th)
    ctx = lm.LmContext(0)
    ctx.set_path(lm.PATH_STAGED)
    ctx.set_profiling(True)
    ts = []
    for it in range(6):
        d_p.copy_(d_p0)
        ctx.get_profile(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.fit_batched_dev(19, lm.DER, B, d_p.data_ptr(), d_meas.data_ptr(), None, d_off.data_ptr(), d_cnt.data_ptr(),
                            512, d_info.data_ptr(), d_ret.data_ptr(), d_extra.data_ptr(), stream=stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    prof = ctx.get_profile(reset=True)
    out = (d_p.cpu().numpy().copy(), d_info.cpu().numpy().copy(), d_ret.cpu().numpy().copy(), d_extra.cpu().numpy().copy())
    if ref is None:
        ref = out
    same = all(np.array_equal(a.view(np.int64) if a.dtype == np.float64 else a, c.view(np.int64) if c.dtype == np.float64 else c)
               for a, c in zip(out, ref))
    ms = min(ts[2:])
    print(f"tail<={th}: {ms:.3f} ms = {B / ms * 1e3 / 1e6:.3f} M fits/s  rounds {ctx.last_rounds() if hasattr(ctx, 'last_rounds') else '?'}  "
          f"profile {{k: (round(v[0], 2), v[1]) for k, v in prof.items()}}  same as first: {same}".replace("{{", "{").replace("}}", "}"), flush=True)
    print("   ", {k: (round(v[0], 2), v[1]) for k, v in prof.items()}, flush=True)
    ctx.close()
