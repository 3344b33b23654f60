#!/usr/bin/env python
"""Times the host-buffer entry point (cps_dlevmar_der_batched: H2D + fits + D2H) on pinned memory."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curvepointslam_b200 import _lib, lm, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
b = synth.make_stereo19_batch_fast(B, K=64, seed=2026)
h_meas = torch.from_numpy(b["meas"]).pin_memory()
h_p0 = torch.from_numpy(b["p0"]).pin_memory()
h_off = torch.from_numpy(b["off"]).pin_memory()
h_cnt = torch.from_numpy(b["counts"]).pin_memory()
h_p = torch.empty_like(h_p0).pin_memory()
h_info = torch.zeros((B, 10), dtype=torch.float64).pin_memory()
h_ret = torch.zeros(B, dtype=torch.int32).pin_memory()
opts = np.array(lm.REFERENCE_OPTS)
ctx = lm.LmContext(0)
if len(sys.argv) > 2:  # force an execution shape: mono | staged
    ctx.set_path(lm.PATH_MONOLITHIC if sys.argv[2] == "mono" else lm.PATH_STAGED)
ctx.set_profiling(True)
fn = _lib.lib().cps_dlevmar_der_batched
ts = []
for it in range(6):
    h_p.copy_(h_p0)
    t0 = time.perf_counter()
    rc = fn(ctx._h, 19, B, h_p.data_ptr(), h_meas.data_ptr(), None, h_off.data_ptr(), h_cnt.data_ptr(), 1000,
            opts.ctypes.data, h_info.data_ptr(), h_ret.data_ptr(), None)
    assert rc == 0
    ts.append((time.perf_counter() - t0) * 1e3)
    prof = ctx.get_profile(reset=True)
print("kernel time of the last call:", {k: (round(v[0], 2), v[1]) for k, v in prof.items()}, "sum %.2f ms" % sum(v[0] for v in prof.values()))
print(f"chunk={os.environ.get('CPS_LM_STAGED_CHUNK', 'default')}: e2e {min(ts[2:]):.2f} ms = {B / min(ts[2:]) * 1e3 / 1e6:.3f} M fits/s  all {['%.1f' % t for t in ts]}")
