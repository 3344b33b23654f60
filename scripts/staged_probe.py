#!/usr/bin/env python
"""Times dlevmar_der on a resident Stereo19 batch in both execution shapes (monolithic / staged) with CUDA events."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curvepointslam_b200 import lm, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
paths = sys.argv[2].split(",") if len(sys.argv) > 2 else ["mono", "staged"]
dev = torch.device("cuda", 0)
b = synth.make_stereo19_batch_fast(B, K=64, seed=2026)
d_meas, d_p0, d_off, d_cnt = (torch.from_numpy(b[k]).to(dev) for k in ("meas", "p0", "off", "counts"))
d_p = torch.empty_like(d_p0)
d_info = torch.zeros((B, 10), dtype=torch.float64, device=dev)
d_ret = torch.zeros(B, dtype=torch.int32, device=dev)
d_extra = torch.zeros((B, 2), dtype=torch.int32, device=dev)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
res = {}
for name in paths:
    ctx = lm.LmContext(0)
    ctx.set_path(lm.PATH_MONOLITHIC if name == "mono" else lm.PATH_STAGED)
    ts = []
    if os.environ.get("PROBE_PROFILE"):
        ctx.set_profiling(True)
    if os.environ.get("PROBE_FMA") and name == "staged":
        ctx.set_arith(True)
    for it in range(6):
        d_p.copy_(d_p0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.fit_batched_dev(19, lm.DER, B, d_p.data_ptr(), d_meas.data_ptr(), None, d_off.data_ptr(), d_cnt.data_ptr(),
                            512, d_info.data_ptr(), d_ret.data_ptr(), d_extra.data_ptr(), stream=stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts[2:])
    if os.environ.get("PROBE_PROFILE"):
        print("profile (6 calls):", ctx.get_profile(), "solve hand-overs of the last call:", ctx.solve_handovers())
    res[name] = (d_p.cpu().numpy().copy(), d_info.cpu().numpy().copy())
    print(f"{name}: {ms:.3f} ms for {B} fits = {B / ms * 1e3 / 1e6:.3f} M fits/s   all {['%.2f' % t for t in ts]}  "
          f"stats {ctx.launch_stats()}", flush=True)
    ctx.close()
if len(res) == 2:
    a, c = res["mono"], res["staged"]
    print("bit-identical p:", np.array_equal(a[0].view(np.int64), c[0].view(np.int64)),
          " info:", np.array_equal(a[1].view(np.int64), c[1].view(np.int64)))
