#!/usr/bin/env python
"""Times dlevmar_der (or dlevmar_dif: third argument "dif") on a resident batch of per-contour m=8 fits in both execution shapes (a warp per fit / a thread
per fit) with CUDA events and checks that they agree bit for bit."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curvepointslam_b200 import lm, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
paths = sys.argv[2].split(",") if len(sys.argv) > 2 else ["mono", "thread"]
variant = lm.DIF if len(sys.argv) > 3 and sys.argv[3] == "dif" else lm.DER
b = synth.make_image8_batch(B, K=64, seed=13)
dev = torch.device("cuda", 0)
d_meas, d_p0 = torch.from_numpy(b["meas"]).to(dev), torch.from_numpy(b["p0"]).to(dev)
d_off, d_cnt = torch.from_numpy(b["off"]).to(dev), torch.from_numpy(b["counts"]).to(dev)
d_p = torch.empty_like(d_p0)
d_info = torch.zeros((B, 10), dtype=torch.float64, device=dev)
d_ret = torch.zeros(B, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream()
res = {}
for name in paths:
    ctx = lm.LmContext(0)
    ctx.set_path(lm.PATH_MONOLITHIC if name == "mono" else lm.PATH_STAGED)
    ts = []
    for it in range(5):
        d_p.copy_(d_p0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        ctx.fit_batched_dev(lm.IMAGE8, variant, B, d_p.data_ptr(), d_meas.data_ptr(), None, d_off.data_ptr(),
                            d_cnt.data_ptr(), 128, d_info.data_ptr(), d_ret.data_ptr(), None, stream=st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res[name] = (d_p.cpu().numpy().copy(), d_info.cpu().numpy().copy())
    ms = min(ts[2:])
    print(f"{name}: {ms:.3f} ms for {B} fits = {B / ms / 1e3:.1f} M contours/s  all {['%.2f' % t for t in ts]}", flush=True)
if len(res) == 2:
    a, c = res["mono"], res["thread"]
    print("bit-identical p:", np.array_equal(a[0].view(np.int64), c[0].view(np.int64)),
          " info:", np.array_equal(a[1].view(np.int64), c[1].view(np.int64)))
