#!/usr/bin/env python
"""m = 8 fits through the host entry (chunk pipeline) in both execution shapes: bit-identical results, timing."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curvepointslam_b200 import lm, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 70000
b = synth.make_image8_batch(B, K=64, seed=21)
for variant, name in ((lm.DER, "der"), (lm.DIF, "dif")):
    res = {}
    for path, pname in ((lm.PATH_MONOLITHIC, "mono"), (lm.PATH_AUTO, "auto")):
        ctx = lm.LmContext(0)
        ctx.set_path(path)
        ctx.fit_batched(lm.IMAGE8, variant, b["p0"], b["meas"], b["off"], b["counts"])
        dt = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            r = ctx.fit_batched(lm.IMAGE8, variant, b["p0"], b["meas"], b["off"], b["counts"])
            dt = min(dt, time.perf_counter() - t0)
        res[pname] = r
        print(f"{name} {pname}: {dt * 1e3:.2f} ms  last launch {ctx.launch_stats()}", flush=True)
        ctx.close()
    a, c = res["mono"], res["auto"]
    same = all(np.array_equal(a[k].view(np.int64) if a[k].dtype == np.float64 else a[k],
                              c[k].view(np.int64) if c[k].dtype == np.float64 else c[k]) for k in ("p", "info", "ret", "extra"))
    print(f"{name}: bit-identical across shapes: {same}")
    assert same
