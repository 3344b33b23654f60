"""Staged vs warp-per-fit crossover for dlevmar_dif / dlevmar_der on resident Stereo19 batches."""
import sys, time
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from curvepointslam_b200 import lm, synth
dev = torch.device("cuda", 0)
for var, name, sizes in ((lm.DIF, "dif", (8192, 16384, 24576, 32768, 49152)), (lm.DER, "der", (2048, 4096, 6144, 8192))):
    for B in sizes:
        b = synth.make_stereo19_batch_fast(B, K=64, seed=5)
        d = {k: torch.from_numpy(b[k]).to(dev) for k in ("meas", "p0", "off", "counts")}
        res = {}
        for pname, path in (("staged", lm.PATH_STAGED), ("warp", lm.PATH_MONOLITHIC)):
            ctx = lm.LmContext(0); ctx.set_path(path)
            best = 1e9
            for rep in range(3):
                p = d["p0"].clone()
                info = torch.zeros(B, 10, dtype=torch.float64, device=dev); ret = torch.zeros(B, dtype=torch.int32, device=dev)
                torch.cuda.synchronize(); t0 = time.perf_counter()
                ctx.fit_batched_dev(19, var, B, p.data_ptr(), d["meas"].data_ptr(), None, d["off"].data_ptr(), d["counts"].data_ptr(), 512, info.data_ptr(), ret.data_ptr(), None)
                ctx.sync(); torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            res[pname] = best; ctx.close()
        print(f"{name} B={B}: staged {res['staged']*1e3:.2f} ms, warp {res['warp']*1e3:.2f} ms", flush=True)
