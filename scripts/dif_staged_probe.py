"""Times dlevmar_dif on resident Stereo19 batches: staged shape vs the warp-per-fit kernel (bit-identity checked)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from curvepointslam_b200 import lm, synth
from oracle_lib import DIF, STEREO19

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
b = synth.make_stereo19_batch_fast(B, K=64, seed=5)
dev = torch.device("cuda", 0)
d = {k: torch.from_numpy(b[k]).to(dev) for k in ("meas", "p0", "off", "counts")}
out = {}
for name, path in (("staged", lm.PATH_STAGED), ("warp", lm.PATH_MONOLITHIC)):
    if name == "warp" and B > 65536:
        continue
    ctx = lm.LmContext(0)
    ctx.set_path(path)
    ctx.set_profiling(True)
    best = 1e9
    for rep in range(3):
        p = d["p0"].clone()
        info = torch.zeros(B, 10, dtype=torch.float64, device=dev)
        ret = torch.zeros(B, dtype=torch.int32, device=dev)
        extra = torch.zeros(B, 2, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.fit_batched_dev(STEREO19, DIF, B, p.data_ptr(), d["meas"].data_ptr(), None, d["off"].data_ptr(),
                            d["counts"].data_ptr(), 512, info.data_ptr(), ret.data_ptr(), extra.data_ptr())
        ctx.sync(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
        if rep == 0:
            print(name, "profile ms [solve, eval, build, tail]", ctx.get_profile(), "stats", ctx.launch_stats())
    out[name] = (p.cpu().numpy(), info.cpu().numpy())
    print(f"{name}: {best*1e3:.1f} ms for {B} fits = {B/best/1e3:.1f} k fits/s; mean iters {out[name][1][:,5].mean():.1f}")
    ctx.close()
if len(out) == 2:
    print("bit-identical:", np.array_equal(out["staged"][0].view(np.int64), out["warp"][0].view(np.int64)),
          np.array_equal(out["staged"][1].view(np.int64), out["warp"][1].view(np.int64)))
