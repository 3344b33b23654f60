"""v4 update-build check: staged vs warp-per-fit with a small itmax (bounded run time even if the sums are wrong)."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from curvepointslam_b200 import lm, synth
from oracle_lib import DIF, STEREO19
B = int(sys.argv[1]); itmax = int(sys.argv[2])
b = synth.make_stereo19_batch_fast(B, K=64, seed=5)
OPTS = [1e-10, 1e-10, 1e-10, 1e-20, 1e-6 * 1e1]
out = {}
for name, path in (("staged", lm.PATH_STAGED), ("warp", lm.PATH_MONOLITHIC)):
    ctx = lm.LmContext(0); ctx.set_path(path)
    out[name] = ctx.fit_batched(STEREO19, DIF, b["p0"], b["meas"], b["off"], b["counts"], OPTS, itmax=itmax)
    ctx.close()
a, c = out["staged"], out["warp"]
same = (a["p"].view(np.int64) == c["p"].view(np.int64)).all(axis=1)
print("B", B, "itmax", itmax, "bit-identical fits", int(same.sum()), "of", B, "first bad", np.nonzero(~same)[0][:10],
      "max rel", np.nanmax(np.abs(a["p"] - c["p"]) / (np.abs(c["p"]) + 1e-300)))
