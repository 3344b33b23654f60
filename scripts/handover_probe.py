import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from curvepointslam_b200 import lm, synth
from oracle_lib import DER, STEREO19
from test_lm_gpu import OPTS
rng = np.random.default_rng(17)
counts = rng.integers(5, 90, size=(700, 4)).astype(np.int32)
b = synth.make_stereo19_batch(700, seed=93, counts=counts, round_pixels=True)
p0 = b["p0"].copy()
p0[::5] += rng.normal(0, 0.5, size=p0[::5].shape)
os.environ["CPS_LM_STAGED_TAIL"] = "0"
for env in (None, "CPS_LM_FORCE_HANDOVER", "CPS_LM_DENSE_SOLVE"):
    os.environ.pop("CPS_LM_FORCE_HANDOVER", None); os.environ.pop("CPS_LM_DENSE_SOLVE", None)
    if env: os.environ[env] = "3"
    c = lm.LmContext(0); c.set_path(lm.PATH_STAGED)
    r = c.fit_batched(STEREO19, DER, p0, b["meas"], b["off"], b["counts"], OPTS)
    print(env, "handovers", c.solve_handovers(), "nlss total", int(r["info"][:, 9].sum()))
    c.close()
