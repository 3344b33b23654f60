#!/usr/bin/env python
"""Generates tests/golden/lm_*.npz with the REFERENCE's own levmar-2.5 (compiled unchanged from
/root/reference/levmar-2.5 into oracle/_ref by `make -C oracle ref`) driving the restated curve models, in the
oracle's libm math mode (glibc sin/cos/pow exactly as the reference calls them) and in the deterministic math
mode the CUDA kernels use.  Runs only where /root/reference exists; the vectors are committed.

    python tests/golden/gen_lm_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from curvepointslam_b200 import synth  # noqa: E402
from oracle_lib import DER, DIF, IMAGE8, MATH_DET, MATH_LIBM, STEREO19, build_oracle, oracle  # noqa: E402

# curveFittingClass.cpp:659-660: opts[4] = LM_DIFF_DELTA*1E1, one ulp below the literal 1e-5; the vectors carry it
OPTS = [1e-10, 1e-10, 1e-10, 1e-20, 1e-6 * 1e1]


def main():
    build_oracle(with_ref=True)
    o = oracle()
    assert o.ref is not None, "oracle/_ref missing: needs /root/reference"
    rng = np.random.default_rng(7)
    counts = rng.integers(6, 102, size=(24, 4)).astype(np.int32)
    counts[0] = (64, 64, 64, 64)
    counts[1] = (3, 2, 3, 2)
    b19 = synth.make_stereo19_batch(24, seed=101, counts=counts, round_pixels=True)
    b8 = synth.make_image8_batch(24, K=64, seed=102)
    out = {"opts": np.array(OPTS)}
    for tag, model, b in (("s19", STEREO19, b19), ("i8", IMAGE8, b8)):
        out[f"{tag}_meas"], out[f"{tag}_off"], out[f"{tag}_counts"], out[f"{tag}_p0"] = b["meas"], b["off"], b["counts"], b["p0"]
        for vname, variant in (("der", DER), ("dif", DIF)):
            for mname, mode in (("det", MATH_DET), ("libm", MATH_LIBM)):
                r = o.fit_batch(model, variant, b["p0"], b["meas"], b["off"], b["counts"], OPTS, math_mode=mode,
                                use_ref=True)
                out[f"{tag}_{vname}_{mname}_p"] = r["p"]
                out[f"{tag}_{vname}_{mname}_info"] = r["info"]
                out[f"{tag}_{vname}_{mname}_ret"] = r["ret"]
    np.savez_compressed(os.path.join(HERE, "lm_reference_levmar.npz"), **out)
    print("written", os.path.join(HERE, "lm_reference_levmar.npz"))


if __name__ == "__main__":
    main()
