"""GPU parity tests of the staged shape of dlevmar_dif on Stereo19 batches (csrc/cps_lm_dif_staged.cuh: solve / eval /
build kernels, lane = fit, four warps per 32 fits in the build): bit-exact against the CPU oracle (itself bit-identical
to the reference's levmar, tests/test_oracle_lm.py), against the committed vectors the reference levmar produced, and
against the warp-per-fit kernel."""
import os

import numpy as np
import pytest

from curvepointslam_b200 import lm, synth
from oracle_lib import DIF, STEREO19, oracle
from test_lm_gpu import OPTS, _bits, assert_same_fit

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def sctx():
    c = lm.LmContext(0)
    c.set_path(lm.PATH_STAGED)
    yield c
    c.close()


def test_dif_staged_dense_bit_exact(sctx):
    b = synth.make_stereo19_batch(200, K=64, seed=71)
    gpu = sctx.fit_batched(STEREO19, DIF, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    cpu = oracle().fit_batch(STEREO19, DIF, b["p0"], b["meas"], b["off"], b["counts"], OPTS, nthreads=8)
    assert_same_fit(gpu, cpu)
    assert np.array_equal(gpu["extra"], cpu["extra"])  # J^T J rebuilds and secant updates, fit by fit
    assert (gpu["info"][:, 6] == 2).all() and gpu["info"][:, 5].mean() > 20


def test_dif_staged_ragged_small_and_mixed(sctx):
    """Ragged segments, fits on both sides of the n*m <= 1024 switch inside one 32-fit group (the unblocked loop walks
    the rows downwards), lengths that are not multiples of four rows, a group with fewer than 32 fits."""
    rng = np.random.default_rng(9)
    counts = rng.integers(5, 102, size=(75, 4)).astype(np.int32)
    counts[0] = (101, 101, 101, 101)
    counts[1] = (3, 2, 3, 2)      # n*m = 380: unblocked
    counts[2] = (7, 7, 6, 7)      # n*m = 1026: blocked, first size above the switch
    counts[3] = (7, 7, 6, 6)      # n*m = 988: unblocked
    counts[4] = (21, 3, 22, 30)
    counts[5] = (16, 16, 16, 16)
    counts[6] = (17, 15, 33, 31)
    counts[40] = (4, 3, 3, 3)     # a small fit in the second group as well
    b = synth.make_stereo19_batch(75, seed=72, counts=counts, round_pixels=True)
    gpu = sctx.fit_batched(STEREO19, DIF, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    cpu = oracle().fit_batch(STEREO19, DIF, b["p0"], b["meas"], b["off"], b["counts"], OPTS, nthreads=8)
    assert_same_fit(gpu, cpu)
    assert np.array_equal(gpu["extra"], cpu["extra"])


def test_dif_staged_explicit_t_opts_itmax(sctx):
    b = synth.make_stereo19_batch(40, K=40, seed=73)
    t = np.tile(np.linspace(0.0, 1.0, 40), 40 * 4)
    gpu = sctx.fit_batched(STEREO19, DIF, b["p0"], b["meas"], b["off"], b["counts"], OPTS, t=t)
    cpu = oracle().fit_batch(STEREO19, DIF, b["p0"], b["meas"], b["off"], b["counts"], OPTS, t=t, nthreads=8)
    assert_same_fit(gpu, cpu)
    for opts, itmax in [(None, 1000), (OPTS, 25), (OPTS, 3), (OPTS, 1), (OPTS, 0),
                        ([1e-3, 1e-2, 1e-15, 1e-20, 1e-6], 1000), ([1e-3, 1e-15, 1e-15, 1e+3, 1e-6], 1000),
                        ([1e-10, 1e-10, 1e-10, 1e-20, -1e-5], 1000)]:  # the last: central differences (warp-per-fit kernel)
        gpu = sctx.fit_batched(STEREO19, DIF, b["p0"], b["meas"], b["off"], b["counts"], opts, itmax=itmax)
        cpu = oracle().fit_batch(STEREO19, DIF, b["p0"], b["meas"], b["off"], b["counts"], opts, itmax=itmax, nthreads=8)
        assert_same_fit(gpu, cpu, check_mu=itmax > 0)


def test_dif_staged_golden_vectors_from_reference_levmar(sctx):
    g = np.load(os.path.join(HERE, "golden", "lm_reference_levmar.npz"))
    gpu = sctx.fit_batched(STEREO19, DIF, g["s19_p0"], g["s19_meas"], g["s19_off"], g["s19_counts"], list(g["opts"]))
    assert np.array_equal(_bits(gpu["p"]), _bits(g["s19_dif_det_p"]))
    assert np.array_equal(_bits(gpu["info"]), _bits(g["s19_dif_det_info"]))
    assert np.array_equal(gpu["ret"], g["s19_dif_det_ret"])


def test_dif_staged_edge_cases(sctx):
    b = synth.make_stereo19_batch(4, seed=16, counts=np.array([[2, 2, 2, 2], [30, 30, 30, 30], [30, 30, 30, 30],
                                                               [1, 1, 1, 1]], dtype=np.int32))
    meas = b["meas"].copy()
    o2 = b["off"][2]
    meas[o2 + 1: o2 + 60: 2] = 200.0  # horizontal segment: t = NaN -> stop 7
    gpu = sctx.fit_batched(STEREO19, DIF, b["p0"], meas, b["off"], b["counts"], OPTS)
    cpu = oracle().fit_batch(STEREO19, DIF, b["p0"], meas, b["off"], b["counts"], OPTS)
    assert np.array_equal(gpu["ret"], cpu["ret"])
    assert gpu["ret"][0] == -1 and gpu["ret"][3] == -1 and gpu["ret"][2] == -1 and gpu["info"][2, 6] == 7
    assert np.array_equal(_bits(gpu["p"][[1]]), _bits(cpu["p"][[1]]))
    assert np.array_equal(_bits(gpu["info"][[1]]), _bits(cpu["info"][[1]]))
    assert np.array_equal(_bits(gpu["p"][[0, 3]]), _bits(b["p0"][[0, 3]]))


def test_dif_staged_equals_warp_per_fit_at_scale(sctx):
    """12 000 fits: every fit of the staged shape is bit-identical to the warp-per-fit kernel's; a sample is checked
    against the oracle; AUTO keeps the warp-per-fit kernel below 14 336 fits per call (the measured crossover is about 12 k)."""
    B = 12000
    b = synth.make_stereo19_batch_fast(B, K=64, seed=75)
    mono = lm.LmContext(0)
    mono.set_path(lm.PATH_MONOLITHIC)
    g1 = sctx.fit_batched(STEREO19, DIF, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    g2 = mono.fit_batched(STEREO19, DIF, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    assert np.array_equal(g1["ret"], g2["ret"]) and np.array_equal(g1["extra"], g2["extra"])
    assert np.array_equal(_bits(g1["p"]), _bits(g2["p"])) and np.array_equal(_bits(g1["info"]), _bits(g2["info"]))
    idx = np.random.default_rng(1).choice(B, 48, replace=False)
    sub_off = np.arange(49, dtype=np.int64) * 512
    sub_meas = np.concatenate([b["meas"][b["off"][i]:b["off"][i + 1]] for i in idx])
    cpu = oracle().fit_batch(STEREO19, DIF, b["p0"][idx], sub_meas, sub_off, b["counts"][idx], OPTS, nthreads=8)
    assert np.array_equal(_bits(g1["p"][idx]), _bits(cpu["p"])) and np.array_equal(_bits(g1["info"][idx]), _bits(cpu["info"]))
    auto = lm.LmContext(0)
    g3 = auto.fit_batched(STEREO19, DIF, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    assert np.array_equal(_bits(g3["p"]), _bits(g1["p"])) and np.array_equal(_bits(g3["info"]), _bits(g1["info"]))
    assert auto.launch_stats()["launches"] >= 1
    mono.close()
    auto.close()
