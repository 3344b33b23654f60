"""GPU parity tests of the linear solver's failure paths (dAx_eq_b_LU_noLapack, Axb_core.c:1044-1132):
normal equations with all-zero rows and zero pivots, in every execution shape.

A stereo fit one of whose curves has no samples has eight all-zero Jacobian columns; with tau = 0 the damping
starts at mu = 0, so the augmented matrix J^T J + mu I keeps its all-zero rows and the reference's LU reports
"Singular matrix A" on every attempt: dlevmar_der doubles nu in its inner loop until the integer wraps (stop
reason 5 after 30 failed solves, lm_core.c:362-371), dlevmar_dif treats each failure as a rejected step
(:717-722).  With the reference's tau the same fits are solvable and take the ordinary path.  The CPU oracle
agrees with the reference's own levmar on these batches bit for bit (tests/test_oracle_lm.py covers the solver,
test_singular_batches_oracle_vs_reference below the batches used here); the GPU must agree with the oracle
bit for bit -- parameters, info[0..9], return codes."""
import numpy as np
import pytest

from curvepointslam_b200 import lm, synth
from oracle_lib import DER, DIF, STEREO19, oracle
from test_lm_gpu import OPTS, assert_same_fit

TAU0 = [0.0] + list(OPTS[1:])
COUNTS = np.array([[40, 0, 40, 0], [0, 40, 0, 40], [40, 40, 40, 40], [30, 0, 0, 30], [64, 64, 0, 0], [1, 1, 1, 1],
                   [1, 0, 1, 0], [101, 0, 101, 0], [12, 12, 12, 12], [3, 2, 3, 2]], dtype=np.int32)


def _batch(reps=4):
    counts = np.tile(COUNTS, (reps, 1))
    return synth.make_stereo19_batch(len(counts), seed=91, counts=counts, round_pixels=True)


@pytest.mark.parametrize("variant", [DER, DIF])
def test_singular_batches_oracle_vs_reference(variant):
    """CPU: the oracle against the reference's own levmar on the batches of this file."""
    if oracle().ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    b = _batch(1)
    for opts in (OPTS, TAU0):
        a = oracle().fit_batch(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], opts, itmax=60)
        r = oracle().fit_batch(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], opts, itmax=60,
                               use_ref=True)
        r["extra"] = a["extra"]  # the reference has no such counters
        assert_same_fit(a, r)
    assert (a["info"][:, 6] == 5).any() and (a["ret"] == -1).sum() == 2  # singular solves; n < m is refused


@pytest.mark.gpu
@pytest.mark.parametrize("path", [lm.PATH_MONOLITHIC, lm.PATH_STAGED])
@pytest.mark.parametrize("variant", [DER, DIF])
def test_singular_normal_equations_every_shape(path, variant):
    ctx = lm.LmContext(0)
    ctx.set_path(path)
    b = _batch()
    try:
        for opts in (OPTS, TAU0):
            gpu = ctx.fit_batched(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], opts, itmax=60)
            cpu = oracle().fit_batch(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], opts, itmax=60,
                                     nthreads=4)
            assert_same_fit(gpu, cpu)
        # the tau = 0 run went through the failing solver: stop reason 5 (der: nu wrapped after 30 failed solves of one
        # iteration; dif: every step rejected), and info[9] counts every attempt
        r5 = gpu["info"][:, 6] == 5
        assert r5.any() and (gpu["info"][r5, 9] >= 30).all()
        assert (gpu["ret"] == -1).sum() == 2 * (len(b["off"]) - 1) // len(COUNTS)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_singular_systems_through_the_dense_handover(monkeypatch):
    """The staged der solve keeps J^T J + mu I in block-arrow form and hands systems it cannot pivot inside a curve's
    own rows to the dense LU (cps_lm_get_solve_handovers).  All-zero rows must take that road too and end like the
    reference; CPS_LM_FORCE_HANDOVER=1 hands every system over, CPS_LM_DENSE_SOLVE uses the dense work matrix from the
    start; CPS_LM_STAGED_TAIL=0 keeps every solve of the batch inside the staged kernels."""
    b = _batch()
    cpu = oracle().fit_batch(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], TAU0, itmax=60, nthreads=4)
    monkeypatch.setenv("CPS_LM_STAGED_TAIL", "0")
    for env in (None, "CPS_LM_FORCE_HANDOVER", "CPS_LM_DENSE_SOLVE"):
        monkeypatch.delenv("CPS_LM_FORCE_HANDOVER", raising=False)
        monkeypatch.delenv("CPS_LM_DENSE_SOLVE", raising=False)
        if env:
            monkeypatch.setenv(env, "1")
        ctx = lm.LmContext(0)
        ctx.set_path(lm.PATH_STAGED)
        try:
            gpu = ctx.fit_batched(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], TAU0, itmax=60)
        finally:
            ctx.close()
        assert_same_fit(gpu, cpu)
