"""Shared acceptance criteria of the parity tests (test infrastructure)."""
import numpy as np

EKF_RTOL = 1e-9    # north star: EKF state and covariance within 1e-9 relative
# Entries below EKF_FLOOR * max|b| are held to the absolute error EKF_RTOL * EKF_FLOOR * max|b| instead of to 1e-9 of
# themselves.  Why 1e-2 and not smaller: the update subtracts K S K^T with cond(S) ~ 5e4 on the SURVEY 8(d) workload
# (landmarks U(-50,50): |H| ~ 50, |S| ~ 3e3 against R ~ 0.04), so ANY float64 evaluation carries an absolute error of
# ~1e-12 * max|P| in every entry -- measured against an 80-bit long-double evaluation of kalmanClass.cpp:863-926 at
# N = 4 512 the C restatement itself (SVD pseudo-inverse, like the reference) is off by 1.1e-12 * max|P| and the
# Cholesky route the CUDA path takes by 5.5e-12 * max|P| -- i.e. a floor of 1e-3 would fail the reference against
# itself.  Both norms are reported by the tests.
EKF_FLOOR = 1e-2


def maxnorm_rel(a, b):
    """max|a - b| / max|b|"""
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def elementwise_rel(a, b, floor=EKF_FLOOR, scale=None):
    """max over entries of |a - b| / max(|b|, floor * max|b|): the per-element relative error, with entries that
    cancel towards zero (cross-covariances far below the diagonal) measured against floor * max|b| instead of against
    themselves.  `scale` overrides max|b| when b is only a block of a larger matrix."""
    s = np.abs(b).max() if scale is None else scale
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor * max(s, 1e-300))).max())


def ekf_close(a, b, what="", rtol=EKF_RTOL, scale=None):
    """Asserts the element-wise criterion and returns both norms (so that a test can report them)."""
    ew, mn = elementwise_rel(a, b, scale=scale), maxnorm_rel(a, b)
    assert ew <= rtol, f"{what}: element-wise relative error {ew:.3e} > {rtol:.0e} (max-norm relative {mn:.3e})"
    return ew, mn
