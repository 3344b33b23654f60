"""Closest point on a planar cubic ("next" row 3; closest_bezier_pt, curveFittingClass.cpp:936-1043).  The reference's
root finder is GSL (absent): the oracle's restated algorithm is checked against numpy.roots and a dense scan of t;
the CUDA kernel is checked against the oracle bit for bit."""
import ctypes as C

import numpy as np
import pytest

from oracle_lib import oracle

dp = C.POINTER(C.c_double)


def bezier(cp, t):
    cp = np.asarray(cp).reshape(4, 2)
    u = 1 - t
    w = np.stack([u ** 3, 3 * u * u * t, 3 * u * t * t, t ** 3], axis=-1)
    return w @ cp


def make_queries(Q, seed):
    rng = np.random.default_rng(seed)
    cp = np.empty((Q, 4, 2))
    cp[:, :, 0] = np.cumsum(rng.uniform(5, 40, (Q, 4)), axis=1) + rng.uniform(0, 100, (Q, 1))
    cp[:, :, 1] = 230 - np.cumsum(rng.uniform(5, 40, (Q, 4)), axis=1) + rng.normal(0, 5, (Q, 4))
    t = rng.uniform(-0.3, 1.3, Q)
    pts = np.stack([bezier(cp[q], np.array([np.clip(t[q], 0, 1)]))[0] for q in range(Q)]) + rng.normal(0, 6, (Q, 2))
    return cp.reshape(Q, 8), pts


def test_polynomial_roots_match_numpy():
    L = oracle().lib
    rng = np.random.default_rng(0)
    for _ in range(500):
        deg = int(rng.integers(1, 6))
        a = rng.standard_normal(deg + 1) * rng.choice([1, 100, 0.01], deg + 1)
        zr, zi = np.zeros(5), np.zeros(5)
        assert L.orc_poly_roots(a.ctypes.data_as(dp), deg, zr.ctypes.data_as(dp), zi.ctypes.data_as(dp)) == 0
        mine = np.sort_complex(zr[:deg] + 1j * zi[:deg])
        ref = np.sort_complex(np.roots(a[::-1]))
        assert np.abs(mine - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())


def test_oracle_closest_point_matches_dense_scan():
    curves, pts = make_queries(200, 1)
    nearest, dist, t, nroots = oracle().closest(curves, pts)
    tt = np.linspace(0.0, 1.0, 20001)
    hit = 0
    for q in range(200):
        if nroots[q] <= 0 or dist[q] > 3000:
            continue
        d = np.linalg.norm(bezier(curves[q], tt) - pts[q], axis=1)
        # the reference only considers an end point when some real root lies outside [0,1]: when it does pick a
        # point, that point must be the global minimiser over [0,1]
        if abs(dist[q] - d.min()) < 1e-3:
            hit += 1
            assert abs(t[q] - tt[d.argmin()]) < 1e-3
            assert np.allclose(nearest[q], bezier(curves[q], np.array([t[q]]))[0], atol=1e-9)
    assert hit > 150


@pytest.mark.gpu
def test_gpu_closest_bit_exact():
    from curvepointslam_b200 import lm
    curves, pts = make_queries(5000, 2)
    gn, gd, gt, gr = lm.closest_bezier_pt(curves, pts)
    on, od, ot, orr = oracle().closest(curves, pts)
    assert np.array_equal(gr, orr)
    b = lambda a: np.ascontiguousarray(a).view(np.int64)
    assert np.array_equal(b(gd), b(od)) and np.array_equal(b(gt), b(ot)) and np.array_equal(b(gn), b(on))
    # shared curve + polynomial-coefficient form (the reference's own argument convention)
    p = np.zeros(8)
    oracle().lib.orc_control2coeffs(np.ascontiguousarray(curves[0]).ctypes.data_as(dp), p.ctypes.data_as(dp))
    gn2, gd2, gt2, _ = lm.closest_bezier_pt(p, pts[:64], is_control_points=False)
    on2, od2, ot2, _ = oracle().closest(p[None], pts[:64], is_control_points=False)
    assert np.array_equal(b(gd2), b(od2)) and np.array_equal(b(gt2), b(ot2))
