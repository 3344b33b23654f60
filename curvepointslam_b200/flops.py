"""Algorithmic work of one LM fit -- the single place both bench.py and DESIGN.md take their roofline
numerators from.  Dense reference-algorithm counts: a multiply-add is 2 flop.

  der:  F = njev*(F_jac + F_JtJ + F_Jte) + nlss*F_solve + nfev*F_func
  dif:  F = n_jtj*(F_JtJ + F_Jte) + nlss*F_solve + nfev*F_func + n_broyden*4nm + njap*2nm
"""
from __future__ import annotations


def f_jtj(n, m):  # upper triangle of J^T J, n*m*(m+1)/2 multiply-adds
    return n * m * (m + 1)


def f_jte(n, m):
    return 2 * n * m


def f_solve(m):  # LU + two triangular solves
    return (2.0 / 3.0) * m ** 3 + 2 * m ** 2


def f_func(n, m):
    ncp = 8 if m == 19 else 0
    return 26 * (n // 2) + 3 * n + 30 * ncp


def f_jac(n, m):
    if m == 19:
        return 64 * (n // 2) + 110 * 8
    return 10 * (n // 2)


def fit_flops_der(n, m, nfev, njev, nlss):
    return njev * (f_jac(n, m) + f_jtj(n, m) + f_jte(n, m)) + nlss * f_solve(m) + nfev * f_func(n, m)


def fit_flops_dif(n, m, nfev, njap, nlss, n_jtj, n_broyden):
    return (n_jtj * (f_jtj(n, m) + f_jte(n, m)) + nlss * f_solve(m) + nfev * f_func(n, m)
            + n_broyden * 4 * n * m + njap * 2 * n * m)


def fit_bytes(n, m):
    """algorithmic HBM bytes of one fit: measurements + parameters in/out + info + segment counts"""
    return 8 * (n + m + m + 10) + 16


def staged_jac_executed_flops(n):
    """FP64 operations the staged shape's jac kernel really issues per Jacobian evaluation of an m=19 fit: it skips
    the 8 structurally zero entries of every row and walks a block twice (pass A / pass B): 223 + 232 operations per
    two rows (cuobjdump of the two row loops, unrolled by two: 122 DMUL + 101 DADD and 130 DMUL + 102 DADD)."""
    return n * 227.5
