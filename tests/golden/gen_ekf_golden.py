#!/usr/bin/env python
"""Generates tests/golden/ekf_*.npz: golden input/output vectors of the EKF hot path.

The reference's kalmanClass.cpp needs the OpenCV 1.x C API and cannot be built here, and it has no tests of
its own.  This script is a LINE-BY-LINE cv2 (4.13, headless) transcription of
KalmanFilter::UpdateNCurvesAndPoints (kalmanClass.cpp:669-945), UpdatePoints (:948-1126) and PredictKF
(:161-272): cv2.gemm / cv2.invert(DECOMP_SVD) / cv2.transpose / cv2.addWeighted are the routines the legacy
cvMatMul / cvInvert(CV_SVD) / cvTranspose / cvAddWeighted wrap.  It runs only in the build container (cv2 is
not a dependency of the tests); the vectors it writes are committed and the C oracle (oracle/orc_ekf.c) is
checked against them by tests/test_oracle_ekf.py.

    python tests/golden/gen_ekf_golden.py
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PI = 3.1415926535
MEAS_COV, PT_MEAS_COV = 0.5, 5.0
PHI_MEAS_COV, THETA_MEAS_COV, Z_MEAS_COV = 0.1, 0.1, 0.2
V_COV = (0.01, 0.01, 0.5)
W_COV = (0.01, 0.01, 0.01)
DT = 0.2
cos, sin = np.cos, np.sin


def mm(a, b):
    return cv2.gemm(np.ascontiguousarray(a), np.ascontiguousarray(b), 1.0, None, 0.0)


def rbe(phi, theta, psi):  # generate_Rbe, common.h:174-186
    R = np.zeros(9)
    R[0] = cos(psi) * cos(theta)
    R[3] = cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi)
    R[6] = cos(psi) * sin(theta) * cos(phi) - sin(psi) * sin(phi)
    R[1] = sin(psi) * cos(theta)
    R[4] = sin(psi) * sin(theta) * sin(phi) + cos(psi) * cos(phi)
    R[7] = sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi)
    R[2] = -sin(theta)
    R[5] = cos(theta) * sin(phi)
    R[8] = cos(theta) * cos(phi)
    return R.reshape(3, 3)


def rbe_derivs(phi, theta, psi):  # get_Rbe_derivs, kalmanClass.cpp:1275-1298 (typo at :1283-1284 kept)
    p, t, s = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((3, 3))
    p[1, 0] = cos(psi) * sin(theta) * cos(phi) + sin(psi) * sin(phi)
    p[2, 0] = -cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi)
    p[1, 1] = sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi)
    p[2, 1] = -sin(psi) * sin(theta) * sin(phi) - cos(psi) * cos(phi)
    p[1, 2] = cos(theta) * cos(phi)
    p[2, 2] = -cos(theta) * sin(phi)
    t[0, 0] = -cos(psi) * sin(theta)
    t[1, 0] = cos(psi) * cos(theta) * sin(phi)
    t[2, 0] = cos(psi) * cos(theta) * cos(phi)
    t[0, 1] = -sin(psi) * sin(theta)
    t[1, 1] = sin(psi) * cos(theta) * sin(phi)
    t[1, 1] = sin(psi) * cos(theta) * cos(phi)
    t[0, 2] = -cos(theta)
    t[1, 2] = -sin(theta) * sin(phi)
    t[2, 2] = -sin(theta) * cos(phi)
    s[0, 0] = -sin(psi) * cos(theta)
    s[1, 0] = -sin(psi) * sin(theta) * sin(phi) - cos(psi) * cos(phi)
    s[2, 0] = -sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi)
    s[0, 1] = cos(psi) * cos(theta)
    s[1, 1] = cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi)
    s[2, 1] = cos(psi) * sin(theta) * cos(phi) + sin(psi) * sin(phi)
    return p, t, s


def rot8(psi):
    Rot, Rd = np.zeros((8, 8)), np.zeros((8, 8))
    for i in range(4):
        Rot[i, i] = cos(psi); Rot[i + 4, i + 4] = cos(psi); Rot[i + 4, i] = -sin(psi); Rot[i, i + 4] = sin(psi)
        Rd[i, i] = -sin(psi); Rd[i + 4, i + 4] = -sin(psi); Rd[i + 4, i] = -cos(psi); Rd[i, i + 4] = cos(psi)
    return Rot, Rd


def predicted_curve(x, A, col):  # GetPredictedMeasurement :1162-1199
    t8 = np.zeros((8, 1))
    t8[:4] = mm(A, x[col:col + 4].reshape(4, 1)) - x[0]
    t8[4:] = mm(A, x[col + 4:col + 8].reshape(4, 1)) - x[1]
    Rot, _ = rot8(x[5])
    return mm(Rot, t8).ravel()


def wrap(x):
    for i in (3, 4, 5):
        while x[i] > PI:
            x[i] -= 2 * PI
        while x[i] < -PI:
            x[i] += 2 * PI


def tail(x, P, H, R, z, zhat):  # :863-926
    Ht = cv2.transpose(H)
    temp = mm(P, Ht)
    S = mm(H, temp)
    S = cv2.add(S, R)
    _, Sinv = cv2.invert(S, flags=cv2.DECOMP_SVD)
    K = mm(Ht, Sinv)
    K = mm(P, K)
    innov = (z - zhat).reshape(-1, 1)
    delx = mm(K, innov)
    x = x + delx.ravel()
    wrap(x)
    Kt = cv2.transpose(K)
    Kt2 = mm(S, Kt)
    delP = mm(K, Kt2)
    P = cv2.subtract(P, delP)
    Pt = cv2.transpose(P)
    P = cv2.addWeighted(P, 0.5, Pt, 0.5, 0.0)
    return x, P, S, K


def point_rows(x, N, H, zhat, row0, point_col):
    R_be = rbe(x[3], x[4], x[5])
    Rd = rbe_derivs(x[3], x[4], x[5])
    for k, c in enumerate(point_col):
        xe = (x[c:c + 3] - x[0:3]).reshape(3, 1)
        xb = mm(R_be, xe).ravel()
        zhat[row0 + 3 * k: row0 + 3 * k + 3] = xb
        for i in range(3):
            for j in range(3):
                H[row0 + 3 * k + i, j] = -R_be[i, j]
                H[row0 + 3 * k + i, c + j] = R_be[i, j]
        for j in range(3):
            t3 = mm(Rd[j], xe).ravel()
            for i in range(3):
                H[row0 + 3 * k + i, 3 + j] = t3[i]


def update_curves_points(x, P, z, A, curve_col, point_meas, point_col):
    n, n_pts = len(curve_col), len(point_col)
    N = P.shape[0]
    cms, M = 8 * n + 3, 8 * n + 3 + 3 * n_pts
    zz = np.concatenate([z, point_meas])
    H = np.zeros((M, N))
    R = np.zeros((M, M))
    for i in range(cms):
        R[i, i] = MEAS_COV ** 2
    R[8 * n, 8 * n] = Z_MEAS_COV ** 2
    R[8 * n + 1, 8 * n + 1] = PHI_MEAS_COV ** 2
    R[8 * n + 2, 8 * n + 2] = THETA_MEAS_COV ** 2
    for i in range(cms, M):
        R[i, i] = PT_MEAS_COV ** 2
    Tx, Ty, psi = x[0], x[1], x[5]
    Rot, Rotderiv = rot8(psi)
    for k in range(n):
        c = curve_col[k]
        temp8 = np.zeros((8, 8))
        temp8[:4, :4] = A[k]
        temp8[4:, 4:] = A[k]
        t81 = x[c:c + 8].reshape(8, 1).copy()
        t81 = mm(temp8, t81)
        t81[:4] -= Tx
        t81[4:] -= Ty
        t81 = mm(Rotderiv, t81)
        temp8 = mm(Rot, temp8)
        H[8 * k:8 * k + 8, c:c + 8] = temp8
        H[8 * k:8 * k + 4, 0] = -cos(psi); H[8 * k:8 * k + 4, 1] = -sin(psi)
        H[8 * k + 4:8 * k + 8, 0] = sin(psi); H[8 * k + 4:8 * k + 8, 1] = -cos(psi)
        H[8 * k:8 * k + 8, 5] = t81.ravel()
    H[8 * n, 2] = 1.0; H[8 * n + 1, 3] = 1.0; H[8 * n + 2, 4] = 1.0
    zhat = np.zeros(M)
    point_rows(x, N, H, zhat, cms, point_col)
    for k in range(n):
        zhat[8 * k:8 * k + 8] = predicted_curve(x, A[k], curve_col[k])
    zhat[8 * n:8 * n + 3] = x[2:5]
    xn, Pn, S, K = tail(x.copy(), P.copy(), H, R, zz, zhat)
    return xn, Pn, H, S, K


def update_points(x, P, point_meas, point_col):
    N, M = P.shape[0], 3 * len(point_col)
    H, R, zhat = np.zeros((M, N)), np.eye(M) * PT_MEAS_COV ** 2, np.zeros(M)
    point_rows(x, N, H, zhat, 0, point_col)
    xn, Pn, _, _ = tail(x.copy(), P.copy(), H, R, point_meas, zhat)
    return xn, Pn


def predict(x, P):  # :161-272
    N = P.shape[0]
    phi, theta, psi = x[3], x[4], x[5]
    Reb = rbe(phi, theta, psi).T.copy()  # Reb->data.db[k] of :171-179 is the transpose layout of generate_Rbe
    Reb = np.array([[cos(psi) * cos(theta), cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi),
                     cos(psi) * sin(theta) * cos(phi) - sin(psi) * sin(phi)],
                    [sin(psi) * cos(theta), sin(psi) * sin(theta) * sin(phi) + cos(psi) * cos(phi),
                     sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi)],
                    [-sin(theta), cos(theta) * sin(phi), cos(theta) * cos(phi)]])
    Q = np.zeros((12, 12))
    t33 = np.diag([v ** 2 for v in V_COV])
    t33 = mm(mm(Reb, t33), cv2.transpose(Reb))
    for i in range(3):
        for j in range(3):
            Q[i, j] = DT ** 3 * t33[i, j]; Q[i, j + 6] = DT ** 2 * t33[i, j]
            Q[j + 6, i] = DT ** 2 * t33[i, j]; Q[i + 6, j + 6] = DT * t33[i, j]
    t33 = np.diag([v ** 2 for v in W_COV])
    for i in range(3):
        for j in range(3):
            Q[i + 3, j + 3] = DT ** 3 * t33[i, j]; Q[i + 3, j + 9] = DT ** 2 * t33[i, j]
            Q[j + 9, i + 3] = DT ** 2 * t33[i, j]; Q[i + 9, j + 9] = DT * t33[i, j]
    F = np.eye(12)
    for i in range(6):
        F[i, i + 6] = DT
    Prr, Pri = P[:12, :12].copy(), P[:12, 12:].copy()
    Pri = mm(F, Pri) if N > 12 else Pri
    Prr = mm(mm(F, Prr), cv2.transpose(F)) + Q
    P = P.copy()
    P[:12, :12] = Prr
    P[:12, 12:] = Pri
    P[12:, :12] = Pri.T
    x = x.copy()
    x[0:6] += x[6:12] * DT
    return x, P


def split_matrices(t):  # GetSplitMatrices :1144-1159
    from math import comb
    A1, A2 = np.zeros((4, 4)), np.zeros((4, 4))
    for i in range(4):
        for j in range(i + 1):
            A1[i, j] = comb(i, j) * t ** j * (1 - t) ** (i - j)
        for j in range(i, 4):
            A2[i, j] = comb(3 - i, j - i) * t ** (j - i) * (1 - t) ** (3 - j)
    return A1, A2


# ---- state augmentation: AddFirstStates (:275-311), AddNewCurve (:314-499), AddNewPoints (:502-666) ----------
def reb(phi, theta, psi):  # generate_Reb, common.h:189-201
    return np.array([[cos(psi) * cos(theta), cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi),
                      cos(psi) * sin(theta) * cos(phi) - sin(psi) * sin(phi)],
                     [sin(psi) * cos(theta), sin(psi) * sin(theta) * sin(phi) + cos(psi) * cos(phi),
                      sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi)],
                     [-sin(theta), cos(theta) * sin(phi), cos(theta) * cos(phi)]])


def reb_derivs(phi, theta, psi):  # get_Reb_derivs, kalmanClass.cpp:1230-1262
    p = np.array([[0.0, cos(psi) * sin(theta) * cos(phi) + sin(psi) * sin(phi), -cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi)],
                  [0.0, sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi), -sin(psi) * sin(theta) * sin(phi) - cos(psi) * cos(phi)],
                  [0.0, cos(theta) * cos(phi), -cos(theta) * sin(phi)]])
    t = np.array([[-cos(psi) * sin(theta), cos(psi) * cos(theta) * sin(phi), cos(psi) * cos(theta) * cos(phi)],
                  [-sin(psi) * sin(theta), sin(psi) * cos(theta) * sin(phi), sin(psi) * cos(theta) * cos(phi)],
                  [-cos(theta), -sin(theta) * sin(phi), -sin(theta) * cos(phi)]])
    s_ = np.array([[-sin(psi) * cos(theta), -sin(psi) * sin(theta) * sin(phi) - cos(psi) * cos(phi), -sin(psi) * sin(theta) * cos(phi) - cos(psi) * sin(phi)],
                   [cos(psi) * cos(theta), cos(psi) * sin(theta) * sin(phi) - sin(psi) * cos(phi), cos(psi) * sin(theta) * cos(phi) + sin(psi) * sin(phi)],
                   [0.0, 0.0, 0.0]])
    return p, t, s_


def add_first_states(z19):
    P, x = np.zeros((28, 28)), np.zeros(28)
    for i in range(16):
        P[i + 12, i + 12] = MEAS_COV ** 2
        x[i + 12] = z19[i]
    x[2], x[3], x[4], x[6] = z19[16], z19[17], z19[18], 1.0
    for i, v in enumerate([1e-7, 1e-7, 0.1, 0.1, 0.1, 1e-7, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1]):
        P[i, i] = v
    return x, P


def add_new_curve(x, P, z8, A):
    N = P.shape[0]
    Tx, Ty, psi = x[0], x[1], x[5]
    Rot, Rotd = np.zeros((8, 8)), np.zeros((8, 8))
    for i in range(4):
        Rot[i, i] = cos(psi); Rot[i + 4, i + 4] = cos(psi); Rot[i + 4, i] = sin(psi); Rot[i, i + 4] = -sin(psi)
        Rotd[i, i] = -sin(psi); Rotd[i + 4, i + 4] = -sin(psi); Rotd[i + 4, i] = cos(psi); Rotd[i, i + 4] = -cos(psi)
    Prr, Pri = P[:12, :12].copy(), P[:12, 12:].copy()
    Pn, xn = np.zeros((N + 8, N + 8)), np.zeros(N + 8)
    Pn[:N, :N] = P
    xn[:N] = x
    _, Ainv = cv2.invert(np.ascontiguousarray(A), flags=cv2.DECOMP_SVD)
    Hinv = np.zeros((8, 8))
    Hinv[:4, :4] = Ainv
    Hinv[4:, 4:] = Ainv
    t81 = mm(Rot, z8.reshape(8, 1))
    t81[:4] += Tx
    t81[4:] += Ty
    xcur = mm(Hinv, t81)
    a1 = mm(Ainv, np.ones((4, 1)))
    Gz = mm(Hinv, Rot)
    t81 = mm(Hinv, mm(Rotd, z8.reshape(8, 1)))
    Gx = np.zeros((8, 12))
    for i in range(4):
        Gx[i, 0] = a1[i, 0]; Gx[i + 4, 1] = a1[i, 0]; Gx[i, 5] = t81[i, 0]; Gx[i + 4, 5] = t81[i + 4, 0]
    Pcur = mm(Gx, mm(Prr, cv2.transpose(Gx)))
    R1 = np.eye(8) * MEAS_COV
    PN1N1 = cv2.add(Pcur, mm(Gz, mm(R1, cv2.transpose(Gz))))
    PN1i = mm(Gx, Pri) if N > 12 else np.zeros((8, 0))
    PN1r = mm(Gx, cv2.transpose(Prr))
    xn[N:] = xcur.ravel()
    Pn[N:, N:] = PN1N1
    Pn[N:, 12:N] = PN1i
    Pn[12:N, N:] = PN1i.T
    Pn[N:, :12] = PN1r
    Pn[:12, N:] = PN1r.T
    return xn, Pn


def add_new_points(x, P, meas):
    N, n_pts = P.shape[0], len(meas) // 3
    Pn, xn = np.zeros((N + 3 * n_pts, N + 3 * n_pts)), np.zeros(N + 3 * n_pts)
    Pn[:N, :N] = P
    xn[:N] = x
    R_eb = reb(x[3], x[4], x[5])
    Rd = reb_derivs(x[3], x[4], x[5])
    Prr, Pri = P[:12, :12].copy(), P[:12, 12:].copy()
    for n in range(n_pts):
        xb = meas[3 * n:3 * n + 3].reshape(3, 1)
        o = N + 3 * n
        xn[o:o + 3] = mm(R_eb, xb).ravel() + xn[0:3]
        Gx = np.zeros((3, 12))
        Gx[0, 0] = Gx[1, 1] = Gx[2, 2] = 1.0
        for j in range(3):
            Gx[:, j + 3] = mm(Rd[j], xb).ravel()
        PN1N1 = mm(Gx, mm(Prr, cv2.transpose(Gx)))
        PN1N1 = cv2.add(PN1N1, mm(R_eb, mm(np.eye(3) * PT_MEAS_COV, cv2.transpose(R_eb))))
        PN1i = mm(Gx, Pri)
        PN1r = mm(Gx, Prr)
        Pn[o:o + 3, o:o + 3] = PN1N1
        Pn[o:o + 3, 12:N] = PN1i
        Pn[12:N, o:o + 3] = PN1i.T
        Pn[o:o + 3, :12] = PN1r
        Pn[:12, o:o + 3] = PN1r.T
    return xn, Pn


def augmentation_case():
    rng = np.random.default_rng(11)
    z19 = np.concatenate([np.linspace(2.0, 5.0, 4), -1.2 + rng.normal(0, 0.05, 4), np.linspace(2.1, 5.2, 4),
                          1.2 + rng.normal(0, 0.05, 4), [-1.0, 0.01, -0.02]])
    x0, P0 = add_first_states(z19)
    x0p, P0p = predict(x0, P0)
    x0p[5] = 0.2  # give the heading something to rotate by
    A1, A2 = split_matrices(0.4)
    z8 = np.concatenate([np.linspace(3.0, 6.0, 4), -1.1 + rng.normal(0, 0.05, 4)])
    x1, P1 = add_new_curve(x0p, P0p, z8, A2)
    pts = rng.uniform(-5, 5, 9)
    x2, P2 = add_new_points(x1, P1, pts)
    return dict(z19=z19, x0=x0, P0=P0, x_in=x0p, P_in=P0p, z8=z8, A=A2, x1=x1, P1=P1, pts=pts, x2=x2, P2=P2)


def make_case(seed, n_curves_state, n_points_state, meas_curves, meas_points, split_t):
    rng = np.random.default_rng(seed)
    N = 12 + 8 * n_curves_state + 3 * n_points_state
    G = rng.standard_normal((N, 16))
    P = G @ G.T / 16.0 + 0.25 * np.eye(N)
    P = 0.5 * (P + P.T)
    x = np.zeros(N)
    x[:12] = (0.3, -0.2, -1.0, 0.01, -0.02, 0.3, 1.0, 0.05, 0.0, 0.01, 0.0, 0.02)
    for c in range(n_curves_state):
        base = 12 + 8 * c
        x[base:base + 4] = np.linspace(2.0, 5.0, 4) + rng.normal(0, 0.1, 4)
        x[base + 4:base + 8] = (-1.2 if c % 2 == 0 else 1.2) + rng.normal(0, 0.05, 4)
    for p in range(n_points_state):
        base = 12 + 8 * n_curves_state + 3 * p
        x[base:base + 3] = rng.uniform(-10, 10, 3)
    curve_col = np.array([12 + 8 * c for c in meas_curves], dtype=np.int32)
    point_col = np.array([12 + 8 * n_curves_state + 3 * p for p in meas_points], dtype=np.int32)
    A = np.stack([split_matrices(t)[k % 2] for k, t in enumerate(split_t)]) if len(meas_curves) else np.zeros((0, 4, 4))
    n = len(meas_curves)
    z = np.zeros(8 * n + 3)
    for k in range(n):
        z[8 * k:8 * k + 8] = predicted_curve(x, A[k], curve_col[k]) + rng.normal(0, 0.5, 8)
    z[8 * n:] = x[2:5] + rng.normal(0, 0.1, 3)
    R_be = rbe(x[3], x[4], x[5])
    pm = np.concatenate([R_be @ (x[c:c + 3] - x[:3]) + rng.normal(0, 2.0, 3) for c in point_col]) if len(point_col) else np.zeros(0)
    return dict(x=x, P=P, z=z, A=A, curve_col=curve_col, point_meas=pm, point_col=point_col)


def main():
    cases = {
        "ekf_curves4": make_case(1, 6, 0, [0, 1, 4, 5], [], [0.3, 0.3, 0.6, 0.6]),
        "ekf_curves2_points3": make_case(2, 3, 5, [2, 0], [4, 0, 2], [0.5, 0.25]),
        "ekf_points_only": make_case(3, 2, 6, [], [1, 5, 3, 0], []),
    }
    for name, c in cases.items():
        out = dict(c)
        if name == "ekf_points_only":
            xn, Pn = update_points(c["x"], c["P"], c["point_meas"], c["point_col"])
            out.update(x_upd=xn, P_upd=Pn)
        else:
            xn, Pn, H, S, K = update_curves_points(c["x"], c["P"], c["z"], c["A"], c["curve_col"], c["point_meas"],
                                                   c["point_col"])
            out.update(x_upd=xn, P_upd=Pn, H=H, S=S, K=K)
        xp, Pp = predict(c["x"], c["P"])
        out.update(x_pred=xp, P_pred=Pp)
        A1, A2 = split_matrices(0.37)
        out.update(A1=A1, A2=A2, split_t=np.array(0.37))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "N", c["P"].shape[0], "written")
    np.savez_compressed(os.path.join(HERE, "ekf_augment.npz"), **augmentation_case())
    print("ekf_augment written")


if __name__ == "__main__":
    main()
