#!/usr/bin/env python
"""Generates tests/golden/lm_init_guess.npz: a numpy/cv2 transcription of fit_curve's initial guess
(curveFittingClass.cpp:516-651: triangulation, centroid, cvSVD of the centred 4x3 matrix, theta/phi/d, rotation into
the ground frame, straight-line control points) on a few synthetic frames.  cv2.SVDecomp stands in for the legacy
cvSVD; the SIGN of its last right singular vector is whatever this OpenCV build produces (the reference's is equally
implementation-defined), so the fixture stores the raw result and the tests accept either orientation.

    python tests/golden/gen_init_golden.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from curvepointslam_b200 import synth  # noqa: E402
from curvepointslam_b200.synth import BASELINE, CX, CY, FX, FY  # noqa: E402


def tri(m, a, b):
    X = FX * BASELINE / (m[2 * a] - m[2 * b])
    Y = 0.5 * (X / FX * (m[2 * a] - CX) - 0.5 * BASELINE) + 0.5 * (X / FX * (m[2 * b] - CX) + 0.5 * BASELINE)
    Z = 0.5 * X / FY * (m[2 * a + 1] - CY) + 0.5 * X / FY * (m[2 * b + 1] - CY)
    return np.array([X, Y, Z])


def init_guess(m, counts):
    nl1, nl2, nr1, nr2 = (int(c) for c in counts)
    npl, npts = nl1 + nl2, nl1 + nl2 + nr1 + nr2
    s1, e1 = tri(m, 0, npl), tri(m, nl1 - 1, npl + nr1 - 1)
    s2, e2 = tri(m, nl1, npl + nr1), tri(m, npl - 1, npts - 1)
    c = 0.25 * (s1 + s2 + e1 + e2)
    M = np.stack([s1 - c, s2 - c, e1 - c, e2 - c])
    _, _, vt = cv2.SVDecomp(M)
    normal = vt[2, :].copy()  # V->data.db[2,5,8]: last column of V
    d = float(normal @ c)
    theta, phi = -np.arcsin(normal[0]), np.arcsin(normal[1])
    origin = d * normal
    Rot = np.array([[np.cos(theta), np.sin(theta) * np.sin(phi), np.sin(theta) * np.cos(phi)],
                    [0.0, np.cos(phi), -np.sin(phi)],
                    [-np.sin(theta), np.cos(theta) * np.sin(phi), np.cos(theta) * np.cos(phi)]])
    s1, e1, s2, e2 = (Rot @ (p - origin) for p in (s1, e1, s2, e2))
    cp1, cp2 = np.zeros(8), np.zeros(8)
    cp1[0:2], cp1[6:8], cp2[0:2], cp2[6:8] = s1[:2], e1[:2], s2[:2], e2[:2]
    for i in (1, 2):
        cp1[2 * i:2 * i + 2] = (3 - i) * cp1[0:2] / 3 + i * cp1[6:8] / 3
        cp2[2 * i:2 * i + 2] = (3 - i) * cp2[0:2] / 3 + i * cp2[6:8] / 3
    return np.concatenate([cp1, cp2, [theta, phi, -d]])


def main():
    rng = np.random.default_rng(3)
    counts = rng.integers(8, 102, size=(12, 4)).astype(np.int32)
    counts[0] = (64, 64, 64, 64)
    b = synth.make_stereo19_batch(12, seed=55, counts=counts, round_pixels=True)
    p = np.stack([init_guess(b["meas"][b["off"][f]:b["off"][f + 1]], b["counts"][f]) for f in range(12)])
    np.savez_compressed(os.path.join(HERE, "lm_init_guess.npz"), meas=b["meas"], off=b["off"], counts=b["counts"], p=p)
    print("written; height signs:", np.sign(p[:, 18]))


if __name__ == "__main__":
    main()
