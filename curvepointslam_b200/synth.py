"""Synthetic stereo contours for the batched curve fits (host side, numpy only).

Follows the reference author's own recipe, ``generateMeasurements`` (curveFittingClass.cpp:1086-1166:
Bezier points generated from a state through ``cp_earth2image`` plus Gaussian pixel noise with
``PIX_STDEV`` 0.75, curveFittingClass.h:18), with a counter-based generator (numpy Philox keyed by the
seed) instead of libc ``rand()`` so that every rank and every run sees the same bytes.  Workload
definition: SURVEY.md section 8(d).

Nothing here is on the measured path: it only manufactures inputs.
"""
from __future__ import annotations

import numpy as np

# camera constants, common.h:70-87,102 (same arithmetic so the doubles are the same)
FX = (521.22 / 2.0 + 527.03 / 2.0) / 2.0
FY = (518.32 / 2.0 + 524.73 / 2.0) / 2.0
CX = (296.08 / 2.0 - 1.0 + 320.95 / 2.0 + 1.0) / 2.0
CY = (236.17 / 2.0 + 228.81 / 2.0) / 2.0
BASELINE = 0.535
PIX_STDEV = 0.75

# the one control-point set the reference mentions (curveFittingClass.cpp:13)
TRUTH_C1 = np.array([2.31, -1.20, 3.03, -1.17, 3.76, -1.14, 4.48, -1.11])


def _rng(seed: int, stream: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[int(seed) & (2**64 - 1), int(stream)]))


def bernstein3(t: np.ndarray) -> np.ndarray:
    u = 1.0 - t
    return np.stack([u * u * u, 3.0 * u * u * t, 3.0 * u * t * t, t * t * t], axis=-1)


def project_control_points(p19: np.ndarray) -> np.ndarray:
    """p19: [B,19] -> image control points [B, 2 curves, 4, 3] holding (u_left, u_right, v)."""
    p19 = np.asarray(p19, dtype=np.float64)
    B = p19.shape[0]
    cp = p19[:, :16].reshape(B, 2, 4, 2)
    th, ph, z = p19[:, 16, None, None], p19[:, 17, None, None], p19[:, 18, None, None]
    st, ct, sp, cph = np.sin(th), np.cos(th), np.sin(ph), np.cos(ph)
    xe, ye = cp[..., 0], cp[..., 1]
    xb0 = ct * xe + st * z
    xb1 = st * sp * xe + cph * ye - ct * sp * z
    xb2 = st * cph * xe - sp * ye - ct * cph * z
    ul = FX * (xb1 + 0.5 * BASELINE) / xb0 + CX
    ur = FX * (xb1 - 0.5 * BASELINE) / xb0 + CX
    v = FY * (xb2 / xb0) + CY
    return np.stack([ul, ur, v], axis=-1)


def segment_t(xy: np.ndarray) -> np.ndarray:
    """t-parameterisation of fit_curve (curveFittingClass.cpp:451-511) for dense [..., K, 2] segments."""
    y = xy[..., 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (y - y[..., :1]) / (y[..., -1:] - y[..., :1])
    t = np.where(t < 0.0, 0.0, t)
    t = np.where(t > 1.0, 1.0, t)
    return t


def make_stereo19_batch(B: int, K: int = 64, seed: int = 0, first: int = 0, noise: float = PIX_STDEV,
                        round_pixels: bool = False, counts=None):
    """B stereo-pair fits (m=19), fit indices first..first+B-1 of the global sequence `seed`.

    Returns dict with
      meas   float64 [sum n]  concatenated (x,y) samples per fit, segment order [L c1|L c2|R c1|R c2]
      off    int64   [B+1]    start of each fit in `meas` (in doubles)
      counts int32   [B,4]    points per segment
      p0     float64 [B,19]   initial parameters
      truth  float64 [B,19]   generating parameters
    `counts` (optional [B,4] or [4]) makes the batch ragged; default is K points in every segment.
    """
    if counts is None:
        counts = np.full((B, 4), K, dtype=np.int32)
    else:
        counts = np.broadcast_to(np.asarray(counts, dtype=np.int32), (B, 4)).copy()
    Kmax = int(counts.max()) if B else 0
    truth = np.empty((B, 19))
    p0 = np.empty((B, 19))
    xy = np.empty((B, 4, max(Kmax, 1), 2))
    # one independent stream per fit index => any shard of the sequence is reproducible on its own
    base = np.concatenate([TRUTH_C1, TRUTH_C1 * np.tile([1.0, -1.0], 4)])
    for b in range(B):
        g = _rng(seed, first + b)
        tr = np.empty(19)
        tr[:16] = base + g.normal(0.0, 0.05, 16)
        tr[16] = g.uniform(-0.03, 0.03)
        tr[17] = g.uniform(-0.03, 0.03)
        tr[18] = -1.0 + g.normal(0.0, 0.02)
        truth[b] = tr
        p0[b, :16] = tr[:16] + g.normal(0.0, 0.1, 16)
        p0[b, 16:] = (0.0, 0.0, -1.02)
        xy[b] = g.normal(0.0, 1.0, (4, max(Kmax, 1), 2))
    img = project_control_points(truth)  # [B,2,4,3]
    seg_curve = np.array([0, 1, 0, 1])
    seg_side = np.array([0, 0, 1, 1])
    pieces, off = [], np.zeros(B + 1, dtype=np.int64)
    for b in range(B):
        rows = []
        for s in range(4):
            k = int(counts[b, s])
            if k == 0:
                continue
            t = np.linspace(0.0, 1.0, k) if k > 1 else np.zeros(1)
            w = bernstein3(t)
            cpi = img[b, seg_curve[s]]
            pts = np.stack([w @ cpi[:, seg_side[s]], w @ cpi[:, 2]], axis=-1)
            pts = pts + noise * xy[b, s, :k]
            if round_pixels:
                pts = np.rint(pts)
            rows.append(pts.reshape(-1))
        flat = np.concatenate(rows) if rows else np.zeros(0)
        pieces.append(flat)
        off[b + 1] = off[b] + flat.size
    meas = np.concatenate(pieces) if pieces else np.zeros(0)
    return {"meas": meas, "off": off, "counts": counts, "p0": p0, "truth": truth}


def make_stereo19_batch_fast(B: int, K: int = 64, seed: int = 0, first: int = 0, noise: float = PIX_STDEV):
    """Vectorised dense variant for benchmark-size batches (one generator for the whole block; the
    block is keyed by (seed, first) so a rank's shard is reproducible)."""
    g = _rng(seed, (1 << 40) + first)
    base = np.concatenate([TRUTH_C1, TRUTH_C1 * np.tile([1.0, -1.0], 4)])
    truth = np.empty((B, 19))
    truth[:, :16] = base + g.normal(0.0, 0.05, (B, 16))
    truth[:, 16] = g.uniform(-0.03, 0.03, B)
    truth[:, 17] = g.uniform(-0.03, 0.03, B)
    truth[:, 18] = -1.0 + g.normal(0.0, 0.02, B)
    p0 = np.empty((B, 19))
    p0[:, :16] = truth[:, :16] + g.normal(0.0, 0.1, (B, 16))
    p0[:, 16:] = (0.0, 0.0, -1.02)
    img = project_control_points(truth)
    w = bernstein3(np.linspace(0.0, 1.0, K))  # [K,4]
    meas = np.empty((B, 4, K, 2))
    for s, (c, side) in enumerate([(0, 0), (1, 0), (0, 1), (1, 1)]):
        meas[:, s, :, 0] = img[:, c, :, side] @ w.T
        meas[:, s, :, 1] = img[:, c, :, 2] @ w.T
    meas += noise * g.standard_normal(meas.shape)
    counts = np.full((B, 4), K, dtype=np.int32)
    off = np.arange(B + 1, dtype=np.int64) * (8 * K)
    return {"meas": meas.reshape(-1), "off": off, "counts": counts, "p0": p0, "truth": truth}


def make_image8_batch(B: int, K: int = 64, seed: int = 0, first: int = 0, noise: float = PIX_STDEV):
    """B per-contour image-space fits (m=8): one segment of K points each; initial guess = straight line
    between the first and last sample (what fit_curve:626-632 does in 3-D)."""
    g = _rng(seed, (2 << 40) + first)
    cp = np.empty((B, 4, 2))
    cp[:, :, 0] = 160.0 + g.normal(0.0, 40.0, (B, 1)) + np.cumsum(g.normal(0.0, 8.0, (B, 4)), axis=1)
    cp[:, :, 1] = np.linspace(230.0, 120.0, 4)[None, :] + g.normal(0.0, 4.0, (B, 4))
    w = bernstein3(np.linspace(0.0, 1.0, K))
    meas = np.einsum("kc,bcd->bkd", w, cp) + noise * g.standard_normal((B, K, 2))
    p0 = np.empty((B, 4, 2))
    for k in range(4):
        p0[:, k] = ((3 - k) * meas[:, 0] + k * meas[:, -1]) / 3.0
    counts = np.full((B, 1), K, dtype=np.int32)
    off = np.arange(B + 1, dtype=np.int64) * (2 * K)
    return {"meas": meas.reshape(-1), "off": off, "counts": counts, "p0": p0.reshape(B, 8),
            "truth": cp.reshape(B, 8)}


def t_from_rows(batch) -> np.ndarray:
    """Concatenated t per sample, computed from the image rows exactly like fit_curve does."""
    meas, off, counts = batch["meas"], batch["off"], batch["counts"]
    out = np.empty(meas.size // 2)
    for b in range(len(off) - 1):
        xy = meas[off[b]:off[b + 1]].reshape(-1, 2)
        k0 = 0
        for c in counts[b]:
            c = int(c)
            if c:
                out[off[b] // 2 + k0: off[b] // 2 + k0 + c] = segment_t(xy[k0:k0 + c])
            k0 += c
    return out


def make_edge_frames(F: int, seed: int = 0, consistent: bool = False):
    """Integer edge sequences like the reference's detector hands to cleanup_and_group_edges (featureDetector.cpp:
    213-288: one (x, y) pixel per image row, bottom-to-top): F frames x [L0, L1, R0, R1].
    consistent=False: image-space sequences that start and end on slightly different rows, some longer than 100 rows,
    some too short -- exercises every trimming loop and the failure returns.
    consistent=True: the projected synthetic ground curves sampled one point per row over a common row range, so that
    every frame survives the trimming and can be fitted.
    Returns pts [P,2] int32, seg_off [4F+1] int64, tracked_endpt_y [F,2]."""
    g = _rng(seed, (3 << 40) + (1 if consistent else 0))
    pts, seg_off, tracked = [], [0], np.zeros((F, 2))
    base = np.concatenate([TRUTH_C1, TRUTH_C1 * np.tile([1.0, -1.0], 4)])
    for f in range(F):
        if consistent:
            tr = np.empty(19)
            tr[:16] = base + g.normal(0.0, 0.05, 16)
            tr[16:] = (g.uniform(-0.03, 0.03), g.uniform(-0.03, 0.03), -1.0 + g.normal(0.0, 0.02))
            img = project_control_points(tr[None])[0]  # [2 curves, 4, (uL, uR, v)]
            w = bernstein3(np.linspace(0.0, 1.0, 400))
            vs = [w @ img[c][:, 2] for c in (0, 1)]
            lo = int(np.ceil(max(v.min() for v in vs))) + 1
            hi = int(np.floor(min(v.max() for v in vs))) - 1
            rows = np.arange(hi, lo - 1, -1)
            for side in (0, 1):
                for c in (0, 1):
                    u = w @ img[c][:, side]
                    order = np.argsort(vs[c])
                    xs = np.rint(np.interp(rows, vs[c][order], u[order])).astype(np.int32)
                    pts.append(np.stack([xs, rows.astype(np.int32)], axis=1))
                    seg_off.append(seg_off[-1] + rows.size)
            tracked[f] = (lo, lo)
        else:
            y0 = 228 + g.integers(-4, 5, 4)
            length = g.integers(90, 170, 4)
            if g.random() < 0.12:
                length[g.integers(0, 4)] = g.integers(3, 12)
            for s in range(4):
                rows = np.arange(y0[s], y0[s] - length[s], -1)
                xs = np.rint(60 + 40 * (s % 2) + 0.4 * (228 - rows) + g.normal(0.0, 0.6, rows.size)).astype(np.int32)
                pts.append(np.stack([xs, rows.astype(np.int32)], axis=1))
                seg_off.append(seg_off[-1] + rows.size)
            tracked[f] = 228 - 100 + g.integers(-6, 30, 2)
    return np.concatenate(pts).astype(np.int32), np.array(seg_off, dtype=np.int64), tracked
