"""CPU tests: the LM oracle against (a) the reference's own levmar-2.5 built in oracle/_ref, bit for bit,
(b) the lmdemo known-answer table (levmar-2.5/lmdemo.c), (c) committed golden vectors produced
by the reference levmar (tests/golden/gen_lm_golden.py), (d) the reference's Jacobian checker dlevmar_chkjac
on the derived analytic Jacobian."""
import ctypes as C
import os

import numpy as np
import pytest

from curvepointslam_b200 import lm, synth
from oracle_lib import DER, DIF, IMAGE8, MATH_DET, MATH_LIBM, STEREO19, oracle

HERE = os.path.dirname(os.path.abspath(__file__))
OPTS = [1e-10, 1e-10, 1e-10, 1e-20, 1e-5]
dp = C.POINTER(C.c_double)
need_ref = pytest.mark.skipif(oracle().ref is None, reason="oracle/_ref not built (needs /root/reference)")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


# problem -> (return, reason, nfev, njev, nlss, ||e||^2, solution)
KAT = {
    0: (1000, 3, 1290, 1000, 1289, 2.08909e-05, [0.9432309, 0.8893884]),
    1: (14, 2, 25, 14, 25, 10000.0, [0.9999992, 0.9999984]),
    2: (198, 6, 208, 198, 207, 8.87028e-21, [-9.00243e-11, -6.719414e-05]),
    3: (113, 6, 158, 11, 113, 2.73165e-24, [1, 1, 1, 1]),
    4: (209, 2, 273, 21, 210, 8.79459e-05, [2.481778, 6.181346, 3.502236]),
    5: (34, 2, 45, 34, 45, 5.46489e-05, [0.3754101, 1.935847, -1.464687, 0.01286753, 0.0221227]),
    6: (9, 6, 10, 9, 9, 3.43421e-25, [1, 3.69e-13, 5.86e-13]),
}


def run_kat(prob, use_ref):
    o = oracle()
    m, n, var = C.c_int(), C.c_int(), C.c_int()
    p, x = np.zeros(8), np.zeros(40)
    f, j = C.c_void_p(), C.c_void_p()
    assert o.lib.orc_kat_problem(prob, C.byref(m), C.byref(n), p.ctypes.data_as(dp), x.ctypes.data_as(dp),
                                 C.byref(var), C.byref(f), C.byref(j)) == 0
    opts = np.array([1e-3, 1e-15, 1e-15, 1e-20, 1e-6])  # lmdemo.c:816-817
    info = np.zeros(10)
    a = (p.ctypes.data_as(dp), x.ctypes.data_as(dp), m.value, n.value, 1000, opts.ctypes.data_as(dp),
         info.ctypes.data_as(dp))
    if use_ref:
        ret = o.levmar.dlevmar_der(f, j, *a, None, None, None) if var.value == 0 else \
            o.levmar.dlevmar_dif(f, *a, None, None, None)
    else:
        ret = o.lib.orc_lm_der(f, j, *a, None, None) if var.value == 0 else o.lib.orc_lm_dif(f, *a, None, None)
    return ret, p[:m.value].copy(), info


@pytest.mark.parametrize("prob", sorted(KAT))
def test_lmdemo_known_answers(prob):
    ret, p, info = run_kat(prob, use_ref=False)
    eret, reason, nfev, njev, nlss, e2, sol = KAT[prob]
    assert (ret, int(info[6]), int(info[7]), int(info[8]), int(info[9])) == (eret, reason, nfev, njev, nlss)
    assert info[1] == pytest.approx(e2, rel=2e-5)
    assert np.allclose(p, sol, rtol=2e-6, atol=1e-12)


@need_ref
@pytest.mark.parametrize("prob", sorted(KAT))
def test_lmdemo_bit_exact_vs_reference_levmar(prob):
    ra, pa, ia = run_kat(prob, use_ref=False)
    rb, pb, ib = run_kat(prob, use_ref=True)
    assert ra == rb and np.array_equal(bits(pa), bits(pb)) and np.array_equal(bits(ia), bits(ib))


@need_ref
@pytest.mark.parametrize("model,variant", [(STEREO19, DER), (STEREO19, DIF), (IMAGE8, DER), (IMAGE8, DIF)])
@pytest.mark.parametrize("mode", [MATH_DET, MATH_LIBM])
def test_batched_fits_bit_exact_vs_reference_levmar(model, variant, mode):
    if model == STEREO19:
        rng = np.random.default_rng(3)
        counts = rng.integers(4, 102, size=(20, 4)).astype(np.int32)
        counts[0] = (2, 3, 3, 2)
        b = synth.make_stereo19_batch(20, seed=21, counts=counts)
    else:
        b = synth.make_image8_batch(20, K=64, seed=22)
    a = oracle().fit_batch(model, variant, b["p0"], b["meas"], b["off"], b["counts"], OPTS, math_mode=mode, nthreads=4)
    r = oracle().fit_batch(model, variant, b["p0"], b["meas"], b["off"], b["counts"], OPTS, math_mode=mode, nthreads=4,
                           use_ref=True)
    assert np.array_equal(a["ret"], r["ret"])
    assert np.array_equal(bits(a["p"]), bits(r["p"])) and np.array_equal(bits(a["info"]), bits(r["info"]))


@pytest.mark.parametrize("tag,model", [("s19", STEREO19), ("i8", IMAGE8)])
@pytest.mark.parametrize("vname,variant", [("der", DER), ("dif", DIF)])
@pytest.mark.parametrize("mname,mode", [("det", MATH_DET), ("libm", MATH_LIBM)])
def test_golden_vectors_from_reference_levmar(tag, model, vname, variant, mname, mode):
    g = np.load(os.path.join(HERE, "golden", "lm_reference_levmar.npz"))
    r = oracle().fit_batch(model, variant, g[f"{tag}_p0"], g[f"{tag}_meas"], g[f"{tag}_off"], g[f"{tag}_counts"],
                           [float(v) for v in g["opts"]], math_mode=mode, nthreads=4)
    key = f"{tag}_{vname}_{mname}"
    assert np.array_equal(r["ret"], g[key + "_ret"])
    if mode == MATH_DET:  # only +,-,*,/ : identical on any IEEE-754 host
        assert np.array_equal(bits(r["p"]), bits(g[key + "_p"]))
        assert np.array_equal(bits(r["info"]), bits(g[key + "_info"]))
    else:  # glibc sin/cos/pow may differ by an ulp between libm builds
        ok = r["ret"] >= 0
        assert np.array_equal(r["info"][:, 6], g[key + "_info"][:, 6])
        assert np.allclose(r["p"][ok], g[key + "_p"][ok], rtol=1e-5, atol=1e-7)


def test_det_vs_libm_distance_der():
    """How far the deterministic math mode (what the GPU computes) is from the reference's glibc calls: the
    north-star tolerances hold on the der path (1e-9 on parameters is met by > 90 % of fits)."""
    b = synth.make_stereo19_batch(48, K=64, seed=23)
    a = oracle().fit_batch(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS, math_mode=MATH_DET, nthreads=4)
    r = oracle().fit_batch(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS, math_mode=MATH_LIBM, nthreads=4)
    assert np.array_equal(a["info"][:, 6], r["info"][:, 6])  # same termination reason
    rel_p = np.abs(a["p"] - r["p"]).max(axis=1) / np.abs(r["p"]).max(axis=1)
    rel_e = np.abs(a["info"][:, 1] - r["info"][:, 1]) / r["info"][:, 1]
    assert np.median(rel_p) < 1e-12 and (rel_p < 1e-9).mean() >= 0.9 and rel_p.max() < 1e-7
    assert rel_e.max() < 1e-10


@need_ref
def test_analytic_jacobian_passes_reference_chkjac():
    """dlevmar_chkjac (misc_core.c:248-319): err > 0.5 means the gradient of that row is probably right."""
    o = oracle()
    b = synth.make_stereo19_batch(1, K=24, seed=24)

    class FitData(C.Structure):
        _fields_ = [("t", C.c_void_p), ("count", C.c_int * 4), ("math_mode", C.c_int)]

    t = synth.t_from_rows(b)
    fd = FitData(t.ctypes.data, (C.c_int * 4)(24, 24, 24, 24), MATH_LIBM)
    n = b["meas"].size
    err = np.zeros(n)
    p = b["p0"][0].copy()
    func = C.cast(o.lib.orc_stereo19_func, C.c_void_p)
    jacf = C.cast(o.lib.orc_stereo19_jac, C.c_void_p)
    o.levmar.dlevmar_chkjac(func, jacf, p.ctypes.data_as(dp), 19, n, C.addressof(fd), err.ctypes.data_as(dp))
    assert err.min() > 0.5, err.min()


@need_ref
def test_building_blocks_vs_reference():
    o = oracle()
    rng = np.random.default_rng(9)
    for m in (3, 8, 19):
        A = rng.standard_normal((m, m))
        A = A @ A.T + np.eye(m) * 1e-3
        if m == 8:
            A[2, :] = 0.0
            A[:, 2] = 0.0
            A[2, 2] = 1.0
        B = rng.standard_normal(m)
        if m == 19:
            B[:5] = 0.0
        xa, xb = np.zeros(m), np.zeros(m)
        ra = o.lib.orc_lu_solve(A.ctypes.data_as(dp), B.ctypes.data_as(dp), xa.ctypes.data_as(dp), m)
        rb = o.levmar.dAx_eq_b_LU_noLapack(A.copy().ctypes.data_as(dp), B.copy().ctypes.data_as(dp), xb.ctypes.data_as(dp), m)
        assert ra == rb == 1 and np.array_equal(bits(xa), bits(xb))
    for n in (1, 7, 8, 9, 64, 515):
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        ea, eb = np.zeros(n), np.zeros(n)
        sa = o.lib.orc_residual_sqnorm(ea.ctypes.data_as(dp), x.ctypes.data_as(dp), y.ctypes.data_as(dp), n)
        sb = o.levmar.dlevmar_L2nrmxmy(eb.ctypes.data_as(dp), x.ctypes.data_as(dp), y.ctypes.data_as(dp), n)
        assert bits(np.array([sa]))[0] == bits(np.array([sb]))[0] and np.array_equal(bits(ea), bits(eb))
    for n, m in ((70, 19), (33, 8), (130, 5)):
        a = rng.standard_normal((n, m))
        ba, bb = np.zeros((m, m)), np.zeros((m, m))
        o.lib.orc_ata_blocked(a.ctypes.data_as(dp), ba.ctypes.data_as(dp), n, m)
        o.levmar.dlevmar_trans_mat_mat_mult(a.ctypes.data_as(dp), bb.ctypes.data_as(dp), n, m)
        assert np.array_equal(bits(ba), bits(bb))


def test_fit_recovers_generating_parameters():
    b = synth.make_stereo19_batch(8, K=64, seed=25, noise=0.0)
    r = oracle().fit_batch(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS,
                           t=np.tile(np.linspace(0, 1, 64), 32))
    assert np.abs(r["p"] - b["truth"]).max() < 1e-6 and (r["info"][:, 1] < 1e-12).all()


def test_postfit_normalisation_matches_oracle():
    rng = np.random.default_rng(4)
    p = rng.uniform(-2, 2, (64, 19))
    p[:16, 18] = np.abs(p[:16, 18])
    p[16:32, 16] = -2.0 - rng.uniform(0, 3, 16)
    p[32:40, 17] = 7.0
    exp = p.copy()
    for row in exp:
        oracle().lib.orc_stereo19_postfit(row.ctypes.data_as(dp))
    assert np.array_equal(bits(lm.postfit_normalise(p)), bits(exp))


def _oracle_init(meas, counts, mode, sign):
    p = np.zeros(19)
    m = np.ascontiguousarray(meas, dtype=np.float64)
    c = np.ascontiguousarray(counts, dtype=np.int32)
    oracle().lib.orc_stereo19_init_guess(m.ctypes.data_as(dp), c.ctypes.data_as(C.POINTER(C.c_int)),
                                         p.ctypes.data_as(dp), mode, sign)
    return p


def test_init_guess_matches_cv2_transcription():
    """fit_curve's initial guess (curveFittingClass.cpp:516-651) against a numpy/cv2 transcription; the sign of the
    plane normal is implementation-defined in the reference, so either orientation of the oracle must match."""
    g = np.load(os.path.join(HERE, "golden", "lm_init_guess.npz"))
    for f in range(len(g["off"]) - 1):
        meas = g["meas"][g["off"][f]:g["off"][f + 1]]
        cands = [_oracle_init(meas, g["counts"][f], MATH_LIBM, s) for s in (1, -1)]
        err = [np.abs(c - g["p"][f]).max() / np.abs(g["p"][f]).max() for c in cands]
        assert min(err) < 1e-9, (f, err)
        assert cands[0][18] <= 0.0  # sign = +1 is the orientation with a non-positive height parameter
        det = _oracle_init(meas, g["counts"][f], MATH_DET, 1)
        assert np.abs(det - cands[0]).max() < 1e-12


def test_cleanup_edges_hand_case():
    """cleanup_and_group_edges (curveFittingClass.cpp:1207-1418) on a hand-made frame: the left view of curve 0 starts 3
    rows lower and the right view ends 2 rows higher -> both are trimmed to the common rows; a sequence shorter than 10
    points makes the reference return false."""
    def seq(x0, y_start, n):
        rows = np.arange(y_start, y_start - n, -1)
        return np.stack([np.full(n, x0), rows], axis=1).astype(np.int32)
    segs = [seq(50, 203, 150), seq(150, 200, 150), seq(40, 200, 149), seq(140, 200, 150)]  # L0 L1 R0 R1
    pts = np.concatenate(segs)
    so = np.concatenate([[0], np.cumsum([len(s) for s in segs])]).astype(np.int64)
    r = oracle().cleanup_edges(pts, so, np.array([[0.0, 0.0]]), reset=True)
    assert r["valid"][0] == 1
    # L0 loses its 3 lowest rows (starts at 200 like R0); R0's end is pulled down to L0's; every sequence is then cut
    # to LENGTH_EACH_CURVE = 100 rows below its start: rows 200..100, 101 points each, all ending on row 100
    m = r["meas"].reshape(-1, 2)
    c = r["counts"][0]
    assert c.tolist() == [101, 101, 101, 101]
    assert m[0].tolist() == [50.0, 200.0] and m[100].tolist() == [50.0, 100.0]
    ends = [m[c[:k + 1].sum() - 1, 1] for k in range(4)]
    assert ends == [100.0] * 4
    # without slack above its end a sequence cannot slide up to meet its partner: the reference returns false
    short = [seq(50, 203, 40), seq(150, 200, 40), seq(40, 200, 39), seq(140, 200, 40)]
    so2 = np.concatenate([[0], np.cumsum([len(s) for s in short])]).astype(np.int64)
    assert oracle().cleanup_edges(np.concatenate(short), so2, np.array([[0.0, 0.0]]), reset=True)["valid"][0] == 0
    segs[1] = seq(150, 200, 9)
    pts = np.concatenate(segs)
    so = np.concatenate([[0], np.cumsum([len(s) for s in segs])]).astype(np.int64)
    assert oracle().cleanup_edges(pts, so, np.array([[0.0, 0.0]]), reset=True)["valid"][0] == 0
