"""GPU parity tests of the batched LM fits: the CUDA path (through the C ABI of include/cps_lm.h) against
the CPU oracle on the same seeded inputs.  Bar: bit-exact parameters, info[] and return codes (the kernels
perform the reference's operation sequence, see DESIGN.md "Bit-exactness")."""
import numpy as np
import pytest

from curvepointslam_b200 import lm, synth
from oracle_lib import DER, DIF, IMAGE8, STEREO19, oracle

pytestmark = pytest.mark.gpu

OPTS = list(lm.REFERENCE_OPTS)


@pytest.fixture(scope="module")
def ctx():
    c = lm.LmContext(0)
    yield c
    c.close()


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


def assert_same_fit(gpu, cpu, check_mu=True):
    assert np.array_equal(gpu["ret"], cpu["ret"]), "return codes differ"
    gi, ci = gpu["info"].copy(), cpu["info"].copy()
    if not check_mu:
        gi[:, 4] = 0
        ci[:, 4] = 0
    bad = np.nonzero((_bits(gi) != _bits(ci)).any(axis=1) | (_bits(gpu["p"]) != _bits(cpu["p"])).any(axis=1))[0]
    assert bad.size == 0, (f"{bad.size} fits differ, first {bad[:5]}: info gpu {gpu['info'][bad[0]]} cpu "
                           f"{cpu['info'][bad[0]]}; max |dp| {np.abs(gpu['p'][bad[0]] - cpu['p'][bad[0]]).max()}")
    assert np.array_equal(gpu["extra"], cpu["extra"])


@pytest.mark.parametrize("variant", [DER, DIF])
def test_stereo19_dense_bit_exact(ctx, variant):
    b = synth.make_stereo19_batch(96, K=64, seed=11)
    gpu = ctx.fit_batched(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    cpu = oracle().fit_batch(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], OPTS, nthreads=8)
    assert_same_fit(gpu, cpu)
    assert (gpu["info"][:, 6] == 2).all()  # the survey's probe: reason 2 in 100 %


@pytest.mark.parametrize("variant", [DER, DIF])
def test_stereo19_ragged_bit_exact(ctx, variant):
    rng = np.random.default_rng(5)
    counts = rng.integers(5, 102, size=(48, 4)).astype(np.int32)  # up to LENGTH_EACH_CURVE+1 points (common.h:112)
    counts[0] = (101, 101, 101, 101)
    counts[1] = (3, 2, 3, 2)  # n = 20 >= m: the unblocked J^T J loop (n*m < 1024)
    counts[2] = (7, 7, 6, 7)  # n*m = 1026, just above the switch
    b = synth.make_stereo19_batch(48, seed=12, counts=counts, round_pixels=True)
    gpu = ctx.fit_batched(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    cpu = oracle().fit_batch(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], OPTS, nthreads=8)
    assert_same_fit(gpu, cpu)


@pytest.mark.parametrize("variant", [DER, DIF])
def test_stereo19_long_contours_bit_exact(ctx, variant):
    """Contours as long as a frame is tall (the reference takes up to 10 000 points per segment,
    curveFittingClass.cpp:1383-1396): n = 3 840 and a ragged neighbour, far beyond the 1 280 measurements the
    warp-per-fit der kernel's compact item list holds -- such batches take the staged kernels whatever their size."""
    counts = np.array([[480, 480, 480, 480], [161, 160, 161, 160], [300, 17, 222, 401], [64, 64, 64, 64]], dtype=np.int32)
    b = synth.make_stereo19_batch(4, seed=27, counts=counts, round_pixels=True)
    gpu = ctx.fit_batched(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    cpu = oracle().fit_batch(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], OPTS, nthreads=4)
    assert (gpu["ret"] >= 0).all()
    assert_same_fit(gpu, cpu)


@pytest.mark.parametrize("variant", [DER, DIF])
@pytest.mark.parametrize("dtype", [np.int32, np.int16])
def test_integer_pixel_samples_equal_the_double_entry(ctx, variant, dtype):
    """cps_dlevmar_batched_pix: CvPoint-style integer samples cross the bus as 4 / 2 bytes per coordinate; results are
    the bits of the double entry on the same values (and so the oracle's), on a ragged batch (warp-per-fit kernel) and
    on one large enough for the staged kernels' pipelined host entry."""
    rng = np.random.default_rng(9)
    counts = rng.integers(5, 102, size=(64, 4)).astype(np.int32)
    b = synth.make_stereo19_batch(64, seed=33, counts=counts, round_pixels=True)
    pix = b["meas"].astype(dtype)
    assert np.array_equal(pix.astype(np.float64), b["meas"])
    a = ctx.fit_batched_pix(STEREO19, variant, b["p0"], pix, b["off"], b["counts"], OPTS)
    c = ctx.fit_batched(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    cpu = oracle().fit_batch(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], OPTS, nthreads=8)
    assert_same_fit(a, c)
    assert_same_fit(a, cpu)
    if variant == DER:
        big = synth.make_stereo19_batch_fast(9000, K=64, seed=34)
        big["meas"] = np.rint(big["meas"])
        a = ctx.fit_batched_pix(STEREO19, DER, big["p0"], big["meas"].astype(dtype), big["off"], big["counts"], OPTS)
        c = ctx.fit_batched(STEREO19, DER, big["p0"], big["meas"], big["off"], big["counts"], OPTS)
        assert_same_fit(a, c)


@pytest.mark.parametrize("variant", [DER, DIF])
def test_stereo19_explicit_t(ctx, variant):
    b = synth.make_stereo19_batch(32, K=40, seed=13)
    t = np.tile(np.linspace(0.0, 1.0, 40), 32 * 4)
    gpu = ctx.fit_batched(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], OPTS, t=t)
    cpu = oracle().fit_batch(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], OPTS, t=t, nthreads=8)
    assert_same_fit(gpu, cpu)
    # t derived on the device from the image rows must equal the host restatement of fit_curve:451-511
    gpu2 = ctx.fit_batched(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], OPTS,
                           t=synth.t_from_rows(b))
    gpu3 = ctx.fit_batched(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    assert np.array_equal(_bits(gpu2["p"]), _bits(gpu3["p"]))


@pytest.mark.parametrize("variant", [DER, DIF])
@pytest.mark.parametrize("K", [64, 63, 4, 5])
def test_image8_bit_exact(ctx, variant, K):
    # K=64: n*m = 1024 sits exactly on the reference's switch: "<" in der (lm_core.c:193), "<=" in dif (:585)
    b = synth.make_image8_batch(64, K=K, seed=14)
    gpu = ctx.fit_batched(IMAGE8, variant, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    cpu = oracle().fit_batch(IMAGE8, variant, b["p0"], b["meas"], b["off"], b["counts"], OPTS, nthreads=8)
    assert_same_fit(gpu, cpu)


def test_default_opts_central_differences_and_itmax(ctx):
    b = synth.make_stereo19_batch(16, K=32, seed=15)
    for variant, opts, itmax in [(DER, None, 1000), (DIF, None, 1000), (DIF, [1e-3, 1e-12, 1e-12, 1e-20, -1e-6], 1000),
                                 (DER, OPTS, 3), (DIF, OPTS, 7), (DIF, OPTS, 0)]:
        gpu = ctx.fit_batched(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], opts, itmax=itmax)
        cpu = oracle().fit_batch(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], opts, itmax=itmax)
        assert_same_fit(gpu, cpu, check_mu=itmax > 0)
        if itmax in (0, 3, 7):
            assert (gpu["info"][:, 6] == 3).all()


def test_edge_cases(ctx):
    """n < m -> LM_ERROR (lm_core.c:117,493); a horizontal segment makes t NaN (fit_curve:451-511 divides
    by y_last - y_first) -> stop reason 7, return -1; an empty batch is a no-op."""
    b = synth.make_stereo19_batch(4, seed=16, counts=np.array([[2, 2, 2, 2], [30, 30, 30, 30], [30, 30, 30, 30],
                                                               [1, 1, 1, 1]], dtype=np.int32))
    meas = b["meas"].copy()
    o2 = b["off"][2]
    meas[o2 + 1: o2 + 60: 2] = 200.0  # segment 0 of fit 2 becomes horizontal
    for variant in (DER, DIF):
        gpu = ctx.fit_batched(STEREO19, variant, b["p0"], meas, b["off"], b["counts"], OPTS)
        cpu = oracle().fit_batch(STEREO19, variant, b["p0"], meas, b["off"], b["counts"], OPTS)
        assert gpu["ret"][0] == -1 and gpu["ret"][3] == -1 and gpu["ret"][2] == -1
        assert gpu["info"][2, 6] == 7
        assert np.array_equal(gpu["ret"], cpu["ret"])
        ok = [1]
        assert np.array_equal(_bits(gpu["p"][ok]), _bits(cpu["p"][ok]))
        assert np.array_equal(_bits(gpu["info"][ok]), _bits(cpu["info"][ok]))
        assert np.array_equal(gpu["info"][2, [5, 6, 7, 8, 9]], cpu["info"][2, [5, 6, 7, 8, 9]])
        assert np.array_equal(_bits(gpu["p"][[0, 3]]), _bits(b["p0"][[0, 3]]))  # untouched
    empty = ctx.fit_batched(STEREO19, DER, np.zeros((0, 19)), np.zeros(0), np.zeros(1, dtype=np.int64),
                            np.zeros((0, 4), dtype=np.int32), OPTS)
    assert empty["p"].shape == (0, 19)


def test_large_batch_properties(ctx):
    """At bench size the oracle is too slow to check every fit; use size-independent properties: a random
    sample of fits is bit-exact, every fit terminates for reason 2, residuals never increase, and the result
    does not depend on how the batch is split."""
    B = 20000
    b = synth.make_stereo19_batch_fast(B, K=64, seed=17)
    gpu = ctx.fit_batched(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    assert (gpu["info"][:, 6] == 2).all() and (gpu["ret"] >= 0).all()
    assert (gpu["info"][:, 1] <= gpu["info"][:, 0]).all()
    idx = np.random.default_rng(0).choice(B, 64, replace=False)
    sub_off = np.arange(65, dtype=np.int64) * 512
    sub_meas = np.concatenate([b["meas"][b["off"][i]:b["off"][i + 1]] for i in idx])
    cpu = oracle().fit_batch(STEREO19, DER, b["p0"][idx], sub_meas, sub_off, b["counts"][idx], OPTS, nthreads=8)
    assert np.array_equal(_bits(gpu["p"][idx]), _bits(cpu["p"]))
    assert np.array_equal(_bits(gpu["info"][idx]), _bits(cpu["info"]))
    half = ctx.fit_batched(STEREO19, DER, b["p0"][B // 2:], b["meas"][b["off"][B // 2]:],
                           b["off"][B // 2:] - b["off"][B // 2], b["counts"][B // 2:], OPTS)
    assert np.array_equal(_bits(half["p"]), _bits(gpu["p"][B // 2:]))


def test_same_signature_shim_and_fit_curve(ctx):
    """cps_dlevmar_dif with the reference's argument list (curveFittingClass.cpp:681) and the
    CurveFittingClass.fit_curve mirror give the batched result; unknown callbacks are refused."""
    import ctypes as C

    from curvepointslam_b200 import _lib

    b = synth.make_stereo19_batch(1, K=48, seed=18)
    L = _lib.lib()

    class MyData(C.Structure):
        _fields_ = [("orig_meas", C.c_void_p), ("t", C.c_void_p), ("curve_num", C.c_void_p), ("num_left1", C.c_int),
                    ("num_right1", C.c_int), ("num_left2", C.c_int), ("num_right2", C.c_int), ("disp", C.c_void_p)]

    t = synth.t_from_rows(b)
    meas = b["meas"].copy()
    p = b["p0"][0].copy()
    info = np.zeros(10)
    opts = np.array(OPTS)
    md = MyData(meas.ctypes.data, t.ctypes.data, None, 48, 48, 48, 48, None)
    token = C.cast(L.cps_modelfunc, C.c_void_p)
    ret = L.cps_dlevmar_dif(token, p.ctypes.data, meas.ctypes.data, 19, meas.size, 1000, opts.ctypes.data,
                            info.ctypes.data, None, None, C.addressof(md))
    ref = ctx.fit_batched(STEREO19, DIF, b["p0"], b["meas"], b["off"], b["counts"], OPTS, t=t)
    assert ret == ref["ret"][0]
    assert np.array_equal(_bits(p), _bits(ref["p"][0])) and np.array_equal(_bits(info), _bits(ref["info"][0]))
    bogus = C.cast(L.cps_last_error, C.c_void_p)
    assert L.cps_dlevmar_dif(bogus, p.ctypes.data, meas.ctypes.data, 19, meas.size, 1000, opts.ctypes.data,
                             info.ctypes.data, None, None, C.addressof(md)) == -3
    # fit_curve mirror: post-fit normalisation equals the oracle's restatement of :684-711
    fitter = lm.CurveFittingClass(ctx, variant=DIF)
    xy = b["meas"].reshape(4, 48, 2)
    out = fitter.fit_curve(b["p0"][0], [xy[0], xy[1]], [xy[2], xy[3]])
    exp = ref["p"][0].copy()
    oracle().lib.orc_stereo19_postfit(exp.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.array_equal(_bits(out), _bits(exp)) and fitter.fitting_error == ref["info"][0, 1]


def test_init_guess_bit_exact_and_feeds_the_fit(ctx):
    """fit_curve's initial guess on the device (curveFittingClass.cpp:516-651) equals the oracle bit for bit, matches
    the cv2 transcription up to the sign of the plane normal, and a fit started from it converges like the
    reference's own pipeline (stop reason 2, small residual)."""
    import ctypes as C
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lm_init_guess.npz"))
    p_gpu = ctx.init_guess(g["meas"], g["off"], g["counts"])
    dp = C.POINTER(C.c_double)
    for f in range(len(g["off"]) - 1):
        m = np.ascontiguousarray(g["meas"][g["off"][f]:g["off"][f + 1]])
        c = np.ascontiguousarray(g["counts"][f], dtype=np.int32)
        p = np.zeros(19)
        oracle().lib.orc_stereo19_init_guess(m.ctypes.data_as(dp), c.ctypes.data_as(C.POINTER(C.c_int)),
                                             p.ctypes.data_as(dp), 0, 1)
        assert np.array_equal(_bits(p_gpu[f]), _bits(p)), f
    b = synth.make_stereo19_batch(64, K=64, seed=41)
    p0 = ctx.init_guess(b["meas"], b["off"], b["counts"])
    assert np.isfinite(p0).all() and (p0[:, 18] <= 0).all()
    gpu = ctx.fit_batched(STEREO19, DIF, p0, b["meas"], b["off"], b["counts"], OPTS)
    cpu = oracle().fit_batch(STEREO19, DIF, p0, b["meas"], b["off"], b["counts"], OPTS, nthreads=8)
    assert_same_fit(gpu, cpu)
    assert (gpu["info"][:, 6] == 2).all()
    assert np.abs(gpu["p"][:, 18] - b["truth"][:, 18]).max() < 0.2
    fitter = lm.CurveFittingClass(ctx, variant=DIF)
    xy = b["meas"][:512].reshape(4, 64, 2)
    out = fitter.fit_curve(None, [xy[0], xy[1]], [xy[2], xy[3]])  # params0=None: device builds the guess
    assert np.allclose(out, lm.postfit_normalise(gpu["p"][:1])[0], rtol=0, atol=0)


@pytest.mark.parametrize("reset", [True, False])
def test_cleanup_and_group_edges_bit_exact(ctx, reset):
    """cleanup_and_group_edges on the device (integer trimming + packing) equals the CPU restatement exactly, frame
    by frame, including the frames the reference rejects."""
    pts, so, tr = synth.make_edge_frames(300, seed=51)
    g = ctx.cleanup_and_group_edges(pts, so, tr, reset_data_assoc=reset)
    o = oracle().cleanup_edges(pts, so, tr, reset)
    assert np.array_equal(g["valid"], o["valid"]) and 0 < g["valid"].sum() < 300
    assert np.array_equal(g["counts"], o["counts"]) and np.array_equal(g["off"], o["off"])
    assert np.array_equal(_bits(g["meas"]), _bits(o["meas"]))
    empty = ctx.cleanup_and_group_edges(np.zeros((0, 2), dtype=np.int32), np.zeros(1, dtype=np.int64), np.zeros((0, 2)))
    assert empty["off"].tolist() == [0]


def test_frame_pipeline_edges_to_fit(ctx):
    """The reference's per-frame chain main.cpp:425-445 -- cleanup_and_group_edges -> fit_curve (initial guess + LM)
    -- entirely through the device entry points, against the same chain of CPU restatements."""
    import ctypes as C
    pts, so, tr = synth.make_edge_frames(24, seed=52, consistent=True)
    g = ctx.cleanup_and_group_edges(pts, so, tr, reset_data_assoc=True)
    assert g["valid"].all()
    p0 = ctx.init_guess(g["meas"], g["off"], g["counts"])
    fit = ctx.fit_batched(STEREO19, DIF, p0, g["meas"], g["off"], g["counts"], OPTS)
    o = oracle().cleanup_edges(pts, so, tr, True)
    dp = C.POINTER(C.c_double)
    p0o = np.zeros((24, 19))
    for f in range(24):
        m = np.ascontiguousarray(o["meas"][o["off"][f]:o["off"][f + 1]])
        c = np.ascontiguousarray(o["counts"][f], dtype=np.int32)
        oracle().lib.orc_stereo19_init_guess(m.ctypes.data_as(dp), c.ctypes.data_as(C.POINTER(C.c_int)),
                                             p0o[f].ctypes.data_as(dp), 0, 1)
    cpu = oracle().fit_batch(STEREO19, DIF, p0o, o["meas"], o["off"], o["counts"], OPTS, nthreads=8)
    assert np.array_equal(_bits(p0), _bits(p0o))
    assert_same_fit(fit, cpu)
    assert (fit["ret"] >= 0).all()


@pytest.mark.parametrize("path", ["monolithic", "staged"])
@pytest.mark.parametrize("vname,variant", [("der", DER), ("dif", DIF)])
@pytest.mark.parametrize("tag", ["s19", "i8"])
def test_golden_vectors_from_reference_levmar_on_gpu(tag, vname, variant, path):
    """The committed vectors the REFERENCE's levmar-2.5 produced (tests/golden/gen_lm_golden.py, deterministic math
    mode) against the CUDA path, bit for bit, in both execution shapes: no oracle in between."""
    import os
    from oracle_lib import IMAGE8
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lm_reference_levmar.npz"))
    c = lm.LmContext(0)
    c.set_path(lm.PATH_MONOLITHIC if path == "monolithic" else lm.PATH_STAGED)
    try:
        r = c.fit_batched(STEREO19 if tag == "s19" else IMAGE8, variant, g[f"{tag}_p0"], g[f"{tag}_meas"],
                          g[f"{tag}_off"], g[f"{tag}_counts"], [float(v) for v in g["opts"]])
    finally:
        c.close()
    key = f"{tag}_{vname}_det"
    assert np.array_equal(r["ret"], g[key + "_ret"])
    assert np.array_equal(_bits(r["p"]), _bits(g[key + "_p"]))
    assert np.array_equal(_bits(r["info"]), _bits(g[key + "_info"]))
