"""GPU parity tests of the staged (phase-split, thread-per-fit) shape of dlevmar_der on Stereo19 batches
(csrc/cps_lm_staged.cuh): bit-exact against the CPU oracle and against the monolithic kernel, on dense, ragged,
mixed-block, tiny and degenerate inputs."""
import numpy as np
import pytest

from curvepointslam_b200 import lm, synth
from oracle_lib import DER, STEREO19, oracle
from test_lm_gpu import OPTS, _bits, assert_same_fit

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sctx():
    c = lm.LmContext(0)
    c.set_path(lm.PATH_STAGED)
    yield c
    c.close()


def test_staged_dense_bit_exact(sctx):
    b = synth.make_stereo19_batch(300, K=64, seed=61)
    gpu = sctx.fit_batched(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    cpu = oracle().fit_batch(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS, nthreads=8)
    assert_same_fit(gpu, cpu)
    assert (gpu["info"][:, 6] == 2).all()


def test_staged_ragged_and_mixed_blocks(sctx):
    """Segment boundaries that fall inside 32-row blocks, a block holding rows of curve 1, 2 and 1 again (a 3-point
    middle segment), the unblocked small-problem loop (n*m < 1024) and the first size above the switch."""
    rng = np.random.default_rng(7)
    counts = rng.integers(5, 102, size=(96, 4)).astype(np.int32)
    counts[0] = (101, 101, 101, 101)
    counts[1] = (3, 2, 3, 2)      # unblocked loop
    counts[2] = (7, 7, 6, 7)      # n*m = 1026
    counts[3] = (21, 3, 22, 30)   # rows of curve 1, curve 2, curve 1 inside one 32-row block
    counts[4] = (10, 2, 2, 50)    # two tiny middle segments in one block
    counts[5] = (16, 16, 16, 16)  # every boundary on a block edge
    counts[6] = (17, 15, 33, 31)
    counts[7] = (2, 40, 2, 40)
    b = synth.make_stereo19_batch(96, seed=62, counts=counts, round_pixels=True)
    gpu = sctx.fit_batched(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    cpu = oracle().fit_batch(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS, nthreads=8)
    assert_same_fit(gpu, cpu)


def test_staged_explicit_t_opts_itmax(sctx):
    b = synth.make_stereo19_batch(32, K=40, seed=63)
    t = np.tile(np.linspace(0.0, 1.0, 40), 32 * 4)
    gpu = sctx.fit_batched(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS, t=t)
    cpu = oracle().fit_batch(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS, t=t, nthreads=8)
    assert_same_fit(gpu, cpu)
    for opts, itmax in [(None, 1000), (OPTS, 3), (OPTS, 1), (OPTS, 0), ([1e-3, 1e-2, 1e-15, 1e-20, 1e-6], 1000),
                        ([1e-3, 1e-15, 1e-15, 1e+3, 1e-6], 1000)]:
        gpu = sctx.fit_batched(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], opts, itmax=itmax)
        cpu = oracle().fit_batch(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], opts, itmax=itmax)
        assert_same_fit(gpu, cpu, check_mu=itmax > 0)


def test_staged_edge_cases(sctx):
    b = synth.make_stereo19_batch(4, seed=16, counts=np.array([[2, 2, 2, 2], [30, 30, 30, 30], [30, 30, 30, 30],
                                                               [1, 1, 1, 1]], dtype=np.int32))
    meas = b["meas"].copy()
    o2 = b["off"][2]
    meas[o2 + 1: o2 + 60: 2] = 200.0  # segment 0 of fit 2 becomes horizontal: t = NaN -> stop 7
    gpu = sctx.fit_batched(STEREO19, DER, b["p0"], meas, b["off"], b["counts"], OPTS)
    cpu = oracle().fit_batch(STEREO19, DER, b["p0"], meas, b["off"], b["counts"], OPTS)
    assert np.array_equal(gpu["ret"], cpu["ret"])
    assert gpu["ret"][0] == -1 and gpu["ret"][3] == -1 and gpu["ret"][2] == -1 and gpu["info"][2, 6] == 7
    assert np.array_equal(_bits(gpu["p"][[1]]), _bits(cpu["p"][[1]]))
    assert np.array_equal(_bits(gpu["info"][[1]]), _bits(cpu["info"][[1]]))
    assert np.array_equal(gpu["info"][2, [5, 6, 7, 8, 9]], cpu["info"][2, [5, 6, 7, 8, 9]])
    assert np.array_equal(_bits(gpu["p"][[0, 3]]), _bits(b["p0"][[0, 3]]))
    # inconsistent segment sizes and an oversized fit get the documented codes, like the monolithic kernel
    bad = synth.make_stereo19_batch(3, K=20, seed=64)
    counts = bad["counts"].copy()
    counts[1, 0] += 1
    mono = lm.LmContext(0)
    mono.set_path(lm.PATH_MONOLITHIC)
    g1 = sctx.fit_batched(STEREO19, DER, bad["p0"], bad["meas"], bad["off"], counts, OPTS)
    g2 = mono.fit_batched(STEREO19, DER, bad["p0"], bad["meas"], bad["off"], counts, OPTS)
    assert g1["ret"][1] == -3 and np.array_equal(g1["ret"], g2["ret"])
    assert np.array_equal(_bits(g1["p"]), _bits(g2["p"])) and np.array_equal(_bits(g1["info"]), _bits(g2["info"]))
    mono.close()


def test_staged_equals_monolithic_at_scale(sctx):
    """40 000 fits: every fit of the staged path is bit-identical to the monolithic kernel's (which the other tests
    pin to the oracle), a random sample is checked against the oracle directly, and AUTO picks the staged shape."""
    B = 40000
    b = synth.make_stereo19_batch_fast(B, K=64, seed=65)
    mono = lm.LmContext(0)
    mono.set_path(lm.PATH_MONOLITHIC)
    g1 = sctx.fit_batched(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    g2 = mono.fit_batched(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    assert np.array_equal(g1["ret"], g2["ret"]) and np.array_equal(g1["extra"], g2["extra"])
    assert np.array_equal(_bits(g1["p"]), _bits(g2["p"])) and np.array_equal(_bits(g1["info"]), _bits(g2["info"]))
    idx = np.random.default_rng(1).choice(B, 64, replace=False)
    sub_off = np.arange(65, dtype=np.int64) * 512
    sub_meas = np.concatenate([b["meas"][b["off"][i]:b["off"][i + 1]] for i in idx])
    cpu = oracle().fit_batch(STEREO19, DER, b["p0"][idx], sub_meas, sub_off, b["counts"][idx], OPTS, nthreads=8)
    assert np.array_equal(_bits(g1["p"][idx]), _bits(cpu["p"])) and np.array_equal(_bits(g1["info"][idx]), _bits(cpu["info"]))
    auto = lm.LmContext(0)
    g3 = auto.fit_batched(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    assert np.array_equal(_bits(g3["p"]), _bits(g1["p"]))
    assert auto.launch_stats()["launches"] > 3  # more than one kernel per chunk: the staged rounds ran
    mono.close()
    auto.close()


def test_staged_device_entry_and_profile(sctx):
    """cps_lm_fit_batched_dev with buffers resident in HBM (torch tensors) through the staged shape: bit-identical to
    the host-buffer entry, and the per-kernel profile of the rounds is filled."""
    import torch
    B = 24000
    b = synth.make_stereo19_batch_fast(B, K=64, seed=66)
    dev = torch.device("cuda", 0)
    d_meas, d_p, d_off, d_cnt = (torch.from_numpy(b[k]).to(dev) for k in ("meas", "p0", "off", "counts"))
    d_info = torch.zeros((B, 10), dtype=torch.float64, device=dev)
    d_ret = torch.zeros(B, dtype=torch.int32, device=dev)
    d_extra = torch.zeros((B, 2), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    sctx.set_profiling(True)
    sctx.get_profile(reset=True)
    sctx.fit_batched_dev(19, lm.DER, B, d_p.data_ptr(), d_meas.data_ptr(), None, d_off.data_ptr(), d_cnt.data_ptr(), 512,
                         d_info.data_ptr(), d_ret.data_ptr(), d_extra.data_ptr(), opts=OPTS)
    sctx.sync()
    torch.cuda.synchronize()
    prof = sctx.get_profile(reset=True)
    sctx.set_profiling(False)
    assert prof["jac"][1] >= 5 and prof["solve"][1] >= 5 and prof["eval"][1] >= 5 and prof["jac"][0] > 0.0
    host = sctx.fit_batched(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    assert np.array_equal(_bits(d_p.cpu().numpy()), _bits(host["p"]))
    assert np.array_equal(_bits(d_info.cpu().numpy()), _bits(host["info"]))
    assert np.array_equal(d_ret.cpu().numpy(), host["ret"]) and np.array_equal(d_extra.cpu().numpy(), host["extra"])


def test_staged_ragged_large_batch_both_jac_variants(sctx):
    """The ragged / mixed-block batch replicated 250 times (24 000 fits): the early rounds run the one-thread-per-fit
    J^T J kernel, the late rounds the two-threads-per-fit variant; every replica must equal the 96-fit batch (which
    runs the two-thread variant throughout and is pinned to the oracle by test_staged_ragged_and_mixed_blocks)."""
    rng = np.random.default_rng(7)
    counts = rng.integers(5, 102, size=(96, 4)).astype(np.int32)
    counts[0] = (101, 101, 101, 101)
    counts[1] = (3, 2, 3, 2)
    counts[2] = (7, 7, 6, 7)
    counts[3] = (21, 3, 22, 30)
    counts[4] = (10, 2, 2, 50)
    counts[5] = (16, 16, 16, 16)
    counts[6] = (17, 15, 33, 31)
    counts[7] = (2, 40, 2, 40)
    b = synth.make_stereo19_batch(96, seed=62, counts=counts, round_pixels=True)
    small = sctx.fit_batched(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    R = 250
    meas = np.tile(b["meas"], R)
    n_per = np.diff(b["off"])
    off = np.concatenate([[0], np.cumsum(np.tile(n_per, R))]).astype(np.int64)
    big = sctx.fit_batched(STEREO19, DER, np.tile(b["p0"], (R, 1)), meas, off, np.tile(b["counts"], (R, 1)), OPTS)
    assert np.array_equal(big["ret"], np.tile(small["ret"], R))
    assert np.array_equal(_bits(big["p"]), _bits(np.tile(small["p"], (R, 1))))
    assert np.array_equal(_bits(big["info"]), _bits(np.tile(small["info"], (R, 1))))


@pytest.mark.parametrize("K", [64, 63, 4, 5, 100])
def test_image8_thread_per_fit_bit_exact(sctx, K):
    """m = 8 der through the thread-per-fit kernel (cps_lm_image8.cuh): bit-exact against the oracle on the blocked
    path, on the reference's switch (K = 64: n*m = 1024), below it (unblocked loop) and on ragged tails."""
    from oracle_lib import IMAGE8
    b = synth.make_image8_batch(200, K=K, seed=71)
    gpu = sctx.fit_batched(IMAGE8, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    st = sctx.launch_stats()
    assert (st["grid"], st["block"]) == (2, 128), "the host entry did not reach the thread-per-fit kernel"
    cpu = oracle().fit_batch(IMAGE8, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS, nthreads=8)
    assert_same_fit(gpu, cpu)
    if K == 64:
        t = np.tile(np.linspace(0.0, 1.0, K), 200)
        for opts, itmax in [(None, 1000), (OPTS, 1), (OPTS, 0)]:
            gpu = sctx.fit_batched(IMAGE8, DER, b["p0"], b["meas"], b["off"], b["counts"], opts, itmax=itmax, t=t)
            cpu = oracle().fit_batch(IMAGE8, DER, b["p0"], b["meas"], b["off"], b["counts"], opts, itmax=itmax, t=t)
            assert_same_fit(gpu, cpu, check_mu=itmax > 0)


@pytest.mark.parametrize("K", [64, 65, 63, 4, 5, 100])
def test_image8_dif_thread_per_fit_bit_exact(sctx, K):
    """m = 8 dif (secant Jacobian) through the thread-per-fit kernel: bit-exact against the oracle on the unblocked
    loop (n*m <= 1024: up to K = 64), the blocked one, ragged tails; forward and central differences; explicit t."""
    from oracle_lib import DIF, IMAGE8
    b = synth.make_image8_batch(200, K=K, seed=72)
    gpu = sctx.fit_batched(IMAGE8, DIF, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    st = sctx.launch_stats()
    assert (st["grid"], st["block"]) == (2, 128), "the host entry did not reach the thread-per-fit kernel"
    cpu = oracle().fit_batch(IMAGE8, DIF, b["p0"], b["meas"], b["off"], b["counts"], OPTS, nthreads=8)
    assert_same_fit(gpu, cpu)
    assert (cpu["info"][:, 8] >= 1).all()
    if K in (64, 100):
        central = list(OPTS)
        central[4] = -central[4]
        t = np.tile(np.linspace(0.0, 1.0, K), 200)
        for opts, itmax in [(central, 1000), (None, 1000), (OPTS, 1), (OPTS, 0), (OPTS, 12)]:
            gpu = sctx.fit_batched(IMAGE8, DIF, b["p0"], b["meas"], b["off"], b["counts"], opts, itmax=itmax, t=t)
            cpu = oracle().fit_batch(IMAGE8, DIF, b["p0"], b["meas"], b["off"], b["counts"], opts, itmax=itmax, t=t)
            assert_same_fit(gpu, cpu, check_mu=itmax > 0)


def test_staged_ragged_with_explicit_t(sctx):
    """Caller-supplied t on a ragged batch: fits whose t stream starts on an odd double (the staging tiles then hold
    the row one element in) must give the same bits as t derived on the device and as the oracle."""
    rng = np.random.default_rng(11)
    counts = rng.integers(5, 60, size=(64, 4)).astype(np.int32)
    counts[0] = (7, 8, 9, 11)  # 35 samples: the next fit's t starts on an odd index
    b = synth.make_stereo19_batch(64, seed=67, counts=counts, round_pixels=False)
    t = synth.t_from_rows(b)
    assert any((int(o) // 2) % 2 == 1 for o in b["off"][:-1])
    gpu_t = sctx.fit_batched(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS, t=t)
    gpu = sctx.fit_batched(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    cpu = oracle().fit_batch(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS, t=t, nthreads=8)
    assert_same_fit(gpu_t, cpu)
    assert np.array_equal(_bits(gpu_t["p"]), _bits(gpu["p"])) and np.array_equal(_bits(gpu_t["info"]), _bits(gpu["info"]))
    from oracle_lib import IMAGE8
    b8 = synth.make_image8_batch(50, K=33, seed=68)  # 33 samples per fit: every second fit starts on an odd t index
    t8 = np.tile(np.linspace(0.0, 1.0, 33), 50)
    g8 = sctx.fit_batched(IMAGE8, DER, b8["p0"], b8["meas"], b8["off"], b8["counts"], OPTS, t=t8)
    c8 = oracle().fit_batch(IMAGE8, DER, b8["p0"], b8["meas"], b8["off"], b8["counts"], OPTS, t=t8, nthreads=8)
    assert_same_fit(g8, c8)


def test_staged_device_entry_unaligned_samples(sctx):
    """Device buffers whose samples start 8 bytes off a 16-byte boundary (a view into a larger allocation): the 16-byte
    staging copies shift the tile row by one element; results must not change."""
    import torch
    B = 400
    b = synth.make_stereo19_batch_fast(B, K=64, seed=69)
    dev = torch.device("cuda", 0)
    big = torch.zeros(b["meas"].size + 1, dtype=torch.float64, device=dev)
    big[1:] = torch.from_numpy(b["meas"]).to(dev)
    assert (big.data_ptr() + 8) % 16 == 8
    d_p, d_off, d_cnt = (torch.from_numpy(b[k]).to(dev) for k in ("p0", "off", "counts"))
    d_info = torch.zeros((B, 10), dtype=torch.float64, device=dev)
    d_ret = torch.zeros(B, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    sctx.fit_batched_dev(19, lm.DER, B, d_p.data_ptr(), big.data_ptr() + 8, None, d_off.data_ptr(), d_cnt.data_ptr(), 512,
                         d_info.data_ptr(), d_ret.data_ptr(), None, opts=OPTS)
    sctx.sync()
    torch.cuda.synchronize()
    host = sctx.fit_batched(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    assert np.array_equal(_bits(d_p.cpu().numpy()), _bits(host["p"]))
    assert np.array_equal(_bits(d_info.cpu().numpy()), _bits(host["info"]))


def test_tail_takeover_is_bit_identical(monkeypatch):
    """The last fits of a staged batch are handed to the warp-per-fit kernel in mid-iteration (their damping state
    travels, the normal equations are rebuilt from p).  With the hand-over forced after the first round, late
    (default) and off, the results must be the same bits, and the oracle's."""
    rng = np.random.default_rng(3)
    counts = rng.integers(5, 80, size=(512, 4)).astype(np.int32)
    b = synth.make_stereo19_batch(512, seed=91, counts=counts, round_pixels=False)
    p0 = b["p0"].copy()
    p0[::7] += rng.normal(0, 0.4, size=p0[::7].shape)  # poor starts: rejected steps, long damping sequences
    cpu = oracle().fit_batch(STEREO19, DER, p0, b["meas"], b["off"], b["counts"], OPTS, nthreads=8)
    outs = []
    for tail in ("0", "100000", None):
        if tail is None:
            monkeypatch.delenv("CPS_LM_STAGED_TAIL", raising=False)
        else:
            monkeypatch.setenv("CPS_LM_STAGED_TAIL", tail)
        c = lm.LmContext(0)
        c.set_path(lm.PATH_STAGED)
        c.set_profiling(True)
        try:
            for itmax in (1000, 3):
                r = c.fit_batched(STEREO19, DER, p0, b["meas"], b["off"], b["counts"], OPTS, itmax=itmax)
                if itmax == 1000:
                    assert_same_fit(r, cpu)
                outs.append((tail, itmax, r))
            taken = c.get_profile_work()[1]
            assert (taken == 0) == (tail == "0")
        finally:
            c.close()
    for tail, itmax, r in outs[2:]:
        ref = outs[0 if itmax == 1000 else 1][2]
        assert np.array_equal(_bits(r["p"]), _bits(ref["p"])) and np.array_equal(_bits(r["info"]), _bits(ref["info"]))
        assert np.array_equal(r["ret"], ref["ret"]) and np.array_equal(r["extra"], ref["extra"])


def test_block_arrow_solve_and_hand_over(monkeypatch):
    """The staged solve works on the block-arrow form of J^T J + mu I (zero blocks not stored) and hands a system to
    the dense solve when the reference's pivot search would leave the structure.  Default, every third system forced
    through the hand-over, and the dense solve for everything: the same bits, and the oracle's (which is the
    reference's dAx_eq_b_LU_noLapack bit for bit, tests/test_oracle_lm.py)."""
    rng = np.random.default_rng(17)
    counts = rng.integers(5, 90, size=(700, 4)).astype(np.int32)
    b = synth.make_stereo19_batch(700, seed=93, counts=counts, round_pixels=True)
    p0 = b["p0"].copy()
    p0[::5] += rng.normal(0, 0.5, size=p0[::5].shape)  # poor starts: rejected steps, large mu, re-solves
    cpu = oracle().fit_batch(STEREO19, DER, p0, b["meas"], b["off"], b["counts"], OPTS, nthreads=8)
    seen = {}
    for env, val in ((None, None), ("CPS_LM_FORCE_HANDOVER", "3"), ("CPS_LM_DENSE_SOLVE", "1")):
        monkeypatch.delenv("CPS_LM_FORCE_HANDOVER", raising=False)
        monkeypatch.delenv("CPS_LM_DENSE_SOLVE", raising=False)
        monkeypatch.setenv("CPS_LM_STAGED_TAIL", "0")  # every solve of the batch in the staged kernels
        if env:
            monkeypatch.setenv(env, val)
        c = lm.LmContext(0)
        c.set_path(lm.PATH_STAGED)
        try:
            r = c.fit_batched(STEREO19, DER, p0, b["meas"], b["off"], b["counts"], OPTS)
            assert_same_fit(r, cpu)
            seen[env] = c.solve_handovers()
        finally:
            c.close()
    # these poor starts do produce systems whose pivot leaves a curve's own rows (about 8 % of the solves; none on the
    # benchmark batches): the hand-over is exercised even without forcing it
    assert seen["CPS_LM_DENSE_SOLVE"] == 0 and seen[None] > 0 and seen["CPS_LM_FORCE_HANDOVER"] > seen[None], seen


@pytest.mark.parametrize("variant", ["der", "dif"])
def test_asynchronous_device_entry_honours_its_stream(variant):
    """cps_lm_set_async: cps_lm_fit_batched_dev returns at once for the staged shapes too (their round loop runs on the
    context's worker thread) and the caller's stream is made to wait on the device, so work enqueued on that stream
    before and after the call is ordered around the fits: the initial guesses are written by a copy enqueued BEFORE the
    call, the results are read by a copy enqueued AFTER it, nothing synchronises in between, and the bits are those of
    the synchronous call."""
    import time
    import torch
    var = lm.DER if variant == "der" else lm.DIF
    B = 40000 if variant == "der" else 6000
    b = synth.make_stereo19_batch_fast(B, K=64, seed=68)
    dev = torch.device("cuda", 0)
    d_meas, d_off, d_cnt = (torch.from_numpy(b[k]).to(dev) for k in ("meas", "off", "counts"))
    h_p0 = torch.from_numpy(b["p0"]).pin_memory()
    outs = {}
    for mode in ("sync", "async"):
        c = lm.LmContext(0)
        c.set_path(lm.PATH_STAGED)
        c.set_async(mode == "async")
        d_p = torch.zeros((B, 19), dtype=torch.float64, device=dev)
        d_info = torch.zeros((B, 10), dtype=torch.float64, device=dev)
        d_ret = torch.zeros(B, dtype=torch.int32, device=dev)
        h_out = torch.zeros((B, 19), dtype=torch.float64).pin_memory()
        h_info = torch.zeros((B, 10), dtype=torch.float64).pin_memory()
        st = torch.cuda.Stream(device=dev)
        torch.cuda.synchronize()
        with torch.cuda.stream(st):
            d_p.copy_(h_p0, non_blocking=True)  # before the call, on the caller's stream
            t0 = time.perf_counter()
            c.fit_batched_dev(19, var, B, d_p.data_ptr(), d_meas.data_ptr(), None, d_off.data_ptr(), d_cnt.data_ptr(), 512,
                              d_info.data_ptr(), d_ret.data_ptr(), None, opts=OPTS, stream=st.cuda_stream)
            host_s = time.perf_counter() - t0
            h_out.copy_(d_p, non_blocking=True)  # after the call, on the caller's stream
            h_info.copy_(d_info, non_blocking=True)
            done = torch.cuda.Event()
            done.record(st)
        done.synchronize()  # only the caller's own event is waited for
        outs[mode] = (h_out.numpy().copy(), h_info.numpy().copy(), host_s)
        c.sync()
        c.close()
    assert np.array_equal(_bits(outs["sync"][0]), _bits(outs["async"][0]))
    assert np.array_equal(_bits(outs["sync"][1]), _bits(outs["async"][1]))
    assert (outs["async"][1][:, 6] == 2).all()
    # the synchronous call blocks for the whole fit; the asynchronous one only queues it
    assert outs["async"][2] < 0.25 * outs["sync"][2], (outs["async"][2], outs["sync"][2])


def test_fused_arithmetic_mode_distribution():
    """cps_lm_set_arith(CPS_ARITH_FMA): the staged der kernels with fused multiply-adds.  Not the reference's operation
    sequence, so not bit-identical -- but as close to the exact results as two CPU builds of the reference are to each
    other (93.5 % of der fits within 1e-9): the same stop reason and ||e||^2 within 1e-10 on every fit,
    parameters within 1e-9 on at least 90 % of the fits and never further than 1e-6.  The default mode is untouched."""
    B = 20000
    b = synth.make_stereo19_batch_fast(B, K=64, seed=69)
    c = lm.LmContext(0)
    c.set_path(lm.PATH_STAGED)
    try:
        exact = c.fit_batched(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
        c.set_arith(True)
        fast = c.fit_batched(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
        c.set_arith(False)
        again = c.fit_batched(STEREO19, DER, b["p0"], b["meas"], b["off"], b["counts"], OPTS)
    finally:
        c.close()
    assert np.array_equal(_bits(exact["p"]), _bits(again["p"]))  # switching back restores the exact path
    assert not np.array_equal(_bits(exact["p"]), _bits(fast["p"]))  # (the fused kernels did run)
    rel = np.abs(fast["p"] - exact["p"]).max(axis=1) / np.abs(exact["p"]).max(axis=1)
    e_rel = np.abs(fast["info"][:, 1] - exact["info"][:, 1]) / exact["info"][:, 1]
    assert (fast["info"][:, 6] == exact["info"][:, 6]).all()
    assert e_rel.max() < 1e-10
    assert (rel < 1e-9).mean() >= 0.90 and rel.max() < 1e-6, ((rel < 1e-9).mean(), rel.max())


@pytest.mark.parametrize("variant", [DER, 1])
def test_length_ordered_list_entry_is_invisible(monkeypatch, variant):
    """Fits of different lengths enter the staged lists ordered by their number of row blocks (a device counting sort
    over 1 024-fit chunks: csrc/cps_lm_staged.cuh, staged_bucket_*) so that the lanes of a warp walk fits of similar
    length; per-fit results do not depend on list order.  3 000 fits of 4 x 5..101 points (three chunks, every bin from
    2 to 26 row blocks, some fits refused): CPS_LM_BUCKETS=0 (index order) and =1 give the same bits, and the oracle's."""
    rng = np.random.default_rng(23)
    counts = rng.integers(5, 102, size=(3000, 4)).astype(np.int32)
    counts[7] = (1, 1, 1, 1)      # n < m: refused by the INIT pass whatever the order
    counts[2048] = (3, 2, 3, 2)   # the unblocked small-problem loop
    b = synth.make_stereo19_batch(3000, seed=64, counts=counts, round_pixels=True)
    out = []
    for flag in ("0", "1"):
        monkeypatch.setenv("CPS_LM_BUCKETS", flag)
        c = lm.LmContext(0)
        c.set_path(lm.PATH_STAGED)
        try:
            out.append(c.fit_batched(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], OPTS, itmax=40))
        finally:
            c.close()
    assert_same_fit(out[0], out[1])
    assert out[1]["ret"][7] == -1
    sample = np.arange(0, 3000, 7)
    sub = synth.make_stereo19_batch(3000, seed=64, counts=counts, round_pixels=True)
    off = sub["off"]
    meas = np.concatenate([sub["meas"][off[f]:off[f + 1]] for f in sample])
    off_s = np.concatenate([[0], np.cumsum([off[f + 1] - off[f] for f in sample])]).astype(np.int64)
    cpu = oracle().fit_batch(STEREO19, variant, sub["p0"][sample], meas, off_s, counts[sample], OPTS, itmax=40, nthreads=8)
    got = {k: out[1][k][sample] for k in ("p", "info", "ret", "extra")}
    assert_same_fit(got, cpu)
