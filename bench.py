#!/usr/bin/env python
"""bench.py -- batched Bezier LM fits/s on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--variant der|dif] [--fits B] [--impl b200|reference]

A step = one pass of the hot path over one batch of synthetic stereo contours: B fits of the m=19 model
(4 contour segments x 64 points each, n=512), i.e. 4*B contours.  Default B = 262144 fits = 1Mi contours per
GPU (BASELINE.json configs[1], the top of the 1k-1M sweep); inputs (1.07 GB) are larger than L2.
`value` times the kernel path with inputs resident in HBM; `e2e` times the C-ABI host-buffer call
(pinned host memory, H2D and D2H inside).  --impl reference times the reference's own levmar (oracle/_ref,
built from /root/reference) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_PTS = 64
M = 19
N_MEAS = 8 * K_PTS


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="der", choices=["der", "dif"])
    ap.add_argument("--fits", type=int, default=262144, help="fits per GPU per step (4 contours each)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --fits per GPU; strong: --contours in total, split over the GPUs (BASELINE configs[4])")
    ap.add_argument("--contours", type=int, default=8388608, help="strong scaling: contours of the whole sequence (4 per fit)")
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def pixel_batch(B, seed, first):
    """The bench's contours: synthetic samples rounded to integer pixel coordinates -- what the reference's detector
    hands to fit_curve (CvPoint sequences, featureDetector.cpp:213-288).  Every arm (resident, end to end, CPU) fits the
    same values; the end-to-end arm ships them as the int16 pairs they are."""
    from curvepointslam_b200 import synth
    b = synth.make_stereo19_batch_fast(B, K=K_PTS, seed=seed, first=first)
    b["meas"] = np.rint(b["meas"])
    assert np.abs(b["meas"]).max() < 32767
    return b


def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path: levmar-2.5 compiled unchanged (oracle/_ref) driven
    by the restated model, all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import DER, DIF, STEREO19, oracle
    from curvepointslam_b200 import lm, synth
    o = oracle()
    use_ref = o.ref is not None
    cores = os.cpu_count() or 1
    variant = DER if args.variant == "der" else DIF
    per_fit_core_s = 0.6e-3 if variant == DER else 5.5e-3
    sample = int(max(256, min(args.fits, 1.5 * cores / per_fit_core_s)))  # about 1.5 s per step
    sample = (sample // 64) * 64
    b = pixel_batch(sample, args.seed, 0)
    opts = list(lm.REFERENCE_OPTS)
    fast = use_ref and o.ref_fast is not None  # the -O3 timing build of the same sources (oracle/Makefile)
    run = lambda: o.fit_batch(STEREO19, variant, b["p0"], b["meas"], b["off"], b["counts"], opts, nthreads=cores,
                              use_ref=use_ref, fast=fast)
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    kind = "reference" if use_ref else "port"
    line = {
        "impl": "reference", "metric": "bezier_lm_fits_per_sec", "value": val, "unit": "fits/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (integer pixel coordinates, like the reference's CvPoint contours)",
        "config": dict(workload_config(args, args.fits), reference_arm_fits_per_step=sample,
                       reference_arm_note=f"a rate metric: the CPU arm runs {sample} fits of the same per-fit workload per "
                                          "step (bounded sample), not the GPU arm's batch"),
        "cpu_baseline": {"value": val, "unit": "fits/s", "cores": cores, "kind": kind,
                         "build": ("-O3 -funroll-loops -march=x86-64-v3 timing build" if fast
                                   else "-O2 -ffp-contract=off checker build"),
                         "sample": f"{sample} fits per step of the same synthetic workload, levmar-2.5 "
                                   f"dlevmar_{args.variant} built from /root/reference (oracle/_ref) + restated model "
                                   "(multiply-based Bernstein weights: faster than the reference's pow() path)"
                         if use_ref else f"{sample} fits per step, oracle port (oracle/_ref not built)"},
        "e2e": {"value": val, "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, fits):
    return {"workload": f"batched Bezier LM fit, m=19 stereo-pair model, {fits} fits/GPU/step = {4 * fits} contours "
                        f"x {K_PTS} pts (n={N_MEAS}), levmar dlevmar_{args.variant}, opts of curveFittingClass.cpp:661",
            "fits_per_gpu": fits, "contours_per_gpu": 4 * fits, "points_per_contour": K_PTS, "variant": args.variant,
            "l2": "inputs larger than L2 (no flush needed)" if fits * N_MEAS * 8 > 126e6 else "L2 flushed between steps"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from curvepointslam_b200 import _lib, flops, lm, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    variant = lm.DER if args.variant == "der" else lm.DIF
    B = args.fits
    first_fit = rank * B
    if args.scaling == "strong":  # a fixed sequence of contours split over the GPUs: contiguous ranges of the fit index
        from curvepointslam_b200 import shard
        lo, hi = shard.shard_range(args.contours // 4, rank, world)
        B, first_fit = hi - lo, lo
    ctx = lm.LmContext(local)
    ctx.set_profiling(True)  # CUDA-event pairs around the staged shape's kernels (about 130 events per step)

    # ---- synthetic shard of this rank (weak scaling: B fits per GPU) ---------------------------------
    b = pixel_batch(B, args.seed, first_fit)
    h_meas = torch.from_numpy(b["meas"]).pin_memory()
    h_pix = torch.from_numpy(b["meas"].astype(np.int16)).pin_memory()  # what the end-to-end arm ships
    h_p0 = torch.from_numpy(b["p0"]).pin_memory()
    h_off = torch.from_numpy(b["off"]).pin_memory()
    h_cnt = torch.from_numpy(b["counts"]).pin_memory()
    d_meas, d_p0, d_off, d_cnt = (x.to(dev) for x in (h_meas, h_p0, h_off, h_cnt))
    d_p = torch.empty_like(d_p0)
    d_info = torch.zeros((B, 10), dtype=torch.float64, device=dev)
    d_ret = torch.zeros(B, dtype=torch.int32, device=dev)
    d_extra = torch.zeros((B, 2), dtype=torch.int32, device=dev)
    # the path's only exchange: the fitted control points go to rank 0 (gather to root), on a side stream, from a
    # double-buffered copy, so that the transfer of step i rides under the kernels of step i + 1; the timed region
    # ends only when the last gather has landed
    Bmax = B
    if world > 1:
        tb = torch.tensor([B], dtype=torch.int64, device=dev)
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        Bmax = int(tb.item())
    out_bufs = [torch.zeros((Bmax, M), dtype=torch.float64, device=dev) for _ in range(2)] if world > 1 else None
    gathered = [torch.empty((Bmax, M), dtype=torch.float64, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    gather_events = []
    flush = None
    if B * N_MEAS * 8 <= 126e6:
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)  # every launch, copy and event of the timed region is on this stream
    torch.cuda.set_stream(stream)

    def step_resident(ev=None):
        if flush is not None:
            flush.zero_()
        d_p.copy_(d_p0)
        if ev:
            ev[0].record(stream)
        ctx.fit_batched_dev(M, variant, B, d_p.data_ptr(), d_meas.data_ptr(), None, d_off.data_ptr(),
                            d_cnt.data_ptr(), N_MEAS, d_info.data_ptr(), d_ret.data_ptr(), d_extra.data_ptr(),
                            stream=stream.cuda_stream)
        if ev:
            ev[1].record(stream)
        if world > 1:
            gather_async(d_p)

    step_no = [0]

    def gather_async(src):
        i = step_no[0]
        step_no[0] += 1
        buf = out_bufs[i % 2]
        if len(gather_events) >= 2:
            stream.wait_event(gather_events[-2])  # the gather that last read this buffer has finished
        buf[:B].copy_(src, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(stream)
        comm.wait_event(ready)
        with torch.cuda.stream(comm):
            dist.gather(buf, gathered, dst=0)
            ev = torch.cuda.Event()
            ev.record(comm)
        gather_events.append(ev)

    def gathers_join():
        if world > 1 and gather_events:
            stream.wait_event(gather_events[-1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM ------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ctx.get_profile(reset=True)
    ctx.get_profile_work(reset=True)
    launches_before = ctx.launch_stats()["launches"]
    e0.record(stream)
    for i in range(args.steps):
        step_resident(kev[i])
    gathers_join()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    kernel_ms = [a.elapsed_time(bb) for a, bb in kev]
    launches_value = ctx.launch_stats()["launches"] - launches_before  # kernels of libcps.so launched in the timed region
    prof = ctx.get_profile(reset=True)  # per-kernel CUDA-event time over the same timed region
    jac_evals, tail_fits = ctx.get_profile_work(reset=True)  # work of staged_jac_kernel / fits the tail kernel took
    staged = prof["jac"][1] > 0

    # ---- e2e: host buffers through the C ABI (H2D + kernel + D2H inside) -----------------------------
    h_p = torch.empty_like(h_p0)
    h_info = torch.zeros((B, 10), dtype=torch.float64).pin_memory()
    h_ret = torch.zeros(B, dtype=torch.int32).pin_memory()
    h_p = h_p.pin_memory()
    fn = _lib.lib().cps_dlevmar_batched_pix
    opts = np.array(lm.REFERENCE_OPTS)

    def step_e2e():
        rc = fn(ctx._h, M, variant, B, h_p.data_ptr(), h_pix.data_ptr(), 1, h_off.data_ptr(), h_cnt.data_ptr(),
                lm.REFERENCE_ITMAX, opts.ctypes.data, h_info.data_ptr(), h_ret.data_ptr(), None)
        _lib.check(rc, "cps_dlevmar_batched_pix")
        if world > 1:
            d_p.copy_(h_p, non_blocking=True)
            gather_async(d_p)
            gathers_join()
            torch.cuda.current_stream().synchronize()

    for _ in range(2):
        h_p.copy_(h_p0)
        step_e2e()
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        h_p.copy_(h_p0)  # p is in/out: the harness restores the initial guesses outside the timed region
        barrier()
        t0 = time.perf_counter()
        step_e2e()
        barrier()
        e2e_s += time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks
    tt = torch.tensor([ms_total, e2e_s * 1e3, sum(kernel_ms) / len(kernel_ms)], dtype=torch.float64, device=dev)
    total_fits = torch.tensor([B], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(total_fits, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms, kernel_ms_avg = (float(x) for x in tt.cpu())
    fits_all = int(total_fits.item())  # fits of the whole job per step

    # ---- roofline numerators from the per-fit counters of the last resident step --------------------
    info = d_info.cpu().numpy()
    extra = d_extra.cpu().numpy()
    nfev, njev, nlss = info[:, 7].sum(), info[:, 8].sum(), info[:, 9].sum()
    if variant == lm.DER:
        fl = flops.fit_flops_der(N_MEAS, M, nfev, njev, nlss)
    else:
        fl = flops.fit_flops_dif(N_MEAS, M, nfev, njev, nlss, extra[:, 0].sum(), extra[:, 1].sum())
    reasons, rc_counts = np.unique(info[:, 6], return_counts=True)

    gate_sharded = None
    if world > 1 and not args.no_extras:
        try:
            import bench_extras
            gate_sharded = bench_extras.gate_sharded_bench(local, rank, world)
        except Exception as ex:
            gate_sharded = {"error": repr(ex)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    flops_jac = njev * (flops.f_jac(N_MEAS, M) + flops.f_jtj(N_MEAS, M) + flops.f_jte(N_MEAS, M))
    pk_fma, pk_ma = _lib.C.c_double(), _lib.C.c_double()
    _lib.check(_lib.lib().cps_measure_fp64_peak(local, _lib.C.byref(pk_fma), _lib.C.byref(pk_ma)), "fp64 peak")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    achieved_tf = fl / (kernel_ms_avg * 1e-3) / 1e12
    value = fits_all * args.steps / (ms_total * 1e-3)
    e2e_val = fits_all * args.steps / (e2e_ms * 1e-3)
    h2d = h_pix.numel() * 2 + h_p0.numel() * 8 + h_off.numel() * 8 + h_cnt.numel() * 4
    d2h = h_p.numel() * 8 + h_info.numel() * 8 + h_ret.numel() * 4
    peak_note = ("measured live by cps_measure_fp64_peak (DFMA stream; MEASURED_PEAKS.json has no FP64 entry); a kernel "
                 f"that may not fuse multiply-adds (bit-exact parity) tops out at {pk_ma.value:.2f} TFLOP/s")
    hbm_part = {"achieved": flops.fit_bytes(N_MEAS, M) * B / (kernel_ms_avg * 1e-3) / 1e9, "peak": hbm_peak,
                "unit": "GB/s", "note": "algorithmic bytes of a whole fit over the step; the path is FP64-bound, not HBM-bound"}
    if staged:
        # dominant kernel of the staged shape: staged_jac_kernel (Jacobian + J^T J + J^T e of every accepted step)
        # Jacobian evaluations the kernel itself performed over all timed steps (counted on the device): the last few
        # thousand fits of a step finish in the warp-per-fit tail kernel, whose evaluations are not credited here
        jac_ms, jac_n = prof["jac"]
        per_eval = flops.f_jac(N_MEAS, M) + flops.f_jtj(N_MEAS, M) + flops.f_jte(N_MEAS, M)
        flops_jac = jac_evals * per_eval / args.steps
        njev_jac = jac_evals / args.steps
        jac_tf = flops_jac * args.steps / (jac_ms * 1e-3) / 1e12  # the profile covers all timed steps
        tot_prof = sum(v[0] for v in prof.values())
        roofline = {
            "bound": "fp64", "kernel": "staged_jac_kernel", "achieved": jac_tf, "peak": pk_fma.value, "unit": "TFLOP/s",
            "frac": jac_tf / pk_fma.value, "peak_source": peak_note, "peak_unfused": pk_ma.value,
            "frac_of_unfused": jac_tf / pk_ma.value,
            "executed": {"achieved": njev_jac * flops.staged_jac_executed_flops(N_MEAS) * args.steps / (jac_ms * 1e-3) / 1e12,
                         "frac_of_unfused": njev_jac * flops.staged_jac_executed_flops(N_MEAS) * args.steps
                         / (jac_ms * 1e-3) / 1e12 / pk_ma.value,
                         "note": "operations the kernel really issues (structural zeros of the Jacobian skipped, "
                                 "no FMA): the algorithmic count above is the dense reference algorithm's"},
            "flops_per_launch": flops_jac * args.steps / max(jac_n, 1), "launches": int(jac_n),
            "avg_launch_ms": jac_ms / max(jac_n, 1),
            "jacobian_evaluations_per_step": {"staged_jac_kernel": njev_jac, "all (sum of info[8])": float(njev)},
            "tail_fits_per_step": tail_fits / args.steps,
            "kernel_share_of_step": {k: v[0] / tot_prof for k, v in prof.items()},
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
            "launches_per_step": {k: v[1] / args.steps for k, v in prof.items()},
            "whole_step": {"achieved": achieved_tf, "frac": achieved_tf / pk_fma.value,
                           "frac_of_unfused": achieved_tf / pk_ma.value, "flops_per_fit": fl / B,
                           "step_ms": kernel_ms_avg,
                           "note": "all kernels of a step (solve + eval + jac rounds): every algorithmic flop of the fits "
                                   "over the device time of the step"},
            "hbm": hbm_part,
            # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of staged_jac_kernel at
            # 262144 fits (profiles/r2_final_ncu_full_staged_jac_details.csv: 2.423 GB + 1.015 GB), per launch
            "traffic": 3.438e9 * (B / 262144.0),
            "traffic_source": "ncu, not this run (a profiler cannot run inside a timed bench): dram__bytes_read.sum + "
                              "dram__bytes_write.sum of the round-2 `ncu --set full` capture of staged_jac_kernel at 262144 fits "
                              "(profiles/r2_final_ncu_full_staged_jac_details.csv, scripts/gpu_r2_profiles.sh), scaled by the "
                              "batch size",
            "traffic_note": "ncu: 2.42 GB read + 1.01 GB written per full-batch launch of staged_jac_kernel at 262144 "
                            "fits (samples 1.6 GB + J^T J / work-area traffic); algorithmic 0.4 GB of J^T J output + 1.6 "
                            "GB of samples",
        }
    else:
        roofline = {
            "bound": "fp64", "kernel": f"lm_fit_kernel<19,{'true' if variant == lm.DIF else 'false'}>",
            "achieved": achieved_tf, "peak": pk_fma.value, "unit": "TFLOP/s", "frac": achieved_tf / pk_fma.value,
            "peak_source": peak_note, "peak_unfused": pk_ma.value, "frac_of_unfused": achieved_tf / pk_ma.value,
            "flops_per_fit": fl / B, "kernel_ms": kernel_ms_avg, "hbm": hbm_part,
            # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel
            # (profiles/r1_ncu_full_lm_der_v2_details.csv: 148.9 MB for 32768 fits), scaled to this launch
            "traffic": 4544.0 * B if variant == lm.DER else None,
            "traffic_note": "ncu-measured 4544 B/fit vs 4496 B/fit algorithmic (der)",
        }
    line = {
        "metric": "bezier_lm_fits_per_sec", "value": value, "unit": "fits/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic (integer pixel coordinates, like the reference's CvPoint contours)",
        "config": dict(workload_config(args, B), exchange="gather of the fitted control points to rank 0 (NCCL), "
                       "asynchronous on a side stream: step i's gather rides under step i+1's kernels; the timed region "
                       "ends when the last gather has landed", total_fits_per_step=fits_all),
        "contours_per_sec": 4 * value,
        "e2e": {"value": e2e_val, "unit": "fits/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "bytes_scope": "per rank (every rank copies its own shard; the job moves n_gpus times as much)",
                "api": f"cps_dlevmar_batched_pix, variant {args.variant} (include/cps_lm.h): the contours' integer pixel "
                       "coordinates as int16 pairs (the double entry cps_dlevmar_*_batched moves 4x the bytes), pinned host buffers, "
                       "the batch streams in while the fits that have arrived iterate"},
        "execution_shape": "staged (solve / eval / jac phase kernels, thread per fit; the last fits finish in the warp-per-fit kernel)" if staged
                           else "monolithic (one persistent kernel, warp per fit)",
        "gpu_launches": launches_value,
        "clocks": clocks,
        "roofline": roofline,
        "fit_stats": {"iters_mean": float(info[:, 5].mean()), "nfev_mean": float(nfev / B),
                      "njev_mean": float(njev / B), "nlss_mean": float(nlss / B),
                      "stop_reasons": {str(int(r)): int(c) for r, c in zip(reasons, rc_counts)}},
    }

    if not args.no_cpu_baseline and world == 1:  # the CPU leg runs at N=1 only
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle_lib import DER, DIF, STEREO19, oracle
        o = oracle()
        use_ref = o.ref is not None
        cores = os.cpu_count() or 1
        per_fit_core_s = 0.6e-3 if variant == lm.DER else 5.5e-3
        sample = int(max(256, min(B, 12.0 * cores / per_fit_core_s)))
        sample = (sample // 64) * 64
        sl = slice(0, sample * N_MEAS)
        t0 = time.perf_counter()
        cpu = o.fit_batch(STEREO19, DER if variant == lm.DER else DIF, b["p0"][:sample], b["meas"][sl],
                          b["off"][:sample + 1], b["counts"][:sample], list(lm.REFERENCE_OPTS), nthreads=cores,
                          use_ref=use_ref)
        dt = time.perf_counter() - t0
        dt_fast = None
        if use_ref and o.ref_fast is not None:  # the same sources at -O3: the number nobody can call sandbagged
            t0 = time.perf_counter()
            o.fit_batch(STEREO19, DER if variant == lm.DER else DIF, b["p0"][:sample], b["meas"][sl], b["off"][:sample + 1],
                        b["counts"][:sample], list(lm.REFERENCE_OPTS), nthreads=cores, use_ref=True, fast=True)
            dt_fast = time.perf_counter() - t0
        # the same sample must agree bit for bit with what the GPU produced for those fits
        gp = d_p.cpu().numpy()[:sample]
        same = bool(np.array_equal(gp.view(np.int64), cpu["p"].view(np.int64)))
        # parity beyond bit-identity with the deterministic-math checker: the distance to the reference's own libm
        # calls (glibc sin/cos/pow exactly where modelfunc calls them; oracle LIBM mode, bit-identical to the reference's
        # compiled modelfunc / fit_curve: tests/test_ref_pinning.py) on a sample of the same fits
        from oracle_lib import MATH_LIBM
        ps = min(sample, 4096)
        lib = o.fit_batch(STEREO19, DER if variant == lm.DER else DIF, b["p0"][:ps], b["meas"][:ps * N_MEAS], b["off"][:ps + 1],
                          b["counts"][:ps], list(lm.REFERENCE_OPTS), nthreads=cores, use_ref=use_ref, math_mode=MATH_LIBM)
        gi = info[:ps]
        prel = np.abs(gp[:ps] - lib["p"]).max(axis=1) / np.abs(lib["p"]).max(axis=1)
        erel = np.abs(gi[:, 1] - lib["info"][:, 1]) / np.maximum(lib["info"][:, 1], 1e-300)
        line["parity"] = {
            "vs_checker": {"what": "reference levmar-2.5 + restated model in the device's deterministic math (DET)",
                           "fits": sample, "bit_identical_params": same},
            "vs_reference_libm": {"what": "reference levmar-2.5 + the reference's libm calls (LIBM mode)", "fits": ps,
                                  "params_within_1e-9": float((prel < 1e-9).mean()), "params_within_1e-10": float((prel < 1e-10).mean()),
                                  "params_rel_median": float(np.median(prel)), "params_rel_max": float(prel.max()),
                                  "sumsq_within_1e-10": float((erel < 1e-10).mean()), "sumsq_rel_max": float(erel.max()),
                                  "same_stop_reason": float((gi[:, 6] == lib["info"][:, 6]).mean()),
                                  "note": "SURVEY D7: two CPU builds of the reference itself agree within 1e-9 on 93.5 % of "
                                          "der fits (31.5 % of dif fits)"}}
        line["cpu_baseline"] = {"value": sample / (dt_fast if dt_fast else dt), "unit": "fits/s", "cores": cores,
                                "build": ("-O3 -funroll-loops -march=x86-64-v3 timing build of the same sources"
                                          if dt_fast else "-O2 -ffp-contract=off checker build"),
                                "value_checker_build_O2": sample / dt,
                                "kind": "reference" if use_ref else "port",
                                "sample": f"first {sample} fits of rank 0's batch, "
                                          + ("levmar-2.5 from /root/reference (oracle/_ref) + restated model"
                                             if use_ref else "oracle port"),
                                "seconds": dt, "gpu_bit_identical_on_sample": same}
    if gate_sharded is not None:
        line["extras"] = {"gate_sharded": gate_sharded}
    if not args.no_extras and world == 1:
        try:
            import bench_extras
            line["extras"] = bench_extras.run(local)
        except Exception as ex:  # extras never take the headline down
            line["extras"] = {"error": repr(ex)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
