"""Builds libcps.so (the CUDA kernels and their C ABI) in-tree with nvcc for sm_100a.

    python -m curvepointslam_b200.build            # incremental
    python -m curvepointslam_b200.build --force

The LM and gate translation units are compiled with -fmad=false: their parity contract is bit-exactness
with the reference's operation order, which a fused multiply-add would break.  The EKF covariance update
is tolerance-checked (1e-9 relative) and keeps FMA contraction.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcps.so")
OUT_LEVMAR = os.path.join(HERE, "libcps_levmar.so")
OBJ_DIR = os.path.join(HERE, "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math"]

# translation unit -> extra flags
UNITS = {
    "cps_lm_api.cu": ["-fmad=false"],
    "cps_lm_fast.cu": [],  # the optional fused-arithmetic build of the staged der kernels (cps_lm_set_arith)
    "cps_gate.cu": ["-fmad=false"],
    "cps_closest.cu": ["-fmad=false"],
    "cps_ekf.cu": [],
    "cps_peaks.cu": [],
}


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libcps.so cannot be built")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers += [os.path.join(HERE, "..", "include", f) for f in os.listdir(os.path.join(HERE, "..", "include"))]
    objs = []
    for unit, extra in UNITS.items():
        src = os.path.join(CSRC, unit)
        if not os.path.exists(src):
            continue
        obj = os.path.join(OBJ_DIR, unit.replace(".cu", ".o"))
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            subprocess.run(cmd, check=True)
        objs.append(obj)
    if force or _stale(OUT, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", OUT, *objs, "-cudart", "static"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    # libcps_levmar.so: dlevmar_der / dlevmar_dif under levmar's own names, forwarding to libcps.so
    shim_src = os.path.join(CSRC, "cps_levmar_shim.c")
    if force or _stale(OUT_LEVMAR, [shim_src, OUT] + headers):
        gcc = shutil.which("gcc") or "gcc"
        cmd = [gcc, "-O2", "-fPIC", "-shared", "-o", OUT_LEVMAR, shim_src, "-L" + HERE, "-lcps", "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
