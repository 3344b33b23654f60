"""CPU test: libcps.so loads and exports every symbol include/*.h declares (no compute without a GPU), and the
product package never reaches into oracle/."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        src = open(os.path.join(ROOT, "include", fn)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names.update(re.findall(r"\b(cps_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    from curvepointslam_b200 import _lib, build
    build.build()
    lib = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def test_no_device_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        return
    from curvepointslam_b200 import _lib
    h = ctypes.c_void_p()
    assert _lib.lib().cps_lm_create(ctypes.byref(h), 0) == -2  # CPS_ERR_CUDA
    _lib.lib().cps_ekf_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int]
    assert _lib.lib().cps_ekf_create(ctypes.byref(h), 0, 64) == -2


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "curvepointslam_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle_lib" not in src and "liborc" not in src, fn
                assert not re.search(r"#include[^\n]*orc_|\borc_[a-z0-9_]+\s*\(|import\s+oracle|from\s+oracle", src), fn
