"""CPU suite: the C-ABI library builds for sm_100a, loads, and exports every symbol include/epc_b200.h declares."""
import ctypes
import importlib
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "epc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(epc_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built_lib):
    lib_mod = importlib.import_module("epc-net_b200._lib")
    syms = _header_symbols()
    assert len(syms) >= 20
    raw = ctypes.CDLL(lib_mod.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), "libepc_b200.so does not export %s" % s
    assert set(syms) == set(lib_mod.PROTOTYPES), "ctypes prototypes and the header disagree"
    assert built_lib.epc_abi_version() == 1


def test_no_cpu_fallback(built_lib):
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert built_lib.epc_device_count() == 0
    models = importlib.import_module("epc-net_b200.models")
    variables = importlib.import_module("epc-net_b200.variables")
    import numpy as np
    V = variables.synthetic_variables("epc-net-l", 0)
    params = {"CLUSTER_SIZE": 64, "FEATURE_OUTPUT_DIM": 256, "KNN": 20, "INPUT_DIM": 3,
              "VARIABLES": variables.VariableStore(V)}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        models.load("epc-net-l").forward(np.zeros((1, 1, 64, 3), np.float32), False, params=params)
    # the C ABI itself refuses host pointers / missing devices instead of computing on the CPU
    buf = (ctypes.c_float * (64 * 3))()
    out = (ctypes.c_int32 * (64 * 20))()
    ws = (ctypes.c_char * 65536)()
    rc = built_lib.epc_knn(ctypes.cast(buf, ctypes.c_void_p), 1, 64, 0, ctypes.cast(out, ctypes.c_void_p), None, None,
                           ctypes.cast(ws, ctypes.c_void_p), 65536, None)
    assert rc < 0 and built_lib.epc_last_error()


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|import_module\(\s*['\"]oracle", re.M)
    for dp, _, fs in os.walk(os.path.join(ROOT, "epc-net_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dp, f)).read()), os.path.join(dp, f)


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_sass_targets_sm100a_and_muladd_not_contracted(built_lib):
    """(1) the library carries sm_100a SASS only; (2) the MULADD kNN kernels keep separately rounded products and
    sums: ptxas 12.9 was seen contracting mul.rn.f32x2+add.rn.f32x2 into FFMA2 (csrc/common.cuh), which would
    silently change the neighbour sets."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from sass_stats import sass_stats
    lib_mod = importlib.import_module("epc-net_b200._lib")
    elf = subprocess.run(["cuobjdump", "-lelf", lib_mod.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", elf))
    assert archs == {"sm_100a"}, archs
    stats = sass_stats(lib_mod.LIB_PATH)
    # The canonical distance (common.cuh canon_dist<ARITH>) is evaluated where exactness matters: knn_finalize_kernel (exact
    # 20th distance + thresholded set) and knn_slow_kernel (mass ties).  Everything else in the two instantiations is
    # identical, so the instruction-count DIFFERENCE isolates it: per call site MULADD = 3 FMUL + 2 FADD + FFMA + FADD,
    # FMA = FMUL + 3 FFMA + FADD.  A contraction of the MULADD products into FFMAs would close the gap.
    for kern in ("knn_finalize_kernel", "knn_slow_kernel", "knn_public_kernel"):
        mul = [v for k, v in stats.items() if kern + "ILi0E" in k]
        fma = [v for k, v in stats.items() if kern + "ILi1E" in k]
        assert len(mul) == 1 and len(fma) == 1, kern
        mul, fma = mul[0], fma[0]
        sites = (mul["FMUL"] - fma["FMUL"]) // 2
        assert sites >= 1 and mul["FMUL"] - fma["FMUL"] == 2 * sites, (kern, dict(mul), dict(fma))
        assert fma["FFMA"] - mul["FFMA"] == 2 * sites and mul["FADD"] - fma["FADD"] == 2 * sites, (kern, dict(mul), dict(fma))
        assert mul["FMUL"] >= 3 * sites and mul.get("FFMA2", 0) == fma.get("FFMA2", 0)
