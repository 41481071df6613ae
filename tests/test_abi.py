"""The C-ABI library loads on a CPU-only box and exports every symbol include/khronos_b200.h
declares; compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import khronos_b200 as kb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "khronos_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(khr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    names = _declared()
    assert len(names) >= 30
    L = kb._lib.lib()
    for n in names:
        assert hasattr(L, n), n
    assert sorted(kb._lib.EXPORTED_SYMBOLS) == names
    out = subprocess.run(["nm", "-D", "--defined-only", kb._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (khr_[a-z0-9_]+)", out))
    assert set(names) <= exported


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", kb._lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        return
    L = kb._lib.lib()
    d = kb._lib.GridDesc()
    d.dtype = 0
    for a in range(3):
        d.n[a] = 8
        d.dl[a] = 0.1
    d.dt, d.z_start, d.nz_local, d.rank, d.nranks = 0.05, 1, 8, 0, 1
    ctx = C.c_void_p()
    assert L.khr_ctx_create(0, C.byref(d), C.byref(ctx)) != 0
    assert b"no CPU fallback" in L.khr_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "khronos.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "khronos_oracle" not in txt and "from bridge" not in txt, f
