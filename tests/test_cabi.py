"""The C-ABI shared library loads and exports every symbol include/femb.h declares; without a GPU
it fails loudly instead of computing on the CPU.  (No compute calls here — CPU suite.)"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "femb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(femb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(libfemb):
    from femb_b200 import _lib

    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(libfemb, n), f"{n} declared in include/femb.h but not exported by libfemb.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert sorted(_lib.SIGNATURES) == names
    assert libfemb.femb_version() >= 100


def test_no_cpu_fallback(libfemb):
    """On a box without a usable B200, femb_create must fail with FEMB_ERR_CUDA and a message."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is exercised on the CPU box")
    h = ctypes.c_void_p()
    st = libfemb.femb_create(0, ctypes.byref(h))
    assert st == 2 and not h.value
    assert b"no CPU fallback" in libfemb.femb_last_error(None)
    from femb_b200 import _lib

    with pytest.raises(_lib.FembError):
        _lib.Context(0)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "extendablefembase.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import|from)\s+\S*oracle|femb_oracle|oracle/", src, flags=re.M), f"{f} uses the oracle"
