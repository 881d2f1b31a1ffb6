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


def _build_c_example(tmp_path):
    import subprocess

    exe = str(tmp_path / "c_abi_poisson")
    pkg = os.path.join(ROOT, "extendablefembase.jl_b200")
    cmd = ["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_poisson.c"),
           "-L", pkg, "-lfemb", "-Wl,-rpath," + pkg, "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True)
    return subprocess.run([exe], capture_output=True, text=True)


def test_plain_c_caller_compiles_against_the_header(libfemb, tmp_path):
    """examples/c_abi_poisson.c uses nothing but include/femb.h and libfemb.so (no torch, no Python); without a GPU it
    reports the library's error and exits with its 'no device' code."""
    import torch

    r = _build_c_example(tmp_path)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 77 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_plain_c_caller_known_answer(libfemb, tmp_path):
    """the two-triangle unit square through the C ABI from C: SURVEY.md Appendix C values, hypotenuse entries filtered"""
    r = _build_c_example(tmp_path)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "nnz = 12"
    vals = {l.split("=")[0].strip(): float(l.split("=")[1]) for l in lines[1:13]}
    assert vals["A[1,1]"] == 1.0 and vals["A[2,1]"] == -0.5 and vals["A[4,4]"] == 1.0 and "A[4,1]" not in vals and "A[1,4]" not in vals
    assert lines[13].split() == ["b", "=", "0.0000", "0.0556", "-0.0556", "0.0000"]


def test_missing_nccl_is_a_status_not_a_crash(tmp_path):
    """without a loadable libnccl, femb_comm_unique_id must return FEMB_ERR_NCCL (5) — in a fresh process,
    because a process that already dlopen'ed NCCL keeps it."""
    import subprocess
    import sys

    from femb_b200 import _lib

    code = (
        "import ctypes, sys\n"
        f"lib = ctypes.CDLL({_lib.LIBPATH!r})\n"
        "buf = ctypes.create_string_buffer(128)\n"
        "lib.femb_comm_unique_id.restype = ctypes.c_int32\n"
        "print(lib.femb_comm_unique_id(buf))\n"
    )
    env = dict(os.environ, FEMB_NCCL_LIB=str(tmp_path / "no_such_libnccl.so"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "5"
