"""Build recipe for libfemb.so (hand-written sm_100a kernels + the C ABI of include/femb.h).

Plain ``nvcc`` — no torch extension machinery, no torch types in the library.  The ``.so`` is
built IN-TREE (``extendablefembase.jl_b200/libfemb.so``) so that it travels with the snapshot to
the GPU box; it is git-ignored.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfemb.so")
SOURCES = ["femb_api.cu", "femb_pattern.cu", "femb_assemble.cu", "femb_generic.cu", "femb_p1.cu", "femb_h1.cu", "femb_p2.cu", "femb_hdiv.cu", "femb_ns.cu", "femb_grid.cu", "femb_system.cu", "femb_comm.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libfemb.so)")


def _stamp() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)) + ["../../include/femb.h"]:
        p = os.path.join(CSRC, name)
        if os.path.isfile(p):
            h.update(name.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(f.replace(ROOT, "<root>") for f in NVCC_FLAGS).encode())  # (the checkout may live under another path on the GPU box)
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link libfemb.so; returns its path."""
    if os.environ.get("FEMB_LIB"):
        return os.environ["FEMB_LIB"]  # an alternative, already built library was requested
    stamp_file = os.path.join(HERE, "build", "stamp")
    stamp = _stamp()

    def fresh() -> bool:
        return os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp

    if not force and fresh():
        return LIB
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    # several ranks of one node may get here at once: one builds, the others wait on the lock and find the fresh stamp
    import fcntl

    with open(os.path.join(bdir, "lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and fresh():
            return LIB
        return _build_locked(bdir, stamp, stamp_file, verbose)


def _build_locked(bdir: str, stamp: str, stamp_file: str, verbose: bool) -> str:
    nvcc = _nvcc()
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out, file=sys.stderr)
    tmp = LIB + f".tmp{os.getpid()}"
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs, "-ldl"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp, LIB)  # never a half-written library under the final name
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
